// kmx_generic.inl -- generic (sort-based) stage 2 and sparse-row stage 3 drivers (included by kmx_api.cu)

static int bit_length(u64 v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

static int count_generic(Lane* ln, uint32_t sample, uint32_t hard_min)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const bool hash = ctx->prm.key_kind == KMX_KEY_HASH;
  const int KW = hash ? 1 : ctx->W;               // words per key
  std::vector<u64> koff(P + 1, 0);
  for (u32 p = 0; p < P; p++) koff[p + 1] = koff[p] + ln->h_kcnt[p];
  const u64 K = koff[P];
  for (u32 p = 0; p < P; p++) ctx->lists[(size_t)sample * P + p] = ListRef();
  if (K == 0) return KMX_OK;
  if (K >= 0xFFFFFFF0ULL) return fail(ln, KMX_ERR_ARG, "sample has %llu k-mers; split it (generic count path sorts < 2^32 keys at a time)", (unsigned long long)K);
  CK(ensure(ln, ln->keys_lo, K * 8)); CK(ensure(ln, ln->keys_lo2, K * 8));
  if (KW == 2) { CK(ensure(ln, ln->keys_hi, K * 8)); CK(ensure(ln, ln->keys_hi2, K * 8)); }
  CK(ensure(ln, ln->tmp_cnt, (size_t)(P + 1) * 8 + P * 4 + 64));
  u64* d_koff = (u64*)ln->tmp_cnt.p; u32* d_kcur = (u32*)(d_koff + P + 1);
  CK(cudaMemcpyAsync(d_koff, koff.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ln->st));
  CK(cudaMemsetAsync(d_kcur, 0, P * 4, ln->st));
  S2Common c; c.W = ctx->W; c.k = (int)ctx->prm.kmer_size; c.P = P; c.records = ln->records.p; c.boff = ln->d_boff;
  c.bcnt = ln->d_cursor; c.max_bcnt = *std::max_element(ln->h_cursor.begin(), ln->h_cursor.end());
  u64 mlo = 0, mhi = 0; const u64 Wb = ctx->prm.window_bits;
  if (hash) fastmod_magic(Wb, mlo, mhi);
  { PROF(KMX_PROF_EXPAND);
  CK(launch_expand_keys(c, hash ? 1 : 0, Wb, Wb, mlo, mhi, d_koff, d_kcur, (u64*)ln->keys_lo.p, (u64*)ln->keys_hi.p, ln->st, &ln->launches)); }
  const size_t wb = std::max(radix_sort_work_bytes(P, koff.data()), rle_work_bytes(P, koff.data()));
  CK(ensure(ln, ln->sort_work, wb));
  const int end_bit = hash ? bit_length(Wb * P - 1) : 2 * (int)ctx->prm.kmer_size;
  int in_alt = 0;
  { PROF(KMX_PROF_SORT);
  CK(segmented_radix_sort(P, koff.data(), (u64*)ln->keys_lo.p, (u64*)ln->keys_hi.p, (u64*)ln->keys_lo2.p, (u64*)ln->keys_hi2.p,
                          KW, 0, end_bit, ln->sort_work.p, &in_alt, ln->st, &ln->launches)); }
  const u64* slo = (const u64*)(in_alt ? ln->keys_lo2.p : ln->keys_lo.p);
  const u64* shi = (const u64*)(in_alt ? ln->keys_hi2.p : ln->keys_hi.p);
  std::vector<u64> toff, soff;
  PROF(KMX_PROF_RLE);
  CK(rle_segments(P, koff.data(), slo, shi, KW, hard_min, ln->sort_work.p, toff, soff, 0, nullptr, nullptr, nullptr, ln->st, &ln->launches));
  const u64 D = soff[P];
  void *kp = nullptr, *hp = nullptr, *cp = nullptr;
  CK(arena_alloc(ctx, D * 8, &kp));
  if (KW == 2) CK(arena_alloc(ctx, D * 8, &hp));
  CK(arena_alloc(ctx, D * 4, &cp));
  CK(rle_segments(P, koff.data(), slo, shi, KW, hard_min, ln->sort_work.p, toff, soff, 1, (u64*)kp, (u64*)hp, (u32*)cp, ln->st, &ln->launches));
  for (u32 p = 0; p < P; p++) {
    ListRef& L = ctx->lists[(size_t)sample * P + p];
    L.lo = (u64*)kp + soff[p]; L.hi = hp ? (u64*)hp + soff[p] : nullptr; L.cnt = (u32*)cp + soff[p]; L.n = soff[p + 1] - soff[p];
  }
  return KMX_OK;
}

static int merge_sparse(Lane* ln, uint32_t partition, const kmx_merge_params* mp, kmx_merge_result* res,
                        const std::vector<MergeList>& hl, u64 max_n, u64 tot_n)
{
  kmx_ctx* ctx = ln->ctx;
  PROF(KMX_PROF_MERGE);
  const u32 N = ctx->prm.nb_samples;
  const bool hash = ctx->prm.key_kind == KMX_KEY_HASH;
  const int KW = hash ? 1 : ctx->W;
  const u32 row_bytes = 8 * KW + (mp->format == KMX_FMT_COUNT ? 4 * N : (N + 7) / 8);
  ctx->last_res.n_rows = 0; ctx->last_res.row_bytes = row_bytes; ctx->last_res.n_union = 0;
  ctx->last_body = (uint8_t*)ctx->body.p;
  if (res) *res = ctx->last_res;
  if (tot_n == 0) return KMX_OK;
  if (tot_n >= 0xFFFFFFF0ULL) return fail(ln, KMX_ERR_ARG, "partition %u holds %llu (key,sample) entries (>= 2^32)", partition, (unsigned long long)tot_n);
  // 1. union of keys: concatenate, sort, unique
  CK(ensure(ln, ctx->uni_lo, tot_n * 8)); CK(ensure(ln, ctx->uni_lo2, tot_n * 8));
  if (KW == 2) { CK(ensure(ln, ctx->uni_hi, tot_n * 8)); CK(ensure(ln, ctx->uni_hi2, tot_n * 8)); }
  u64 o = 0;
  for (u32 s = 0; s < N; s++) if (hl[s].n) {
    CK(cudaMemcpyAsync((u64*)ctx->uni_lo.p + o, hl[s].lo, hl[s].n * 8, cudaMemcpyDeviceToDevice, ln->st));
    if (KW == 2) CK(cudaMemcpyAsync((u64*)ctx->uni_hi.p + o, hl[s].hi, hl[s].n * 8, cudaMemcpyDeviceToDevice, ln->st));
    o += hl[s].n;
  }
  u64 seg[2] = {0, tot_n};
  const size_t wb = std::max(radix_sort_work_bytes(1, seg), rle_work_bytes(1, seg));
  CK(ensure(ln, ln->sort_work, wb));
  const int end_bit = hash ? bit_length(ctx->prm.window_bits * ctx->prm.nb_partitions - 1) : 2 * (int)ctx->prm.kmer_size;
  int in_alt = 0;
  CK(segmented_radix_sort(1, seg, (u64*)ctx->uni_lo.p, (u64*)ctx->uni_hi.p, (u64*)ctx->uni_lo2.p, (u64*)ctx->uni_hi2.p,
                          KW, 0, end_bit, ln->sort_work.p, &in_alt, ln->st, &ln->launches));
  const u64* slo = (const u64*)(in_alt ? ctx->uni_lo2.p : ctx->uni_lo.p);
  const u64* shi = (const u64*)(in_alt ? ctx->uni_hi2.p : ctx->uni_hi.p);
  std::vector<u64> toff, soff;
  CK(rle_segments(1, seg, slo, shi, KW, 1, ln->sort_work.p, toff, soff, 0, nullptr, nullptr, nullptr, ln->st, &ln->launches));
  const u64 nu = soff[1];
  CK(ensure(ln, ln->keys_lo, nu * 8));
  if (KW == 2) CK(ensure(ln, ln->keys_hi, nu * 8));
  CK(ensure(ln, ln->tmp_cnt, nu * 4 + 64));
  u64* ulo = (u64*)ln->keys_lo.p; u64* uhi = KW == 2 ? (u64*)ln->keys_hi.p : nullptr;
  CK(rle_segments(1, seg, slo, shi, KW, 1, ln->sort_work.p, toff, soff, 1, ulo, uhi, (u32*)ln->tmp_cnt.p, ln->st, &ln->launches));
  // 2. solid_in per row, keep flags, output row numbers
  CK(ensure(ln, ctx->solid_in, nu * 4)); CK(ensure(ln, ctx->keep, nu * 4)); CK(ensure(ln, ctx->out_row, nu * 4 + 16));
  CK(ensure(ln, ctx->scan_work, scan_u32_work_bytes(nu)));
  CK(cudaMemsetAsync(ctx->solid_in.p, 0, nu * 4, ln->st));
  CK(launch_sparse_solid((const MergeList*)ctx->d_lists.p, N, (const u32*)ctx->d_soft.p, ulo, uhi, nu, KW, (u32*)ctx->solid_in.p, max_n, ln->st, &ln->launches));
  CK(launch_row_keep((const u32*)ctx->solid_in.p, nu, mp->recurrence_min, mp->emit_all, (u32*)ctx->keep.p, ln->st, &ln->launches));
  CK(cudaMemcpyAsync(ctx->out_row.p, ctx->keep.p, nu * 4, cudaMemcpyDeviceToDevice, ln->st));
  u32* d_tot = (u32*)ctx->out_row.p + nu;
  CK(scan_u32_inplace((u32*)ctx->out_row.p, nu, d_tot, ctx->scan_work.p, ln->st, &ln->launches));
  u32 n_rows = 0;
  CK(cudaMemcpyAsync(&n_rows, d_tot, 4, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  // 3. rows
  const size_t body_bytes = (size_t)n_rows * row_bytes;
  CK(ensure(ln, ctx->body, body_bytes + 8));
  CK(cudaMemsetAsync(ctx->body.p, 0, body_bytes + 8, ln->st));
  if (mp->emit_all) CK(ensure(ln, ctx->row_keep, (size_t)n_rows + 8));
  CK(launch_sparse_emit((const MergeList*)ctx->d_lists.p, N, (const u32*)ctx->d_soft.p, mp->recurrence_min, mp->share_min, mp->emit_all,
                        ulo, uhi, nu, KW, (const u32*)ctx->solid_in.p, (const u32*)ctx->keep.p, (const u32*)ctx->out_row.p,
                        mp->format == KMX_FMT_COUNT ? 0 : 1, (uint8_t*)ctx->body.p, row_bytes, (uint8_t*)ctx->row_keep.p,
                        (u64*)ctx->stats.p, max_n, ln->st, &ln->launches));
  ctx->last_body = (uint8_t*)ctx->body.p;
  ctx->last_res.n_rows = n_rows; ctx->last_res.row_bytes = row_bytes; ctx->last_res.n_union = nu;
  if (res) *res = ctx->last_res;
  return KMX_OK;
}
