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
  S2Common c; c.W = ctx->W; c.k = (int)ctx->prm.kmer_size; c.P = P; c.records = ln->records.p; c.boff = ln->d_boff; c.kcnt = ln->d_kcnt;
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
  CK(segmented_radix_sort(P, koff.data(), nullptr, (u64*)ln->keys_lo.p, (u64*)ln->keys_hi.p, (u64*)ln->keys_lo2.p, (u64*)ln->keys_hi2.p,
                          KW, 0, end_bit, ln->sort_work.p, &in_alt, ln->st, &ln->launches)); }
  const u64* slo = (const u64*)(in_alt ? ln->keys_lo2.p : ln->keys_lo.p);
  const u64* shi = (const u64*)(in_alt ? ln->keys_hi2.p : ln->keys_hi.p);
  std::vector<u64> toff, soff;
  PROF(KMX_PROF_RLE);
  CK(rle_segments(P, koff.data(), slo, shi, KW, hard_min, ln->sort_work.p, toff, soff, 0, nullptr, nullptr, nullptr, ln->st, &ln->launches));
  const u64 D = soff[P];
  void *kp = nullptr, *hp = nullptr, *cp = nullptr;
  CK(list_alloc(ln, D * 8, &kp));
  if (KW == 2) CK(list_alloc(ln, D * 8, &hp));
  CK(list_alloc(ln, D * 4, &cp));
  CK(rle_segments(P, koff.data(), slo, shi, KW, hard_min, ln->sort_work.p, toff, soff, 1, (u64*)kp, (u64*)hp, (u32*)cp, ln->st, &ln->launches));
  for (u32 p = 0; p < P; p++) {
    ListRef& L = ctx->lists[(size_t)sample * P + p];
    L.lo = (u64*)kp + soff[p]; L.hi = hp ? (u64*)hp + soff[p] : nullptr; L.cnt = (u32*)cp + soff[p]; L.n = soff[p + 1] - soff[p];
  }
  return KMX_OK;
}

static u64 next_pow2(u64 v) { u64 p = 1; while (p < v) p <<= 1; return p; }

// k-mer keys: open-addressed hash-count in HBM (64-bit slots for k <= 32, 128-bit slots claimed by a
// 128-bit CAS for k <= 63), then sort only the distinct survivors
static int count_kmer_ht(Lane* ln, uint32_t sample, uint32_t hard_min)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const int KW = ctx->W;
  double f;
  { std::lock_guard<std::mutex> g(ctx->mu); f = ctx->ht_factor; }
  std::vector<u64> toff(P + 1, 0), tcap(P);
  u64 K = 0, max_cap = 0;
  for (u32 p = 0; p < P; p++) {
    tcap[p] = next_pow2(std::max<u64>(1024, (u64)((double)ln->h_kcnt[p] * f) + 1));
    toff[p + 1] = toff[p] + tcap[p]; K += ln->h_kcnt[p]; max_cap = std::max(max_cap, tcap[p]);
  }
  for (u32 p = 0; p < P; p++) ctx->lists[(size_t)sample * P + p] = ListRef();
  if (K == 0) return KMX_OK;
  const u64 TS = toff[P];
  if (TS >= 0xFFFFFFF0ULL) return KMX_HT_FALLBACK;
  CK(ensure(ln, ln->ht_keys, TS * 8 * KW)); CK(ensure(ln, ln->ht_cnts, TS * 4));
  CK(ensure(ln, ln->keys_lo, TS * 8)); CK(ensure(ln, ln->keys_lo2, TS * 8));
  if (KW == 2) { CK(ensure(ln, ln->keys_hi, TS * 8)); CK(ensure(ln, ln->keys_hi2, TS * 8)); }
  // device meta: u64 toff[P] | tcap[P] | oo[P] ; u32 pcnt[P] | overflow
  const size_t mb = (size_t)P * 24 + (size_t)P * 4 + 64;
  CK(ensure(ln, ln->tmp_cnt, mb));
  CK(ensure_pin(ln, mb + (size_t)P * 32 + 256));
  u64* d_toff = (u64*)ln->tmp_cnt.p; u64* d_tcap = d_toff + P; u64* d_oo = d_tcap + P;
  u32* d_pcnt = (u32*)(d_oo + P); u32* d_ovf = d_pcnt + P;
  u64* hp = (u64*)ln->h_pin;
  memcpy(hp, toff.data(), P * 8); memcpy(hp + P, tcap.data(), P * 8);
  CK(cudaMemcpyAsync(d_toff, hp, (size_t)P * 16, cudaMemcpyHostToDevice, ln->st));
  CK(cudaMemsetAsync(d_pcnt, 0, (size_t)P * 4 + 4, ln->st));
  { PROF(KMX_PROF_FILL);
    CK(cudaMemsetAsync(ln->ht_keys.p, 0xFF, TS * 8 * KW, ln->st));
    CK(cudaMemsetAsync(ln->ht_cnts.p, 0, TS * 4, ln->st)); }
  S2Common c; c.W = KW; c.k = (int)ctx->prm.kmer_size; c.P = P; c.records = ln->records.p; c.boff = ln->d_boff; c.kcnt = ln->d_kcnt;
  c.bcnt = ln->d_cursor; c.max_bcnt = *std::max_element(ln->h_cursor.begin(), ln->h_cursor.end());
  { PROF(KMX_PROF_EXPAND);
    if (KW == 1) CK(launch_ht_insert_records(c, (u64*)ln->ht_keys.p, (u32*)ln->ht_cnts.p, d_toff, d_tcap, d_ovf, ln->st, &ln->launches));
    else CK(launch_ht2_insert_records(c, ln->ht_keys.p, (u32*)ln->ht_cnts.p, d_toff, d_tcap, d_ovf, ln->st, &ln->launches)); }
  { PROF(KMX_PROF_RLE);
    if (KW == 1) CK(launch_ht_compact(P, max_cap, (const u64*)ln->ht_keys.p, (const u32*)ln->ht_cnts.p, d_toff, d_tcap, hard_min,
                                      (u64*)ln->keys_lo.p, d_pcnt, ln->st, &ln->launches));
    else CK(launch_ht2_compact(P, max_cap, ln->ht_keys.p, (const u32*)ln->ht_cnts.p, d_toff, d_tcap, hard_min,
                               (u64*)ln->keys_lo.p, (u64*)ln->keys_hi.p, d_pcnt, ln->st, &ln->launches)); }
  u32* h_pc = (u32*)(ln->h_pin + (size_t)P * 16);
  CK(cudaMemcpyAsync(h_pc, d_pcnt, (size_t)P * 4 + 4, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  if (h_pc[P]) {                                   // a table filled up: grow the factor, redo this sample on the sort path
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->ht_factor = std::min(4.0, std::max(ctx->ht_factor, f) * 2.0);
    return KMX_HT_FALLBACK;
  }
  std::vector<u32> pcnt(h_pc, h_pc + P);
  std::vector<u64> sb(P), se(P), oo(P + 1, 0);
  u32 max_n = 0;
  for (u32 p = 0; p < P; p++) { sb[p] = toff[p]; se[p] = toff[p] + pcnt[p]; oo[p + 1] = oo[p] + pcnt[p]; max_n = std::max(max_n, pcnt[p]); }
  const u64 D = oo[P];
  CK(ensure(ln, ln->sort_work, radix_sort_work_bytes(P, sb.data(), se.data())));
  int in_alt = 0;
  { PROF(KMX_PROF_SORT);
    CK(segmented_radix_sort(P, sb.data(), se.data(), (u64*)ln->keys_lo.p, KW == 2 ? (u64*)ln->keys_hi.p : nullptr, (u64*)ln->keys_lo2.p,
                            KW == 2 ? (u64*)ln->keys_hi2.p : nullptr, KW, 0, 2 * (int)ctx->prm.kmer_size, ln->sort_work.p, &in_alt, ln->st, &ln->launches)); }
  void *kp = nullptr, *khp = nullptr, *cp = nullptr;
  CK(list_alloc(ln, D * 8, &kp));
  if (KW == 2) CK(list_alloc(ln, D * 8, &khp));
  CK(list_alloc(ln, D * 4, &cp));
  memcpy(hp, oo.data(), P * 8);
  CK(cudaMemcpyAsync(d_oo, hp, (size_t)P * 8, cudaMemcpyHostToDevice, ln->st));
  { PROF(KMX_PROF_RLE);
    if (KW == 1) CK(launch_ht_lookup(P, max_n, (const u64*)ln->ht_keys.p, (const u32*)ln->ht_cnts.p, d_toff, d_tcap,
                                     (const u64*)(in_alt ? ln->keys_lo2.p : ln->keys_lo.p), d_toff, d_pcnt, d_oo, (u64*)kp, (u32*)cp, ln->st, &ln->launches));
    else CK(launch_ht2_lookup(P, max_n, ln->ht_keys.p, (const u32*)ln->ht_cnts.p, d_toff, d_tcap,
                              (const u64*)(in_alt ? ln->keys_lo2.p : ln->keys_lo.p), (const u64*)(in_alt ? ln->keys_hi2.p : ln->keys_hi.p),
                              d_toff, d_pcnt, d_oo, (u64*)kp, (u64*)khp, (u32*)cp, ln->st, &ln->launches)); }
  CK(cudaStreamSynchronize(ln->st));               // pinned staging is reused by the next call
  for (u32 p = 0; p < P; p++) {
    ListRef& L = ctx->lists[(size_t)sample * P + p];
    L.lo = (u64*)kp + oo[p]; L.hi = khp ? (u64*)khp + oo[p] : nullptr; L.cnt = (u32*)cp + oo[p]; L.n = pcnt[p];
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
  // 1. union of keys.  k <= 32: open-addressed hash SET of all N lists' keys, then sort only the distinct
  //    keys; otherwise (or if the set overflows): concatenate, sort, unique
  u64* ulo = nullptr; u64* uhi = nullptr; u64 nu = 0;
  bool have_union = false;
  // The set is sized from the largest list times a factor learned from the partitions merged so far (samples that share
  // most keys: ~1.3; samples with private keys, e.g. hard-min 1: hundreds), never beyond what tot_n entries can need; an
  // overflow raises the factor and retries, so that one dissimilar partition does not send all the later ones to the
  // concatenate + sort path.
  for (int attempt = 0; ctx->ht_union_ok && !have_union && attempt < 3; attempt++) {
    double factor;
    { std::lock_guard<std::mutex> g(ctx->mu); factor = ctx->union_factor; }
    const u64 want = (u64)(factor * (double)max_n);
    const u64 cap = next_pow2(std::max<u64>(4096, std::min<u64>(want, 2 * tot_n)));
    if (cap >= 0x7FFFFFFFULL) break;
    {
      CK(ensure(ln, ctx->uni_lo2, cap * 8 * KW));                    // table
      CK(ensure(ln, ctx->uni_lo, cap * 8));                          // distinct keys, unordered
      CK(ensure(ln, ln->keys_lo, cap * 8)); CK(ensure(ln, ln->keys_lo2, cap * 8));
      if (KW == 2) { CK(ensure(ln, ln->keys_hi, cap * 8)); CK(ensure(ln, ln->keys_hi2, cap * 8)); }
      CK(ensure(ln, ln->tmp_cnt, 64));
      u32* d_cnt = (u32*)ln->tmp_cnt.p; u32* d_ovf = d_cnt + 1;
      CK(cudaMemsetAsync(d_cnt, 0, 8, ln->st));
      CK(cudaMemsetAsync(ctx->uni_lo2.p, 0xFF, cap * 8 * KW, ln->st));
      if (KW == 1) CK(launch_ht_union((const MergeList*)ctx->d_lists.p, N, max_n, (u64*)ctx->uni_lo2.p, nullptr, cap, d_ovf, (u64*)ln->keys_lo.p, d_cnt, ln->st, &ln->launches));
      else CK(launch_ht2_union((const MergeList*)ctx->d_lists.p, N, max_n, ctx->uni_lo2.p, cap, d_ovf, (u64*)ln->keys_lo.p, (u64*)ln->keys_hi.p, d_cnt, ln->st, &ln->launches));
      u32* hres = (u32*)ln->h_pin;                                   // lists/soft staging was consumed by the copy kernel above
      { SmallCopyBatch b(ln); b.add(hres, d_cnt, 8); CK(b.go()); }
      CK(cudaStreamSynchronize(ln->st));
      if (hres[1]) {                                                 // overflow: a bigger set next time (and now, if it can still grow)
        std::lock_guard<std::mutex> g(ctx->mu);
        ctx->union_factor = std::max(ctx->union_factor, factor * 4.0);
        if (cap >= next_pow2(2 * tot_n)) break;
        continue;
      }
      {
        nu = hres[0];
        { std::lock_guard<std::mutex> g(ctx->mu); ctx->union_factor = std::max(ctx->union_factor, 2.2 * (double)nu / (double)std::max<u64>(max_n, 1)); }
        u64 seg1[2] = {0, nu};
        CK(ensure(ln, ln->sort_work, radix_sort_work_bytes(1, seg1)));
        int alt = 0;
        CK(segmented_radix_sort(1, seg1, nullptr, (u64*)ln->keys_lo.p, KW == 2 ? (u64*)ln->keys_hi.p : nullptr, (u64*)ln->keys_lo2.p,
                                KW == 2 ? (u64*)ln->keys_hi2.p : nullptr, KW, 0,
                                hash ? bit_length(ctx->prm.window_bits * ctx->prm.nb_partitions - 1) : 2 * (int)ctx->prm.kmer_size,
                                ln->sort_work.p, &alt, ln->st, &ln->launches));
        ulo = (u64*)(alt ? ln->keys_lo2.p : ln->keys_lo.p);
        if (KW == 2) uhi = (u64*)(alt ? ln->keys_hi2.p : ln->keys_hi.p);
        have_union = true;
      }
    }
  }
  if (!have_union) {
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
  CK(segmented_radix_sort(1, seg, nullptr, (u64*)ctx->uni_lo.p, (u64*)ctx->uni_hi.p, (u64*)ctx->uni_lo2.p, (u64*)ctx->uni_hi2.p,
                          KW, 0, end_bit, ln->sort_work.p, &in_alt, ln->st, &ln->launches));
  const u64* slo = (const u64*)(in_alt ? ctx->uni_lo2.p : ctx->uni_lo.p);
  const u64* shi = (const u64*)(in_alt ? ctx->uni_hi2.p : ctx->uni_hi.p);
  std::vector<u64> toff, soff;
  CK(rle_segments(1, seg, slo, shi, KW, 1, ln->sort_work.p, toff, soff, 0, nullptr, nullptr, nullptr, ln->st, &ln->launches));
  nu = soff[1];
  CK(ensure(ln, ln->keys_lo, nu * 8));
  if (KW == 2) CK(ensure(ln, ln->keys_hi, nu * 8));
  CK(ensure(ln, ln->tmp_cnt, nu * 4 + 64));
  ulo = (u64*)ln->keys_lo.p; uhi = KW == 2 ? (u64*)ln->keys_hi.p : nullptr;
  CK(rle_segments(1, seg, slo, shi, KW, 1, ln->sort_work.p, toff, soff, 1, ulo, uhi, (u32*)ln->tmp_cnt.p, ln->st, &ln->launches));
  }
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
