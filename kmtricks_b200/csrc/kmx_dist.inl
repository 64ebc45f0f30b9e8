// kmx_dist.inl -- multi-GPU: one process per GPU; samples shard over ranks for stage 1, partitions
// shard over ranks (contiguous blocks) for stages 2-4, and ONE exchange in between moves every
// sample's super-k-mer bucket regions to the partitions' owners: an all-to-all-v of
// bucket bytes = grouped ncclSend/ncclRecv over NVLink, preceded by a tiny ncclAllGather of the
// bucket geometry.  NCCL is resolved at run time from the libnccl.so.2 the process already has
// (torch's) -- torch.distributed only does the rendezvous (unique-id broadcast).
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static const char* nccl_load()
{
  std::lock_guard<std::mutex> g(g_nccl_mu);
  if (g_nccl.h) return nullptr;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return "libnccl.so.2 not found";
#define KMX_SYM(f, name) g_nccl.f = (decltype(g_nccl.f))dlsym(h, name); if (!g_nccl.f) return "missing NCCL symbol " name
  KMX_SYM(GetUniqueId, "ncclGetUniqueId"); KMX_SYM(CommInitRank, "ncclCommInitRank"); KMX_SYM(CommDestroy, "ncclCommDestroy");
  KMX_SYM(GroupStart, "ncclGroupStart"); KMX_SYM(GroupEnd, "ncclGroupEnd"); KMX_SYM(Send, "ncclSend"); KMX_SYM(Recv, "ncclRecv");
  KMX_SYM(AllGather, "ncclAllGather"); KMX_SYM(GetErrorString, "ncclGetErrorString");
#undef KMX_SYM
  g_nccl.h = h;
  return nullptr;
}
}  // namespace

struct KmxDist {
  int rank = 0, world = 1;
  u32 use_lanes = 0;                      // 0 = all
  u32 warmed_lanes = 0;                   // lanes whose first sample has been through alone (see kmx_dist_run_batch)
  std::vector<ncclComm_t> comms;          // one per lane
  std::vector<DBuf> recv, meta_dev;       // per lane
};

#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(ln, KMX_ERR_CUDA, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)

static inline u32 part_first(u32 P, int world, int g) { return (u32)(((u64)g * P) / (u64)world); }

extern "C" int kmx_dist_unique_id(uint8_t* out128)
{
  if (!out128) return KMX_ERR_ARG;
  if (nccl_load()) return KMX_ERR_CUDA;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return KMX_ERR_CUDA;
  memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return KMX_OK;
}

extern "C" int kmx_dist_init(kmx_ctx* ctx, int rank, int world, uint32_t nlanes, const uint8_t* ids)
{
  if (!ctx || !ids || world < 1 || rank < 0 || rank >= world || nlanes < 1 || nlanes > 8) return KMX_ERR_ARG;
  LANE0;
  if (const char* e = nccl_load()) return fail(ln, KMX_ERR_CUDA, "NCCL: %s", e);
  cudaSetDevice(ctx->device);
  while (ctx->lanes.size() < nlanes) { int rc = lane_create(ctx, (int)ctx->lanes.size()); if (rc) return rc; }
  ctx->dist.reset(new KmxDist());
  KmxDist* d = ctx->dist.get();
  d->rank = rank; d->world = world;
  d->recv.resize(nlanes); d->meta_dev.resize(nlanes);
  for (uint32_t t = 0; t < nlanes; t++) {
    ncclUniqueId id; memcpy(id.internal, ids + (size_t)t * NCCL_UNIQUE_ID_BYTES, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t c;
    NK(g_nccl.CommInitRank(&c, world, id, rank));
    d->comms.push_back(c);
  }
  return KMX_OK;
}

extern "C" int kmx_dist_set_lanes(kmx_ctx* ctx, uint32_t nlanes)
{
  if (!ctx || !ctx->dist || nlanes < 1 || nlanes > ctx->dist->comms.size()) return KMX_ERR_ARG;
  ctx->dist->use_lanes = nlanes;
  return KMX_OK;
}

extern "C" int kmx_dist_owner(const kmx_ctx* ctx, uint32_t partition, int world)
{
  if (!ctx || world < 1 || partition >= ctx->prm.nb_partitions) return -1;
  const u32 P = ctx->prm.nb_partitions;
  for (int g = 0; g < world; g++) if (partition >= part_first(P, world, g) && partition < part_first(P, world, g + 1)) return g;
  return -1;
}

// one batch on one lane: local sample -> buckets -> exchange -> count every rank's sample on my partitions
// slot of rank g's sample i_local of this batch: g * n_local (samples per rank over the whole run) + i_local (slot_base included)
static int dist_batch(Lane* ln, u32 lane_idx, u32 i_local, u32 n_local, const char* text, size_t nbytes, int on_device,
                      u32 hard_min, uint64_t* pinfo_out)
{
  kmx_ctx* ctx = ln->ctx;
  KmxDist* d = ctx->dist.get();
  const u32 P = ctx->prm.nb_partitions;
  const int G = d->world, me = d->rank;
  const size_t rec = ctx->W == 1 ? 16 : 32;
  ncclComm_t comm = d->comms[lane_idx];
  int rc = superk_begin(ln);
  if (!rc) rc = superk_push_fastq(ln, text, nbytes, on_device);
  if (!rc) rc = superk_end(ln, pinfo_out);
  // a rank that failed still takes part in the geometry all-gather (with its status), so that all ranks stop together
  // instead of the others waiting forever in the exchange
  const int rc_local = rc;
  // ---- geometry all-gather: [boff[P+1] | cursor[P] | kcnt[P] | hard_min | status] as u64
  const size_t ML = (size_t)3 * P + 3;
  CK(ensure_pin(ln, (size_t)(G + 1) * ML * 8 + 256));
  CK(ensure(ln, d->meta_dev[lane_idx], (size_t)(G + 1) * ML * 8));
  u64* hm = (u64*)ln->h_pin;
  for (u32 p = 0; p < P; p++) { hm[p] = ln->h_boff[p]; hm[P + 1 + p] = ln->h_cursor[p]; hm[2 * P + 1 + p] = ln->h_kcnt[p]; }
  hm[P] = P ? ln->h_boff[P - 1] + ln->h_bcap[P - 1] : 0;
  hm[3 * P + 1] = hard_min;
  hm[3 * P + 2] = (u64)(u32)rc_local;
  if (rc_local) for (size_t i = 0; i < (size_t)3 * P + 1; i++) hm[i] = 0;
  u64* dm = (u64*)d->meta_dev[lane_idx].p;
  { SmallCopyBatch b(ln); b.add(dm, hm, ML * 8); CK(b.go()); }
  NK(g_nccl.AllGather(dm, dm + ML, ML * 8, ncclUint8, comm, ln->st));
  u64* all = hm + ML;
  { SmallCopyBatch b(ln); b.add(all, dm + ML, (size_t)G * ML * 8); CK(b.go()); }
  CK(cudaStreamSynchronize(ln->st));
  std::vector<u64> meta(all, all + (size_t)G * ML);      // the pinned scratch is reused below
  if (rc_local) return rc_local;
  for (int g = 0; g < G; g++) if (meta[(size_t)g * ML + 3 * P + 2]) return fail(ln, KMX_ERR_STATE, "rank %d failed in stage 1 of this batch (status %u)", g, (unsigned)meta[(size_t)g * ML + 3 * P + 2]);
  // ---- payload: my slab region of g's partitions -> g ; g's region of my partitions -> me
  const u32 myf = part_first(P, G, me), myl = part_first(P, G, me + 1);
  std::vector<u64> roff(G + 1, 0);
  for (int g = 0; g < G; g++) { const u64* mg = &meta[(size_t)g * ML]; roff[g + 1] = roff[g] + (mg[myl] - mg[myf]); }
  CK(ensure(ln, d->recv[lane_idx], (size_t)((double)(roff[G] * rec) * 1.1) + 256));
  char* rbuf = (char*)d->recv[lane_idx].p;
  const u64* mm = &meta[(size_t)me * ML];
  {
    PROF(KMX_PROF_EXCHANGE);
    NK(g_nccl.GroupStart());
    for (int g = 0; g < G; g++) {
      const u32 gf = part_first(P, G, g), gl = part_first(P, G, g + 1);
      const u64 sbytes = (mm[gl] - mm[gf]) * rec, rbytes = (roff[g + 1] - roff[g]) * rec;
      if (sbytes) NK(g_nccl.Send((const char*)ln->records.p + mm[gf] * rec, sbytes, ncclUint8, g, comm, ln->st));
      if (g != me) ctx->stat[KMX_STAT_EXCH_BYTES] += sbytes;
      if (rbytes) NK(g_nccl.Recv(rbuf + roff[g] * rec, rbytes, ncclUint8, g, comm, ln->st));
    }
    NK(g_nccl.GroupEnd());
  }
  // ---- stage 2 for every rank's sample of this batch, restricted to my partitions
  DBuf save_rec = ln->records;
  std::vector<u64> save_boff = ln->h_boff, save_kcnt = ln->h_kcnt; std::vector<u32> save_bcap = ln->h_bcap, save_cur = ln->h_cursor;
  const u32 Pown = myl - myf;
  const bool binned = hash_binned_usable(ctx);
  const bool hist_path = ctx->prm.key_kind == KMX_KEY_HASH && (ctx->hist_ok == 1 || binned);
  if (hist_path && (u64)G * Pown <= P && Pown > 0) {
    // ONE pass: histogram window v = g*Pown + (p - myf) holds partition p of rank g's sample; all
    // hard-mins of a batch are equal in practice, otherwise fall through to the per-source passes
    bool same_hm = true;
    for (int g = 1; g < G; g++) if (meta[(size_t)g * ML + 3 * P + 1] != meta[3 * P + 1]) same_hm = false;
    if (same_hm) {
      std::vector<u32> wsmp(P, 0), wprt(P, 0);
      for (u32 v = 0; v < P; v++) { ln->h_boff[v] = 0; ln->h_cursor[v] = 0; ln->h_bcap[v] = 0; ln->h_kcnt[v] = 0; }
      for (int g = 0; g < G; g++) {
        const u64* mg = &meta[(size_t)g * ML];
        for (u32 p = myf; p < myl; p++) {
          const u32 v = (u32)g * Pown + (p - myf);
          ln->h_boff[v] = roff[g] + (mg[p] - mg[myf]);
          ln->h_cursor[v] = (u32)mg[P + 1 + p]; ln->h_bcap[v] = ln->h_cursor[v];
          ln->h_kcnt[v] = mg[2 * P + 1 + p];
          wsmp[v] = (u32)g * n_local + i_local; wprt[v] = p;
        }
      }
      ln->records.p = rbuf;
      rc = upload_bucket_meta(ln);
      ln->sample_ready = true;
      for (int g = 0; g < G; g++) {
        const size_t slot = (size_t)g * n_local + i_local;
        if (slot < ctx->prm.nb_samples) for (u32 p = 0; p < P; p++) ctx->lists[slot * P + p] = ListRef();
        else for (u32 p = myf; p < myl; p++) if (meta[(size_t)g * ML + P + 1 + p]) return fail(ln, KMX_ERR_ARG, "rank %d sent data for slot %zu >= nb_samples", g, slot);
      }
      if (!rc) {
        rc = binned ? count_hash_binned(ln, 0, (u32)meta[3 * P + 1], wsmp.data(), wprt.data()) : KMX_BIN_FALLBACK;
        if (rc == KMX_BIN_FALLBACK && ctx->hist_ok == 1) rc = count_hash_hist(ln, 0, (u32)meta[3 * P + 1], wsmp.data(), wprt.data());
      }
      if (rc != KMX_BIN_FALLBACK) {
        ln->records = save_rec; ln->h_boff = save_boff; ln->h_kcnt = save_kcnt; ln->h_bcap = save_bcap; ln->h_cursor = save_cur;
        return rc;
      }
      rc = KMX_OK;                                       // neither single-pass path took it: per-source passes below
    }
  }
  for (int g = 0; g < G && !rc; g++) {
    const u64* mg = &meta[(size_t)g * ML];
    for (u32 p = 0; p < P; p++) {
      const bool mine = p >= myf && p < myl;
      ln->h_boff[p] = mine ? mg[p] - mg[myf] : 0;
      ln->h_cursor[p] = mine ? (u32)mg[P + 1 + p] : 0;
      ln->h_bcap[p] = ln->h_cursor[p];
      ln->h_kcnt[p] = mine ? mg[2 * P + 1 + p] : 0;
    }
    ln->records.p = rbuf + roff[g] * rec;
    rc = upload_bucket_meta(ln);
    ln->sample_ready = true;
    if (!rc && (size_t)g * n_local + i_local < ctx->prm.nb_samples) rc = count_sample(ln, (u32)g * n_local + i_local, (u32)mg[3 * P + 1]);
  }
  ln->records = save_rec; ln->h_boff = save_boff; ln->h_kcnt = save_kcnt; ln->h_bcap = save_bcap; ln->h_cursor = save_cur;
  return rc;
}

extern "C" int kmx_dist_run_samples(kmx_ctx* ctx, uint32_t n_local, const char* const* texts, const size_t* nbytes, int on_device,
                                    const uint32_t* hard_min, uint64_t* kmers_per_partition)
{
  return kmx_dist_run_batch(ctx, n_local, texts, nbytes, on_device, hard_min, 0, n_local, kmers_per_partition);
}

extern "C" int kmx_dist_run_batch(kmx_ctx* ctx, uint32_t n_batch, const char* const* texts, const size_t* nbytes, int on_device,
                                  const uint32_t* hard_min, uint32_t slot_base, uint32_t n_local, uint64_t* kmers_per_partition)
{
  if (!ctx || !ctx->dist || (n_batch && (!texts || !nbytes || !hard_min)) || (u64)slot_base + n_batch > n_local) return KMX_ERR_ARG;
  KmxDist* d = ctx->dist.get();
  const u32 P = ctx->prm.nb_partitions;
  const u32 nlanes = d->use_lanes ? std::min<u32>(d->use_lanes, (u32)d->comms.size()) : (u32)d->comms.size();
  {
    LANE0;
    // slots >= nb_samples may only hold empty padding samples (the last rank of a run whose sample count is not a multiple of the world)
    for (u32 i = 0; i < n_batch; i++)
      if ((u64)d->rank * n_local + slot_base + i >= ctx->prm.nb_samples && nbytes[i])
        return fail(ln, KMX_ERR_ARG, "sample slot %llu >= nb_samples %u", (unsigned long long)((u64)d->rank * n_local + slot_base + i), ctx->prm.nb_samples);
    if (ctx->prm.key_kind == KMX_KEY_HASH && ctx->hist_ok < 0) {
      size_t free_b = 0, tot_b = 0;
      CK(cudaMemGetInfo(&free_b, &tot_b));
      ctx->hist_ok = (size_t)ctx->prm.window_bits * 4 * nlanes < (free_b / 4) ? 1 : 0;
    }
  }
  ctx->active_lanes = (int)nlanes;
  // First call of a context: the first sample of every lane goes through ALONE, one lane after the other, on every rank alike.
  // That is where each lane's buffers are allocated (and outgrown the first few times); allocations and frees are device-wide
  // synchronisation points, and taken while another lane's NCCL kernel waits for a peer they can close a cycle over lanes and
  // ranks (the first step hung at 8 GPUs).  Afterwards the lanes run concurrently and nothing is allocated in the steady state.
  u32 start = 0;
  if (d->warmed_lanes < nlanes) {
    const u32 npro = std::min<u32>(nlanes, n_batch);
    cudaSetDevice(ctx->device);
    for (u32 i = 0; i < npro; i++) {
      Lane* ln = ctx->lanes[i].get();
      int rc = dist_batch(ln, i, slot_base + i, n_local, texts[i], nbytes[i], on_device, hard_min[i],
                          kmers_per_partition ? kmers_per_partition + (size_t)i * P : nullptr);
      if (!rc && cudaStreamSynchronize(ln->st) != cudaSuccess) rc = fail(ln, KMX_ERR_CUDA, "stream synchronisation failed");
      if (rc) return rc;
    }
    d->warmed_lanes = std::max(d->warmed_lanes, npro); start = npro;
  }
  std::atomic<int> first_err(0);
  auto work = [&](u32 t) {
    cudaSetDevice(ctx->device);
    Lane* ln = ctx->lanes[t].get();
    // every rank walks the same (lane, sample) schedule, so the collectives of a lane's communicator match up
    for (u32 i = start + t; i < n_batch; i += nlanes) {
      int rc = first_err.load() ? first_err.load() : dist_batch(ln, t, slot_base + i, n_local, texts[i], nbytes[i], on_device, hard_min[i],
                                                                kmers_per_partition ? kmers_per_partition + (size_t)i * P : nullptr);
      if (rc) { int z = 0; first_err.compare_exchange_strong(z, rc); return; }
    }
    cudaStreamSynchronize(ln->st);
  };
  if (nlanes == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (u32 t = 0; t < nlanes; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  if (!first_err.load()) reap(ctx);               // every lane synchronised its stream: none of this rank's kernels is in flight
  return first_err.load();
}

static void dist_destroy(kmx_ctx* ctx)
{
  if (!ctx->dist) return;
  for (auto c : ctx->dist->comms) if (g_nccl.CommDestroy) g_nccl.CommDestroy(c);
  for (auto& b : ctx->dist->recv) release(ctx, b);
  for (auto& b : ctx->dist->meta_dev) release(ctx, b);
  ctx->dist.reset();
}
