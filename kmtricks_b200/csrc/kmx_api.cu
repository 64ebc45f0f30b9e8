// kmx_api.cu -- the C ABI (include/kmx.h): context, lanes, device memory, stage drivers.
#include "../../include/kmx.h"
#include "common.cuh"
#include "kmx_internal.h"
#include "s1_v5.cuh"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace kmx;

namespace kmx {
size_t scan_u32_work_bytes(u64 n);
cudaError_t scan_u32_inplace(u32* d, u64 n, u32* d_total, void* work, cudaStream_t st, u64* launches);
size_t radix_sort_work_bytes(u32 nseg, const u64* h_seg_off, const u64* h_seg_end = nullptr);
cudaError_t segmented_radix_sort(u32 nseg, const u64* h_seg_off, const u64* h_seg_end, u64* lo, u64* hi, u64* lo_alt, u64* hi_alt,
                                 int W, int begin_bit, int end_bit, void* d_work, int* result_in_alt,
                                 cudaStream_t st, u64* launches);
size_t rle_work_bytes(u32 nseg, const u64* h_seg_off);
cudaError_t rle_segments(u32 nseg, const u64* h_seg_off, const u64* lo, const u64* hi, int W, u32 hard_min,
                         void* d_work, std::vector<u64>& h_tile_off, std::vector<u64>& h_seg_out_off,
                         int phase, u64* out_lo, u64* out_hi, u32* out_cnt, cudaStream_t st, u64* launches);
cudaError_t launch_ht_insert_records(const S2Common& c, u64* keys, u32* cnts, const u64* toff, const u64* tcap, u32* overflow,
                                     cudaStream_t st, u64* launches);
cudaError_t launch_ht_compact(u32 P, u64 max_cap, const u64* keys, const u32* cnts, const u64* toff, const u64* tcap, u32 hard_min,
                              u64* out, u32* pcnt, cudaStream_t st, u64* launches);
cudaError_t launch_ht_lookup(u32 P, u32 max_n, const u64* keys, const u32* cnts, const u64* toff, const u64* tcap, const u64* skeys,
                             const u64* soff, const u32* pcnt, const u64* oo, u64* out_keys, u32* out_cnt, cudaStream_t st, u64* launches);
cudaError_t launch_ht_union(const MergeList* d_lists, u32 N, u64 max_n, u64* keys, u32* cnts, u64 cap, u32* overflow,
                            u64* out, u32* count, cudaStream_t st, u64* launches);
cudaError_t launch_ht2_insert_records(const S2Common& c, void* keys, u32* cnts, const u64* toff, const u64* tcap, u32* overflow,
                                      cudaStream_t st, u64* launches);
cudaError_t launch_ht2_compact(u32 P, u64 max_cap, const void* keys, const u32* cnts, const u64* toff, const u64* tcap, u32 hard_min,
                               u64* out_lo, u64* out_hi, u32* pcnt, cudaStream_t st, u64* launches);
cudaError_t launch_ht2_lookup(u32 P, u32 max_n, const void* keys, const u32* cnts, const u64* toff, const u64* tcap, const u64* slo, const u64* shi,
                              const u64* soff, const u32* pcnt, const u64* oo, u64* out_lo, u64* out_hi, u32* out_cnt, cudaStream_t st, u64* launches);
cudaError_t launch_ht2_union(const MergeList* d_lists, u32 N, u64 max_n, void* keys, u64 cap, u32* overflow,
                             u64* out_lo, u64* out_hi, u32* count, cudaStream_t st, u64* launches);
cudaError_t launch_sparse_solid(const MergeList* d_lists, u32 N, const u32* d_soft, const u64* ulo, const u64* uhi,
                                u64 nu, int W, u32* solid_in, u64 max_n, cudaStream_t st, u64* launches);
cudaError_t launch_sparse_emit(const MergeList* d_lists, u32 N, const u32* d_soft, u32 rmin, u32 share, u32 emit_all,
                               const u64* ulo, const u64* uhi, u64 nu, int W, const u32* solid_in,
                               const u32* keep, const u32* out_row, int fmt, uint8_t* body, u32 row_bytes,
                               uint8_t* row_keep, u64* stats, u64 max_n, cudaStream_t st, u64* launches);
}

namespace {

struct DBuf { void* p = nullptr; size_t cap = 0; };

struct ListRef {            // one (sample, partition) list in HBM
  u64* lo = nullptr; u64* hi = nullptr; u32* cnt = nullptr; u64 n = 0;
};

struct ArenaBlock { char* p; size_t cap, used; };

struct ProfSpan { cudaEvent_t a, b; int kind; };

}  // namespace

struct kmx_ctx;
struct KmxDist;

// A lane owns everything one in-flight sample needs (stream, text staging, line index, bucket
// slab, histogram, sort scratch).  Lane 0 also serves the single-sample API and the merges;
// kmx_run_samples drives several lanes from host threads so that the H2D copy and the small
// host round-trips of one sample overlap the kernels of the others.
struct Lane {
  kmx_ctx* ctx = nullptr; int id = 0;
  cudaStream_t st = nullptr;
  std::string err;
  u64 launches = 0;
  // ---- stage 1
  DBuf text, seq_start, seq_len, tile_counts, tile_prefix, nlmask, cta_tile;
  u64* d_total = nullptr;          // 1 u64
  u32* d_flags = nullptr;          // [0] fmt error [1] max len [2] overflow
  DBuf records;                    // bucket slab
  u64* d_boff = nullptr; u32* d_bcap = nullptr; u32* d_cursor = nullptr; u64* d_kcnt = nullptr;
  std::vector<u64> h_boff; std::vector<u32> h_bcap, h_cursor; std::vector<u64> h_kcnt;
  bool in_sample = false, sample_ready = false;
  char* h_pin = nullptr; size_t h_pin_cap = 0;     // pinned scratch for small read-backs
  // ---- stage 2
  DBuf hist, sub_counts, sub_off, bitmap;
  std::vector<ArenaBlock> sarena;  // lists of a pass whose output is thrown away (the all-keys pass behind --hist)
  bool to_scratch = false;
  DBuf hist_dev;                   // histogram pass: list table + counters
  DBuf binbuf, binmeta;            // binned hash counting: 16-bit slot offsets per (window, bin) + cursors / look-back words / layout
  u64 d_est = 0;                   // expected surviving (key,count) pairs per sample (grows with what was seen)
  DBuf keys_lo, keys_hi, keys_lo2, keys_hi2, sort_work, tmp_cnt, ht_keys, ht_cnts;
  // ---- profiling
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
};

struct kmx_ctx {
  int device = 0;
  kmx_params prm{};
  int W = 1;                       // words per k-mer
  int wlen = 0, max_nk = 0;
  std::string err;
  std::mutex mu;                   // arena, err, dev_bytes
  std::mutex lanes_mu;             // lane creation
  // Buffers replaced by bigger ones are NOT freed on the spot: cudaFree / cudaFreeHost wait for every kernel on the device, and in
  // a multi-GPU run another lane's NCCL kernel may be spinning for a peer whose own lane thread sits in the same kind of wait on
  // its device -- a cycle over lanes and ranks that hung the first step at 8 GPUs.  They are freed at points where none of this
  // context's kernels can be in flight (after all lanes were synchronised).
  std::mutex grave_mu;
  std::vector<std::pair<void*, size_t>> grave_dev;
  std::vector<void*> grave_host;
  u64 dev_bytes = 0;
  std::vector<void*> user_allocs;
  int hist_ok = -1;
  bool hist16 = true;              // hash histogram with 16-bit counters until one wraps (count_hash_hist / count_hash_binned)
  double bin_slack = 1.25;         // bin-region capacity over the mean bin load (doubles when a bin overflowed; > 8: L2-histogram path)
  int active_lanes = 1;            // lanes running concurrently (sizes the L2-resident histogram groups)
  std::atomic<u64> stat[KMX_STAT_KINDS];
  std::vector<double> rec_rate;    // [P] bucket records per input byte, the most seen so far: sizes the bucket regions of the next sample
  u32 s1_len_hint = 0;             // longest FASTQ read seen so far: geometry of the self-indexing stage-1 launch (0 = none yet)
  double ht_factor = 0.5;          // table slots per k-mer occurrence (doubles after an overflow)
  bool ht_union_ok = true;
  double union_factor = 2.5;       // slots of the merge's key-union set per entry of the largest list (grows with what the partitions show)
  bool prof_on = false;
  double prof_ms[KMX_PROF_KINDS] = {0};
  u64 prof_cnt[KMX_PROF_KINDS] = {0};
  void* merge_out = nullptr; size_t merge_out_cap = 0;
  uint16_t* d_repart = nullptr;
  u64* d_mload = nullptr;          // minimizer loads (kmx_minimizer_load_enable)
  std::vector<std::unique_ptr<Lane>> lanes;
  std::vector<ArenaBlock> arena;
  std::vector<ListRef> lists;      // [N*P]
  // ---- stage 3 (on lane 0's stream)
  DBuf d_lists, d_soft, solid_in, body, body2, stats, keep, out_row, row_keep, uni_lo, uni_hi, uni_lo2, uni_hi2, scan_work;
  kmx_merge_result last_res{};
  u32 last_emit_all = 0;
  uint8_t* last_body = nullptr;    // where the final body lives
  std::shared_ptr<KmxDist> dist;   // multi-GPU state (kmx_dist.inl)
};

// ---------------------------------------------------------------------------------------
static int fail(Lane* ln, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (ln) {
    ln->err = buf;
    std::lock_guard<std::mutex> g(ln->ctx->mu);
    ln->ctx->err = buf;
  }
  return code;
}

static cudaEvent_t prof_event(Lane* ln)
{
  cudaEvent_t e;
  if (!ln->prof_pool.empty()) { e = ln->prof_pool.back(); ln->prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  Lane* ln; ProfSpan s; bool on;
  ProfScope(Lane* l, int kind) : ln(l), on(l->ctx->prof_on) { if (on) { s.kind = kind; s.a = prof_event(ln); s.b = prof_event(ln); cudaEventRecord(s.a, ln->st); } }
  ~ProfScope() { if (on) { cudaEventRecord(s.b, ln->st); ln->prof_spans.push_back(s); } }
};
#define PROF(kind) ProfScope prof_scope_##kind(ln, kind)
static void prof_collect(kmx_ctx* c)
{
  for (auto& lp : c->lanes) {
    Lane* ln = lp.get();
    if (ln->prof_spans.empty()) continue;
    cudaStreamSynchronize(ln->st);
    for (auto& s : ln->prof_spans) {
      float ms = 0; cudaEventElapsedTime(&ms, s.a, s.b);
      c->prof_ms[s.kind] += ms; c->prof_cnt[s.kind] += 1;
      ln->prof_pool.push_back(s.a); ln->prof_pool.push_back(s.b);
    }
    ln->prof_spans.clear();
  }
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ln, e_ == cudaErrorMemoryAllocation ? KMX_ERR_NOMEM : KMX_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// Small host<->device transfers by a KERNEL that reads / writes the lane's pinned scratch directly (cudaMallocHost memory
// is device-accessible under unified addressing).  A cudaMemcpyAsync of a few hundred bytes queues on the copy engine
// behind the 315 MB FASTQ copies of the other lanes and stalls its own lane for milliseconds; a kernel does not.
struct SmallCopy { void* dst[4]; const void* src[4]; u32 n[4]; };
__global__ void __launch_bounds__(256) kmx_small_copy_kernel(SmallCopy c)
{
  for (int s = 0; s < 4; s++) {
    const u32 n = c.n[s];
    if (!n) continue;
    char* d = (char*)c.dst[s]; const char* a = (const char*)c.src[s];
    if ((((uintptr_t)d | (uintptr_t)a | n) & 7u) == 0) { for (u32 i = threadIdx.x; i < n / 8; i += 256) ((u64*)d)[i] = ((const volatile u64*)a)[i]; }
    else for (u32 i = threadIdx.x; i < n; i += 256) d[i] = ((const volatile char*)a)[i];
  }
}
struct SmallCopyBatch {
  Lane* ln; SmallCopy c; int k = 0;
  explicit SmallCopyBatch(Lane* l) : ln(l) { memset(&c, 0, sizeof c); }
  void add(void* dst, const void* src, size_t n) { c.dst[k] = dst; c.src[k] = src; c.n[k] = (u32)n; k++; }
  cudaError_t go();
};

static void add_bytes(kmx_ctx* ctx, long long d) { std::lock_guard<std::mutex> g(ctx->mu); ctx->dev_bytes = (u64)((long long)ctx->dev_bytes + d); }

cudaError_t SmallCopyBatch::go()
{
  if (!k) return cudaSuccess;
  kmx_small_copy_kernel<<<1, 256, 0, ln->st>>>(c);
  ln->launches += 1; k = 0; memset(&c, 0, sizeof c);
  return cudaGetLastError();
}

static void defer_free(kmx_ctx* ctx, void* dev, size_t cap) { std::lock_guard<std::mutex> g(ctx->grave_mu); ctx->grave_dev.push_back({dev, cap}); }
// frees what was set aside; the caller guarantees that no kernel of this context is in flight (all lanes synchronised)
static void reap(kmx_ctx* ctx)
{
  std::vector<std::pair<void*, size_t>> dv; std::vector<void*> hv;
  { std::lock_guard<std::mutex> g(ctx->grave_mu); dv.swap(ctx->grave_dev); hv.swap(ctx->grave_host); }
  for (auto& x : dv) { cudaFree(x.first); add_bytes(ctx, -(long long)x.second); }
  for (void* h : hv) cudaFreeHost(h);
}
static cudaError_t ensure(Lane* ln, DBuf& b, size_t bytes)
{
  if (bytes <= b.cap) return cudaSuccess;
  size_t ncap = std::max(bytes, b.cap + b.cap / 2);
  ncap = (ncap + 255) & ~(size_t)255;
  void* np = nullptr;
  cudaError_t e = cudaMalloc(&np, ncap);
  if (e == cudaErrorMemoryAllocation) {            // out of memory with buffers set aside: give them back (this lane only waits for itself first)
    cudaGetLastError();
    cudaStreamSynchronize(ln->st);
    reap(ln->ctx);
    e = cudaMalloc(&np, ncap);
  }
  if (e != cudaSuccess) return e;
  if (b.p) defer_free(ln->ctx, b.p, b.cap);        // kernels already queued on the old buffer stay valid
  b.p = np; b.cap = ncap; add_bytes(ln->ctx, (long long)ncap);
  return cudaSuccess;
}
static void release(kmx_ctx* ctx, DBuf& b) { if (b.p) { cudaFree(b.p); add_bytes(ctx, -(long long)b.cap); } b.p = nullptr; b.cap = 0; }

static cudaError_t ensure_pin(Lane* ln, size_t bytes)
{
  if (bytes <= ln->h_pin_cap) return cudaSuccess;
  if (ln->h_pin) { std::lock_guard<std::mutex> g(ln->ctx->grave_mu); ln->ctx->grave_host.push_back(ln->h_pin); ln->h_pin = nullptr; ln->h_pin_cap = 0; }
  size_t cap = std::max(bytes + bytes / 2, (size_t)1 << 20);
  cudaError_t e = cudaMallocHost((void**)&ln->h_pin, cap);
  if (e == cudaSuccess) ln->h_pin_cap = cap;
  return e;
}

static cudaError_t arena_alloc(kmx_ctx* ctx, size_t bytes, void** out)
{
  std::lock_guard<std::mutex> g(ctx->mu);
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  for (auto& b : ctx->arena) if (b.cap - b.used >= bytes) { *out = b.p + b.used; b.used += bytes; return cudaSuccess; }
  size_t cap = std::max(bytes, (size_t)256 << 20);
  ArenaBlock nb; nb.cap = cap; nb.used = bytes;
  cudaError_t e = cudaMalloc((void**)&nb.p, cap);
  if (e != cudaSuccess) return e;
  ctx->dev_bytes += cap;
  ctx->arena.push_back(nb);
  *out = nb.p;
  return cudaSuccess;
}
// output space of a counting pass: the context's arena (lists that stay), or the lane's scratch blocks
static cudaError_t list_alloc(Lane* ln, size_t bytes, void** out)
{
  if (!ln->to_scratch) return arena_alloc(ln->ctx, bytes, out);
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  for (auto& b : ln->sarena) if (b.cap - b.used >= bytes) { *out = b.p + b.used; b.used += bytes; return cudaSuccess; }
  ArenaBlock nb; nb.cap = std::max(bytes, (size_t)64 << 20); nb.used = bytes;
  cudaError_t e = cudaMalloc((void**)&nb.p, nb.cap);
  if (e != cudaSuccess) return e;
  add_bytes(ln->ctx, (long long)nb.cap);
  ln->sarena.push_back(nb);
  *out = nb.p;
  return cudaSuccess;
}
static void arena_clear(kmx_ctx* ctx)
{
  for (auto& b : ctx->arena) { cudaFree(b.p); ctx->dev_bytes -= b.cap; }
  ctx->arena.clear();
}

static void fastmod_magic(u64 d, u64& mlo, u64& mhi)
{
  unsigned __int128 M = (~(unsigned __int128)0) / d + 1;
  mlo = (u64)M; mhi = (u64)(M >> 64);
}

// lanes are created on demand; lane 0 at kmx_create
static int lane_create(kmx_ctx* ctx, int id)
{
  std::unique_ptr<Lane> up(new Lane());
  Lane* ln = up.get();
  ln->ctx = ctx; ln->id = id;
  ctx->lanes.push_back(std::move(up));
  const u32 P = ctx->prm.nb_partitions;
  CK(cudaStreamCreateWithFlags(&ln->st, cudaStreamNonBlocking));
  CK(cudaMalloc(&ln->d_total, 8)); CK(cudaMalloc(&ln->d_flags, 16));
  CK(cudaMalloc(&ln->d_boff, P * 8)); CK(cudaMalloc(&ln->d_bcap, P * 4));
  CK(cudaMalloc(&ln->d_cursor, P * 4)); CK(cudaMalloc(&ln->d_kcnt, P * 8));
  ln->h_boff.assign(P, 0); ln->h_bcap.assign(P, 0); ln->h_cursor.assign(P, 0); ln->h_kcnt.assign(P, 0);
  CK(ensure_pin(ln, (size_t)P * 1024 + ((size_t)1 << 20)));      // big enough for every later use at this P: no pinned reallocation mid-run
  return KMX_OK;
}
static void lane_destroy(Lane* ln)
{
  kmx_ctx* ctx = ln->ctx;
  if (ln->st) cudaStreamSynchronize(ln->st);
  DBuf* bufs[] = {&ln->text, &ln->seq_start, &ln->seq_len, &ln->tile_counts, &ln->tile_prefix, &ln->nlmask, &ln->records, &ln->hist, &ln->binbuf, &ln->binmeta,
                  &ln->sub_counts, &ln->sub_off, &ln->bitmap, &ln->cta_tile, &ln->keys_lo, &ln->keys_hi, &ln->keys_lo2, &ln->keys_hi2, &ln->sort_work, &ln->tmp_cnt, &ln->ht_keys, &ln->ht_cnts};
  for (DBuf* b : bufs) release(ctx, *b);
  release(ctx, ln->hist_dev);
  for (auto& b : ln->sarena) { cudaFree(b.p); add_bytes(ctx, -(long long)b.cap); }
  ln->sarena.clear();
  void* singles[] = {ln->d_total, ln->d_flags, ln->d_boff, ln->d_bcap, ln->d_cursor, ln->d_kcnt};
  for (void* p : singles) if (p) cudaFree(p);
  if (ln->h_pin) cudaFreeHost(ln->h_pin);
  for (auto e : ln->prof_pool) cudaEventDestroy(e);
  for (auto& s : ln->prof_spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  if (ln->st) cudaStreamDestroy(ln->st);
}

// ---------------------------------------------------------------------------------------
extern "C" int kmx_create(int device, const kmx_params* prm, kmx_ctx** out)
{
  if (!prm || !out) return KMX_ERR_ARG;
  *out = nullptr;
  kmx_ctx* ctx = new kmx_ctx();
  for (auto& v : ctx->stat) v = 0;
  *out = ctx;                                   // returned even on failure so the caller can read the error
  ctx->device = device; ctx->prm = *prm;
  Lane tmp; tmp.ctx = ctx; Lane* ln = &tmp;     // error sink until lane 0 exists
  const u32 k = prm->kmer_size, m = prm->minim_size;
  if (k < 8 || k > 63) return fail(ln, KMX_ERR_ARG, "kmer_size %u unsupported (8..63)", k);
  if (m < 4 || m > 12 || m > k) return fail(ln, KMX_ERR_ARG, "minim_size %u unsupported (4..12, <= k)", m);
  if (prm->nb_partitions < 1 || prm->nb_partitions > 65535) return fail(ln, KMX_ERR_ARG, "nb_partitions %u unsupported", prm->nb_partitions);
  if (!prm->repart_table) return fail(ln, KMX_ERR_ARG, "repart_table is NULL");
  if (prm->nb_samples < 1) return fail(ln, KMX_ERR_ARG, "nb_samples must be >= 1");
  if (prm->key_kind == KMX_KEY_HASH && (prm->window_bits == 0 || prm->window_bits % 64)) return fail(ln, KMX_ERR_ARG, "window_bits must be a positive multiple of 64");
  if (prm->key_kind > KMX_KEY_HASH) return fail(ln, KMX_ERR_ARG, "bad key_kind");
  ctx->W = (k + 31) / 32;
  ctx->wlen = (int)(k - m + 1);
  ctx->max_nk = (ctx->W == 1 ? KMX_REC1_MAXN : KMX_REC2_MAXN) - (int)k + 1;
  if (s1_smem_bytes(128, 2048, ctx->wlen, prm->nb_partitions) > 200 * 1024)
    return fail(ln, KMX_ERR_ARG, "nb_partitions %u too large for the stage-1 staging layout", prm->nb_partitions);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= device) return fail(ln, KMX_ERR_CUDA, "no usable CUDA device %d (%s)", device, cudaGetErrorString(e));
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(ln, KMX_ERR_CUDA, "device %d is sm_%d%d; libkmx_sm100 needs sm_100", device, prop.major, prop.minor);
  const size_t tn = (size_t)1 << (2 * m);
  for (size_t i = 0; i < tn; i++) if (prm->repart_table[i] >= prm->nb_partitions) return fail(ln, KMX_ERR_ARG, "repart_table[%zu]=%u >= nb_partitions", i, prm->repart_table[i]);
  CK(cudaMalloc(&ctx->d_repart, tn * 2)); ctx->dev_bytes += tn * 2;
  CK(cudaMemcpy(ctx->d_repart, prm->repart_table, tn * 2, cudaMemcpyHostToDevice));
  ctx->prm.repart_table = nullptr;
  ctx->lists.assign((size_t)prm->nb_samples * prm->nb_partitions, ListRef());
  ctx->rec_rate.assign(prm->nb_partitions, 0.0);
  int rc = lane_create(ctx, 0);
  if (rc) return rc;
  return KMX_OK;
}

static void dist_destroy(kmx_ctx* ctx);

extern "C" void kmx_destroy(kmx_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& lp : ctx->lanes) if (lp->st) cudaStreamSynchronize(lp->st);
  reap(ctx);
  dist_destroy(ctx);
  for (auto& lp : ctx->lanes) lane_destroy(lp.get());
  DBuf* bufs[] = {&ctx->d_lists, &ctx->d_soft, &ctx->solid_in, &ctx->body, &ctx->body2, &ctx->stats, &ctx->keep, &ctx->out_row,
                  &ctx->row_keep, &ctx->uni_lo, &ctx->uni_hi, &ctx->uni_lo2, &ctx->uni_hi2, &ctx->scan_work};
  for (DBuf* b : bufs) release(ctx, *b);
  arena_clear(ctx);
  for (void* p : ctx->user_allocs) cudaFree(p);
  if (ctx->d_repart) cudaFree(ctx->d_repart);
  if (ctx->d_mload) cudaFree(ctx->d_mload);
  delete ctx;
}

#define LANE0 Lane* ln = ctx->lanes.empty() ? nullptr : ctx->lanes[0].get(); if (!ln) return KMX_ERR_STATE; cudaSetDevice(ctx->device)

extern "C" const char* kmx_last_error(const kmx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" uint64_t kmx_launch_count(const kmx_ctx* ctx)
{
  if (!ctx) return 0;
  u64 t = 0; for (auto& l : ctx->lanes) t += l->launches;
  return t;
}
extern "C" uint64_t kmx_device_bytes(const kmx_ctx* ctx) { return ctx ? ctx->dev_bytes : 0; }
extern "C" uint64_t kmx_stat(const kmx_ctx* ctx, int which) { return (ctx && which >= 0 && which < KMX_STAT_KINDS) ? ctx->stat[which].load() : 0; }
extern "C" void* kmx_stream(kmx_ctx* ctx) { return (ctx && !ctx->lanes.empty()) ? (void*)ctx->lanes[0]->st : nullptr; }
extern "C" int kmx_sync(kmx_ctx* ctx)
{
  if (!ctx) return KMX_ERR_ARG;
  for (auto& lp : ctx->lanes) { Lane* ln = lp.get(); CK(cudaStreamSynchronize(ln->st)); }
  reap(ctx);                                      // nothing of this context is in flight now
  return KMX_OK;
}

extern "C" int kmx_reset(kmx_ctx* ctx)
{
  if (!ctx) return KMX_ERR_ARG;
  int rc = kmx_sync(ctx);
  if (rc) return rc;
  for (auto& b : ctx->arena) b.used = 0;          // keep the blocks: no cudaFree/cudaMalloc per step
  for (auto& l : ctx->lists) l = ListRef();
  for (auto& lp : ctx->lanes) lp->in_sample = lp->sample_ready = false;
  return KMX_OK;
}

// ---------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------
static int upload_bucket_meta(Lane* ln)
{
  const u32 P = ln->ctx->prm.nb_partitions;
  // stage through the pinned scratch so the copies are truly asynchronous
  char* hp = ln->h_pin;
  memcpy(hp, ln->h_boff.data(), P * 8); memcpy(hp + P * 8, ln->h_kcnt.data(), P * 8);
  memcpy(hp + P * 16, ln->h_bcap.data(), P * 4); memcpy(hp + P * 20, ln->h_cursor.data(), P * 4);
  SmallCopyBatch b(ln);
  b.add(ln->d_boff, hp, P * 8); b.add(ln->d_kcnt, hp + P * 8, P * 8);
  b.add(ln->d_bcap, hp + P * 16, P * 4); b.add(ln->d_cursor, hp + P * 20, P * 4);
  CK(b.go());
  CK(cudaStreamSynchronize(ln->st));             // the scratch is reused right away
  return KMX_OK;
}

// make room for need[p] records in every partition (keeps what is there)
static int grow_buckets(Lane* ln, const std::vector<u64>& need)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const size_t rec = ctx->W == 1 ? 16 : 32;
  bool ok = true;
  for (u32 p = 0; p < P; p++) if (need[p] > ln->h_bcap[p]) ok = false;
  if (ok) return KMX_OK;
  std::vector<u64> nboff(P); std::vector<u32> ncap(P);
  u64 tot = 0;
  for (u32 p = 0; p < P; p++) {
    u64 c = std::max<u64>(need[p], ln->h_bcap[p]);
    c = (c + 63) & ~(u64)63;
    if (c > 0xFFFFFFF0ULL) return fail(ln, KMX_ERR_NOMEM, "partition %u needs %llu records (> 2^32)", p, (unsigned long long)c);
    nboff[p] = tot; ncap[p] = (u32)c; tot += c;
  }
  DBuf nb;
  cudaError_t e = cudaMalloc(&nb.p, tot * rec + 256);
  if (e != cudaSuccess) return fail(ln, KMX_ERR_NOMEM, "bucket slab of %llu bytes: %s", (unsigned long long)(tot * rec), cudaGetErrorString(e));
  nb.cap = tot * rec + 256; add_bytes(ctx, (long long)nb.cap);
  if (ln->records.p) {
    for (u32 p = 0; p < P; p++) if (ln->h_cursor[p])
      CK(cudaMemcpyAsync((char*)nb.p + nboff[p] * rec, (char*)ln->records.p + ln->h_boff[p] * rec,
                         (size_t)std::min<u64>(ln->h_cursor[p], ln->h_bcap[p]) * rec, cudaMemcpyDeviceToDevice, ln->st));
    defer_free(ctx, ln->records.p, ln->records.cap);      // the copies above read it; freed when nothing is in flight
    ln->records.p = nullptr; ln->records.cap = 0;
  }
  ln->records = nb; ln->h_boff = nboff; ln->h_bcap = ncap;
  return KMX_OK;
}

// a sample's first push: lay the regions out afresh, need[p] records each (nothing to keep), reusing the slab when it is big enough
static int layout_buckets(Lane* ln, const std::vector<u64>& need)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const size_t rec = ctx->W == 1 ? 16 : 32;
  u64 tot = 0;
  for (u32 p = 0; p < P; p++) {
    const u64 c = (need[p] + 63) & ~(u64)63;
    if (c > 0xFFFFFFF0ULL) return fail(ln, KMX_ERR_NOMEM, "partition %u needs %llu records (> 2^32)", p, (unsigned long long)c);
    ln->h_boff[p] = tot; ln->h_bcap[p] = (u32)c; tot += c;
  }
  if (tot * rec + 256 > ln->records.cap) {
    if (ln->records.p) { defer_free(ctx, ln->records.p, ln->records.cap); ln->records.p = nullptr; ln->records.cap = 0; }
    const size_t cap = (size_t)((double)(tot * rec) * 1.1) + 256;
    cudaError_t e = cudaMalloc(&ln->records.p, cap);
    if (e != cudaSuccess) { ln->records.p = nullptr; return fail(ln, KMX_ERR_NOMEM, "bucket slab of %zu bytes: %s", cap, cudaGetErrorString(e)); }
    ln->records.cap = cap; add_bytes(ctx, (long long)cap);
  }
  return KMX_OK;
}

static int superk_begin(Lane* ln)
{
  std::fill(ln->h_cursor.begin(), ln->h_cursor.end(), 0u);
  std::fill(ln->h_kcnt.begin(), ln->h_kcnt.end(), 0ull);
  int rc = upload_bucket_meta(ln);             // device cursors start at zero even with no push
  if (rc) return rc;
  ln->in_sample = true; ln->sample_ready = false;
  return KMX_OK;
}

// run stage 1 over nseg segments described by (d_start, d_len) into the buckets; retries
// with exact capacities when a bucket overflowed.
static const int KMX_S1_FALLBACK = -1001;   // internal: the self-indexing launch met a longer read or a format problem; redo with the line index
static int run_s1(Lane* ln, const uint8_t* d_text, u64 text_bytes, const u32* d_start, const u32* d_len, u64 nseg, u64 est_kmers, u32 max_len,
                  const S1Idx* fi = nullptr)
{
  if (max_len > 2048) return fail(ln, KMX_ERR_FORMAT, "sequence of %u bases: stage 1 takes segments of <= 2048 bases (kmx_superk_push_reads splits long sequences)", max_len);
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  // Region capacities.  From the densest sample seen so far (records per input byte and partition, +15 %): the regions are
  // then barely larger than what is written, which is also what the multi-GPU exchange ships.  Before any sample was seen:
  // ~1 record per 8 k-mers, 30 % head-room for partition imbalance.  Too small -> the cursors keep counting, one retry.
  std::vector<u64> need(P);
  bool fresh = true;
  for (u32 p = 0; p < P; p++) if (ln->h_cursor[p]) fresh = false;
  const std::vector<u32> cur0 = ln->h_cursor;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    for (u32 p = 0; p < P; p++) {
      const double r = kmx_env_flag("KMX_S1_LOOSE") ? 0.0 : ctx->rec_rate[p];
      need[p] = (u64)ln->h_cursor[p] + (r > 0 ? (u64)(r * (double)text_bytes * 1.15) + 512 : (u64)((double)est_kmers / 8.0 / P * 1.3) + 1024);
    }
  }
  for (int attempt = 0; attempt < 3; attempt++) {
    int rc = fresh ? layout_buckets(ln, need) : grow_buckets(ln, need);
    if (rc) return rc;
    rc = upload_bucket_meta(ln);
    if (rc) return rc;
    if (fi) CK(cudaMemsetAsync(ln->d_flags, 0, 16, ln->st));
    else CK(cudaMemsetAsync(ln->d_flags + 2, 0, 4, ln->st));
    S1Args a;
    a.text = d_text; a.text_bytes = text_bytes; a.seg_start = d_start; a.seg_len = d_len; a.nseg = nseg;
    a.k = (int)ctx->prm.kmer_size; a.m = (int)ctx->prm.minim_size; a.wlen = ctx->wlen; a.max_nk = ctx->max_nk;
    a.P = P; a.repart = ctx->d_repart; a.records = ln->records.p; a.boff = ln->d_boff; a.bcap = ln->d_bcap;
    a.cursor = ln->d_cursor; a.kcnt = ln->d_kcnt; a.overflow = ln->d_flags + 2;
    a.pack_words = (max_len + 15) / 16; if (a.pack_words < 1) a.pack_words = 1;
    a.stage_cap = 2048; a.flush_thr = 2048 - 1152;
    a.mload = ctx->d_mload;
    { PROF(KMX_PROF_S1);
      s1v5::Geo geo; size_t smem5 = 0;
      if (fi) {
        if (!s1_v5_usable(max_len, a.k, a.m, P, &geo, &smem5) || geo.R != S1_FUSED_R) return KMX_S1_FALLBACK;
        CK(launch_s1_v5(ctx->W, a, geo, smem5, fi, ln->st, &ln->launches));
      } else if (s1_v5_usable(max_len, a.k, a.m, P, &geo, &smem5)) CK(launch_s1_v5(ctx->W, a, geo, smem5, nullptr, ln->st, &ln->launches));
      else CK(launch_s1(ctx->W, a, ln->st, &ln->launches)); }
    u64* kc = (u64*)ln->h_pin; u32* cur = (u32*)(ln->h_pin + P * 8); u32* ovf = (u32*)(ln->h_pin + P * 12);
    u32* fl4 = (u32*)(ln->h_pin + P * 12 + 16);
    { SmallCopyBatch b(ln);
      b.add(kc, ln->d_kcnt, P * 8); b.add(cur, ln->d_cursor, P * 4); b.add(ovf, ln->d_flags + 2, 4);
      if (fi) b.add(fl4, ln->d_flags, 16);
      CK(b.go()); }
    CK(cudaStreamSynchronize(ln->st));
    if (fi && (fl4[0] || fl4[3])) return KMX_S1_FALLBACK;     // the device cursors are restored from the host snapshot by the next upload
    if (!*ovf) {
      ln->h_cursor.assign(cur, cur + P); ln->h_kcnt.assign(kc, kc + P);
      if (text_bytes >= 4096) {
        std::lock_guard<std::mutex> g(ctx->mu);
        for (u32 p = 0; p < P; p++) ctx->rec_rate[p] = std::max(ctx->rec_rate[p], (double)(cur[p] - cur0[p]) / (double)text_bytes);
      }
      return KMX_OK;
    }
    ctx->stat[KMX_STAT_S1_RETRY]++;
    // exact sizes are now known (the cursors kept counting); redo this push from the snapshot
    for (u32 p = 0; p < P; p++) need[p] = (u64)cur[p] + 64;
  }
  return fail(ln, KMX_ERR_CUDA, "stage-1 bucket overflow persisted after resize");
}

static int superk_push_fastq(Lane* ln, const char* text, size_t nbytes, int on_device)
{
  if (!ln->in_sample) return fail(ln, KMX_ERR_STATE, "kmx_superk_push_fastq outside begin/end");
  if (nbytes == 0) return KMX_OK;
  if (nbytes >= 0xFFFFFFF0ULL) return fail(ln, KMX_ERR_ARG, "text block must be < 4 GiB; split at record boundaries");
  const uint8_t* d_text;
  if (on_device) d_text = (const uint8_t*)text;
  else {
    CK(ensure(ln, ln->text, nbytes + 64));
    CK(cudaMemcpyAsync(ln->text.p, text, nbytes, cudaMemcpyHostToDevice, ln->st));
    d_text = (const uint8_t*)ln->text.p;
  }
  const u64 ntiles = fq_num_tiles(d_text, nbytes);
  CK(ensure(ln, ln->tile_counts, ntiles * 4));
  CK(ensure(ln, ln->tile_prefix, ntiles * 8));
  CK(ensure(ln, ln->nlmask, ntiles * 256 * 8));             // one bit per text byte, whole 16 KiB tiles
  { PROF(KMX_PROF_INDEX); CK(launch_fq_index(d_text, nbytes, (u32*)ln->tile_counts.p, (u64*)ln->tile_prefix.p, (u64*)ln->nlmask.p, ln->d_total, nullptr, nullptr, 0, nullptr, 0, ln->st, &ln->launches)); }
  u64* nl = (u64*)ln->h_pin; uint8_t* last = (uint8_t*)(ln->h_pin + 8);
  { SmallCopyBatch b(ln); b.add(nl, ln->d_total, 8); b.add(last, d_text + nbytes - 1, 1); CK(b.go()); }
  CK(cudaStreamSynchronize(ln->st));
  u64 nlines = *nl + (*last != '\n' ? 1 : 0);
  if (nlines % 4) return fail(ln, KMX_ERR_FORMAT, "FASTQ block has %llu lines (not a multiple of 4)", (unsigned long long)nlines);
  const u64 nrec = nlines / 4;
  if (nrec == 0) return KMX_OK;
  // Self-indexing launch: once the read length of this run is known (from an earlier block), stage 1 finds its reads in the
  // newline masks itself -- no line-index pass, no seq_start / seq_len arrays.  A longer read or a format problem sends the
  // block through the indexed path below (which also reports the error).
  u32 hint;
  { std::lock_guard<std::mutex> g(ln->ctx->mu); hint = ln->ctx->s1_len_hint; }
  if (hint && !kmx_env_flag("KMX_S1_NOFUSE")) {
    const u64 ncta = (nrec + S1_FUSED_R - 1) / S1_FUSED_R;
    CK(ensure(ln, ln->cta_tile, ncta * 4));
    { PROF(KMX_PROF_INDEX); CK(launch_fq_cta_pos((const u64*)ln->nlmask.p, (const u64*)ln->tile_prefix.p, ntiles, (u32*)ln->cta_tile.p, ncta, ln->st, &ln->launches)); }
    S1Idx fi;
    const uintptr_t ta = reinterpret_cast<uintptr_t>(d_text);
    fi.nlmask64 = (const u64*)ln->nlmask.p; fi.cta_pos = (const u32*)ln->cta_tile.p;
    fi.ntiles = ntiles; fi.lead = ta & 15; fi.tot = fi.lead + nbytes; fi.flags = ln->d_flags; fi.geo_maxlen = hint;
    const int rc = run_s1(ln, d_text, nbytes, nullptr, nullptr, nrec, nbytes / 2, hint, &fi);
    if (rc != KMX_S1_FALLBACK) { if (rc == KMX_OK) ln->ctx->stat[KMX_STAT_S1_SELF_INDEXED]++; return rc; }
  }
  CK(ensure(ln, ln->seq_start, nrec * 4));
  CK(ensure(ln, ln->seq_len, nrec * 4));
  CK(cudaMemsetAsync(ln->d_flags, 0, 8, ln->st));
  { PROF(KMX_PROF_INDEX); CK(launch_fq_index(d_text, nbytes, (u32*)ln->tile_counts.p, (u64*)ln->tile_prefix.p, (u64*)ln->nlmask.p, ln->d_total,
                     (u32*)ln->seq_start.p, (u32*)ln->seq_len.p, nrec, ln->d_flags, 1, ln->st, &ln->launches)); }
  u32* fl = (u32*)ln->h_pin;
  { SmallCopyBatch b(ln); b.add(fl, ln->d_flags, 8); CK(b.go()); }
  CK(cudaStreamSynchronize(ln->st));
  if (fl[0]) return fail(ln, KMX_ERR_FORMAT, "text is not strict 4-line FASTQ");
  const u32 block_max = fl[1];
  ln->ctx->stat[KMX_STAT_S1_INDEXED]++;
  { std::lock_guard<std::mutex> g(ln->ctx->mu); ln->ctx->s1_len_hint = std::max(ln->ctx->s1_len_hint, block_max); }
  return run_s1(ln, d_text, nbytes, (const u32*)ln->seq_start.p, (const u32*)ln->seq_len.p, nrec, nbytes / 2, block_max);
}

static int superk_push_reads(Lane* ln, const char* seqs, const uint64_t* off, size_t nseq)
{
  if (!ln->in_sample) return fail(ln, KMX_ERR_STATE, "kmx_superk_push_reads outside begin/end");
  if (nseq == 0) return KMX_OK;
  const u32 k = ln->ctx->prm.kmer_size;
  const u64 CH = 1024;                    // k-mers per segment for long sequences
  size_t i = 0;
  while (i < nseq) {                      // blocks of < 2 GiB of sequence
    u64 b0 = off[i];
    size_t j = i;
    std::vector<u32> st, sl;
    u64 kmers = 0;
    while (j < nseq && off[j + 1] - b0 < 0x7FFFFFFFULL) {
      u64 s = off[j] - b0, len = off[j + 1] - off[j];
      if (len >= k) {
        if (len <= CH + k - 1) { st.push_back((u32)s); sl.push_back((u32)len); }
        else for (u64 c = 0; c + k <= len; c += CH) { st.push_back((u32)(s + c)); sl.push_back((u32)std::min<u64>(CH + k - 1, len - c)); }
        kmers += len - k + 1;
      }
      j++;
    }
    if (j == i) return fail(ln, KMX_ERR_ARG, "sequence %zu is longer than 2 GiB", i);
    u64 nb = off[j] - b0;
    if (!st.empty()) {
      CK(ensure(ln, ln->text, nb + 64));
      CK(cudaMemcpyAsync(ln->text.p, seqs + b0, nb, cudaMemcpyHostToDevice, ln->st));
      CK(ensure(ln, ln->seq_start, st.size() * 4));
      CK(ensure(ln, ln->seq_len, st.size() * 4));
      CK(cudaMemcpyAsync(ln->seq_start.p, st.data(), st.size() * 4, cudaMemcpyHostToDevice, ln->st));
      CK(cudaMemcpyAsync(ln->seq_len.p, sl.data(), sl.size() * 4, cudaMemcpyHostToDevice, ln->st));
      CK(cudaStreamSynchronize(ln->st));
      u32 mx = 0; for (u32 v : sl) mx = std::max(mx, v);
      int rc = run_s1(ln, (const uint8_t*)ln->text.p, nb, (const u32*)ln->seq_start.p, (const u32*)ln->seq_len.p, st.size(), kmers, mx);
      if (rc) return rc;
    }
    i = j;
  }
  return KMX_OK;
}

static int superk_end(Lane* ln, uint64_t* kmers_per_partition)
{
  if (!ln->in_sample) return fail(ln, KMX_ERR_STATE, "kmx_superk_end without begin");
  if (kmers_per_partition) memcpy(kmers_per_partition, ln->h_kcnt.data(), ln->ctx->prm.nb_partitions * 8);
  ln->in_sample = false; ln->sample_ready = true;
  return KMX_OK;
}

extern "C" int kmx_superk_begin(kmx_ctx* ctx) { if (!ctx) return KMX_ERR_ARG; LANE0; return superk_begin(ln); }
extern "C" int kmx_superk_push_fastq(kmx_ctx* ctx, const char* text, size_t nbytes, int on_device)
{
  if (!ctx || (!text && nbytes)) return KMX_ERR_ARG;
  LANE0; return superk_push_fastq(ln, text, nbytes, on_device);
}
extern "C" int kmx_superk_push_reads(kmx_ctx* ctx, const char* seqs, const uint64_t* off, size_t nseq)
{
  if (!ctx || (nseq && (!seqs || !off))) return KMX_ERR_ARG;
  LANE0; return superk_push_reads(ln, seqs, off, nseq);
}
extern "C" int kmx_superk_end(kmx_ctx* ctx, uint64_t* kmers_per_partition) { if (!ctx) return KMX_ERR_ARG; LANE0; return superk_end(ln, kmers_per_partition); }

// ---------------------------------------------------------------------------------------
// stage 2
// ---------------------------------------------------------------------------------------
static int count_generic(Lane* ln, uint32_t sample, uint32_t hard_min);
static int count_kmer_ht(Lane* ln, uint32_t sample, uint32_t hard_min);
static const int KMX_HT_FALLBACK = -1000;   // internal: table overflowed, use the sort path

// Hash keys, histogram path.  Partitions are processed in GROUPS of `gp` windows whose histogram
// (gp x W x u32, <= 32 MB) is reused group after group and therefore stays L2-resident: fill (RED),
// count survivors per sub-chunk, device-side scan + bump allocation of the output space, ordered
// emit + re-zero -- no DRAM streaming of a P x W histogram and no host round trip per group.
// win_sample / win_part (host, [P], optional): window v holds partition win_part[v] of sample slot
// win_sample[v] (multi-GPU: one pass counts the same partitions of several samples).
static int count_hash_hist(Lane* ln, uint32_t sample, uint32_t hard_min, const u32* win_sample = nullptr, const u32* win_part = nullptr)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const u64 Wb = ctx->prm.window_bits;
  // up to 1 GiB of histogram per lane: all partitions in one group (fewest launches, measured fastest);
  // beyond that (big Bloom filters) groups of ~100 MB per in-flight lane, reused group after group
  u32 gp = P;
  if ((u64)P * Wb * 4 > ((u64)1 << 30)) {
    const u64 budget = ((u64)100 << 20) / (u64)std::max(1, ctx->active_lanes);
    gp = (u32)std::max<u64>(1, std::min<u64>(P, budget / (Wb * 4)));
  }
  // 16-bit counters (half the histogram bytes) until a sample makes one wrap; from then on 32-bit counters
  bool h16;
  { std::lock_guard<std::mutex> g(ctx->mu); h16 = ctx->hist16 && !kmx_env_flag("KMX_HIST32"); }
  auto ensure_hist = [&](bool half) -> int {
    const size_t hist_bytes = (size_t)gp * Wb * (half ? 2 : 4);
    if (ln->hist.cap < hist_bytes) {
      CK(ensure(ln, ln->hist, hist_bytes));
      CK(cudaMemsetAsync(ln->hist.p, 0, ln->hist.cap, ln->st));
    }
    return KMX_OK;
  };
  { int rc = ensure_hist(h16); if (rc) return rc; }
  // device meta: chunk_counts u32[gp*CW] | slice_counts u32[gp*CW*8] ; chunk_off u64[gp*CW] | list_off u64[P] | meta u64[4] |
  // flags u32[2] | win_part u32[P] ; staging: one (u16 slot offset, u32 count) entry per slot of the group (runs per 1024-slot slice)
  const size_t n_chunks = (size_t)gp * hash_sweep_chunks_per_window(Wb);
  const size_t sc_words = (n_chunks * 9 + (size_t)gp + 2 + 1) & ~(size_t)1;      // u32 words before the (8-byte aligned) window sums
  CK(ensure(ln, ln->sub_counts, sc_words * 4 + (size_t)gp * 8));
  CK(ensure(ln, ln->sub_off, n_chunks * 8 + (size_t)P * 8 + 32 + 8 + (size_t)P * 4 + 64));
  CK(ensure(ln, ln->bitmap, n_chunks * HIST_SUB * 6));
  SweepStage stage; stage.cnt = (u32*)ln->bitmap.p; stage.idx = (uint16_t*)(stage.cnt + n_chunks * HIST_SUB);
  stage.slice_counts = (u32*)ln->sub_counts.p + n_chunks; stage.done = stage.slice_counts + n_chunks * 8;
  stage.win_sum = (u64*)((u32*)ln->sub_counts.p + sc_words);
  u64* d_coff = (u64*)ln->sub_off.p; u64* d_loff = d_coff + n_chunks; u64* d_meta = d_loff + P;
  u32* d_flags = (u32*)(d_meta + 4); u32* d_wpart = d_flags + 2;
  CK(ensure_pin(ln, (size_t)P * 8 + 64 + (size_t)P * 32 + 256));
  S2Common c; c.W = ctx->W; c.k = (int)ctx->prm.kmer_size; c.P = P; c.records = ln->records.p; c.boff = ln->d_boff;
  c.bcnt = ln->d_cursor; c.kcnt = ln->d_kcnt; c.max_bcnt = *std::max_element(ln->h_cursor.begin(), ln->h_cursor.end());
  u64 mlo, mhi; fastmod_magic(Wb, mlo, mhi);
  const u32* dwp = nullptr;
  if (win_part) {
    u32* hp = (u32*)(ln->h_pin + (size_t)P * 8 + 64);
    memcpy(hp, win_part, (size_t)P * 4);
    { SmallCopyBatch b(ln); b.add(d_wpart, hp, (size_t)P * 4); CK(b.go()); }
    dwp = d_wpart;
  }
  u64 cap = std::max<u64>(4096, ln->d_est);
  for (int attempt = 0; attempt < 4; attempt++) {
    void* kp = nullptr; void* cp = nullptr;
    CK(list_alloc(ln, cap * 8, &kp));
    CK(list_alloc(ln, cap * 4, &cp));
    u64* hm = (u64*)(ln->h_pin + (size_t)P * 8);             // staging for meta = {cursor, cursor', capacity, -} + flags
    hm[0] = 0; hm[1] = 0; hm[2] = cap; hm[3] = 0; hm[4] = 0;
    { SmallCopyBatch b(ln); b.add(d_meta, hm, 40); CK(b.go()); }            // meta[4] + flags[2]
    u32 ngroups = 0;
    for (u32 p0 = 0; p0 < P; p0 += gp, ngroups++) {
      const u32 g = std::min(gp, P - p0);
      { PROF(KMX_PROF_HASH_HIST);
        CK(launch_hash_group(c, Wb, Wb, mlo, mhi, (u32*)ln->hist.p, hard_min, p0, g, ngroups, (u32*)ln->sub_counts.p, d_coff, stage, d_loff,
                             d_meta, d_flags, (u64*)kp, (u32*)cp, dwp, ln->st, &ln->launches, 0, h16)); }
      { PROF(KMX_PROF_HASH_EMIT);
        CK(launch_hash_group(c, Wb, Wb, mlo, mhi, (u32*)ln->hist.p, hard_min, p0, g, ngroups, (u32*)ln->sub_counts.p, d_coff, stage, d_loff,
                             d_meta, d_flags, (u64*)kp, (u32*)cp, dwp, ln->st, &ln->launches, 1, h16)); }
    }
    u64* h_l = (u64*)ln->h_pin;                               // list_off[P] then meta[4], flags[2]
    { SmallCopyBatch b(ln); b.add(h_l, d_loff, (size_t)P * 8 + 40); CK(b.go()); }
    CK(cudaStreamSynchronize(ln->st));
    const u64 D = h_l[P + (ngroups & 1u)];
    const u32 ovf = *(u32*)(h_l + P + 4), wrapped = *((u32*)(h_l + P + 4) + 1);
    if (h16 && wrapped) {                                     // a 16-bit counter wrapped: the lists are void, the histogram is all-zero again
      { std::lock_guard<std::mutex> g(ctx->mu); ctx->hist16 = false; }
      h16 = false;
      int rc = ensure_hist(false);
      if (rc) return rc;
      continue;
    }
    ln->d_est = std::max<u64>(ln->d_est, D + D / 4 + 1024);
    if (ovf) { cap = D + 1024; continue; }                   // the cursor kept counting: exact size now known, histogram is all-zero again
    for (u32 v = 0; v < P; v++) {
      if (win_part && ln->h_cursor[v] == 0) continue;       // unused window
      const u32 smp = win_sample ? win_sample[v] : sample, prt = win_part ? win_part[v] : v;
      ListRef& L = ctx->lists[(size_t)smp * P + prt];
      const u64 end = v + 1 < P ? h_l[v + 1] : D;
      L.lo = (u64*)kp + h_l[v]; L.hi = nullptr; L.cnt = (u32*)cp + h_l[v]; L.n = end - h_l[v];
    }
    return KMX_OK;
  }
  return fail(ln, KMX_ERR_CUDA, "hash-count output space overflowed repeatedly");
}

// Hash keys, k <= 32, binned path (s2_bin.cu): pass A hashes every k-mer once and appends its 16-bit in-bin offset to the
// region of its (window, bin); pass B counts each bin in shared memory and emits the survivors in order.  Window v is
// partition v of `sample`, or (win_sample[v], win_part[v]) in the multi-GPU single pass.  Returns KMX_BIN_FALLBACK when
// the data defeats the uniform bin regions (one slot holding a large share of a window): the caller takes the L2 path.
static const int KMX_BIN_FALLBACK = -1002;
static bool hash_binned_usable(const kmx_ctx* ctx)
{
  if (ctx->prm.key_kind != KMX_KEY_HASH || ctx->W != 1) return false;
  if (kmx_env_flag("KMX_HASH_NOBIN") || kmx_env_flag("KMX_HIST_NOROLL") || kmx_env_flag("KMX_HIST_FUSE")) return false;   // A/B: L2-histogram kernels
  const u64 Wb = ctx->prm.window_bits;
  if (Wb >= (1ULL << 30)) return false;                    // hb_hash_mod: r < 3 d must fit 32 bits
  const u64 NB = (Wb + (1ULL << 14) - 1) >> 14;            // bins per window with 32-bit counters (the finer split)
  return NB <= hash_bin_max_bins() && (u64)ctx->prm.nb_partitions * NB <= (1ULL << 22);
}

static int count_hash_binned(Lane* ln, uint32_t sample, uint32_t hard_min, const u32* win_sample = nullptr, const u32* win_part = nullptr)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  const u64 Wb = ctx->prm.window_bits;
  const u32 TR = hash_bin_tile_records();
  bool h16; double slack;
  { std::lock_guard<std::mutex> g(ctx->mu); h16 = ctx->hist16 && !kmx_env_flag("KMX_HIST32"); slack = ctx->bin_slack; }
  CK(ensure_pin(ln, (size_t)P * 32 + 512));
  char* hp = ln->h_pin;
  u64* r_loff = (u64*)hp;                                   // read-back: list_off[P] | meta[4] | flags[4]
  u64* u_meta = (u64*)(hp + (size_t)P * 8 + 64);            // upload: meta[4] | flags[4]
  u64* u_base = (u64*)(hp + (size_t)P * 8 + 128);           // upload: win_base[P] | tile_pref[P+1] | win_cap[P] | win_part[P]
  u32* u_tpref = (u32*)(u_base + P); u32* u_cap = u_tpref + (P + 1); u32* u_wpart = u_cap + P;
  const size_t up_bytes = (((size_t)P * 8 + ((size_t)3 * P + 1) * 4) + 7) & ~(size_t)7;
  u64 cap = std::max<u64>(4096, ln->d_est);
  for (int attempt = 0; attempt < 8; attempt++) {
    const u32 bs_log = h16 ? 15u : 14u;
    const u32 NB = (u32)((Wb + (1ULL << bs_log) - 1) >> bs_log);
    const size_t items = (size_t)P * NB;
    u64 ent = 0, tiles = 0;
    for (u32 v = 0; v < P; v++) {
      const u64 kc = ln->h_kcnt[v];
      u64 c = kc ? (u64)((double)kc / NB * slack) + 512 : 0;
      c = (c + 7) & ~(u64)7;
      if (c * NB > 0xFFFFFFF0ULL) return KMX_BIN_FALLBACK;    // in-window entry offsets are 32-bit
      u_base[v] = ent; u_cap[v] = (u32)c; ent += c * NB;
      u_tpref[v] = (u32)tiles; tiles += ((u64)ln->h_cursor[v] + TR - 1) / TR;
    }
    if (tiles >= 0x7FFFFFF0ULL) return KMX_BIN_FALLBACK;
    u_tpref[P] = (u32)tiles;
    if (win_part) memcpy(u_wpart, win_part, (size_t)P * 4);
    CK(ensure(ln, ln->binbuf, (size_t)((double)ent * 2.2) + 256));   // 10 % head-room: the next samples differ by a few per cent, no regrowth
    // device meta: [status u64[items] | bin_cursor u32[items] | tickets u32[2]] zeroed per sample; layout upload; results
    const size_t zbytes = (items * 12 + 8 + 15) & ~(size_t)15;
    CK(ensure(ln, ln->binmeta, zbytes + up_bytes + (size_t)P * 8 + 64));
    char* dm = (char*)ln->binmeta.p;
    void* kp = nullptr; void* cp = nullptr;
    CK(list_alloc(ln, cap * 8, &kp));
    CK(list_alloc(ln, cap * 4, &cp));
    HashBinArgs a;
    a.records = ln->records.p; a.boff = ln->d_boff; a.bcnt = ln->d_cursor;
    a.k = (int)ctx->prm.kmer_size; a.nwin = P; a.Wbits = Wb; a.NB = NB; a.bs_log = bs_log;
    a.status = (u64*)dm; a.bin_cursor = (u32*)(dm + items * 8); a.tickets = a.bin_cursor + items;
    a.win_base = (const u64*)(dm + zbytes); a.tile_pref = (const u32*)(a.win_base + P); a.win_cap = a.tile_pref + (P + 1);
    a.win_part = win_part ? a.win_cap + P : nullptr;
    a.list_off = (u64*)(dm + zbytes + up_bytes); a.meta = a.list_off + P; a.flags = (u32*)(a.meta + 4);
    a.binbuf = (uint16_t*)ln->binbuf.p; a.out_keys = (u64*)kp; a.out_counts = (u32*)cp; a.hard_min = hard_min ? hard_min : 1;
    u_meta[0] = 0; u_meta[1] = 0; u_meta[2] = cap; u_meta[3] = 0; u_meta[4] = 0; u_meta[5] = 0;
    { PROF(KMX_PROF_HASH_HIST);
      CK(cudaMemsetAsync(dm, 0, zbytes, ln->st));
      { SmallCopyBatch b(ln); b.add(dm + zbytes, u_base, up_bytes); b.add(a.meta, u_meta, 48); CK(b.go()); }
      CK(launch_hash_binned(a, (u32)tiles, 0, h16, ln->st, &ln->launches)); }
    { PROF(KMX_PROF_HASH_EMIT);
      CK(launch_hash_binned(a, (u32)tiles, 1, h16, ln->st, &ln->launches)); }
    { SmallCopyBatch b(ln); b.add(r_loff, a.list_off, (size_t)P * 8 + 48); CK(b.go()); }
    CK(cudaStreamSynchronize(ln->st));
    const u64 D = r_loff[P];
    const u32* fl = (const u32*)(r_loff + P + 4);
    if (fl[2]) {                                            // a bin region overflowed: more room, then the L2-histogram path
      slack *= 2;
      if (slack > 8.0) return KMX_BIN_FALLBACK;
      { std::lock_guard<std::mutex> g(ctx->mu); ctx->bin_slack = std::max(ctx->bin_slack, slack); }
      continue;
    }
    if (h16 && fl[1]) {                                     // a 16-bit counter wrapped: 32-bit counters from now on
      { std::lock_guard<std::mutex> g(ctx->mu); ctx->hist16 = false; }
      h16 = false;
      continue;
    }
    ln->d_est = std::max<u64>(ln->d_est, D + D / 4 + 1024);
    if (fl[0]) { cap = D + 1024; continue; }                // exact size now known
    for (u32 v = 0; v < P; v++) {
      if (win_part && ln->h_cursor[v] == 0) continue;       // unused window
      const u32 smp = win_sample ? win_sample[v] : sample, prt = win_part ? win_part[v] : v;
      ListRef& L = ctx->lists[(size_t)smp * P + prt];
      const u64 end = v + 1 < P ? r_loff[v + 1] : D;
      L.lo = (u64*)kp + r_loff[v]; L.hi = nullptr; L.cnt = (u32*)cp + r_loff[v]; L.n = end - r_loff[v];
    }
    ctx->stat[KMX_STAT_HASH_BINNED]++;
    return KMX_OK;
  }
  return fail(ln, KMX_ERR_CUDA, "binned hash-count did not converge");
}

static int count_sample(Lane* ln, uint32_t sample, uint32_t hard_min)
{
  kmx_ctx* ctx = ln->ctx;
  if (sample >= ctx->prm.nb_samples) return fail(ln, KMX_ERR_ARG, "sample %u >= nb_samples", sample);
  if (!ln->sample_ready) return fail(ln, KMX_ERR_STATE, "kmx_count_sample needs a sample finished by kmx_superk_end");
  if (hash_binned_usable(ctx)) {
    int rc = count_hash_binned(ln, sample, hard_min);
    if (rc != KMX_BIN_FALLBACK) return rc;
  }
  if (ctx->prm.key_kind == KMX_KEY_HASH) {
    size_t hist_bytes = (size_t)ctx->prm.nb_partitions * ctx->prm.window_bits * 4;
    int ok;
    {
      std::lock_guard<std::mutex> g(ctx->mu);
      if (ctx->hist_ok < 0) {                      // decide once: histogram path if it fits comfortably
        size_t free_b = 0, tot_b = 0;
        cudaMemGetInfo(&free_b, &tot_b);
        (void)hist_bytes;
        ctx->hist_ok = ctx->prm.window_bits * 4 < (free_b / 8) ? 1 : 0;   // one window must fit comfortably
      }
      ok = ctx->hist_ok;
    }
    if (ok == 1) return count_hash_hist(ln, sample, hard_min);
  }
  if (ctx->prm.key_kind == KMX_KEY_KMER && !kmx_env_flag("KMX_NO_HT")) {
    int rc = count_kmer_ht(ln, sample, hard_min);
    if (rc != KMX_HT_FALLBACK) return rc;
  }
  return count_generic(ln, sample, hard_min);
}

extern "C" int kmx_count_sample(kmx_ctx* ctx, uint32_t sample, uint32_t hard_min) { if (!ctx) return KMX_ERR_ARG; LANE0; return count_sample(ln, sample, hard_min); }

// ---- many samples, pipelined over lanes ---------------------------------------------------
extern "C" int kmx_run_samples(kmx_ctx* ctx, uint32_t n, const char* const* texts, const size_t* nbytes, int on_device,
                               const uint32_t* sample_ids, const uint32_t* hard_min, uint32_t nlanes,
                               uint64_t* kmers_per_partition)
{
  if (!ctx || (n && (!texts || !nbytes || !hard_min))) return KMX_ERR_ARG;
  if (n == 0) return KMX_OK;
  if (nlanes < 1) nlanes = 1;
  if (nlanes > 8) nlanes = 8;
  if (nlanes > n) nlanes = n;
  const u32 P = ctx->prm.nb_partitions;
  cudaSetDevice(ctx->device);
  while (ctx->lanes.size() < nlanes) { int rc = lane_create(ctx, (int)ctx->lanes.size()); if (rc) return rc; }
  {                                                 // settle the hist/sort decision before threads start
    Lane* ln = ctx->lanes[0].get();
    if (ctx->prm.key_kind == KMX_KEY_HASH && ctx->hist_ok < 0) {
      size_t free_b = 0, tot_b = 0;
      CK(cudaMemGetInfo(&free_b, &tot_b));
      size_t hist_bytes = (size_t)P * ctx->prm.window_bits * 4;
      (void)hist_bytes;
      ctx->hist_ok = ctx->prm.window_bits * 4 * nlanes < (free_b / 4) ? 1 : 0;
    }
  }
  ctx->active_lanes = (int)nlanes;
  std::atomic<int> first_err(0);
  auto work = [&](u32 t) {
    cudaSetDevice(ctx->device);
    Lane* ln = ctx->lanes[t].get();
    for (u32 i = t; i < n && !first_err.load(); i += nlanes) {
      const u32 sid = sample_ids ? sample_ids[i] : i;
      int rc = superk_begin(ln);
      if (!rc) rc = superk_push_fastq(ln, texts[i], nbytes[i], on_device);
      if (!rc) rc = superk_end(ln, kmers_per_partition ? kmers_per_partition + (size_t)i * P : nullptr);
      if (!rc) rc = count_sample(ln, sid, hard_min[i]);
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
  ctx->active_lanes = 1;
  reap(ctx);                                      // every lane synchronised its stream before its thread ended
  return first_err.load();
}

// ---- k-mer abundance histogram of a sample (--hist; KHist::inc is called for EVERY distinct key, before the hard-min
// test: count_processor.hpp:61-70,135-146, histogram.hpp:52-71).  The survivor lists do not hold the keys below hard-min, so
// the sample is first counted with hard-min 1 into lane scratch (all distinct keys), a kernel bins the counts, the scratch is
// dropped, and the sample is counted again with its real hard-min.  A side output, not on the timed path.
static const u32 KMX_HIST_MAXBINS = 1024;
__global__ void __launch_bounds__(256) count_hist_kernel(const MergeList* __restrict__ lists, u32 lower, u32 upper, u64* __restrict__ out)
{
  __shared__ u32 s_u[KMX_HIST_MAXBINS];
  __shared__ unsigned long long s_n[KMX_HIST_MAXBINS], s_acc[6];
  const u32 nb = upper - lower + 1u;
  for (u32 i = threadIdx.x; i < nb; i += 256) { s_u[i] = 0; s_n[i] = 0; }
  if (threadIdx.x < 6) s_acc[threadIdx.x] = 0;
  __syncthreads();
  const MergeList L = lists[blockIdx.y];
  unsigned long long a[6] = {0, 0, 0, 0, 0, 0};            // uniq, total, oob_lu, oob_ln, oob_uu, oob_un
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < L.n; i += (u64)gridDim.x * 256) {
    const u32 c = L.cnt[i];
    a[0]++; a[1] += c;
    if (c < lower) { a[2]++; a[3] += c; }
    else if (c > upper) { a[4]++; a[5] += c; }
    else { atomicAdd(&s_u[c - lower], 1u); atomicAdd(&s_n[c - lower], (unsigned long long)c); }
  }
  for (int q = 0; q < 6; q++) if (a[q]) atomicAdd(&s_acc[q], a[q]);
  __syncthreads();
  if (threadIdx.x < 6 && s_acc[threadIdx.x]) atomicAdd((unsigned long long*)out + threadIdx.x, s_acc[threadIdx.x]);
  for (u32 i = threadIdx.x; i < nb; i += 256) {
    if (s_u[i]) atomicAdd((unsigned long long*)out + 6 + i, (unsigned long long)s_u[i]);
    if (s_n[i]) atomicAdd((unsigned long long*)out + 6 + nb + i, s_n[i]);
  }
}

static int count_sample_hist(Lane* ln, uint32_t sample, uint32_t hard_min, uint32_t lower, uint32_t upper, uint64_t* out)
{
  kmx_ctx* ctx = ln->ctx;
  const u32 P = ctx->prm.nb_partitions;
  if (!out || lower < 1 || upper < lower || upper - lower + 1 > KMX_HIST_MAXBINS) return fail(ln, KMX_ERR_ARG, "histogram range [%u, %u] unsupported (1 <= lower <= upper, at most %u bins)", lower, upper, KMX_HIST_MAXBINS);
  if (sample >= ctx->prm.nb_samples) return fail(ln, KMX_ERR_ARG, "sample %u >= nb_samples", sample);
  const u32 nb = upper - lower + 1u;
  const bool one_pass = hard_min <= 1;                       // the lists already hold every distinct key
  ln->to_scratch = !one_pass;
  const u64 d_est0 = ln->d_est;
  int rc = count_sample(ln, sample, 1);
  ln->to_scratch = false;
  if (!one_pass) ln->d_est = d_est0;                         // the all-keys pass must not inflate the output estimate of the real passes
  if (!rc) {
    const size_t tab = (size_t)P * sizeof(MergeList), nout = (size_t)(6 + 2 * nb) * 8;
    cudaError_t e = ensure(ln, ln->hist_dev, tab + nout);
    if (e == cudaSuccess) e = ensure_pin(ln, std::max(tab, nout) + 256);
    if (e != cudaSuccess) rc = fail(ln, KMX_ERR_NOMEM, "histogram buffers: %s", cudaGetErrorString(e));
    if (!rc) {
      MergeList* hl = (MergeList*)ln->h_pin;
      u64 max_n = 0;
      for (u32 p = 0; p < P; p++) { const ListRef& L = ctx->lists[(size_t)sample * P + p]; hl[p].lo = L.lo; hl[p].hi = L.hi; hl[p].cnt = L.cnt; hl[p].n = L.n; max_n = std::max(max_n, L.n); }
      u64* d_out = (u64*)((char*)ln->hist_dev.p + tab);
      SmallCopyBatch b(ln); b.add(ln->hist_dev.p, hl, tab);
      e = b.go();
      if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, nout, ln->st);
      if (e == cudaSuccess && max_n) {
        const unsigned gx = (unsigned)std::min<u64>(64, (max_n + 255) / 256);
        count_hist_kernel<<<dim3(gx, P), 256, 0, ln->st>>>((const MergeList*)ln->hist_dev.p, lower, upper, d_out);
        ln->launches += 1;
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(ln->st);              // the table staging is free again
      if (e == cudaSuccess) { SmallCopyBatch b2(ln); b2.add(ln->h_pin, d_out, nout); e = b2.go(); }
      if (e == cudaSuccess) e = cudaStreamSynchronize(ln->st);
      if (e != cudaSuccess) rc = fail(ln, KMX_ERR_CUDA, "histogram pass: %s", cudaGetErrorString(e));
      else memcpy(out, ln->h_pin, nout);
    }
  }
  for (auto& b : ln->sarena) b.used = 0;                     // the all-keys lists are dropped (their blocks are kept for the next sample)
  if (!one_pass) {
    for (u32 p = 0; p < P; p++) ctx->lists[(size_t)sample * P + p] = ListRef();
    if (!rc) rc = count_sample(ln, sample, hard_min);
  }
  return rc;
}

// ---- lane-addressed entry points (one host thread per lane)
extern "C" int kmx_lanes(kmx_ctx* ctx, uint32_t n)
{
  if (!ctx || n < 1 || n > 8) return KMX_ERR_ARG;
  cudaSetDevice(ctx->device);
  std::lock_guard<std::mutex> g(ctx->lanes_mu);
  while (ctx->lanes.size() < n) { int rc = lane_create(ctx, (int)ctx->lanes.size()); if (rc) return rc; }
  ctx->active_lanes = std::max(ctx->active_lanes, (int)n);
  if (ctx->prm.key_kind == KMX_KEY_HASH && ctx->hist_ok < 0) {
    size_t free_b = 0, tot_b = 0;
    if (cudaMemGetInfo(&free_b, &tot_b) != cudaSuccess) return KMX_ERR_CUDA;
    ctx->hist_ok = ctx->prm.window_bits * 4 * n < (free_b / 4) ? 1 : 0;
  }
  return KMX_OK;
}
#define LANE_N(idx) if (!ctx || (idx) >= ctx->lanes.size()) return KMX_ERR_ARG; cudaSetDevice(ctx->device); Lane* ln = ctx->lanes[idx].get()
extern "C" int kmx_lane_superk_begin(kmx_ctx* ctx, uint32_t lane) { LANE_N(lane); return superk_begin(ln); }
extern "C" int kmx_lane_superk_push_fastq(kmx_ctx* ctx, uint32_t lane, const char* text, size_t nbytes, int on_device)
{
  if (!text && nbytes) return KMX_ERR_ARG;
  LANE_N(lane); return superk_push_fastq(ln, text, nbytes, on_device);
}
extern "C" int kmx_lane_superk_push_reads(kmx_ctx* ctx, uint32_t lane, const char* seqs, const uint64_t* off, size_t nseq)
{
  if (nseq && (!seqs || !off)) return KMX_ERR_ARG;
  LANE_N(lane); return superk_push_reads(ln, seqs, off, nseq);
}
extern "C" int kmx_lane_superk_end(kmx_ctx* ctx, uint32_t lane, uint64_t* kmers_per_partition) { LANE_N(lane); return superk_end(ln, kmers_per_partition); }
extern "C" int kmx_lane_count_sample(kmx_ctx* ctx, uint32_t lane, uint32_t sample, uint32_t hard_min) { LANE_N(lane); return count_sample(ln, sample, hard_min); }
extern "C" int kmx_lane_count_sample_hist(kmx_ctx* ctx, uint32_t lane, uint32_t sample, uint32_t hard_min, uint32_t lower, uint32_t upper, uint64_t* out)
{
  LANE_N(lane);
  if (!ln->sample_ready) return fail(ln, KMX_ERR_STATE, "kmx_lane_count_sample_hist needs a sample finished by kmx_lane_superk_end");
  return count_sample_hist(ln, sample, hard_min, lower, upper, out);
}

extern "C" int kmx_counts_size(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint64_t* n)
{
  if (!ctx || !n || sample >= ctx->prm.nb_samples || partition >= ctx->prm.nb_partitions) return KMX_ERR_ARG;
  *n = ctx->lists[(size_t)sample * ctx->prm.nb_partitions + partition].n;
  return KMX_OK;
}

extern "C" int kmx_counts_get(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint64_t* keys, uint32_t* counts)
{
  if (!ctx || sample >= ctx->prm.nb_samples || partition >= ctx->prm.nb_partitions) return KMX_ERR_ARG;
  LANE0;
  int rc = kmx_sync(ctx);                           // lists may have been produced on another lane
  if (rc) return rc;
  const ListRef& L = ctx->lists[(size_t)sample * ctx->prm.nb_partitions + partition];
  if (L.n == 0) return KMX_OK;
  const bool two = (ctx->prm.key_kind == KMX_KEY_KMER && ctx->W == 2);
  if (keys) {
    if (!two) CK(cudaMemcpyAsync(keys, L.lo, L.n * 8, cudaMemcpyDeviceToHost, ln->st));
    else {
      CK(cudaMemcpy2DAsync(keys, 16, L.lo, 8, 8, L.n, cudaMemcpyDeviceToHost, ln->st));
      CK(cudaMemcpy2DAsync(keys + 1, 16, L.hi, 8, 8, L.n, cudaMemcpyDeviceToHost, ln->st));
    }
  }
  if (counts) CK(cudaMemcpyAsync(counts, L.cnt, L.n * 4, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}

extern "C" int kmx_counts_put(kmx_ctx* ctx, uint32_t sample, uint32_t partition, const uint64_t* keys,
                              const uint32_t* counts, uint64_t n)
{
  if (!ctx || sample >= ctx->prm.nb_samples || partition >= ctx->prm.nb_partitions || (n && (!keys || !counts))) return KMX_ERR_ARG;
  LANE0;
  ListRef& L = ctx->lists[(size_t)sample * ctx->prm.nb_partitions + partition];
  L = ListRef();
  if (n == 0) return KMX_OK;
  const bool two = (ctx->prm.key_kind == KMX_KEY_KMER && ctx->W == 2);
  void* kp = nullptr; void* hp = nullptr; void* cp = nullptr;
  CK(arena_alloc(ctx, n * 8, &kp));
  CK(arena_alloc(ctx, n * 4, &cp));
  if (!two) CK(cudaMemcpyAsync(kp, keys, n * 8, cudaMemcpyHostToDevice, ln->st));
  else {
    CK(arena_alloc(ctx, n * 8, &hp));
    CK(cudaMemcpy2DAsync(kp, 8, keys, 16, 8, n, cudaMemcpyHostToDevice, ln->st));
    CK(cudaMemcpy2DAsync(hp, 8, keys + 1, 16, 8, n, cudaMemcpyHostToDevice, ln->st));
  }
  CK(cudaMemcpyAsync(cp, counts, n * 4, cudaMemcpyHostToDevice, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  L.lo = (u64*)kp; L.hi = (u64*)hp; L.cnt = (u32*)cp; L.n = n;
  return KMX_OK;
}

extern "C" int kmx_counts_vector(kmx_ctx* ctx, uint32_t sample, uint32_t partition, uint8_t* bits)
{
  if (!ctx || !bits || sample >= ctx->prm.nb_samples || partition >= ctx->prm.nb_partitions) return KMX_ERR_ARG;
  LANE0;
  if (ctx->prm.key_kind != KMX_KEY_HASH) return fail(ln, KMX_ERR_ARG, "kmx_counts_vector needs hash keys");
  int rc = kmx_sync(ctx);
  if (rc) return rc;
  const u64 Wb = ctx->prm.window_bits;
  const ListRef& L = ctx->lists[(size_t)sample * ctx->prm.nb_partitions + partition];
  CK(ensure(ln, ctx->body, Wb / 8 + 8));
  CK(cudaMemsetAsync(ctx->body.p, 0, Wb / 8 + 8, ln->st));
  CK(launch_hash_vector(L.lo, L.n, Wb * partition, (uint8_t*)ctx->body.p, ln->st, &ln->launches));
  CK(cudaMemcpyAsync(bits, ctx->body.p, Wb / 8, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}

// ---------------------------------------------------------------------------------------
// stage 3 / 4
// ---------------------------------------------------------------------------------------
static int merge_sparse(Lane* ln, uint32_t partition, const kmx_merge_params* mp, kmx_merge_result* res,
                        const std::vector<MergeList>& hl, u64 max_n, u64 tot_n);

extern "C" int kmx_merge_partition(kmx_ctx* ctx, uint32_t partition, const kmx_merge_params* mp, kmx_merge_result* res)
{
  if (!ctx || !mp || !mp->soft_min || partition >= ctx->prm.nb_partitions) return KMX_ERR_ARG;
  LANE0;
  const u32 N = ctx->prm.nb_samples, P = ctx->prm.nb_partitions;
  const bool hash = ctx->prm.key_kind == KMX_KEY_HASH;
  if ((mp->format == KMX_FMT_BF || mp->format == KMX_FMT_BFT) && !hash) return fail(ln, KMX_ERR_ARG, "bf/bft rows need hash keys");
  if (mp->format > KMX_FMT_BFT) return fail(ln, KMX_ERR_ARG, "bad format");
  if (mp->emit_all && mp->format != KMX_FMT_COUNT) return fail(ln, KMX_ERR_ARG, "emit_all needs KMX_FMT_COUNT");
  for (size_t i = 1; i < ctx->lanes.size(); i++) { Lane* o = ctx->lanes[i].get(); cudaError_t e = cudaStreamSynchronize(o->st); if (e != cudaSuccess) return fail(ln, KMX_ERR_CUDA, "lane sync: %s", cudaGetErrorString(e)); }
  CK(ensure_pin(ln, N * sizeof(MergeList) + N * 4 + 256));
  MergeList* hl_pin = (MergeList*)ln->h_pin;
  u32* soft_pin = (u32*)(ln->h_pin + N * sizeof(MergeList));
  std::vector<MergeList> hl(N);
  u64 max_n = 0, tot_n = 0;
  for (u32 s = 0; s < N; s++) {
    const ListRef& L = ctx->lists[(size_t)s * P + partition];
    hl[s].lo = L.lo; hl[s].hi = L.hi; hl[s].cnt = L.cnt; hl[s].n = L.n;
    max_n = std::max(max_n, L.n); tot_n += L.n;
  }
  memcpy(hl_pin, hl.data(), N * sizeof(MergeList)); memcpy(soft_pin, mp->soft_min, N * 4);
  CK(ensure(ln, ctx->d_lists, N * sizeof(MergeList)));
  CK(ensure(ln, ctx->d_soft, N * 4));
  CK(ensure(ln, ctx->stats, (size_t)6 * N * 8));
  { SmallCopyBatch b(ln); b.add(ctx->d_lists.p, hl_pin, N * sizeof(MergeList)); b.add(ctx->d_soft.p, soft_pin, N * 4); CK(b.go()); }
  CK(cudaMemsetAsync(ctx->stats.p, 0, (size_t)6 * N * 8, ln->st));
  ctx->last_emit_all = mp->emit_all;
  if (mp->format == KMX_FMT_COUNT || mp->format == KMX_FMT_PA) return merge_sparse(ln, partition, mp, res, hl, max_n, tot_n);
  // dense Bloom slab
  const u64 Wb = ctx->prm.window_bits;
  const u32 rb = (N + 7) / 8;
  const size_t slab_bytes = (size_t)Wb * rb;
  uint8_t* slab = nullptr;
  if (ctx->merge_out && mp->format == KMX_FMT_BF) {
    if (ctx->merge_out_cap < slab_bytes + 8) return fail(ln, KMX_ERR_ARG, "merge output buffer too small (%zu < %zu)", ctx->merge_out_cap, slab_bytes + 8);
    slab = (uint8_t*)ctx->merge_out;
  } else { CK(ensure(ln, ctx->body, slab_bytes + 8)); slab = (uint8_t*)ctx->body.p; }
  { PROF(KMX_PROF_FILL); CK(cudaMemsetAsync(slab, 0, slab_bytes + 8, ln->st)); }
  const bool need_si = mp->share_min != 0 || mp->recurrence_min > 1;
  u32* si = nullptr;
  {
    PROF(KMX_PROF_MERGE);
    if (need_si) {
      CK(ensure(ln, ctx->solid_in, Wb * 4));
      CK(cudaMemsetAsync(ctx->solid_in.p, 0, Wb * 4, ln->st));
      si = (u32*)ctx->solid_in.p;
      CK(launch_dense_solid((const MergeList*)ctx->d_lists.p, N, (const u32*)ctx->d_soft.p, Wb * partition, si, max_n, ln->st, &ln->launches));
    }
    CK(launch_dense_emit((const MergeList*)ctx->d_lists.p, N, (const u32*)ctx->d_soft.p, mp->recurrence_min, mp->share_min,
                         Wb * partition, si, slab, rb, (u64*)ctx->stats.p, max_n, ln->st, &ln->launches));
  }
  ctx->last_body = slab;
  ctx->last_res.n_rows = Wb; ctx->last_res.row_bytes = rb; ctx->last_res.n_union = 0;
  if (mp->format == KMX_FMT_BFT) {
    // W x (8*rb) bits -> (8*rb) x W bits
    uint8_t* tout;
    if (ctx->merge_out) {
      if (ctx->merge_out_cap < slab_bytes + 8) return fail(ln, KMX_ERR_ARG, "merge output buffer too small");
      tout = (uint8_t*)ctx->merge_out;
    } else { CK(ensure(ln, ctx->body2, slab_bytes + 8)); tout = (uint8_t*)ctx->body2.p; }
    { PROF(KMX_PROF_TRANSPOSE); CK(launch_transpose_bits(slab, Wb, (u64)rb * 8, tout, ln->st, &ln->launches)); }
    ctx->last_body = tout;
    ctx->last_res.n_rows = (u64)rb * 8; ctx->last_res.row_bytes = Wb / 8;
  }
  // the pinned staging (lists, soft_min) is reused by the next merge: its H2D copies must be done.
  CK(cudaStreamSynchronize(ln->st));
  if (res) *res = ctx->last_res;
  return KMX_OK;
}

extern "C" int kmx_merge_get(kmx_ctx* ctx, void* body, uint64_t* stats, uint8_t* row_keep)
{
  if (!ctx) return KMX_ERR_ARG;
  LANE0;
  const u64 nb = ctx->last_res.n_rows * ctx->last_res.row_bytes;
  if (body && nb) CK(cudaMemcpyAsync(body, ctx->last_body, nb, cudaMemcpyDeviceToHost, ln->st));
  if (stats) CK(cudaMemcpyAsync(stats, ctx->stats.p, (size_t)6 * ctx->prm.nb_samples * 8, cudaMemcpyDeviceToHost, ln->st));
  if (row_keep && ctx->last_emit_all && ctx->last_res.n_rows) CK(cudaMemcpyAsync(row_keep, ctx->row_keep.p, ctx->last_res.n_rows, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}

extern "C" const void* kmx_merge_body_device(kmx_ctx* ctx) { return ctx ? ctx->last_body : nullptr; }

extern "C" int kmx_transpose_bits(kmx_ctx* ctx, const uint8_t* in, uint64_t nrows, uint64_t ncols, uint8_t* out)
{
  if (!ctx || !in || !out || (nrows % 8) || (ncols % 8)) return KMX_ERR_ARG;
  LANE0;
  const size_t nb = (size_t)(nrows * ncols / 8);
  if (nb == 0) return KMX_OK;
  CK(ensure(ln, ctx->body, nb + 8));
  CK(ensure(ln, ctx->body2, nb + 8));
  CK(cudaMemcpyAsync(ctx->body.p, in, nb, cudaMemcpyHostToDevice, ln->st));
  { PROF(KMX_PROF_TRANSPOSE); CK(launch_transpose_bits((const uint8_t*)ctx->body.p, nrows, ncols, (uint8_t*)ctx->body2.p, ln->st, &ln->launches)); }
  CK(cudaMemcpyAsync(out, ctx->body2.p, nb, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}

// ---------------------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------------------
extern "C" int kmx_synth_fastq(kmx_ctx* ctx, uint64_t seed, uint32_t sample, uint64_t first_read, uint64_t R,
                               uint32_t L, uint64_t G, double d, double e, int revcomp, char* dev_out)
{
  if (!ctx || !dev_out || G < L) return KMX_ERR_ARG;
  LANE0;
  auto thr = [](double p) { double v = p * 4294967296.0; return v >= 4294967295.0 ? 0xFFFFFFFFu : (u32)v; };
  u64 dummy = 0;                                  // the generator is a test utility, not a hot-path launch
  CK(launch_synth_fastq(seed, sample, first_read, R, L, G, thr(d), thr(e), revcomp, dev_out, ln->st, &dummy));
  return KMX_OK;
}

extern "C" int kmx_dev_alloc(kmx_ctx* ctx, size_t nbytes, void** dev_ptr)
{
  if (!ctx || !dev_ptr) return KMX_ERR_ARG;
  LANE0;
  CK(cudaMalloc(dev_ptr, nbytes ? nbytes : 1));
  ctx->user_allocs.push_back(*dev_ptr);
  return KMX_OK;
}
extern "C" int kmx_dev_free(kmx_ctx* ctx, void* dev_ptr)
{
  if (!ctx) return KMX_ERR_ARG;
  LANE0;
  auto it = std::find(ctx->user_allocs.begin(), ctx->user_allocs.end(), dev_ptr);
  if (it == ctx->user_allocs.end()) return KMX_ERR_ARG;
  ctx->user_allocs.erase(it);
  int rc = kmx_sync(ctx);
  if (rc) return rc;
  CK(cudaFree(dev_ptr));
  return KMX_OK;
}
extern "C" int kmx_memcpy_d2h(kmx_ctx* ctx, void* host, const void* dev, size_t nbytes)
{
  if (!ctx) return KMX_ERR_ARG;
  LANE0;
  CK(cudaMemcpyAsync(host, dev, nbytes, cudaMemcpyDeviceToHost, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}
extern "C" int kmx_memcpy_h2d(kmx_ctx* ctx, void* dev, const void* host, size_t nbytes)
{
  if (!ctx) return KMX_ERR_ARG;
  LANE0;
  CK(cudaMemcpyAsync(dev, host, nbytes, cudaMemcpyHostToDevice, ln->st));
  CK(cudaStreamSynchronize(ln->st));
  return KMX_OK;
}
extern "C" int kmx_host_alloc(size_t nbytes, void** host_ptr)
{
  if (!host_ptr) return KMX_ERR_ARG;
  return cudaMallocHost(host_ptr, nbytes ? nbytes : 1) == cudaSuccess ? KMX_OK : KMX_ERR_NOMEM;
}
extern "C" int kmx_host_free(void* host_ptr) { return cudaFreeHost(host_ptr) == cudaSuccess ? KMX_OK : KMX_ERR_CUDA; }

extern "C" int kmx_set_merge_output(kmx_ctx* ctx, void* dev_ptr, size_t cap_bytes)
{
  if (!ctx) return KMX_ERR_ARG;
  ctx->merge_out = dev_ptr; ctx->merge_out_cap = dev_ptr ? cap_bytes : 0;
  return KMX_OK;
}

// ---- repartition estimate (RepartTask, task.hpp:170-222 -> RepartitionAlgorithm.cpp:395-492 samples the banks on the CPU)
extern "C" int kmx_minimizer_load_enable(kmx_ctx* ctx, int on)
{
  if (!ctx) return KMX_ERR_ARG;
  LANE0;
  int rc = kmx_sync(ctx);
  if (rc) return rc;
  const size_t tn = (size_t)1 << (2 * ctx->prm.minim_size);
  if (on) {
    if (!ctx->d_mload) { CK(cudaMalloc(&ctx->d_mload, tn * 8)); add_bytes(ctx, (long long)(tn * 8)); }
    CK(cudaMemsetAsync(ctx->d_mload, 0, tn * 8, ln->st));
    CK(cudaStreamSynchronize(ln->st));
  } else if (ctx->d_mload) { CK(cudaFree(ctx->d_mload)); ctx->d_mload = nullptr; add_bytes(ctx, -(long long)(tn * 8)); }
  return KMX_OK;
}
extern "C" int kmx_minimizer_load_get(kmx_ctx* ctx, uint64_t* load)
{
  if (!ctx || !load) return KMX_ERR_ARG;
  LANE0;
  if (!ctx->d_mload) return fail(ln, KMX_ERR_STATE, "kmx_minimizer_load_get without kmx_minimizer_load_enable");
  int rc = kmx_sync(ctx);
  if (rc) return rc;
  CK(cudaMemcpy(load, ctx->d_mload, ((size_t)1 << (2 * ctx->prm.minim_size)) * 8, cudaMemcpyDeviceToHost));
  return KMX_OK;
}

extern "C" int kmx_profile_enable(kmx_ctx* ctx, int on)
{
  if (!ctx) return KMX_ERR_ARG;
  prof_collect(ctx);
  ctx->prof_on = on != 0;
  return KMX_OK;
}
extern "C" int kmx_profile_reset(kmx_ctx* ctx)
{
  if (!ctx) return KMX_ERR_ARG;
  prof_collect(ctx);
  for (int i = 0; i < KMX_PROF_KINDS; i++) { ctx->prof_ms[i] = 0; ctx->prof_cnt[i] = 0; }
  return KMX_OK;
}
extern "C" int kmx_profile_get(kmx_ctx* ctx, int kind, double* total_ms, uint64_t* count)
{
  if (!ctx || kind < 0 || kind >= KMX_PROF_KINDS) return KMX_ERR_ARG;
  prof_collect(ctx);
  if (total_ms) *total_ms = ctx->prof_ms[kind];
  if (count) *count = ctx->prof_cnt[kind];
  return KMX_OK;
}

#include "kmx_generic.inl"
#include "kmx_dist.inl"
