// s2_hash.cu -- stage 2, hash keys, histogram path.
//
// Replaces (behaviour, not code) ReadSuperkHash / HashSort / HashPartCounter::executeDump
// (include/kmtricks/gatb/sorting_count.hpp:387-470,525-528,971-990), KmXXHash (:346-363) and
// HashCountProcessor (include/kmtricks/gatb/count_processor.hpp:61-70).
//
// The reference sorts all hashes of a (sample, partition) and run-length counts them.  Hash
// keys live in a dense window [W*p, W*(p+1)), so here each super-k-mer record is expanded in
// registers (every k-mer extracted independently from the packed record, canonical via
// bit-reversal, XXH64, exact multiply-shift modulo) and counted with one L2 atomic into a
// W-slot u32 histogram of the partition.  The add that lifts a slot to `hard_min` also bumps
// the survivor counter of the slot's 64K sub-chunk, so a single ordered sweep (hash_emit) can
// compact the survivors into the ascending (key,count) list and zero the histogram again.
#include "common.cuh"
#include "kmx_internal.h"
#include "records.cuh"

#include <algorithm>
#include <cstdlib>

namespace kmx {

// debugging / A-B switches read once from the environment (e.g. KMX_HIST_NOROLL=1)
bool kmx_env_flag(const char* name)
{
  const char* v = getenv(name);
  return v && *v && *v != '0';
}

static constexpr int HH_THREADS = 256;
static constexpr int HH_WARPS = HH_THREADS / 32;

// grid: (x tiles, P).  Each warp takes 32 records at a time, stages them in shared memory with
// the prefix sum of their k-mer counts, and then walks the k-mers of those records with all 32
// lanes busy (lane t -> k-mer t, t+32, ...; the owning record is found by a 5-step binary
// search in the warp's prefix array).  Every lane runs the same instruction stream (a rolling
// variant with per-lane contiguous chunks was measured 1.9x SLOWER: lanes cross record
// boundaries at different steps, so the warp pays extract + roll at every step).
// The histogram update is a fire-and-forget RED; the modulo is an exact 32-bit Barrett when W < 2^32.
// Counter width.  H16: two 16-bit counters per 32-bit word (slot s = half s&1 of word s>>1), incremented by a RED of
// 1 << 16*(s&1).  A counter that reaches 65536 wraps (and, from the low half, carries into its neighbour): the sweep
// detects ANY such event exactly, because it makes the sum of all fields of the window fall short of the number of
// k-mers inserted (known from stage 1); the sample is then redone with 32-bit counters.  Half the histogram bytes.
template <bool H16>
__device__ __forceinline__ void hist_inc(u32* __restrict__ h, u64 key)
{
  if (H16) atomicAdd(h + (key >> 1), 1u << (16u * (u32)(key & 1ULL)));   // result unused -> RED.ADD
  else atomicAdd(h + key, 1u);
}

template <int W, bool D32, bool H16>
__global__ void __launch_bounds__(HH_THREADS)
hash_hist_kernel(const uint4* __restrict__ recs, const u64* __restrict__ boff, const u32* __restrict__ bcnt,
                 int k, u64 Wbits, FastMod64 fm, FastMod32 fm32, u32* __restrict__ hist, u32 p0)
{
  __shared__ uint4 s_rec[HH_WARPS][32 * W];
  __shared__ u32 s_pref[HH_WARPS][33];
  const u32 p = p0 + blockIdx.y;                 // window blockIdx.y of the group buffer holds partition p
  const u32 n = bcnt[p];
  const u64 b0 = boff[p];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32* __restrict__ h = hist + (u64)blockIdx.y * (Wbits >> (H16 ? 1 : 0));
  for (u32 r0 = (blockIdx.x * HH_WARPS + w) * 32; r0 < n; r0 += gridDim.x * HH_THREADS) {
    const u32 r = r0 + lane;
    u32 nk = 0;
    if (r < n) {
      if (W == 1) {
        uint4 v = __ldg(recs + b0 + r);
        s_rec[w][lane] = v;
        nk = (v.w >> 24) - k + 1;
      } else {
        uint4 a = __ldg(recs + 2 * (b0 + r)), b = __ldg(recs + 2 * (b0 + r) + 1);
        s_rec[w][2 * lane] = a; s_rec[w][2 * lane + 1] = b;
        nk = (b.w >> 24) - k + 1;
      }
    }
    u32 x = nk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
    s_pref[w][lane + 1] = x;
    if (lane == 0) s_pref[w][0] = 0;
    const u32 T = __shfl_sync(0xffffffffu, x, 31);
    __syncwarp();
    for (u32 t = lane; t < T; t += 32) {
      // largest q in [0,32) with pref[q] <= t  (amortising the search over 4 k-mers per lane was measured slower)
      u32 q = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1) if (s_pref[w][q + st] <= t) q += st;
      const int j = (int)(t - s_pref[w][q]);
      u64 key;
      if (W == 1) {
        uint4 v = s_rec[w][q];
        Rec1 rec; rec.lo = (u64)v.x | ((u64)v.y << 32);
        u64 hh = (u64)v.z | ((u64)v.w << 32);
        rec.n = (int)(hh >> 56); rec.hi = hh & 0x00FFFFFFFFFFFFFFULL;
        u64 c; canon1(rec, k, j, c);
        { const u64 hv = xxh64_8(c); key = D32 ? (u64)fastmod64_d32(hv, fm32) : fastmod64(hv, fm); }
      } else {
        uint4 a = s_rec[w][2 * q], b = s_rec[w][2 * q + 1];
        Rec2 rec;
        rec.v0 = (u64)a.x | ((u64)a.y << 32); rec.v1 = (u64)a.z | ((u64)a.w << 32);
        rec.v2 = (u64)b.x | ((u64)b.y << 32);
        u64 hh = (u64)b.z | ((u64)b.w << 32);
        rec.n = (int)(hh >> 56); rec.v3 = hh & 0x00FFFFFFFFFFFFFFULL;
        u64 clo, chi; canon2(rec, k, j, clo, chi);
        { const u64 hv = xxh64_16(clo, chi); key = D32 ? (u64)fastmod64_d32(hv, fm32) : fastmod64(hv, fm); }
      }
      hist_inc<H16>(h, key);
    }
    __syncwarp();
  }
}

static constexpr int HC_THREADS = 256;
static constexpr int HC_WARPS = HC_THREADS / 32;
static constexpr u32 HC_SLICE = HIST_SUB / HC_WARPS;                   // 1024 slots per warp slice
static constexpr int HC_ROWS_PER_WARP = HC_SLICE / 128;                // 8
static_assert(HC_ROWS_PER_WARP * HC_WARPS * 128 == (int)HIST_SUB, "HIST_SUB must be 8 warps x whole rows");

struct SweepArgs {
  u32 CW;                 // chunks per window
  u32 hmin;
  uint16_t* st_idx; u32* st_cnt; u32* slice_counts; u32* chunk_counts;
  u32* done;              // fused kernel only: [gp] tiles finished per window
  u64* win_sum;           // 16-bit counters only: [gp] sum of all fields of the window (overflow check)
};

// chunk c (HIST_SUB slots of window c / CW) by one CTA of HC_THREADS threads; s_agg: HC_WARPS words of shared memory.
// L2ONLY: the window was just filled by REDs and is expected in L2 (ld.cg); otherwise a streaming read.
template <bool L2ONLY>
__device__ __forceinline__ void compact_chunk(u32 c, u64 Wbits, u32* __restrict__ hist, const SweepArgs& sa, bool touched, u32* s_agg)
{
  const u32 wl = c / sa.CW, sub = c - wl * sa.CW;
  const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (!touched) {                                     // untouched window: all-zero
    if (lane == 0) sa.slice_counts[(u64)c * HC_WARPS + w] = 0;
    if (threadIdx.x == 0) sa.chunk_counts[c] = 0;
    return;
  }
  const u32 hmin = sa.hmin;
  const u32 ltmask = (1u << lane) - 1u;
  uint4* __restrict__ h4 = reinterpret_cast<uint4*>(hist + (u64)wl * Wbits);       // W multiple of 64 -> aligned
  const u64 qend = Wbits / 4;
  const u64 q0 = ((u64)sub * HIST_SUB + (u64)w * HC_SLICE) / 4 + lane;
  uint4 v[HC_ROWS_PER_WARP];
#pragma unroll
  for (int j = 0; j < HC_ROWS_PER_WARP; j++) {
    const u64 q = q0 + (u64)j * 32;
    v[j] = make_uint4(0, 0, 0, 0);
    if (q < qend) v[j] = L2ONLY ? __ldcg(h4 + q) : __ldcs(h4 + q);
  }
  const u64 sbase = ((u64)c * HC_WARPS + w) * HC_SLICE;  // this slice's staging run
  u32 run = 0;
#pragma unroll
  for (int j = 0; j < HC_ROWS_PER_WARP; j++) {
    const bool sx = v[j].x >= hmin, sy = v[j].y >= hmin, sz = v[j].z >= hmin, sw = v[j].w >= hmin;
    const u32 bx = __ballot_sync(0xffffffffu, sx), by = __ballot_sync(0xffffffffu, sy);
    const u32 bz = __ballot_sync(0xffffffffu, sz), bw = __ballot_sync(0xffffffffu, sw);
    if (bx | by | bz | bw) {
      u32 o = run + __popc(bx & ltmask) + __popc(by & ltmask) + __popc(bz & ltmask) + __popc(bw & ltmask);
      const u32 si = (u32)j * 128 + lane * 4;           // slot offset inside the slice
      if (sx) { sa.st_idx[sbase + o] = (uint16_t)si; sa.st_cnt[sbase + o] = v[j].x; o++; }
      if (sy) { sa.st_idx[sbase + o] = (uint16_t)(si + 1); sa.st_cnt[sbase + o] = v[j].y; o++; }
      if (sz) { sa.st_idx[sbase + o] = (uint16_t)(si + 2); sa.st_cnt[sbase + o] = v[j].z; o++; }
      if (sw) { sa.st_idx[sbase + o] = (uint16_t)(si + 3); sa.st_cnt[sbase + o] = v[j].w; o++; }
      run += __popc(bx) + __popc(by) + __popc(bz) + __popc(bw);
    }
    if (v[j].x | v[j].y | v[j].z | v[j].w) h4[q0 + (u64)j * 32] = make_uint4(0, 0, 0, 0);
  }
  if (lane == 0) { s_agg[w] = run; sa.slice_counts[(u64)c * HC_WARPS + w] = run; }
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < HC_WARPS; i++) t += s_agg[i]; sa.chunk_counts[c] = t; }
}

// 16-bit counters: a uint4 holds 8 slots, a row of 32 lanes 256 slots, a 1024-slot slice 4 rows.  Besides compacting,
// every warp adds the sum of all its fields to win_sum[window] (the overflow check, see hist_inc).
static constexpr int HC16_ROWS_PER_WARP = HC_SLICE / 256;              // 4
__device__ __forceinline__ void compact_chunk16(u32 c, u64 Wbits, u32* __restrict__ hist, const SweepArgs& sa, bool touched, u32* s_agg)
{
  const u32 wl = c / sa.CW, sub = c - wl * sa.CW;
  const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (!touched) {                                     // untouched window: all-zero
    if (lane == 0) sa.slice_counts[(u64)c * HC_WARPS + w] = 0;
    if (threadIdx.x == 0) sa.chunk_counts[c] = 0;
    return;
  }
  const u32 hmin = sa.hmin;
  const u32 ltmask = (1u << lane) - 1u;
  uint4* __restrict__ h4 = reinterpret_cast<uint4*>(hist + (u64)wl * (Wbits >> 1));   // W multiple of 64 -> 128-byte aligned
  const u64 qend = Wbits / 8;
  const u64 q0 = ((u64)sub * HIST_SUB + (u64)w * HC_SLICE) / 8 + lane;
  uint4 v[HC16_ROWS_PER_WARP];
#pragma unroll
  for (int j = 0; j < HC16_ROWS_PER_WARP; j++) {
    const u64 q = q0 + (u64)j * 32;
    v[j] = make_uint4(0, 0, 0, 0);
    if (q < qend) v[j] = __ldcs(h4 + q);
  }
  const u64 sbase = ((u64)c * HC_WARPS + w) * HC_SLICE;  // this slice's staging run
  u32 run = 0, fsum = 0;
#pragma unroll
  for (int j = 0; j < HC16_ROWS_PER_WARP; j++) {
    const u32 wv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
    u32 f[8];
    u32 smask = 0;                                      // bit q: field q of this lane survives
#pragma unroll
    for (int q = 0; q < 8; q++) {
      f[q] = (q & 1) ? (wv[q >> 1] >> 16) : (wv[q >> 1] & 0xFFFFu);
      fsum += f[q];
      smask |= (u32)(f[q] >= hmin) << q;
    }
    if (__any_sync(0xffffffffu, smask != 0)) {
      // rank = survivors of the lower lanes (lane-major slot order) + own earlier fields: one shuffle scan of the
      // per-lane counts instead of one ballot per field
      const u32 cl = __popc(smask);
      u32 x = cl;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
      u32 o = run + x - cl;
      const u32 si = (u32)j * 256 + lane * 8;            // slot offset inside the slice
#pragma unroll
      for (int q = 0; q < 8; q++)
        if ((smask >> q) & 1u) { sa.st_idx[sbase + o] = (uint16_t)(si + q); sa.st_cnt[sbase + o] = f[q]; o++; }
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (wv[0] | wv[1] | wv[2] | wv[3]) h4[q0 + (u64)j * 32] = make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
  if (lane == 0) {
    s_agg[w] = run; sa.slice_counts[(u64)c * HC_WARPS + w] = run;
    if (fsum) atomicAdd((unsigned long long*)(sa.win_sum + wl), (unsigned long long)fsum);
  }
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < HC_WARPS; i++) t += s_agg[i]; sa.chunk_counts[c] = t; }
}

// ---- rolled variant (k <= 32) ---------------------------------------------------------------
// The search + 128-bit extraction + bit-reversal of the kernel above cost ~60 of its ~155
// instructions per k-mer.  Here a CTA stages a tile of HR_TILE records in shared memory and
// counting-sorts them by their number of k-mers (one shared atomic per record), so that each warp
// then owns 32 records of (nearly) EQUAL length: lane = record, the forward k-mer and its reverse
// complement are rolled base by base (~15 instructions), and the warp stays converged because all
// lanes finish together.  grid = (tiles, windows): blockIdx.x runs fastest, so only a few windows
// are being filled at any time and they stay L2-resident.
static constexpr int HR_THREADS = 256;
static constexpr u32 HR_MAXWIN = 1024;         // windows per launch the persistent kernel can index (power of two)

static constexpr u32 HR_DELAY = 2;             // fused sweep: window z is compacted after the tiles of window z + HR_DELAY were handed out

__device__ __forceinline__ u32 ld_acquire_u32(const u32* p)
{
  u32 v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// FUSED: the item stream interleaves, after the tiles of window b, the compact chunks (compact_chunk)
// of window b - HR_DELAY, whose fill is complete by then (checked on its done counter): the window is
// swept while it is still L2-resident and the sweep's instructions hide under the RED-bound fill.
template <bool D32, bool TAIL64, int HR_TILE, bool FUSED, bool H16>
__global__ void __launch_bounds__(HR_THREADS)
hash_hist_roll_kernel(const uint4* __restrict__ recs, const u64* __restrict__ boff, const u32* __restrict__ bcnt,
                      int k, u64 Wbits, FastMod64 fm, FastMod32 fm32, u32* __restrict__ hist, u32 p0, u32 gp, u32* __restrict__ ticket,
                      SweepArgs sa)
{
  __shared__ u32 s_agg[HC_WARPS];
  __shared__ uint4 s_rec[HR_TILE];
  __shared__ uint16_t s_perm[HR_TILE];
  __shared__ u32 s_cnt[64], s_start[64];
  __shared__ u32 s_pref[HR_MAXWIN + 1];
  __shared__ u32 s_next;
  const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  constexpr int HR_PER = HR_TILE / HR_THREADS;
  const u64 kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
  const int rcsh = 2 * (k - 1);
  // persistent with in-order tickets: the (window, tile) items are numbered window-major and
  // handed out by an atomic counter, so the tiles in flight are always consecutive -- all CTAs
  // work on the same one or two windows (L2-resident REDs) -- and a grid capped at a few CTAs per
  // SM leaves room on every SM for the issue-bound stage-1 kernel of another lane.
  for (u32 i = tid; i < HR_MAXWIN; i += HR_THREADS) s_pref[i + 1] = i < gp ? (bcnt[p0 + i] + HR_TILE - 1) / HR_TILE : 0u;
  if (tid == 0) s_pref[0] = 0;
  __syncthreads();
  if (w == 0) {                                  // inclusive scan of s_pref[1..HR_MAXWIN] by warp 0
    u32 carry = 0;
    for (u32 base = 1; base <= HR_MAXWIN; base += 32) {
      u32 x = s_pref[base + lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
      s_pref[base + lane] = x + carry;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  const u32 ntile_all = s_pref[HR_MAXWIN];
  const u32 CWf = FUSED ? sa.CW : 0u;
  const u32 total = ntile_all + CWf * gp;
  // item stream: block b in [0, gp + HR_DELAY) = tiles of window b (b < gp), then chunks of window b - HR_DELAY (b >= HR_DELAY)
  auto bstart = [&](u32 b) -> u32 { return s_pref[min(b, gp)] + CWf * (b > HR_DELAY ? b - HR_DELAY : 0u); };
  u32 nxt = 0;
  if (tid == 0) { nxt = atomicAdd(ticket, 1u); s_next = nxt; }
  for (;;) {
  __syncthreads();                               // previous tile fully consumed, s_next visible
  const u32 item = s_next;
  if (item >= total) break;
  u32 y = 0;                                     // largest block y with bstart(y) <= item
  if (FUSED) {
    u32 lo = 0, hi = gp + HR_DELAY;              // bstart(lo) <= item < bstart(hi) = total
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (bstart(mid) <= item) lo = mid; else hi = mid; }
    y = lo;
    const u32 off = item - bstart(y);
    const u32 ty = y < gp ? s_pref[y + 1] - s_pref[y] : 0u;
    if (off >= ty) {                             // ---- a compact chunk of window z
      const u32 z = y - HR_DELAY, tz = s_pref[z + 1] - s_pref[z];
      if (tid == 0) {
        nxt = atomicAdd(ticket, 1u);
        while (ld_acquire_u32(sa.done + z) < tz) __nanosleep(64);   // its tiles hold earlier tickets: they are running
      }
      __syncthreads();
      compact_chunk<true>(z * sa.CW + (off - ty), Wbits, hist, sa, tz != 0, s_agg);
      if (tid == 0) s_next = nxt;
      continue;
    }
  } else {
#pragma unroll
    for (u32 st = HR_MAXWIN / 2; st > 0; st >>= 1) if (s_pref[y + st] <= item) y += st;
  }
  const u32 p = p0 + y;
  const u32 n = bcnt[p];
  const u64 b0 = boff[p];
  u32* __restrict__ h = hist + (u64)y * (Wbits >> (H16 ? 1 : 0));
  const u32 tile0 = (item - (FUSED ? bstart(y) : s_pref[y])) * HR_TILE;
  const u32 nt = min((u32)HR_TILE, n - tile0);
  if (tid < 64) s_cnt[tid] = 0;
  __syncthreads();
  if (tid == 0) nxt = atomicAdd(ticket, 1u);     // next ticket: its latency hides behind this tile (stored at the end)
  u32 nkr[HR_PER], rank[HR_PER];
#pragma unroll
  for (int i = 0; i < HR_PER; i++) {
    const u32 r = tid + HR_THREADS * i;
    nkr[i] = 0; rank[i] = 0;
    if (r < nt) {
      const uint4 v = __ldg(recs + b0 + tile0 + r);
      s_rec[r] = v;
      nkr[i] = ((v.w >> 24) - (u32)k + 1u) & 63u;
      rank[i] = atomicAdd(&s_cnt[nkr[i]], 1u);
    }
  }
  __syncthreads();
  if (tid < 32) {
    const u32 c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
    u32 x = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= (u32)o) x += y; }
    s_start[2 * tid] = x - c0 - c1; s_start[2 * tid + 1] = x - c1;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < HR_PER; i++) {
    const u32 r = tid + HR_THREADS * i;
    if (r < nt) s_perm[s_start[nkr[i]] + rank[i]] = (uint16_t)r;
  }
  __syncthreads();
  for (u32 g0 = w * 32; g0 < nt; g0 += (HR_THREADS / 32) * 32) {
    const u32 idx = g0 + lane;
    u32 nk = 0; u64 f = 0, rc = 0, tlo = 0, thi = 0;
    if (idx < nt) {
      const uint4 v = s_rec[s_perm[idx]];
      const u64 lo = (u64)v.x | ((u64)v.y << 32);
      const u64 hh = (u64)v.z | ((u64)v.w << 32);
      const int nb = (int)(hh >> 56);
      const u64 hi = hh & 0x00FFFFFFFFFFFFFFULL;
      nk = (u32)(nb - k + 1);
      f = rec1_kmer(lo, hi, nb, k, 0);
      rc = revcomp64(f, k);
      // the nb-k bases after the first k-mer, aligned to the top of (thi:tlo)
      const int tb = 2 * (nb - k);                     // 0 .. 2*(max_nk-1)
      if (TAIL64) { thi = tb ? (lo << (64 - tb)) : 0ULL; }
      else {
        // 128-bit left shift of (hi:lo) by 128 - tb
        const int sh = 128 - tb;                       // 128 >= sh > 0
        if (sh >= 128) { thi = 0; tlo = 0; }
        else if (sh >= 64) { thi = lo << (sh - 64); tlo = 0; }
        else { thi = (hi << sh) | (lo >> (64 - sh)); tlo = lo << sh; }
      }
    }
    const u32 maxnk = __reduce_max_sync(0xffffffffu, nk);
    for (u32 j = 0; j < maxnk; j++) {
      const u64 c = f < rc ? f : rc;
      const u64 hv = xxh64_8(c);
      const u64 key = D32 ? (u64)fastmod64_d32(hv, fm32) : fastmod64(hv, fm);
      if (j < nk) hist_inc<H16>(h, key);
      const u64 b = thi >> 62;
      if (TAIL64) thi <<= 2;
      else { thi = (thi << 2) | (tlo >> 62); tlo <<= 2; }
      f = ((f << 2) | b) & kmask;
      rc = (rc >> 2) | ((b ^ 2ULL) << rcsh);
    }
  }
  if (FUSED) {                                   // this tile's REDs are performed before the window counts it as done
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(sa.done + y, 1u);
  }
  if (tid == 0) s_next = nxt;
  }   // items
}

// ---- ordered sweep: compact (+re-zero) -> scan -> copy ---------------------------------------
// The histogram is sparse (a few % of the slots survive hard-min) and every access below is a
// coalesced stream -- no gathers, no look-back chain:
//   hash_compact_kernel  streams the windows once (128-bit coalesced loads, DRAM-bound).  Warp w of
//       a CTA owns a 1024-slot slice of the CTA's chunk; per row of 128 slots the four component
//       ballots rank the survivors, which are appended IN SLOT ORDER to the slice's own staging
//       run (slot offset u16 + count u32; capacity = the slice, so it cannot overflow).  Non-zero
//       words are overwritten with zeros in the same pass, so the histogram is clean again.
//   hash_scan_kernel     one CTA: exclusive prefix of the chunk counts on top of the running cursor
//       (bump allocation of the output space, overflow flag, list offsets per window).
//   hash_copy_kernel     moves every slice's run to its final place as (key u64, count u32).
template <bool H16>
__global__ void __launch_bounds__(HC_THREADS, 5)
hash_compact_kernel(u64 Wbits, u32 nchunks, u32* __restrict__ hist, SweepArgs sa, const u32* __restrict__ bcnt, u32 p0)
{
  __shared__ u32 s_agg[HC_WARPS];
  for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {       // grid = nchunks, or a persistent grid striding over the chunks
    if (H16) compact_chunk16(c, Wbits, hist, sa, bcnt[p0 + c / sa.CW] != 0, s_agg);
    else compact_chunk<false>(c, Wbits, hist, sa, bcnt[p0 + c / sa.CW] != 0, s_agg);
    __syncthreads();
  }
}

// One CTA: exclusive prefix of the group's chunk counts on top of the running cursor *base_in.  Every
// thread owns a contiguous range: one round of independent loads, one block scan, one round of stores.
__global__ void __launch_bounds__(1024)
hash_scan_kernel(u32 CW, u32 nchunks, u32 p0, const u32* __restrict__ chunk_counts, u64* __restrict__ chunk_off,
                 u64* __restrict__ list_off, const u64* __restrict__ base_in, u64* __restrict__ total_out,
                 const u64* __restrict__ cap_p, u32* __restrict__ flags,
                 const u64* __restrict__ win_sum /* NULL unless 16-bit counters */, const u64* __restrict__ kcnt, u32 gp)
{
  __shared__ u64 s_warp[32];
  if (win_sum)                                    // a window whose fields do not add up to its k-mers saw a counter wrap
    for (u32 wdw = threadIdx.x; wdw < gp; wdw += 1024) if (win_sum[wdw] != kcnt[p0 + wdw]) flags[1] = 1u;
  const u32 per = (nchunks + 1023u) / 1024u;
  const u32 i0 = min(nchunks, threadIdx.x * per), i1 = min(nchunks, i0 + per);
  u64 sum = 0;
  for (u32 i = i0; i < i1; i += 8) {              // 8 independent loads in flight, then the adds
    u32 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = i + q < i1 ? chunk_counts[i + q] : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) sum += v[q];
  }
  u64 x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    const u64 wv = s_warp[threadIdx.x]; u64 xw = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, xw, o); if (threadIdx.x >= (u32)o) xw += y; }
    s_warp[threadIdx.x] = xw - wv;
  }
  __syncthreads();
  u64 run = *base_in + s_warp[threadIdx.x >> 5] + x - sum;
  for (u32 i = i0; i < i1; i += 8) {
    u32 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = i + q < i1 ? chunk_counts[i + q] : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      if (i + q < i1) {
        chunk_off[i + q] = run;
        if ((i + q) % CW == 0) list_off[p0 + (i + q) / CW] = run;
        run += v[q];
      }
    }
  }
  if (threadIdx.x == 1023) {                      // its range ends at nchunks: run is the grand total
    *total_out = run;
    if (run > *cap_p) flags[0] = 1u;              // output space ran out: nothing is copied, the host retries
  }
}

// one warp per chunk: copies the chunk's 8 staging runs to their final place as (key u64, count u32)
__global__ void __launch_bounds__(HC_THREADS)
hash_copy_kernel(u64 Wbits, u32 CW, u32 nchunks, const uint16_t* __restrict__ st_idx, const u32* __restrict__ st_cnt,
                 const u32* __restrict__ slice_counts, const u32* __restrict__ chunk_counts, const u64* __restrict__ chunk_off,
                 u64* __restrict__ out_keys, u32* __restrict__ out_counts,
                 const u32* __restrict__ win_part /* NULL: window p holds partition p */, u32 p0, const u32* __restrict__ flags)
{
  const u32 lane = threadIdx.x & 31u;
  const u32 c = blockIdx.x * HC_WARPS + (threadIdx.x >> 5);
  if (c >= nchunks || flags[0]) return;
  const u32 mine = lane < (u32)HC_WARPS ? slice_counts[(u64)c * HC_WARPS + lane] : 0u;
  if (__ballot_sync(0xffffffffu, mine != 0) == 0) return;
  const u32 wl = c / CW, sub = c - wl * CW, p = p0 + wl;
  u32 x = mine;                                    // inclusive scan over the 8 slices
#pragma unroll
  for (int o = 1; o < HC_WARPS; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
  const u64 obase = chunk_off[c];
  const u64 kb = (u64)(win_part ? win_part[p] : p) * Wbits + (u64)sub * HIST_SUB;
#pragma unroll
  for (int w = 0; w < HC_WARPS; w++) {
    const u32 n = __shfl_sync(0xffffffffu, mine, w);
    const u64 o = obase + __shfl_sync(0xffffffffu, x - mine, w);
    const u64 sbase = ((u64)c * HC_WARPS + w) * HC_SLICE;
    for (u32 i = lane; i < n; i += 32) {
      out_keys[o + i] = kb + (u64)w * HC_SLICE + st_idx[sbase + i];
      out_counts[o + i] = st_cnt[sbase + i];
    }
  }
}

// phase 0: histogram fill of windows [p0, p0+gp).  phase 1: the ordered sweep of the same windows.
// meta (device): [0],[1] ping-pong running cursor (group g reads [g&1], writes [(g+1)&1]), [2] capacity.
cudaError_t launch_hash_group(const S2Common& c, u64 Wbits, u64 mod_d, u64 mod_mlo, u64 mod_mhi, u32* hist, u32 hard_min,
                              u32 p0, u32 gp, u32 group_idx, u32* chunk_counts, u64* chunk_off, SweepStage stage, u64* list_off, u64* meta, u32* flags,
                              u64* out_keys, u32* out_counts, const u32* win_part, cudaStream_t st, u64* launches, int phase, bool h16)
{
  const u32 hmin = hard_min ? hard_min : 1;
  const u32 CW = hash_sweep_chunks_per_window(Wbits);
  // rolled persistent fill for k <= 32; fused with the compact pass unless there is nothing to fill
  const bool roll = c.W == 1 && gp <= HR_MAXWIN && !kmx_env_flag("KMX_HIST_NOROLL");
  // (opt-in: measured SLOWER than fill + separate compact -- 1.37 vs 1.02 ms per 1.2e8 k-mers -- because every
  //  tile needs a __threadfence before its window can be declared filled; kept for experiments)
  const bool fused = roll && c.max_bcnt != 0 && !h16 && kmx_env_flag("KMX_HIST_FUSE");
  SweepArgs sa; sa.CW = CW; sa.hmin = hmin; sa.st_idx = stage.idx; sa.st_cnt = stage.cnt; sa.slice_counts = stage.slice_counts;
  sa.chunk_counts = chunk_counts; sa.done = stage.done; sa.win_sum = stage.win_sum;
  if (phase == 0) {
    if (c.max_bcnt) {
      FastMod64 fm; fm.d = mod_d; fm.mlo = mod_mlo; fm.mhi = mod_mhi;
      FastMod32 f32; f32.d = (u32)mod_d; f32.m64 = mod_d >= 2 ? (~0ULL) / mod_d : 0;
      const bool d32 = mod_d >= 2 && mod_d < (1ULL << 31);
      unsigned gx = (c.max_bcnt + HH_THREADS - 1) / HH_THREADS;
      if (gx > 592) gx = 592;                 // 4 waves of 148 SMs per partition row at most
      dim3 grid(gx, gp);
      const uint4* recs = (const uint4*)c.records;
      if (roll) {
        const bool tail64 = 2 * (KMX_REC1_MAXN - c.k) <= 64;
        static const int tile = []{ const char* v = getenv("KMX_HR_TILE"); int t = v ? atoi(v) : 256; return (t == 512 || t == 1024) ? t : 256; }();
        static const int cap = []{ const char* v = getenv("KMX_HIST_CAP"); int t = v ? atoi(v) : 4; return (t >= 1 && t <= 8) ? t : 4; }();
        const u64 max_items = (u64)((c.max_bcnt + 255) / 256) * gp;
        const unsigned rgrid = (unsigned)std::min<u64>(max_items, (u64)148 * cap);
        u32* hist_ticket = stage.done + gp;
        cudaError_t me = cudaMemsetAsync(stage.done, 0, ((size_t)gp + 1) * 4, st);    // done[gp] | ticket
        if (me != cudaSuccess) return me;
#define KMX_ROLL(D, T, TILE, F, H) hash_hist_roll_kernel<D, T, TILE, F, H><<<rgrid, HR_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0, gp, hist_ticket, sa)
#define KMX_ROLL_F(D, T, TILE) do { if (fused) KMX_ROLL(D, T, TILE, true, false); else if (h16) KMX_ROLL(D, T, TILE, false, true); else KMX_ROLL(D, T, TILE, false, false); } while (0)
#define KMX_ROLL_T(TILE) do { if (d32 && tail64) KMX_ROLL_F(true, true, TILE); else if (d32) KMX_ROLL_F(true, false, TILE); else if (tail64) KMX_ROLL_F(false, true, TILE); else KMX_ROLL_F(false, false, TILE); } while (0)
        if (tile == 512) KMX_ROLL_T(512); else if (tile == 1024) KMX_ROLL_T(1024); else KMX_ROLL_T(256);
#undef KMX_ROLL_T
#undef KMX_ROLL_F
#undef KMX_ROLL
      }
      else {
#define KMX_HH(WW, D, H) hash_hist_kernel<WW, D, H><<<grid, HH_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0)
#define KMX_HH_H(WW, D) do { if (h16) KMX_HH(WW, D, true); else KMX_HH(WW, D, false); } while (0)
        if (c.W == 1 && d32) KMX_HH_H(1, true); else if (c.W == 1) KMX_HH_H(1, false); else if (d32) KMX_HH_H(2, true); else KMX_HH_H(2, false);
#undef KMX_HH_H
#undef KMX_HH
      }
      *launches += 1;
    }
  } else {
    const u64 nchunks64 = (u64)gp * CW;
    if (nchunks64 >= 0x7FFFFFF0ULL) return cudaErrorInvalidValue;
    const u32 nchunks = (u32)nchunks64;
    if (!fused) {
      static const unsigned ccap = []{ const char* v = getenv("KMX_COMPACT_CAP"); int t = v ? atoi(v) : 0; return (unsigned)((t >= 0 && t <= 8) ? t : 0); }();
      const unsigned cgrid = ccap ? std::min<unsigned>(nchunks, 148u * ccap) : nchunks;
      if (h16) {
        cudaError_t me = cudaMemsetAsync(stage.win_sum, 0, (size_t)gp * 8, st);
        if (me != cudaSuccess) return me;
        hash_compact_kernel<true><<<cgrid, HC_THREADS, 0, st>>>(Wbits, nchunks, hist, sa, c.bcnt, p0);
      } else hash_compact_kernel<false><<<cgrid, HC_THREADS, 0, st>>>(Wbits, nchunks, hist, sa, c.bcnt, p0);
      *launches += 1;
    }
    hash_scan_kernel<<<1, 1024, 0, st>>>(CW, nchunks, p0, chunk_counts, chunk_off, list_off, meta + (group_idx & 1u),
                                         meta + ((group_idx + 1u) & 1u), meta + 2, flags, h16 ? stage.win_sum : nullptr, c.kcnt, gp);
    hash_copy_kernel<<<(nchunks + HC_WARPS - 1) / HC_WARPS, HC_THREADS, 0, st>>>(Wbits, CW, nchunks, stage.idx, stage.cnt, stage.slice_counts, chunk_counts, chunk_off,
                                                     out_keys, out_counts, win_part, p0, flags);
    *launches += 2;
  }
  return cudaGetLastError();
}

u32 hash_sweep_chunks_per_window(u64 Wbits) { return (u32)((Wbits + HIST_SUB - 1) / HIST_SUB); }

}  // namespace kmx
