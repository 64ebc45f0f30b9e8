// s2_sort.cu -- stage 2, generic path: expand super-k-mer records to canonical keys, segmented
// LSD radix sort (one segment per partition), run-length count + hard-min filter.
//
// Replaces (behaviour, not code) ReadSuperk / KmerSort / KmerPartCounter::executeDump
// (include/kmtricks/gatb/sorting_count.hpp:141-312,498-508,694-884) and KmerCountProcessor
// (include/kmtricks/gatb/count_processor.hpp:135-146).  The reference's kx-mer radix bins and
// 453-way heap are a CPU memory optimisation; the contract is only "ascending distinct
// canonical k-mers with counts >= hard_min" (SURVEY §3.3).
#include "common.cuh"
#include "kmx_internal.h"
#include "records.cuh"
#include <algorithm>
#include <vector>

namespace kmx {

// ---------------------------------------------------------------------------------------
// multi-CTA exclusive scan of a u32 array, in place (sum must fit in u32)
// ---------------------------------------------------------------------------------------
static constexpr int SC_THREADS = 256;
static constexpr int SC_ITEMS = 16;
static constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const u32* __restrict__ d, u64 n, u32* __restrict__ block_sums)
{
  __shared__ u32 s[SC_THREADS / 32];
  u64 base = (u64)blockIdx.x * SC_TILE;
  u32 sum = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { u64 j = base + (u64)i * SC_THREADS + threadIdx.x; if (j < n) sum += d[j]; }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < SC_THREADS / 32; i++) t += s[i]; block_sums[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) scan_blocksums_kernel(u32* __restrict__ bs, u64 nb, u32* __restrict__ total)
{
  __shared__ u32 s_warp[32];
  __shared__ u32 s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u64 base = 0; base < nb; base += 1024) {
    u64 i = base + threadIdx.x;
    u32 v = (i < nb) ? bs[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      u32 w = s_warp[threadIdx.x], xw = w;
      for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, xw, o); if (threadIdx.x >= o) xw += y; }
      s_warp[threadIdx.x] = xw - w;
    }
    __syncthreads();
    u32 excl = s_carry + s_warp[threadIdx.x >> 5] + x - v;
    if (i < nb) bs[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = s_carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(u32* __restrict__ d, u64 n, const u32* __restrict__ block_sums)
{
  __shared__ u32 s_warp[SC_THREADS / 32];
  // blocked arrangement: thread t owns items [t*16, t*16+16) of the tile
  u64 base = (u64)blockIdx.x * SC_TILE + (u64)threadIdx.x * SC_ITEMS;
  u32 v[SC_ITEMS];
  u32 sum = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { u64 j = base + i; v[i] = (j < n) ? d[j] : 0; sum += v[i]; }
  u32 x = sum;
  for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  u32 wb = 0;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) wb += s_warp[i];
  u32 run = block_sums[blockIdx.x] + wb + x - sum;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { u64 j = base + i; if (j < n) d[j] = run; run += v[i]; }
}

size_t scan_u32_work_bytes(u64 n) { return (size_t)((n + SC_TILE - 1) / SC_TILE + 1) * 4; }

// d[0..n) -> exclusive prefix in place; *d_total (device) = sum.  work: scan_u32_work_bytes(n)
cudaError_t scan_u32_inplace(u32* d, u64 n, u32* d_total, void* work, cudaStream_t st, u64* launches)
{
  if (n == 0) { if (d_total) return cudaMemsetAsync(d_total, 0, 4, st); return cudaSuccess; }
  u64 nb = (n + SC_TILE - 1) / SC_TILE;
  u32* bs = (u32*)work;
  scan_reduce_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(d, n, bs);
  scan_blocksums_kernel<<<1, 1024, 0, st>>>(bs, nb, d_total);
  scan_apply_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(d, n, bs);
  *launches += 3;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// expand records -> keys
// ---------------------------------------------------------------------------------------
static constexpr int EX_THREADS = 256;

template <int W, int KIND>
__global__ void __launch_bounds__(EX_THREADS)
expand_keys_kernel(const uint4* __restrict__ recs, const u64* __restrict__ boff, const u32* __restrict__ bcnt,
                   int k, u64 Wbits, FastMod64 fm, const u64* __restrict__ koff, u32* __restrict__ kcursor,
                   u64* __restrict__ keys_lo, u64* __restrict__ keys_hi)
{
  __shared__ u32 s_warp[EX_THREADS / 32];
  __shared__ u32 s_base;
  const u32 p = blockIdx.y;
  const u32 n = bcnt[p];
  const u64 b0 = boff[p];
  const u64 k0 = koff[p];
  const u64 hbase = Wbits * p;
  for (u32 r0 = blockIdx.x * EX_THREADS; r0 < n; r0 += gridDim.x * EX_THREADS) {
    const u32 r = r0 + threadIdx.x;
    Rec1 r1; Rec2 r2; int nk = 0;
    if (r < n) {
      if (W == 1) { r1 = load_rec1(recs, b0 + r); nk = r1.n - k + 1; }
      else { r2 = load_rec2(recs, b0 + r); nk = r2.n - k + 1; }
    }
    u32 x = (u32)nk;
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    u32 wb = 0, tot = 0;
    for (int i = 0; i < EX_THREADS / 32; i++) { u32 t = s_warp[i]; if (i < (int)(threadIdx.x >> 5)) wb += t; tot += t; }
    if (threadIdx.x == 0) s_base = atomicAdd(&kcursor[p], tot);
    __syncthreads();
    u64 o = k0 + s_base + wb + x - (u32)nk;
    for (int j = 0; j < nk; j++) {
      if (W == 1) {
        u64 c; canon1(r1, k, j, c);
        keys_lo[o + j] = (KIND == 1) ? fastmod64(xxh64_8(c), fm) + hbase : c;
      } else {
        u64 clo, chi; canon2(r2, k, j, clo, chi);
        if (KIND == 1) keys_lo[o + j] = fastmod64(xxh64_16(clo, chi), fm) + hbase;
        else { keys_lo[o + j] = clo; keys_hi[o + j] = chi; }
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_expand_keys(const S2Common& c, int key_kind, u64 Wbits, u64 mod_d, u64 mod_mlo, u64 mod_mhi,
                               const u64* koff, u32* kcursor, u64* keys_lo, u64* keys_hi,
                               cudaStream_t st, u64* launches)
{
  if (c.max_bcnt == 0) return cudaSuccess;
  FastMod64 fm; fm.d = mod_d ? mod_d : 1; fm.mlo = mod_mlo; fm.mhi = mod_mhi;
  unsigned gx = (c.max_bcnt + EX_THREADS - 1) / EX_THREADS;
  if (gx > 2048) gx = 2048;
  dim3 grid(gx, c.P);
  const uint4* recs = (const uint4*)c.records;
  if (c.W == 1 && key_kind == 0) expand_keys_kernel<1, 0><<<grid, EX_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, koff, kcursor, keys_lo, keys_hi);
  else if (c.W == 1) expand_keys_kernel<1, 1><<<grid, EX_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, koff, kcursor, keys_lo, keys_hi);
  else if (key_kind == 0) expand_keys_kernel<2, 0><<<grid, EX_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, koff, kcursor, keys_lo, keys_hi);
  else expand_keys_kernel<2, 1><<<grid, EX_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, koff, kcursor, keys_lo, keys_hi);
  *launches += 1;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// segmented LSD radix sort, 8-bit digits, tiles of 4096 keys that never cross a segment
// ---------------------------------------------------------------------------------------
static constexpr int RS_THREADS = 256;
static constexpr int RS_ITEMS = 16;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 4096

struct RsTile { u64 begin; u32 len; u32 dstride; u64 cbase; u64 adj; };   // cbase: index of (tile, digit 0) in counts; adj = segment begin - keys in earlier segments

__device__ __forceinline__ u32 rs_digit(u64 lo, u64 hi, int shift, u32 mask)
{
  // shifts are multiples of 8, so a digit never straddles the two words
  u64 v = (shift < 64) ? (lo >> shift) : (hi >> (shift - 64));
  return (u32)v & mask;
}

template <int W>
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const RsTile* __restrict__ tiles, const u64* __restrict__ lo, const u64* __restrict__ hi,
               int shift, u32 mask, u32* __restrict__ counts)
{
  __shared__ u32 s_h[256];
  const RsTile t = tiles[blockIdx.x];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  for (u32 i = threadIdx.x; i < t.len; i += RS_THREADS) {
    u64 l = lo[t.begin + i], h = (W == 2) ? hi[t.begin + i] : 0;
    atomicAdd(&s_h[rs_digit(l, h, shift, mask)], 1u);
  }
  __syncthreads();
  counts[t.cbase + (u64)threadIdx.x * t.dstride] = s_h[threadIdx.x];
}

template <int W>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const RsTile* __restrict__ tiles, const u64* __restrict__ lo, const u64* __restrict__ hi,
                  u64* __restrict__ olo, u64* __restrict__ ohi, int shift, u32 mask,
                  const u32* __restrict__ scanned)
{
  __shared__ u32 s_wcnt[RS_THREADS / 32][256];
  __shared__ u32 s_gpos[256];
  const RsTile t = tiles[blockIdx.x];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = 0; i < RS_THREADS / 32; i++) s_wcnt[i][threadIdx.x] = 0;
  __syncthreads();
  u64 kl[RS_ITEMS], kh[RS_ITEMS];
  u32 dg[RS_ITEMS], rk[RS_ITEMS];
  const u32 wbase = w * (32 * RS_ITEMS);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    u32 idx = wbase + i * 32 + lane;
    bool valid = idx < t.len;
    kl[i] = valid ? lo[t.begin + idx] : 0;
    kh[i] = (W == 2 && valid) ? hi[t.begin + idx] : 0;
    dg[i] = valid ? rs_digit(kl[i], kh[i], shift, mask) : (0xFFFFFF00u + lane);
  }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 d = dg[i];
    const u32 peers = __match_any_sync(0xffffffffu, d);
    const bool valid = d < 256;
    u32 base = valid ? s_wcnt[w][d] : 0;
    rk[i] = base + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && lane == (u32)(__ffs(peers) - 1)) s_wcnt[w][d] = base + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    const u32 d = threadIdx.x;
    u32 run = 0;
#pragma unroll
    for (int i = 0; i < RS_THREADS / 32; i++) { u32 c = s_wcnt[i][d]; s_wcnt[i][d] = run; run += c; }
    s_gpos[d] = scanned[t.cbase + (u64)d * t.dstride];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 d = dg[i];
    if (d < 256) {
      u64 pos = t.adj + s_gpos[d] + s_wcnt[w][d] + rk[i];
      olo[pos] = kl[i];
      if (W == 2) ohi[pos] = kh[i];
    }
  }
}

// Host: build the tile table for segments [h_seg_off[i], h_seg_off[i+1]).
// Work buffer layout: [RsTile ntiles][u32 counts 256*ntiles][scan work]
// Segments: [so[s], se ? se[s] : so[s+1]).  With explicit ends the segments may have gaps.
static u64 rs_count_tiles(u32 nseg, const u64* so, const u64* se = nullptr)
{
  u64 nt = 0;
  for (u32 s = 0; s < nseg; s++) nt += ((se ? se[s] : so[s + 1]) - so[s] + RS_TILE - 1) / RS_TILE;
  return nt;
}

size_t radix_sort_work_bytes(u32 nseg, const u64* h_seg_off, const u64* h_seg_end)
{
  u64 nt = rs_count_tiles(nseg, h_seg_off, h_seg_end);
  size_t a = (size_t)nt * sizeof(RsTile);
  a = (a + 255) & ~(size_t)255;
  size_t b = (size_t)nt * 256 * 4;
  b = (b + 255) & ~(size_t)255;
  return a + b + scan_u32_work_bytes(nt * 256) + 256;
}

cudaError_t segmented_radix_sort(u32 nseg, const u64* h_seg_off, const u64* h_seg_end, u64* lo, u64* hi, u64* lo_alt, u64* hi_alt,
                                 int W, int begin_bit, int end_bit, void* d_work, int* result_in_alt,
                                 cudaStream_t st, u64* launches)
{
  *result_in_alt = 0;
  const u64 nt = rs_count_tiles(nseg, h_seg_off, h_seg_end);
  if (nt == 0 || end_bit <= begin_bit) return cudaSuccess;
  u64 total = 0;
  for (u32 s = 0; s < nseg; s++) total += (h_seg_end ? h_seg_end[s] : h_seg_off[s + 1]) - h_seg_off[s];
  if (total >= 0xFFFFFFFFULL) return cudaErrorInvalidValue;
  // pinned staging of the tile table, per host thread.  It only grows, with head-room, and an outgrown buffer is NOT freed here:
  // cudaFreeHost waits for the whole device, and another lane's NCCL kernel may be waiting for a peer at that moment (see
  // kmx_api.cu: defer_free); the few hundred KB are kept until the process ends.
  static thread_local RsTile* h_tiles = nullptr; static thread_local u64 h_cap = 0;
  if (h_cap < nt) {
    const u64 ncap = std::max<u64>(2 * nt, 2048);
    RsTile* np = nullptr;
    cudaError_t e = cudaMallocHost((void**)&np, ncap * sizeof(RsTile));
    if (e != cudaSuccess) return e;
    h_tiles = np; h_cap = ncap;
  }
  u64 ti = 0, before = 0;
  for (u32 s = 0; s < nseg; s++) {
    u64 b = h_seg_off[s], e = h_seg_end ? h_seg_end[s] : h_seg_off[s + 1];
    u64 n = (e - b + RS_TILE - 1) / RS_TILE;
    for (u64 j = 0; j < n; j++) {
      RsTile& t = h_tiles[ti + j];
      t.begin = b + j * RS_TILE; t.len = (u32)std::min<u64>(RS_TILE, e - t.begin);
      t.dstride = (u32)n; t.cbase = ti * 256 + j; t.adj = b - before;
    }
    ti += n; before += e - b;
  }
  char* wp = (char*)d_work;
  RsTile* d_tiles = (RsTile*)wp; wp += ((size_t)nt * sizeof(RsTile) + 255) & ~(size_t)255;
  u32* d_counts = (u32*)wp; wp += ((size_t)nt * 256 * 4 + 255) & ~(size_t)255;
  void* d_scanw = wp;
  cudaError_t e = cudaMemcpyAsync(d_tiles, h_tiles, nt * sizeof(RsTile), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  u64 *cl = lo, *ch = hi, *al = lo_alt, *ah = hi_alt;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    int bits = std::min(8, end_bit - shift);
    u32 mask = (1u << bits) - 1u;
    if (W == 1) rs_hist_kernel<1><<<(unsigned)nt, RS_THREADS, 0, st>>>(d_tiles, cl, ch, shift, mask, d_counts);
    else rs_hist_kernel<2><<<(unsigned)nt, RS_THREADS, 0, st>>>(d_tiles, cl, ch, shift, mask, d_counts);
    *launches += 1;
    e = scan_u32_inplace(d_counts, nt * 256, nullptr, d_scanw, st, launches);
    if (e != cudaSuccess) return e;
    if (W == 1) rs_scatter_kernel<1><<<(unsigned)nt, RS_THREADS, 0, st>>>(d_tiles, cl, ch, al, ah, shift, mask, d_counts);
    else rs_scatter_kernel<2><<<(unsigned)nt, RS_THREADS, 0, st>>>(d_tiles, cl, ch, al, ah, shift, mask, d_counts);
    *launches += 1;
    std::swap(cl, al); std::swap(ch, ah);
    *result_in_alt ^= 1;
  }
  // the pinned tile table is reused by the next call: make sure the copy has been consumed
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// run-length + hard-min over sorted segments
// ---------------------------------------------------------------------------------------
static constexpr int RL_THREADS = 256;
static constexpr int RL_ITEMS = 16;          // same 4096-key tiles as the sort

template <int W>
__device__ __forceinline__ bool key_eq(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 a, u64 l, u64 h)
{
  return lo[a] == l && (W == 1 || hi[a] == h);
}

// first index in (j, se] whose key differs from key[j] (galloping + binary search; sorted input)
template <int W>
__device__ u64 run_end(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 j, u64 se, u64 l, u64 h)
{
  u64 step = 1, a = j;            // key[a] == key[j]
  while (a + step < se && key_eq<W>(lo, hi, a + step, l, h)) { a += step; step <<= 1; }
  u64 b = (a + step < se) ? a + step : se;     // key[b] differs or b == se
  while (a + 1 < b) { u64 m = (a + b) >> 1; if (key_eq<W>(lo, hi, m, l, h)) a = m; else b = m; }
  return b;
}

// phase 0: tile_counts[t] = survivors whose run STARTS in tile t ; phase 1: write them.
template <int W, int PHASE>
__global__ void __launch_bounds__(RL_THREADS)
rle_kernel(const RsTile* __restrict__ tiles, const u64* __restrict__ seg_end_of_tile,
           const u64* __restrict__ seg_begin_of_tile,
           const u64* __restrict__ lo, const u64* __restrict__ hi, u32 hmin,
           u32* __restrict__ tile_counts, const u64* __restrict__ tile_off,
           u64* __restrict__ out_lo, u64* __restrict__ out_hi, u32* __restrict__ out_cnt)
{
  __shared__ u32 s_warp[RL_THREADS / 32];
  const RsTile t = tiles[blockIdx.x];
  const u64 sb = seg_begin_of_tile[blockIdx.x], se = seg_end_of_tile[blockIdx.x];
  const u32 i0 = threadIdx.x * RL_ITEMS;
  u32 nsurv = 0;
  u64 sl[RL_ITEMS], sh[RL_ITEMS]; u32 sc[RL_ITEMS];
  if (i0 < t.len) {
    const u64 j0 = t.begin + i0;
    const u32 cnt = min((u32)RL_ITEMS, t.len - i0);
    u64 pl = 0, ph = 0; bool have_prev = j0 > sb;
    if (have_prev) { pl = lo[j0 - 1]; ph = (W == 2) ? hi[j0 - 1] : 0; }
    u32 i = 0;
    while (i < cnt) {
      u64 l = lo[j0 + i], h = (W == 2) ? hi[j0 + i] : 0;
      bool head = !have_prev || l != pl || (W == 2 && h != ph);
      if (head) {
        // scan forward inside the chunk first, then gallop
        u32 e = i + 1;
        while (e < cnt && key_eq<W>(lo, hi, j0 + e, l, h)) e++;
        u64 end = j0 + e;
        if (e == cnt && end < se && key_eq<W>(lo, hi, end, l, h)) end = run_end<W>(lo, hi, end, se, l, h);
        u64 c = end - (j0 + i);
        if (c >= hmin) {
          if (PHASE == 1) { sl[nsurv] = l; sh[nsurv] = h; sc[nsurv] = c > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (u32)c; }
          nsurv++;
        }
        i = e;
      } else {
        // inside a run that started earlier: skip to its end within the chunk
        u32 e = i + 1;
        while (e < cnt && key_eq<W>(lo, hi, j0 + e, l, h)) e++;
        i = e;
      }
      pl = l; ph = h; have_prev = true;
    }
  }
  u32 x = nsurv;
  for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  u32 wb = 0, tot = 0;
  for (int i = 0; i < RL_THREADS / 32; i++) { u32 v = s_warp[i]; if (i < (int)(threadIdx.x >> 5)) wb += v; tot += v; }
  if (PHASE == 0) { if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot; }
  else {
    u64 o = tile_off[blockIdx.x] + wb + x - nsurv;
    for (u32 q = 0; q < nsurv; q++) { out_lo[o + q] = sl[q]; if (W == 2) out_hi[o + q] = sh[q]; out_cnt[o + q] = sc[q]; }
  }
}

// device tile table must already be in d_work (left there by segmented_radix_sort) -- to keep
// the two independent, rle builds its own small per-tile arrays.
size_t rle_work_bytes(u32 nseg, const u64* h_seg_off)
{
  u64 nt = rs_count_tiles(nseg, h_seg_off);
  return (size_t)nt * (sizeof(RsTile) + 8 + 8 + 4 + 8) + 1024;
}

// phase 0 returns per-tile survivor counts in h_tile_off form (exclusive prefix, nt+1 entries).
cudaError_t rle_segments(u32 nseg, const u64* h_seg_off, const u64* lo, const u64* hi, int W, u32 hard_min,
                         void* d_work, std::vector<u64>& h_tile_off, std::vector<u64>& h_seg_out_off,
                         int phase, u64* out_lo, u64* out_hi, u32* out_cnt, cudaStream_t st, u64* launches)
{
  const u64 nt = rs_count_tiles(nseg, h_seg_off);
  if (phase == 0) { h_tile_off.assign(nt + 1, 0); h_seg_out_off.assign(nseg + 1, 0); }
  if (nt == 0) return cudaSuccess;
  char* wp = (char*)d_work;
  RsTile* d_tiles = (RsTile*)wp; wp += (size_t)nt * sizeof(RsTile);
  u64* d_sb = (u64*)wp; wp += nt * 8;
  u64* d_se = (u64*)wp; wp += nt * 8;
  u64* d_toff = (u64*)wp; wp += nt * 8;
  u32* d_tcnt = (u32*)wp;
  const u32 hmin = hard_min ? hard_min : 1;
  if (phase == 0) {
    std::vector<RsTile> tiles(nt); std::vector<u64> sb(nt), se(nt);
    u64 ti = 0;
    for (u32 s = 0; s < nseg; s++) {
      u64 b = h_seg_off[s], e = h_seg_off[s + 1];
      u64 n = (e - b + RS_TILE - 1) / RS_TILE;
      for (u64 j = 0; j < n; j++) {
        RsTile& t = tiles[ti + j];
        t.begin = b + j * RS_TILE; t.len = (u32)std::min<u64>(RS_TILE, e - t.begin); t.dstride = 0; t.cbase = 0; t.adj = 0;
        sb[ti + j] = b; se[ti + j] = e;
      }
      ti += n;
    }
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_tiles, tiles.data(), nt * sizeof(RsTile), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(d_sb, sb.data(), nt * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(d_se, se.data(), nt * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if (W == 1) rle_kernel<1, 0><<<(unsigned)nt, RL_THREADS, 0, st>>>(d_tiles, d_se, d_sb, lo, hi, hmin, d_tcnt, nullptr, nullptr, nullptr, nullptr);
    else rle_kernel<2, 0><<<(unsigned)nt, RL_THREADS, 0, st>>>(d_tiles, d_se, d_sb, lo, hi, hmin, d_tcnt, nullptr, nullptr, nullptr, nullptr);
    *launches += 1;
    std::vector<u32> tc(nt);
    if ((e = cudaMemcpyAsync(tc.data(), d_tcnt, nt * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;      // also protects the host staging vectors
    for (u64 i = 0; i < nt; i++) h_tile_off[i + 1] = h_tile_off[i] + tc[i];
    ti = 0;
    for (u32 s = 0; s < nseg; s++) {
      h_seg_out_off[s] = h_tile_off[ti];
      ti += (h_seg_off[s + 1] - h_seg_off[s] + RS_TILE - 1) / RS_TILE;
    }
    h_seg_out_off[nseg] = h_tile_off[nt];
    return cudaGetLastError();
  }
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_toff, h_tile_off.data(), nt * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if (W == 1) rle_kernel<1, 1><<<(unsigned)nt, RL_THREADS, 0, st>>>(d_tiles, d_se, d_sb, lo, hi, hmin, nullptr, d_toff, out_lo, out_hi, out_cnt);
  else rle_kernel<2, 1><<<(unsigned)nt, RL_THREADS, 0, st>>>(d_tiles, d_se, d_sb, lo, hi, hmin, nullptr, d_toff, out_lo, out_hi, out_cnt);
  *launches += 1;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace kmx
