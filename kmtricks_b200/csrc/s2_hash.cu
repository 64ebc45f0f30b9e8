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

namespace kmx {

static constexpr int HH_THREADS = 256;
static constexpr int HH_WARPS = HH_THREADS / 32;

// grid: (x tiles, P).  Each warp takes 32 records at a time, stages them in shared memory with
// the prefix sum of their k-mer counts, and then walks the k-mers of those records with all 32
// lanes busy (lane t -> k-mer t, t+32, ...; the owning record is found by a 5-step binary
// search in the warp's prefix array).  Every lane runs the same instruction stream (a rolling
// variant with per-lane contiguous chunks was measured 1.9x SLOWER: lanes cross record
// boundaries at different steps, so the warp pays extract + roll at every step).
// The histogram update is a fire-and-forget RED; the modulo is an exact 32-bit Barrett when W < 2^32.
template <int W, bool D32>
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
  u32* __restrict__ h = hist + (u64)blockIdx.y * Wbits;
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
      atomicAdd(h + key, 1u);       // result unused -> RED.ADD
    }
    __syncwarp();
  }
}

// survivors per 64K-slot sub-chunk: grid = P*S CTAs
static constexpr int HC_THREADS = 256;
__global__ void __launch_bounds__(HC_THREADS)
hash_count_kernel(u64 Wbits, u32 S, const u32* __restrict__ hist, u32 hmin, u32* __restrict__ sub_counts,
                  const u32* __restrict__ bcnt, u32 p0)
{
  __shared__ u32 s_warp[HC_THREADS / 32];
  const u32 wl = blockIdx.x / S, s = blockIdx.x % S, p = p0 + wl;
  if (bcnt[p] == 0) { if (threadIdx.x == 0) sub_counts[(u64)p * S + s] = 0; return; }   // untouched window
  const u64 slot0 = (u64)s * HIST_SUB;
  const u64 slot1 = min(Wbits, slot0 + HIST_SUB);
  const uint4* __restrict__ h4 = reinterpret_cast<const uint4*>(hist + (u64)wl * Wbits);
  u32 c = 0;
  for (u64 q = slot0 / 4 + threadIdx.x; q < slot1 / 4; q += HC_THREADS) {
    uint4 v = h4[q];
    c += (v.x >= hmin) + (v.y >= hmin) + (v.z >= hmin) + (v.w >= hmin);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < HC_THREADS / 32; i++) t += s_warp[i]; sub_counts[(u64)p * S + s] = t; }
}

// One CTA per group: exclusive prefix of the group's sub-chunk survivor counts, output space
// bump-allocated from a device cursor (no host round trip per group).  meta: [0] cursor (u64),
// [1] capacity; flags[0] = overflow.  list_off / list_n per partition.
__global__ void __launch_bounds__(256)
hash_group_scan_kernel(u32 S, u32 p0, u32 gp, const u32* __restrict__ sub_counts, u64* __restrict__ sub_off,
                       u64* __restrict__ list_off, u64* __restrict__ list_n, u64* __restrict__ meta, u32* __restrict__ flags)
{
  __shared__ u64 s_warp[8];
  __shared__ u64 s_carry, s_base;
  const u64 n = (u64)gp * S, first = (u64)p0 * S;
  // total
  u64 tot = 0;
  for (u64 i = threadIdx.x; i < n; i += 256) tot += sub_counts[first + i];
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 t = 0; for (int i = 0; i < 8; i++) t += s_warp[i];
    const u64 base = atomicAdd((unsigned long long*)&meta[0], (unsigned long long)t);
    if (base + t > meta[1]) flags[0] = 1u;
    s_base = base; s_carry = 0;
  }
  __syncthreads();
  for (u64 b = 0; b < n; b += 256) {
    const u64 i = b + threadIdx.x;
    const u64 v = i < n ? sub_counts[first + i] : 0;
    u64 x = v;
    for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    u64 wb = 0, all = 0;
    for (int q = 0; q < 8; q++) { u64 t = s_warp[q]; if (q < (int)(threadIdx.x >> 5)) wb += t; all += t; }
    const u64 excl = s_base + s_carry + wb + x - v;
    if (i < n) {
      sub_off[first + i] = excl;
      if (i % S == 0) list_off[p0 + i / S] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += all;
    __syncthreads();
  }
  // list sizes: difference of consecutive partition starts (last one ends at base + total)
  __syncthreads();
  for (u32 q = threadIdx.x; q < gp; q += 256) {
    const u64 st = sub_off[first + (u64)q * S];
    const u64 en = (q + 1 < gp) ? sub_off[first + (u64)(q + 1) * S] : s_base + s_carry;
    list_n[p0 + q] = en - st;
  }
}

// grid = P*S CTAs; CTA (p, s) sweeps slots [s*HIST_SUB, min(W, (s+1)*HIST_SUB)) in order.
static constexpr int HE_THREADS = 256;
__global__ void __launch_bounds__(HE_THREADS)
hash_emit_kernel(u64 Wbits, u32 S, u32* __restrict__ hist, u32 hmin, const u64* __restrict__ sub_off,
                 const u32* __restrict__ sub_counts, u64* __restrict__ out_keys, u32* __restrict__ out_counts,
                 const u32* __restrict__ bcnt, const u32* __restrict__ win_part /* NULL: window p holds partition p */,
                 u32 p0, const u32* __restrict__ flags)
{
  __shared__ u32 s_warp[HE_THREADS / 32];
  __shared__ u32 s_run;
  const u32 wl = blockIdx.x / S, s = blockIdx.x % S, p = p0 + wl;
  if (bcnt[p] == 0) return;                               // window never touched: already all-zero
  const u64 slot0 = (u64)s * HIST_SUB;
  const u64 slot1 = min(Wbits, slot0 + HIST_SUB);
  uint4* __restrict__ h4 = reinterpret_cast<uint4*>(hist + (u64)wl * Wbits);   // W multiple of 64 -> aligned
  const u64 key_base = (u64)(win_part ? win_part[p] : p) * Wbits;
  const u64 obase = sub_off[(u64)p * S + s];
  if (sub_counts[(u64)p * S + s] == 0 || flags[0]) {     // nothing survives here, or the output space ran out: only re-zero
    // nothing survives here: still has to clear non-zero (below hard-min) slots
    for (u64 q = slot0 / 4 + threadIdx.x; q < slot1 / 4; q += HE_THREADS) {
      uint4 v = h4[q];
      if (v.x | v.y | v.z | v.w) h4[q] = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (u64 q0 = slot0 / 4; q0 < slot1 / 4; q0 += HE_THREADS) {
    u64 q = q0 + threadIdx.x;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < slot1 / 4) {
      v = h4[q];
      if (v.x | v.y | v.z | v.w) h4[q] = make_uint4(0, 0, 0, 0);
    }
    u32 c = (v.x >= hmin) + (v.y >= hmin) + (v.z >= hmin) + (v.w >= hmin);
    u32 x = c;
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    u32 wb = 0, tot = 0;
    for (int i = 0; i < HE_THREADS / 32; i++) { u32 t = s_warp[i]; if (i < (int)(threadIdx.x >> 5)) wb += t; tot += t; }
    u32 run = s_run;
    u64 o = obase + run + wb + x - c;
    if (c) {
      u64 kb = key_base + q * 4;
      if (v.x >= hmin) { out_keys[o] = kb; out_counts[o] = v.x; o++; }
      if (v.y >= hmin) { out_keys[o] = kb + 1; out_counts[o] = v.y; o++; }
      if (v.z >= hmin) { out_keys[o] = kb + 2; out_counts[o] = v.z; o++; }
      if (v.w >= hmin) { out_keys[o] = kb + 3; out_counts[o] = v.w; o++; }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_run = run + tot;
    __syncthreads();
  }
}

cudaError_t launch_hash_group(const S2Common& c, u64 Wbits, u64 mod_d, u64 mod_mlo, u64 mod_mhi, u32* hist, u32 hard_min,
                              u32 S, u32 p0, u32 gp, u32* sub_counts, u64* sub_off, u64* list_off, u64* list_n, u64* meta, u32* flags,
                              u64* out_keys, u32* out_counts, const u32* win_part, cudaStream_t st, u64* launches, int phase)
{
  const u32 hmin = hard_min ? hard_min : 1;
  if (phase == 0) {
    if (c.max_bcnt) {
      FastMod64 fm; fm.d = mod_d; fm.mlo = mod_mlo; fm.mhi = mod_mhi;
      FastMod32 f32; f32.d = (u32)mod_d; f32.m64 = mod_d >= 2 ? (~0ULL) / mod_d : 0;
      const bool d32 = mod_d >= 2 && mod_d < (1ULL << 32);
      unsigned gx = (c.max_bcnt + HH_THREADS - 1) / HH_THREADS;
      if (gx > 592) gx = 592;                 // 4 waves of 148 SMs per partition row at most
      dim3 grid(gx, gp);
      const uint4* recs = (const uint4*)c.records;
      if (c.W == 1 && d32) hash_hist_kernel<1, true><<<grid, HH_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0);
      else if (c.W == 1) hash_hist_kernel<1, false><<<grid, HH_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0);
      else if (d32) hash_hist_kernel<2, true><<<grid, HH_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0);
      else hash_hist_kernel<2, false><<<grid, HH_THREADS, 0, st>>>(recs, c.boff, c.bcnt, c.k, Wbits, fm, f32, hist, p0);
      *launches += 1;
    }
    hash_count_kernel<<<gp * S, HC_THREADS, 0, st>>>(Wbits, S, hist, hmin, sub_counts, c.bcnt, p0);
    *launches += 1;
  } else {
    hash_group_scan_kernel<<<1, 256, 0, st>>>(S, p0, gp, sub_counts, sub_off, list_off, list_n, meta, flags);
    hash_emit_kernel<<<gp * S, HE_THREADS, 0, st>>>(Wbits, S, hist, hmin, sub_off, sub_counts, out_keys, out_counts, c.bcnt, win_part, p0, flags);
    *launches += 2;
  }
  return cudaGetLastError();
}

}  // namespace kmx
