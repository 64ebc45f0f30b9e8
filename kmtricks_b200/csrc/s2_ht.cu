// s2_ht.cu -- open-addressed hash-count in HBM for k-mer keys (k <= 32, one 64-bit word).
//
// Replaces (behaviour, not code) KmerSort + KmerPartCounter::executeDump
// (include/kmtricks/gatb/sorting_count.hpp:498-508,694-884): instead of sorting every k-mer
// occurrence of a (sample, partition), occurrences are counted in a per-partition linear-probing
// table (u64 key + u32 count, power-of-two capacity), and only the DISTINCT keys that survive
// hard-min are sorted afterwards (the output contract is still "ascending distinct canonical
// k-mers with counts").  At 30x coverage that is ~20x fewer keys through the radix sort.
//
//   ht_insert_records : warp stages 32 super-k-mer records, lanes walk their k-mers (same
//                       schedule as hash_hist_kernel), canonical k-mer, probe: plain L2 read
//                       first (most occurrences are repeats), CAS only on an empty slot,
//                       RED.ADD on the count.
//   ht_insert_keys    : same table filled from key lists (stage 3: union of N samples' keys).
//   ht_compact        : survivors (count >= hard_min) of each partition's table, unordered,
//                       block-aggregated append.
//   ht_lookup         : after the survivors are sorted: count of each key by probing again.
#include "common.cuh"
#include "kmx_internal.h"
#include "records.cuh"
#include <algorithm>

namespace kmx {

static constexpr u64 HT_EMPTY = ~0ULL;
static constexpr int HT_THREADS = 256;
static constexpr int HT_WARPS = HT_THREADS / 32;
static constexpr u32 HT_MAX_PROBE = 128;

__device__ __forceinline__ u64 ht_mix(u64 x)
{
  x ^= x >> 31; x *= 0x7FB5D329728EA185ULL; x ^= x >> 27; x *= 0x81DADEF4BC2DD44DULL; x ^= x >> 33;
  return x;
}

// returns false when the probe limit is hit (table too full)
__device__ __forceinline__ bool ht_add(u64* __restrict__ keys, u32* __restrict__ cnts, u64 mask, u64 c, u32 inc)
{
  u64 slot = ht_mix(c) & mask;
  for (u32 probe = 0; probe < HT_MAX_PROBE; probe++) {
    u64 cur = __ldcg(keys + slot);
    if (cur == HT_EMPTY) cur = atomicCAS((unsigned long long*)(keys + slot), HT_EMPTY, c);
    if (cur == HT_EMPTY || cur == c) { if (inc) atomicAdd(cnts + slot, inc); return true; }
    slot = (slot + 1) & mask;
  }
  return false;
}

// grid (x, P); partition p's table = keys/cnts + toff[p], capacity tcap[p] (power of two)
__global__ void __launch_bounds__(HT_THREADS)
ht_insert_records(const uint4* __restrict__ recs, const u64* __restrict__ boff, const u32* __restrict__ bcnt, int k,
                  u64* __restrict__ keys, u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
                  u32* __restrict__ overflow)
{
  __shared__ uint4 s_rec[HT_WARPS][32];
  __shared__ u32 s_pref[HT_WARPS][33];
  const u32 p = blockIdx.y;
  const u32 n = bcnt[p];
  const u64 b0 = boff[p];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u64* __restrict__ tk = keys + toff[p];
  u32* __restrict__ tc = cnts + toff[p];
  const u64 mask = tcap[p] - 1;
  bool ok = true;
  for (u32 r0 = (blockIdx.x * HT_WARPS + w) * 32; r0 < n; r0 += gridDim.x * HT_THREADS) {
    const u32 r = r0 + lane;
    u32 nk = 0;
    if (r < n) { uint4 v = __ldg(recs + b0 + r); s_rec[w][lane] = v; nk = (v.w >> 24) - k + 1; }
    u32 x = nk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
    s_pref[w][lane + 1] = x;
    if (lane == 0) s_pref[w][0] = 0;
    const u32 T = __shfl_sync(0xffffffffu, x, 31);
    __syncwarp();
    for (u32 t = lane; t < T; t += 32) {
      u32 q = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1) if (s_pref[w][q + st] <= t) q += st;
      const int j = (int)(t - s_pref[w][q]);
      const uint4 v = s_rec[w][q];
      Rec1 rec; rec.lo = (u64)v.x | ((u64)v.y << 32);
      const u64 hh = (u64)v.z | ((u64)v.w << 32);
      rec.n = (int)(hh >> 56); rec.hi = hh & 0x00FFFFFFFFFFFFFFULL;
      u64 c; canon1(rec, k, j, c);
      ok &= ht_add(tk, tc, mask, c, 1u);
    }
    __syncwarp();
  }
  if (!ok) *overflow = 1u;
}

// stage 3: union of key lists (counts unused: inc = 0 keeps cnts untouched, presence only)
__global__ void __launch_bounds__(HT_THREADS)
ht_insert_keys(const MergeList* __restrict__ lists, u64* __restrict__ keys, u32* __restrict__ cnts, u64 cap, u32* __restrict__ overflow)
{
  const MergeList L = lists[blockIdx.y];
  const u64 mask = cap - 1;
  bool ok = true;
  for (u64 i = (u64)blockIdx.x * HT_THREADS + threadIdx.x; i < L.n; i += (u64)gridDim.x * HT_THREADS)
    ok &= ht_add(keys, cnts, mask, L.lo[i], 0u);
  if (!ok) *overflow = 1u;
}

// survivors of partition p -> out[ooff[p] + ...] (unordered); pcnt[p] = how many
__global__ void __launch_bounds__(HT_THREADS)
ht_compact(const u64* __restrict__ keys, const u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
           u32 hmin, u64* __restrict__ out, u32* __restrict__ pcnt)
{
  __shared__ u32 s_warp[HT_WARPS];
  __shared__ u32 s_base;
  const u32 p = blockIdx.y;
  const u64 cap = tcap[p], t0 = toff[p];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u64 i0 = (u64)blockIdx.x * HT_THREADS; i0 < cap; i0 += (u64)gridDim.x * HT_THREADS) {
    const u64 i = i0 + threadIdx.x;
    u64 key = HT_EMPTY; u32 c = 0;
    if (i < cap) { key = keys[t0 + i]; c = cnts[t0 + i]; }
    const bool surv = key != HT_EMPTY && c >= hmin;
    const u32 bal = __ballot_sync(0xffffffffu, surv);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 tot = 0;
      for (int q = 0; q < HT_WARPS; q++) { u32 v = s_warp[q]; s_warp[q] = tot; tot += v; }
      s_base = tot ? atomicAdd(&pcnt[p], tot) : 0;
    }
    __syncthreads();
    if (surv) out[t0 + s_base + s_warp[w] + __popc(bal & ((1u << lane) - 1u))] = key;
    __syncthreads();
  }
}

// sorted survivors of partition p at skeys[soff[p] .. +pcnt[p]) -> final lists at oo[p]
__global__ void __launch_bounds__(HT_THREADS)
ht_lookup(const u64* __restrict__ keys, const u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
          const u64* __restrict__ skeys, const u64* __restrict__ soff, const u32* __restrict__ pcnt, const u64* __restrict__ oo,
          u64* __restrict__ out_keys, u32* __restrict__ out_cnt)
{
  const u32 p = blockIdx.y;
  const u64 mask = tcap[p] - 1, t0 = toff[p];
  const u32 n = pcnt[p];
  for (u32 i = blockIdx.x * HT_THREADS + threadIdx.x; i < n; i += gridDim.x * HT_THREADS) {
    const u64 c = skeys[soff[p] + i];
    u64 slot = ht_mix(c) & mask;
    while (keys[t0 + slot] != c) slot = (slot + 1) & mask;      // the key is in the table
    out_keys[oo[p] + i] = c;
    out_cnt[oo[p] + i] = cnts[t0 + slot];
  }
}

// compaction of a single table (stage 3 union): distinct keys, unordered
__global__ void __launch_bounds__(HT_THREADS)
ht_compact_keys(const u64* __restrict__ keys, u64 cap, u64* __restrict__ out, u32* __restrict__ count)
{
  __shared__ u32 s_warp[HT_WARPS];
  __shared__ u32 s_base;
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u64 i0 = (u64)blockIdx.x * HT_THREADS; i0 < cap; i0 += (u64)gridDim.x * HT_THREADS) {
    const u64 i = i0 + threadIdx.x;
    const u64 key = i < cap ? keys[i] : HT_EMPTY;
    const bool surv = key != HT_EMPTY;
    const u32 bal = __ballot_sync(0xffffffffu, surv);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 tot = 0;
      for (int q = 0; q < HT_WARPS; q++) { u32 v = s_warp[q]; s_warp[q] = tot; tot += v; }
      s_base = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    if (surv) out[s_base + s_warp[w] + __popc(bal & ((1u << lane) - 1u))] = key;
    __syncthreads();
  }
}

// ---- 128-bit keys (32 < k <= 63): same scheme, key slots are 16-byte (lo, hi) pairs claimed with one
// 128-bit CAS (ATOMG.CAS.128).  The high word of a k-mer never reaches 2^62, so hi == ~0 marks an empty slot.
__device__ __forceinline__ void cas128(ulonglong2* p, u64 clo, u64 chi, u64 nlo, u64 nhi, u64& olo, u64& ohi)
{
  asm volatile("{\n.reg .b128 c, s, d;\nmov.b128 c, {%2, %3};\nmov.b128 s, {%4, %5};\natom.global.relaxed.gpu.cas.b128 d, [%6], c, s;\nmov.b128 {%0, %1}, d;\n}"
               : "=l"(olo), "=l"(ohi) : "l"(clo), "l"(chi), "l"(nlo), "l"(nhi), "l"(p) : "memory");
}
__device__ __forceinline__ u64 ht2_slot(u64 lo, u64 hi) { return ht_mix(lo ^ (hi * 0x9E3779B97F4A7C15ULL)); }

__device__ __forceinline__ bool ht2_add(ulonglong2* __restrict__ keys, u32* __restrict__ cnts, u64 mask, u64 lo, u64 hi, u32 inc)
{
  u64 slot = ht2_slot(lo, hi) & mask;
  for (u32 probe = 0; probe < HT_MAX_PROBE; probe++) {
    ulonglong2 cur = __ldcg(keys + slot);
    // empty-looking slot (or a key whose low word equals the filler, where a torn view could fake a match): let the CAS decide
    if (cur.y == HT_EMPTY || lo == HT_EMPTY) {
      u64 olo, ohi;
      cas128(keys + slot, HT_EMPTY, HT_EMPTY, lo, hi, olo, ohi);
      if (olo == HT_EMPTY && ohi == HT_EMPTY) { cur.x = lo; cur.y = hi; } else { cur.x = olo; cur.y = ohi; }
    }
    if (cur.x == lo && cur.y == hi) { if (inc) atomicAdd(cnts + slot, inc); return true; }
    slot = (slot + 1) & mask;
  }
  return false;
}

__global__ void __launch_bounds__(HT_THREADS)
ht2_insert_records(const uint4* __restrict__ recs, const u64* __restrict__ boff, const u32* __restrict__ bcnt, int k,
                   ulonglong2* __restrict__ keys, u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
                   u32* __restrict__ overflow)
{
  __shared__ uint4 s_rec[HT_WARPS][64];
  __shared__ u32 s_pref[HT_WARPS][33];
  const u32 p = blockIdx.y;
  const u32 n = bcnt[p];
  const u64 b0 = boff[p];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  ulonglong2* __restrict__ tk = keys + toff[p];
  u32* __restrict__ tc = cnts + toff[p];
  const u64 mask = tcap[p] - 1;
  bool ok = true;
  for (u32 r0 = (blockIdx.x * HT_WARPS + w) * 32; r0 < n; r0 += gridDim.x * HT_THREADS) {
    const u32 r = r0 + lane;
    u32 nk = 0;
    if (r < n) {
      const uint4 a = __ldg(recs + 2 * (b0 + r)), b = __ldg(recs + 2 * (b0 + r) + 1);
      s_rec[w][2 * lane] = a; s_rec[w][2 * lane + 1] = b;
      nk = (b.w >> 24) - k + 1;
    }
    u32 x = nk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += y; }
    s_pref[w][lane + 1] = x;
    if (lane == 0) s_pref[w][0] = 0;
    const u32 T = __shfl_sync(0xffffffffu, x, 31);
    __syncwarp();
    for (u32 t = lane; t < T; t += 32) {
      u32 q = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1) if (s_pref[w][q + st] <= t) q += st;
      const int j = (int)(t - s_pref[w][q]);
      const uint4 a = s_rec[w][2 * q], b = s_rec[w][2 * q + 1];
      Rec2 rec;
      rec.v0 = (u64)a.x | ((u64)a.y << 32); rec.v1 = (u64)a.z | ((u64)a.w << 32);
      rec.v2 = (u64)b.x | ((u64)b.y << 32);
      const u64 hh = (u64)b.z | ((u64)b.w << 32);
      rec.n = (int)(hh >> 56); rec.v3 = hh & 0x00FFFFFFFFFFFFFFULL;
      u64 clo, chi; canon2(rec, k, j, clo, chi);
      ok &= ht2_add(tk, tc, mask, clo, chi, 1u);
    }
    __syncwarp();
  }
  if (!ok) *overflow = 1u;
}

// stage 3: union of 128-bit key lists (presence only)
__global__ void __launch_bounds__(HT_THREADS)
ht2_insert_keys(const MergeList* __restrict__ lists, ulonglong2* __restrict__ keys, u64 cap, u32* __restrict__ overflow)
{
  const MergeList L = lists[blockIdx.y];
  const u64 mask = cap - 1;
  bool ok = true;
  for (u64 i = (u64)blockIdx.x * HT_THREADS + threadIdx.x; i < L.n; i += (u64)gridDim.x * HT_THREADS)
    ok &= ht2_add(keys, nullptr, mask, L.lo[i], L.hi[i], 0u);
  if (!ok) *overflow = 1u;
}

// survivors of partition p -> out_lo/out_hi[t0 + ...] (unordered); pcnt[p] = how many.  cnts == NULL: every occupied slot
__global__ void __launch_bounds__(HT_THREADS)
ht2_compact(const ulonglong2* __restrict__ keys, const u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
            u64 one_cap, u32 hmin, u64* __restrict__ out_lo, u64* __restrict__ out_hi, u32* __restrict__ pcnt)
{
  __shared__ u32 s_warp[HT_WARPS];
  __shared__ u32 s_base;
  const u32 p = blockIdx.y;
  const u64 cap = tcap ? tcap[p] : one_cap, t0 = toff ? toff[p] : 0;
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u64 i0 = (u64)blockIdx.x * HT_THREADS; i0 < cap; i0 += (u64)gridDim.x * HT_THREADS) {
    const u64 i = i0 + threadIdx.x;
    ulonglong2 key = make_ulonglong2(HT_EMPTY, HT_EMPTY); u32 c = hmin;
    if (i < cap) { key = keys[t0 + i]; if (cnts) c = cnts[t0 + i]; }
    const bool surv = key.y != HT_EMPTY && c >= hmin;
    const u32 bal = __ballot_sync(0xffffffffu, surv);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 tot = 0;
      for (int q = 0; q < HT_WARPS; q++) { u32 v = s_warp[q]; s_warp[q] = tot; tot += v; }
      s_base = tot ? atomicAdd(&pcnt[p], tot) : 0;
    }
    __syncthreads();
    if (surv) { const u64 o = t0 + s_base + s_warp[w] + __popc(bal & ((1u << lane) - 1u)); out_lo[o] = key.x; out_hi[o] = key.y; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(HT_THREADS)
ht2_lookup(const ulonglong2* __restrict__ keys, const u32* __restrict__ cnts, const u64* __restrict__ toff, const u64* __restrict__ tcap,
           const u64* __restrict__ slo, const u64* __restrict__ shi, const u64* __restrict__ soff, const u32* __restrict__ pcnt,
           const u64* __restrict__ oo, u64* __restrict__ out_lo, u64* __restrict__ out_hi, u32* __restrict__ out_cnt)
{
  const u32 p = blockIdx.y;
  const u64 mask = tcap[p] - 1, t0 = toff[p];
  const u32 n = pcnt[p];
  for (u32 i = blockIdx.x * HT_THREADS + threadIdx.x; i < n; i += gridDim.x * HT_THREADS) {
    const u64 lo = slo[soff[p] + i], hi = shi[soff[p] + i];
    u64 slot = ht2_slot(lo, hi) & mask;
    for (;;) { const ulonglong2 cur = keys[t0 + slot]; if (cur.x == lo && cur.y == hi) break; slot = (slot + 1) & mask; }   // the key is in the table
    out_lo[oo[p] + i] = lo; out_hi[oo[p] + i] = hi;
    out_cnt[oo[p] + i] = cnts[t0 + slot];
  }
}

cudaError_t launch_ht2_insert_records(const S2Common& c, void* keys, u32* cnts, const u64* toff, const u64* tcap, u32* overflow,
                                      cudaStream_t st, u64* launches)
{
  if (c.max_bcnt == 0) return cudaSuccess;
  unsigned gx = (c.max_bcnt + HT_THREADS - 1) / HT_THREADS;
  if (gx > 592) gx = 592;
  ht2_insert_records<<<dim3(gx, c.P), HT_THREADS, 0, st>>>((const uint4*)c.records, c.boff, c.bcnt, c.k, (ulonglong2*)keys, cnts, toff, tcap, overflow);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_ht2_compact(u32 P, u64 max_cap, const void* keys, const u32* cnts, const u64* toff, const u64* tcap, u32 hard_min,
                               u64* out_lo, u64* out_hi, u32* pcnt, cudaStream_t st, u64* launches)
{
  unsigned gx = (unsigned)std::min<u64>((max_cap + HT_THREADS - 1) / HT_THREADS, 592);
  ht2_compact<<<dim3(gx ? gx : 1, P), HT_THREADS, 0, st>>>((const ulonglong2*)keys, cnts, toff, tcap, 0, hard_min ? hard_min : 1, out_lo, out_hi, pcnt);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_ht2_lookup(u32 P, u32 max_n, const void* keys, const u32* cnts, const u64* toff, const u64* tcap, const u64* slo, const u64* shi,
                              const u64* soff, const u32* pcnt, const u64* oo, u64* out_lo, u64* out_hi, u32* out_cnt, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  unsigned gx = (unsigned)std::min<u32>((max_n + HT_THREADS - 1) / HT_THREADS, 592);
  ht2_lookup<<<dim3(gx, P), HT_THREADS, 0, st>>>((const ulonglong2*)keys, cnts, toff, tcap, slo, shi, soff, pcnt, oo, out_lo, out_hi, out_cnt);
  *launches += 1;
  return cudaGetLastError();
}

// union of N 128-bit key lists: distinct keys (unordered) -> out_lo/out_hi, *count = how many
cudaError_t launch_ht2_union(const MergeList* d_lists, u32 N, u64 max_n, void* keys, u64 cap, u32* overflow,
                             u64* out_lo, u64* out_hi, u32* count, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  unsigned gx = (unsigned)std::min<u64>((max_n + HT_THREADS * 4 - 1) / (HT_THREADS * 4), 1024);
  ht2_insert_keys<<<dim3(gx ? gx : 1, N), HT_THREADS, 0, st>>>(d_lists, (ulonglong2*)keys, cap, overflow);
  unsigned gc = (unsigned)std::min<u64>((cap + HT_THREADS - 1) / HT_THREADS, 1184);
  ht2_compact<<<dim3(gc ? gc : 1, 1), HT_THREADS, 0, st>>>((const ulonglong2*)keys, nullptr, nullptr, nullptr, cap, 1, out_lo, out_hi, count);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_ht_insert_records(const S2Common& c, u64* keys, u32* cnts, const u64* toff, const u64* tcap, u32* overflow,
                                     cudaStream_t st, u64* launches)
{
  if (c.max_bcnt == 0) return cudaSuccess;
  unsigned gx = (c.max_bcnt + HT_THREADS - 1) / HT_THREADS;
  if (gx > 592) gx = 592;
  ht_insert_records<<<dim3(gx, c.P), HT_THREADS, 0, st>>>((const uint4*)c.records, c.boff, c.bcnt, c.k, keys, cnts, toff, tcap, overflow);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_ht_compact(u32 P, u64 max_cap, const u64* keys, const u32* cnts, const u64* toff, const u64* tcap, u32 hard_min,
                              u64* out, u32* pcnt, cudaStream_t st, u64* launches)
{
  unsigned gx = (unsigned)std::min<u64>((max_cap + HT_THREADS - 1) / HT_THREADS, 592);
  ht_compact<<<dim3(gx ? gx : 1, P), HT_THREADS, 0, st>>>(keys, cnts, toff, tcap, hard_min ? hard_min : 1, out, pcnt);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_ht_lookup(u32 P, u32 max_n, const u64* keys, const u32* cnts, const u64* toff, const u64* tcap, const u64* skeys,
                             const u64* soff, const u32* pcnt, const u64* oo, u64* out_keys, u32* out_cnt, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  unsigned gx = (unsigned)std::min<u32>((max_n + HT_THREADS - 1) / HT_THREADS, 592);
  ht_lookup<<<dim3(gx, P), HT_THREADS, 0, st>>>(keys, cnts, toff, tcap, skeys, soff, pcnt, oo, out_keys, out_cnt);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_ht_union(const MergeList* d_lists, u32 N, u64 max_n, u64* keys, u32* cnts, u64 cap, u32* overflow,
                            u64* out, u32* count, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  unsigned gx = (unsigned)std::min<u64>((max_n + HT_THREADS * 4 - 1) / (HT_THREADS * 4), 1024);
  ht_insert_keys<<<dim3(gx ? gx : 1, N), HT_THREADS, 0, st>>>(d_lists, keys, cnts, cap, overflow);
  unsigned gc = (unsigned)std::min<u64>((cap + HT_THREADS - 1) / HT_THREADS, 1184);
  ht_compact_keys<<<gc ? gc : 1, HT_THREADS, 0, st>>>(keys, cap, out, count);
  *launches += 2;
  return cudaGetLastError();
}

}  // namespace kmx
