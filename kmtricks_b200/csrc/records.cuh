// records.cuh -- decoding of the super-k-mer bucket records written by stage 1 (common.cuh
// describes the layout) and canonical k-mer extraction.  Every k-mer of a record is extracted
// independently (shift + mask of the packed big number), so there is no rolling dependency.
#pragma once
#include "common.cuh"

namespace kmx {

// extract the k-mer starting at base j of a w=1 record (n bases, V big-endian in 120 bits)
__device__ __forceinline__ u64 rec1_kmer(u64 lo, u64 hi, int n, int k, int j)
{
  int sh = 2 * (n - k - j);                 // right shift of the 128-bit value
  u64 v;
  if (sh == 0) v = lo;
  else if (sh < 64) v = (lo >> sh) | (hi << (64 - sh));
  else v = hi >> (sh - 64);
  return (k == 32) ? v : (v & ((1ULL << (2 * k)) - 1ULL));
}

// 256-bit right shift by sh (< 256), return low 128 bits
__device__ __forceinline__ void shr256_lo128(u64 v0, u64 v1, u64 v2, u64 v3, int sh, u64& lo, u64& hi)
{
  int ws = sh >> 6, bs = sh & 63;
  u64 a[6] = {v0, v1, v2, v3, 0, 0};
  u64 x0 = a[ws], x1 = a[ws + 1], x2 = a[ws + 2];
  if (bs == 0) { lo = x0; hi = x1; }
  else { lo = (x0 >> bs) | (x1 << (64 - bs)); hi = (x1 >> bs) | (x2 << (64 - bs)); }
}

__device__ __forceinline__ void rec2_kmer(u64 v0, u64 v1, u64 v2, u64 v3, int n, int k, int j, u64& lo, u64& hi)
{
  // select by branches instead of a dynamically indexed local array
  int sh = 2 * (n - k - j);
  int ws = sh >> 6, bs = sh & 63;
  u64 x0, x1, x2;
  if (ws == 0) { x0 = v0; x1 = v1; x2 = v2; }
  else if (ws == 1) { x0 = v1; x1 = v2; x2 = v3; }
  else if (ws == 2) { x0 = v2; x1 = v3; x2 = 0; }
  else { x0 = v3; x1 = 0; x2 = 0; }
  if (bs == 0) { lo = x0; hi = x1; }
  else { lo = (x0 >> bs) | (x1 << (64 - bs)); hi = (x1 >> bs) | (x2 << (64 - bs)); }
  int hb = 2 * (k - 32);                    // k > 32 here
  if (hb < 64) hi &= ((1ULL << hb) - 1ULL);
}

struct Rec1 { u64 lo, hi; int n; };
struct Rec2 { u64 v0, v1, v2, v3; int n; };

__device__ __forceinline__ Rec1 load_rec1(const uint4* __restrict__ recs, u64 ridx)
{
  uint4 r = __ldg(recs + ridx);
  Rec1 o;
  o.lo = (u64)r.x | ((u64)r.y << 32);
  u64 h = (u64)r.z | ((u64)r.w << 32);
  o.n = (int)(h >> 56);
  o.hi = h & 0x00FFFFFFFFFFFFFFULL;
  return o;
}
__device__ __forceinline__ Rec2 load_rec2(const uint4* __restrict__ recs, u64 ridx)
{
  uint4 a = __ldg(recs + 2 * ridx), b = __ldg(recs + 2 * ridx + 1);
  Rec2 o;
  o.v0 = (u64)a.x | ((u64)a.y << 32); o.v1 = (u64)a.z | ((u64)a.w << 32);
  o.v2 = (u64)b.x | ((u64)b.y << 32);
  u64 h = (u64)b.z | ((u64)b.w << 32);
  o.n = (int)(h >> 56);
  o.v3 = h & 0x00FFFFFFFFFFFFFFULL;
  return o;
}

__device__ __forceinline__ void canon1(const Rec1& r, int k, int j, u64& c)
{
  u64 f = rec1_kmer(r.lo, r.hi, r.n, k, j);
  u64 rc = revcomp64(f, k);
  c = f < rc ? f : rc;
}
__device__ __forceinline__ void canon2(const Rec2& r, int k, int j, u64& clo, u64& chi)
{
  u64 flo, fhi, rlo, rhi;
  rec2_kmer(r.v0, r.v1, r.v2, r.v3, r.n, k, j, flo, fhi);
  revcomp128(flo, fhi, k, rlo, rhi);
  bool fl = (fhi < rhi) || (fhi == rhi && flo < rlo);
  clo = fl ? flo : rlo; chi = fl ? fhi : rhi;
}


}  // namespace kmx
