// common.cuh -- shared device helpers for libkmx_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace kmx {

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- nucleotide code / validity: (c>>1)&3 => A0 C1 T2 G3 (gatb Data.hpp:179), only ACGTacgt valid
__device__ __forceinline__ u32 nt_code(u32 c) { return (c >> 1) & 3u; }
__device__ __forceinline__ bool nt_valid(u32 c)
{
  // letters have (c & 0xC0) == 0x40; index by low 5 bits: A=1 C=3 G=7 T=20
  const u32 mask = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);
  return ((c & 0xC0u) == 0x40u) && ((mask >> (c & 31u)) & 1u);
}

// reverse the order of the 32 2-bit groups of x
__device__ __forceinline__ u64 rev2(u64 x)
{
  x = __brevll(x);
  return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}
// reverse complement of a k-mer (k <= 32) in the low 2k bits; complement = code ^ 2
__device__ __forceinline__ u64 revcomp64(u64 x, int k)
{
  return (rev2(x) ^ 0xAAAAAAAAAAAAAAAAULL) >> (64 - 2 * k);
}
// reverse complement for 32 < k <= 64: value = (hi,lo)
__device__ __forceinline__ void revcomp128(u64 lo, u64 hi, int k, u64& rlo, u64& rhi)
{
  u64 a = rev2(lo) ^ 0xAAAAAAAAAAAAAAAAULL;   // becomes the high word
  u64 b = rev2(hi) ^ 0xAAAAAAAAAAAAAAAAULL;   // becomes the low word
  int sh = 128 - 2 * k;                        // 0 <= sh < 64
  if (sh == 0) { rlo = b; rhi = a; }
  else { rlo = (b >> sh) | (a << (64 - sh)); rhi = a >> sh; }
}

// ---- XXH64 for 8- and 16-byte inputs, seed 0 (xxHash 0.8.3 xxhash.h:3454-3673 short path)
#define KMX_P1 0x9E3779B185EBCA87ULL
#define KMX_P2 0xC2B2AE3D27D4EB4FULL
#define KMX_P3 0x165667B19E3779F9ULL
#define KMX_P4 0x85EBCA77C2B2AE63ULL
#define KMX_P5 0x27D4EB2F165667C5ULL
__device__ __forceinline__ u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 xxh64_word(u64 h, u64 w)
{
  u64 k1 = rotl64(w * KMX_P2, 31) * KMX_P1;
  return rotl64(h ^ k1, 27) * KMX_P1 + KMX_P4;
}
__device__ __forceinline__ u64 xxh64_fin(u64 h)
{
  h ^= h >> 33; h *= KMX_P2; h ^= h >> 29; h *= KMX_P3; h ^= h >> 32;
  return h;
}
__device__ __forceinline__ u64 xxh64_8(u64 w0) { return xxh64_fin(xxh64_word(KMX_P5 + 8ULL, w0)); }
__device__ __forceinline__ u64 xxh64_16(u64 w0, u64 w1)
{
  return xxh64_fin(xxh64_word(xxh64_word(KMX_P5 + 16ULL, w0), w1));
}

// exact x % d for a loop-invariant 64-bit d: q = floor(x * M / 2^128)-style (Lemire fastmod,
// 128-bit magic M = floor(2^128 / d) + 1 held as (mhi, mlo)).  Exact for all 64-bit x, d >= 1.
struct FastMod64 { u64 d, mlo, mhi; };
__device__ __forceinline__ u64 fastmod64(u64 x, const FastMod64& f)
{
  // lowbits = (M * x) mod 2^128 ; result = (lowbits * d) >> 128
  u64 l_lo = f.mlo * x;
  u64 l_hi = f.mhi * x + __umul64hi(f.mlo, x);
  // (l_hi:l_lo) * d >> 128
  u64 t = __umul64hi(l_lo, f.d);
  u64 p_lo = l_hi * f.d;
  u64 p_hi = __umul64hi(l_hi, f.d);
  u64 s = p_lo + t;
  return p_hi + (s < p_lo ? 1ULL : 0ULL);
}

// exact x % d for 2 <= d < 2^31 (the usual case: a partition's window has far fewer slots):
// one Barrett step.  m64 = floor((2^64-1) / d), host-computed.  q = floor(x*m64 / 2^64) is the
// true quotient or one less (x*m64/2^64 > x/d - 1), so r = x - q*d < 2d < 2^32 and the 32-bit
// arithmetic below is exact; min(r, r-d) is the conditional subtract (r-d wraps when r < d).
struct FastMod32 { u32 d; u64 m64; };
__device__ __forceinline__ u32 fastmod64_d32(u64 x, const FastMod32& f)
{
  const u32 q = (u32)__umul64hi(x, f.m64);
  const u32 r = (u32)x - q * f.d;
  return min(r, r - f.d);
}

// ---- super-k-mer bucket record (own format; SURVEY F6 allows any) -----------------------
// w=1 (k<=32):  16 bytes: 128-bit big number V (first base most significant) in bits [0,120),
//               n = number of bases (k <= n <= 60) in bits [120,128).
// w=2 (k<=64):  32 bytes: 256-bit V in bits [0,248), n (<= 124) in bits [248,256).
// A record with n bases holds n-k+1 consecutive k-mers of one partition.
#define KMX_REC1_MAXN 60
#define KMX_REC2_MAXN 124

__device__ __forceinline__ u32 warp_lane() { return threadIdx.x & 31u; }

}  // namespace kmx
