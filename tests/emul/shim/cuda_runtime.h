// Host stand-in for <cuda_runtime.h>, used ONLY by the CPU emulation tests (tests/emul/*.cpp, include path
// -Itests/emul/shim): the device arithmetic headers of the product (kmtricks_b200/csrc/common.cuh, records.cuh)
// are compiled unmodified with g++ and checked against the CPU restatement.  Test infrastructure, not product code.
#pragma once
#include <stdint.h>
#include <stddef.h>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }
struct KmxEmulDim3 { unsigned x, y, z; };
static KmxEmulDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline unsigned long long __brevll(unsigned long long x)
{
  unsigned long long r = 0;
  for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1ULL); x >>= 1; }
  return r;
}
static inline unsigned __brev(unsigned x) { return (unsigned)(__brevll(x) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
