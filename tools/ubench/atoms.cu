// Micro-benchmarks that decide the stage-2 design: shared-memory atomic throughput at full occupancy
// (spread addresses), L2 RED throughput, and the pure XXH64 + modulo issue cost.  Build + run: tools/ubench/run.sh
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../kmtricks_b200/csrc/common.cuh"
using namespace kmx;

__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>   // 0: ATOMS no return, 1: ATOMS with return, 2: 16-bit packed add no return, 3: plain STS (floor), 4: mix only
__global__ void __launch_bounds__(256) k_atoms(u32* out, int iters, u32 words)
{
  extern __shared__ u32 sm[];
  for (u32 i = threadIdx.x; i < words; i += 256) sm[i] = 0;
  __syncthreads();
  u32 x = blockIdx.x * 256 + threadIdx.x, acc = 0;
  for (int i = 0; i < iters; i++) {
    x = mix(x + i);
    const u32 a = x % words;
    if (MODE == 0) atomicAdd(&sm[a], 1u);
    else if (MODE == 1) acc += atomicAdd(&sm[a], 1u);
    else if (MODE == 2) atomicAdd(&sm[a], 1u << (16u * ((x >> 31) & 1u)));
    else if (MODE == 3) sm[a] = x;
    else acc += a;
  }
  __syncthreads();
  if (MODE != 0 && MODE != 2) { if (acc == 0x12345) out[0] = acc; }
  if (threadIdx.x == 0) out[1 + (blockIdx.x & 1023)] = sm[x % words];
}

__global__ void __launch_bounds__(256) k_red(u32* hist, int iters, u32 words)
{
  u32 x = blockIdx.x * 256 + threadIdx.x;
  for (int i = 0; i < iters; i++) { x = mix(x + i); atomicAdd(&hist[x % words], 1u); }
}

__global__ void __launch_bounds__(256) k_hash(u32* out, int iters, FastMod32 fm)
{
  u64 c = (u64)(blockIdx.x * 256 + threadIdx.x) * 0x9E3779B97F4A7C15ULL;
  u32 acc = 0;
  for (int i = 0; i < iters; i++) { c = (c << 2) | (acc & 3); acc += fastmod64_d32(xxh64_8(c), fm); }
  if (acc == 0x12345) out[0] = acc;
}

template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main()
{
  u32* out; cudaMalloc(&out, 1 << 20);
  u32* hist; cudaMalloc(&hist, 64 << 20); cudaMemset(hist, 0, 64 << 20);
  const int iters = 4096;
  const char* names[] = {"ATOMS.ADD no return", "ATOMS.ADD with return", "ATOMS.ADD 16-bit packed", "STS random", "address math only"};
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    if (ctas == 8) ctas = 3;                      // 3 x 64 KB fits one SM
    const u32 words = 16384; const size_t smem = words * 4;
    const int grid = 148 * ctas;
    const double ops = (double)grid * 256 * iters;
    float ms[5];
    cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_atoms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_atoms<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_atoms<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_atoms<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ms[0] = timeit([&] { k_atoms<0><<<grid, 256, smem>>>(out, iters, words); });
    ms[1] = timeit([&] { k_atoms<1><<<grid, 256, smem>>>(out, iters, words); });
    ms[2] = timeit([&] { k_atoms<2><<<grid, 256, smem>>>(out, iters, words); });
    ms[3] = timeit([&] { k_atoms<3><<<grid, 256, smem>>>(out, iters, words); });
    ms[4] = timeit([&] { k_atoms<4><<<grid, 256, smem>>>(out, iters, words); });
    for (int m = 0; m < 5; m++) printf("smem 64KB table, %d CTA/SM x 256 thr: %-26s %8.3f ms  %.3e ops/s  (1.2e8 ops = %.3f ms)\n", ctas, names[m], ms[m], ops / ms[m] * 1e3, 1.2e8 / (ops / ms[m] * 1e3) * 1e3);
    if (ctas == 3) break;
  }
  for (u32 mb : {4u, 12u, 48u}) {
    const u32 words = mb << 18; const int grid = 148 * 8; const double ops = (double)grid * 256 * iters;
    float ms = timeit([&] { k_red<<<grid, 256>>>(hist, iters, words); });
    printf("L2 RED.ADD random over %u MB: %8.3f ms  %.3e ops/s (1.2e8 ops = %.3f ms)\n", mb, ms, ops / ms * 1e3, 1.2e8 / (ops / ms * 1e3) * 1e3);
  }
  {
    FastMod32 fm; fm.d = 3125056; fm.m64 = (~0ULL) / fm.d;
    const int grid = 148 * 8; const double ops = (double)grid * 256 * iters;
    float ms = timeit([&] { k_hash<<<grid, 256>>>(out, iters, fm); });
    printf("XXH64(8 B) + Barrett mod, dependent chain per thread, 8 CTA/SM: %8.3f ms  %.3e hashes/s (1.2e8 = %.3f ms)\n", ms, ops / ms * 1e3, 1.2e8 / (ops / ms * 1e3) * 1e3);
  }
  return 0;
}
