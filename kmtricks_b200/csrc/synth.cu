// synth.cu -- device twin of kmtricks_b200/synth.py (counter-based synthetic FASTQ).
// One thread per output byte; identical bytes to the numpy generator (tests/test_synth.py).
#include "common.cuh"
#include "kmx_internal.h"

namespace kmx {

__device__ __forceinline__ u64 splitmix64(u64 x)
{
  u64 z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ u64 rnd(u64 stream, u64 idx) { return splitmix64(stream * 0xD1342543DE82EF95ULL + idx); }

__global__ void synth_fastq_kernel(u64 base, u32 sample, u64 first_read, u64 R, u32 L, u64 G, u32 thr_d, u32 thr_e,
                                   int revcomp, char* __restrict__ out)
{
  const u64 rb = 2ULL * L + 15;
  const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= R * rb) return;
  const u64 rl = gid / rb; const u32 o = (u32)(gid % rb);
  const u64 ridx = first_read + rl;
  char ch;
  if (o == 0) ch = '@';
  else if (o == 1) ch = 'r';
  else if (o < 10) { u64 v = ridx; for (u32 d = 9; d > o; d--) v /= 10; ch = (char)('0' + (v % 10)); }
  else if (o == 10) ch = '\n';
  else if (o < 11 + L) {
    const u32 j = o - 11;
    const u64 s_genome = base + 1, s_snp = base + 1000 + 4ULL * sample, s_start = s_snp + 1, s_err = s_snp + 2, s_strand = s_snp + 3;
    const u64 start = rnd(s_start, ridx) % (G - L + 1);
    bool rc = revcomp && (rnd(s_strand, ridx) & 1ULL);
    const u64 pos = rc ? start + (L - 1) - j : start + j;
    u32 b = (u32)(rnd(s_genome, pos) & 3ULL);
    u64 u = rnd(s_snp, pos);
    if ((u32)(u >> 32) < thr_d) b = (u32)((b + 1 + (u & 0xFFFFULL) % 3ULL) & 3ULL);
    if (rc) b = 3 - b;
    u = rnd(s_err, ridx * L + j);
    if ((u32)(u >> 32) < thr_e) b = (u32)((b + 1 + (u & 0xFFFFULL) % 3ULL) & 3ULL);
    ch = "ACGT"[b];
  }
  else if (o == 11 + L) ch = '\n';
  else if (o == 12 + L) ch = '+';
  else if (o == 13 + L) ch = '\n';
  else if (o < 14 + 2 * L) ch = 'I';
  else ch = '\n';
  out[gid] = ch;
}

cudaError_t launch_synth_fastq(u64 seed, u32 sample, u64 first_read, u64 R, u32 L, u64 G, u32 thr_d, u32 thr_e,
                               int revcomp, char* out, cudaStream_t st, u64* launches)
{
  if (!R) return cudaSuccess;
  u64 base = (seed * 1000003ULL) & 0x7FFFFFFFFFFFULL;
  u64 total = R * (2ULL * L + 15);
  synth_fastq_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(base, sample, first_read, R, L, G, thr_d, thr_e, revcomp, out);
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace kmx
