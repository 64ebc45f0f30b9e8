// s4_bits.cu -- stage 4: bit-matrix transpose (hash:bft) and per-sample Bloom window vector.
//
// Replaces (behaviour, not code) BitMatrix::transpose / __sse_trans
// (include/kmtricks/bitmatrix.hpp:209-214,238-289: SSE2 movemask on 16x8 blocks) and
// HashVecProcessor (include/kmtricks/gatb/count_processor.hpp:84-120).
//
// Transpose: in = nrows x ncols bits (rows of ncols/8 bytes, LSB-first), out = ncols x nrows
// bits; out[c][r] = in[r][c].  Each warp owns a 32x32-bit block: lane l holds the 32 column
// bits of row r0+l, 32 warp ballots turn them into the 32 output words, which are staged in
// shared memory so that every output row is written as one coalesced 128-byte run.
#include "common.cuh"
#include "kmx_internal.h"

#include <algorithm>

namespace kmx {

__global__ void __launch_bounds__(1024)
transpose_bits_kernel(const uint8_t* __restrict__ in, u64 nrows, u64 ncols, uint8_t* __restrict__ out, int rows_in_x)
{
  __shared__ u32 s_t[32][33];
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u64 in_rb = ncols >> 3, out_rb = nrows >> 3;
  // the longer dimension goes to gridDim.x (gridDim.y stops at 65535): a Bloom window has millions of rows, its transpose millions of columns
  const u64 c0 = (u64)(rows_in_x ? blockIdx.y : blockIdx.x) * 32;      // first column (bit) of this CTA
  const u64 rblk0 = (u64)(rows_in_x ? blockIdx.x : blockIdx.y) * 32;   // first 32-row block of this CTA
  const u64 r = (rblk0 + w) * 32 + lane;                 // input row of this lane
  u32 word = 0;
  if (r < nrows) {
    const uint8_t* src = in + r * in_rb + (c0 >> 3);
#pragma unroll
    for (int b = 0; b < 4; b++)
      if ((c0 >> 3) + b < in_rb) word |= (u32)__ldg(src + b) << (8 * b);
  }
  u32 mine = 0;
#pragma unroll
  for (int b = 0; b < 32; b++) {
    u32 v = __ballot_sync(0xffffffffu, (word >> b) & 1u);
    if (lane == (u32)b) mine = v;
  }
  s_t[lane][w] = mine;                                   // out row c0+lane, word index rblk0+w
  __syncthreads();
  // warp w now writes out row c0+w: 32 consecutive words starting at word rblk0
  const u64 oc = c0 + w;
  if (oc < ncols) {
    u32 v = s_t[w][lane];
    u64 byte0 = (rblk0 + lane) * 4;                      // byte offset within the out row
    uint8_t* dst = out + oc * out_rb + byte0;
    if (byte0 + 4 <= out_rb && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) *reinterpret_cast<u32*>(dst) = v;
    else {
#pragma unroll
      for (int b = 0; b < 4; b++) if (byte0 + b < out_rb) dst[b] = (uint8_t)(v >> (8 * b));
    }
  }
}

cudaError_t launch_transpose_bits(const uint8_t* in, u64 nrows, u64 ncols, uint8_t* out, cudaStream_t st, u64* launches)
{
  if (!nrows || !ncols) return cudaSuccess;
  const u64 gc = (ncols + 31) / 32, gr = (nrows + 1023) / 1024;
  const int rows_in_x = gr >= gc;
  if (std::min(gc, gr) > 65535 || std::max(gc, gr) > 0x7FFFFFFFULL) return cudaErrorInvalidValue;
  dim3 grid((unsigned)(rows_in_x ? gr : gc), (unsigned)(rows_in_x ? gc : gr));
  transpose_bits_kernel<<<grid, 1024, 0, st>>>(in, nrows, ncols, out, rows_in_x);
  *launches += 1;
  return cudaGetLastError();
}

// bits must be zero-filled, padded to a multiple of 4 bytes
__global__ void hash_vector_kernel(const u64* __restrict__ keys, u64 n, u64 lower, u32* __restrict__ bits)
{
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { u64 b = keys[i] - lower; atomicOr(&bits[b >> 5], 1u << (b & 31)); }
}

cudaError_t launch_hash_vector(const u64* keys, u64 n, u64 lower, uint8_t* bits, cudaStream_t st, u64* launches)
{
  if (!n) return cudaSuccess;
  hash_vector_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, lower, reinterpret_cast<u32*>(bits));
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace kmx
