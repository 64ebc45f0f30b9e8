// s3_merge.cu -- stage 3/4: N-sample merge with soft-min / share-min rescue / recurrence-min,
// the six MergeStatistics vectors, and row emission (count, pa, dense Bloom slab).
//
// Replaces (behaviour, not code) KmerMerger::next / HashMerger::next
// (include/kmtricks/merge.hpp:183-260,441-517), MergeStatistics (:49-100) and
// write_as_bin / write_as_pa / write_as_bf (:262-286,519-600).
//
// The reference walks N sorted streams with an O(N) head scan per row.  Here a row is
// addressed directly: row = key - W*p for the dense Bloom slab, or the rank of the key in the
// sorted union of the partition's keys (binary search, L2 resident) for count / pa rows.
// Pass 1 accumulates solid_in[row] with atomics (only needed when rescue or recurrence > 1 is
// on), pass 2 classifies every (sample, key, count) entry, accumulates the statistics with
// one atomic per CTA and stat, and scatters the count / presence bit into the zero-filled
// output body.
#include "common.cuh"
#include "kmx_internal.h"

namespace kmx {

static constexpr int MG_THREADS = 256;

// ---------------------------------------------------------------------------------------
// row resolution
// ---------------------------------------------------------------------------------------
struct RowDense { u64 lower; __device__ __forceinline__ u64 row(u64 lo, u64) const { return lo - lower; } };

struct RowSparse {
  const u64* ulo; const u64* uhi; u64 nu; int W;
  __device__ __forceinline__ u64 row(u64 lo, u64 hi) const
  {
    // lower_bound in the sorted distinct keys (the key is always present)
    u64 a = 0, b = nu;
    while (a < b) {
      u64 mid = (a + b) >> 1;
      u64 ml = ulo[mid];
      bool less;
      if (W == 1) less = ml < lo;
      else { u64 mh = uhi[mid]; less = (mh < hi) || (mh == hi && ml < lo); }
      if (less) a = mid + 1; else b = mid;
    }
    return a;
  }
};

// ---------------------------------------------------------------------------------------
// pass 1: solid_in[row] += (count >= soft_min[s])
// grid (x, N)
// ---------------------------------------------------------------------------------------
template <class Row>
__global__ void __launch_bounds__(MG_THREADS)
merge_solid_kernel(const MergeList* __restrict__ lists, const u32* __restrict__ soft, Row rr, u32* __restrict__ solid_in)
{
  const u32 s = blockIdx.y;
  const MergeList L = lists[s];
  const u32 sm = soft[s];
  for (u64 i = (u64)blockIdx.x * MG_THREADS + threadIdx.x; i < L.n; i += (u64)gridDim.x * MG_THREADS) {
    u32 c = L.cnt[i];
    if (c >= sm) atomicAdd(&solid_in[rr.row(L.lo[i], L.hi ? L.hi[i] : 0)], 1u);
  }
}

// keep[row] = solid_in[row] >= rmin  (or 1 when emit_all)
__global__ void row_keep_kernel(const u32* __restrict__ solid_in, u64 nrows, u32 rmin, u32 emit_all, u32* __restrict__ keep)
{
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) keep[i] = (emit_all || solid_in[i] >= rmin) ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------
// pass 2: classify + stats + scatter
// fmt: 0 count rows, 1 pa rows, 2 dense bf slab
// ---------------------------------------------------------------------------------------
struct EmitArgs {
  const u32* soft; u32 rmin, share, emit_all;
  const u32* solid_in;       // NULL => not needed (share==0 && rmin<=1)
  const u32* out_row;        // sparse: exclusive scan of keep flags (NULL for dense)
  const u32* keep;           // sparse: keep flag per row (NULL for dense)
  uint8_t* body; u32 row_bytes; u32 key_bytes; int fmt;
  uint8_t* row_keep;         // emit_all: default keep decision per emitted row
  u64* stats; u32 N;
};

template <class Row>
__global__ void __launch_bounds__(MG_THREADS)
merge_emit_kernel(const MergeList* __restrict__ lists, Row rr, EmitArgs a)
{
  __shared__ u64 s_red[6][MG_THREADS / 32];
  const u32 s = blockIdx.y;
  const MergeList L = lists[s];
  const u32 sm = a.soft[s];
  u64 st_ns = 0, st_rd = 0, st_uwo = 0, st_uw = 0, st_two = 0, st_tw = 0;
  for (u64 i = (u64)blockIdx.x * MG_THREADS + threadIdx.x; i < L.n; i += (u64)gridDim.x * MG_THREADS) {
    const u32 c = L.cnt[i];
    const u64 lo = L.lo[i], hi = L.hi ? L.hi[i] : 0;
    const bool solid = c >= sm;
    u64 row = 0; bool have_row = false;
    u32 si = solid ? 1u : 0u;
    if (a.solid_in) { row = rr.row(lo, hi); have_row = true; si = a.solid_in[row]; }
    const bool rescued = !solid && a.share && si >= a.share;
    const bool keep = si >= a.rmin;
    const bool nz = solid || rescued;
    if (solid) { st_uwo++; st_uw++; st_two += c; st_tw += c; }
    else { st_ns++; if (rescued) { st_rd++; st_uw++; st_tw += c; } }
    if (a.fmt == 2) {
      if (keep && nz) {
        if (!have_row) row = rr.row(lo, hi);
        u64 byte = row * a.row_bytes + (s >> 3);
        u32* w = reinterpret_cast<u32*>(a.body + (byte & ~(u64)3));
        atomicOr(w, 1u << (8 * (u32)(byte & 3) + (s & 7)));
      }
    } else {
      if (!have_row) row = rr.row(lo, hi);
      if (a.keep[row]) {
        u64 orow = a.out_row[row];
        uint8_t* dst = a.body + orow * a.row_bytes + a.key_bytes;
        if (nz) {
          if (a.fmt == 0) reinterpret_cast<u32*>(dst)[s] = c;     // row_bytes, key_bytes multiples of 4
          else {
            u64 byte = (u64)(dst - a.body) + (s >> 3);
            u32* w = reinterpret_cast<u32*>(a.body + (byte & ~(u64)3));
            atomicOr(w, 1u << (8 * (u32)(byte & 3) + (s & 7)));
          }
        }
      }
    }
  }
  // block reduce the six statistics
  u64 v[6] = {st_ns, st_rd, st_uwo, st_uw, st_two, st_tw};
#pragma unroll
  for (int q = 0; q < 6; q++) {
    u64 x = v[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    u64 t = 0;
    for (int i = 0; i < MG_THREADS / 32; i++) t += s_red[threadIdx.x][i];
    if (t) atomicAdd(&a.stats[(u64)threadIdx.x * a.N + s], t);
  }
}

// sparse rows: write the key words of every kept row (+ default keep decision for emit_all)
__global__ void sparse_keys_kernel(const u64* __restrict__ ulo, const u64* __restrict__ uhi, u64 nu, int W,
                                   const u32* __restrict__ keep, const u32* __restrict__ out_row,
                                   const u32* __restrict__ solid_in, u32 rmin,
                                   uint8_t* __restrict__ body, u32 row_bytes, uint8_t* __restrict__ row_keep)
{
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu || !keep[i]) return;
  u64 o = out_row[i];
  // rows are 4-byte aligned only (row_bytes = 8W + 4N or 8W + ceil(N/8) padded by caller? no:
  // pa rows may have any byte length) -> byte-wise key store
  uint8_t* dst = body + o * row_bytes;
  u64 lo = ulo[i];
#pragma unroll
  for (int b = 0; b < 8; b++) dst[b] = (uint8_t)(lo >> (8 * b));
  if (W == 2) {
    u64 hi = uhi[i];
#pragma unroll
    for (int b = 0; b < 8; b++) dst[8 + b] = (uint8_t)(hi >> (8 * b));
  }
  if (row_keep) row_keep[o] = (solid_in[i] >= rmin) ? 1 : 0;
}

static unsigned grid_x_for(u64 max_n)
{
  u64 g = (max_n + MG_THREADS * 4 - 1) / (MG_THREADS * 4);
  if (g < 1) g = 1;
  if (g > 1024) g = 1024;
  return (unsigned)g;
}

cudaError_t launch_dense_solid(const MergeList* d_lists, u32 N, const u32* d_soft, u64 lower, u32* solid_in,
                               u64 max_n, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  RowDense rr; rr.lower = lower;
  merge_solid_kernel<RowDense><<<dim3(grid_x_for(max_n), N), MG_THREADS, 0, st>>>(d_lists, d_soft, rr, solid_in);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_dense_emit(const MergeList* d_lists, u32 N, const u32* d_soft, u32 rmin, u32 share,
                              u64 lower, const u32* solid_in, uint8_t* slab, u32 row_bytes, u64* stats,
                              u64 max_n, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  RowDense rr; rr.lower = lower;
  EmitArgs a;
  a.soft = d_soft; a.rmin = rmin; a.share = share; a.emit_all = 0; a.solid_in = solid_in;
  a.out_row = nullptr; a.keep = nullptr; a.body = slab; a.row_bytes = row_bytes; a.key_bytes = 0; a.fmt = 2;
  a.row_keep = nullptr; a.stats = stats; a.N = N;
  merge_emit_kernel<RowDense><<<dim3(grid_x_for(max_n), N), MG_THREADS, 0, st>>>(d_lists, rr, a);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_sparse_solid(const MergeList* d_lists, u32 N, const u32* d_soft, const u64* ulo, const u64* uhi,
                                u64 nu, int W, u32* solid_in, u64 max_n, cudaStream_t st, u64* launches)
{
  if (!max_n) return cudaSuccess;
  RowSparse rr; rr.ulo = ulo; rr.uhi = uhi; rr.nu = nu; rr.W = W;
  merge_solid_kernel<RowSparse><<<dim3(grid_x_for(max_n), N), MG_THREADS, 0, st>>>(d_lists, d_soft, rr, solid_in);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_row_keep(const u32* solid_in, u64 nrows, u32 rmin, u32 emit_all, u32* keep_flag, cudaStream_t st, u64* launches)
{
  if (!nrows) return cudaSuccess;
  row_keep_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(solid_in, nrows, rmin, emit_all, keep_flag);
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_sparse_emit(const MergeList* d_lists, u32 N, const u32* d_soft, u32 rmin, u32 share, u32 emit_all,
                               const u64* ulo, const u64* uhi, u64 nu, int W, const u32* solid_in,
                               const u32* keep, const u32* out_row, int fmt, uint8_t* body, u32 row_bytes,
                               uint8_t* row_keep, u64* stats, u64 max_n, cudaStream_t st, u64* launches)
{
  if (!nu) return cudaSuccess;
  RowSparse rr; rr.ulo = ulo; rr.uhi = uhi; rr.nu = nu; rr.W = W;
  sparse_keys_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(ulo, uhi, nu, W, keep, out_row, solid_in, rmin,
                                                                   body, row_bytes, emit_all ? row_keep : nullptr);
  EmitArgs a;
  a.soft = d_soft; a.rmin = rmin; a.share = share; a.emit_all = emit_all; a.solid_in = solid_in;
  a.out_row = out_row; a.keep = keep; a.body = body; a.row_bytes = row_bytes; a.key_bytes = 8 * W; a.fmt = fmt;
  a.row_keep = row_keep; a.stats = stats; a.N = N;
  merge_emit_kernel<RowSparse><<<dim3(grid_x_for(max_n), N), MG_THREADS, 0, st>>>(d_lists, rr, a);
  *launches += 2;
  return cudaGetLastError();
}

}  // namespace kmx
