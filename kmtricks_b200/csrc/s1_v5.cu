// s1_v5.cu -- stage 1, position-parallel kernel (phase bodies and their description: s1_v5.cuh).
// Used for launches whose longest read fits the per-CTA shared arrays (short reads); longer
// sequences keep the streaming kernel of s1_superk.cu.
#include "common.cuh"
#include "kmx_internal.h"
#include "s1_v5.cuh"

namespace kmx {

using namespace s1v5;

static constexpr int V5_THREADS = 256;
static constexpr int V5_WARPS = V5_THREADS / 32;

template <int W>
__global__ void __launch_bounds__(V5_THREADS)
s1_superk_v5(const S1Args a, const Geo geo)
{
  extern __shared__ __align__(16) u32 smem5[];
  __shared__ u32 s_end;
  Cta x;
  x.k = a.k; x.m = a.m; x.w = a.wlen; x.max_nk = (u32)a.max_nk;
  x.mmask = (u32)((1ull << (2 * a.m)) - 1ull);
  x.ban_mask = 0x55555555u & ((1u << (2 * (a.m - 2))) - 1u);
  x.g = geo;
  carve(x, smem5, a.P);
  const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const u32 R = geo.R;
  const u64 seg0 = (u64)blockIdx.x * R;

  for (u32 p = tid; p < a.P; p += V5_THREADS) { x.hist[p] = 0; x.kc[p] = 0; }
  for (u32 r = tid; r < R; r += V5_THREADS) {
    const u64 seg = seg0 + r;
    u32 len = 0, st = 0;
    if (seg < a.nseg) { len = a.seg_len[seg]; st = a.seg_start[seg]; }
    if (len < (u32)a.k) len = 0;                  // Sequence2SuperKmer.hpp:143-144
    x.len[r] = len; x.start[r] = st; x.inval[r] = 0;
  }
  __syncthreads();

  // ---- P0: pack
  {
    const u32* wend = reinterpret_cast<const u32*>((reinterpret_cast<uintptr_t>(a.text) + a.text_bytes + 3) & ~(uintptr_t)3);
    const u32 ntask = R * geo.nch;
    u32 r = tid / geo.nch, c = tid - r * geo.nch;           // (read, chunk) of this thread's item, advanced incrementally
    const u32 dr = V5_THREADS / geo.nch, dc = V5_THREADS - dr * geo.nch;
    for (u32 task = tid; task < ntask; task += V5_THREADS) {
      p0_pack(x, r, c, a.text + x.start[r], x.len[r], wend);
      r += dr; c += dc;
      if (c >= geo.nch) { c -= geo.nch; r++; }
    }
  }
  __syncthreads();

  // ---- P1: lut values, one warp per read, lanes over the m-mers
  for (u32 r = wid; r < R; r += V5_WARPS) {
    const u32 len = x.len[r];
    if (!len) continue;
    p1_row(x, r, lane, len - (u32)a.m + 1u);
  }
  __syncthreads();

  // ---- P2: sliding minimum, lane = read (odd row pitch: conflict-free), warp = block
  for (u32 g = wid; g < geo.nblk; g += V5_WARPS)
    for (u32 r = lane; r < R; r += 32) p2_block(x, r, g, x.len[r]);
  __syncthreads();

  // ---- P3a per read, P3b count per item (read, block)
  const u32 ntask = R * geo.nblk;
  for (u32 r = tid; r < R; r += V5_THREADS) p3a_read(x, r, x.len[r]);
  __syncthreads();
  for (u32 t = tid; t < ntask; t += V5_THREADS) { const u32 r = item_read(geo, t); x.pfx[t + 1] = p3_count(x, r, t - r * geo.nblk, x.len[r]); }
  __syncthreads();
  if (wid == 0) {                                   // counts at pfx[1..ntask] -> inclusive prefix in place (pfx[t] = events before item t)
    const u32 per = (ntask + 31) / 32;
    const u32 i0 = min(ntask, lane * per), i1 = min(ntask, i0 + per);
    u32 sum = 0;
    for (u32 i = i0; i < i1; i++) sum += x.pfx[i + 1];
    u32 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += y; }
    u32 run = inc - sum;
    for (u32 i = i0; i < i1; i++) { run += x.pfx[i + 1]; x.pfx[i + 1] = run; }
    if (lane == 0) x.pfx[0] = 0;
  }
  __syncthreads();

  u32 first = 0;
  while (first < ntask) {
    // items [first, end) are flushed in this round: the longest prefix whose events fit the queue
    const u32 base = x.pfx[first];
    u32 end = ntask;
    if (x.pfx[ntask] - base > geo.evcap) {          // CTA-uniform; the usual case is one round, no search
      for (u32 t = first + tid; t < ntask; t += V5_THREADS) {
        const bool fits = x.pfx[t + 1] - base <= geo.evcap;
        if (fits && x.pfx[t + 2] - base > geo.evcap) s_end = t + 1;      // exactly one writer (prefix sums are monotone; t + 1 < ntask here)
      }
      if (tid == 0 && x.pfx[first + 1] - base > geo.evcap) s_end = first; // no item fits: cannot happen (evcap >= events of any item)
      __syncthreads();
      end = s_end;
      if (end <= first) { if (tid == 0) *a.overflow = 1u; return; }       // fail loudly through the host's retry limit
    }
    const u32 nev = x.pfx[end] - base;
    // ---- P3c: the events of one item per thread, then (P4 pass 1) their partition, per-partition rank and k-mer totals
    for (u32 t = first + tid; t < end; t += V5_THREADS) {
      const u32 s0 = x.pfx[t] - base, n = x.pfx[t + 1] - x.pfx[t];
      if (!n) continue;
      const u32 r = item_read(geo, t);
      if (!x.inval[r]) p3_emit_item(x, r, t - r * geo.nblk, s0);
      else p3_slow<true>(x, r, x.len[r], s0);       // reads with invalid bases: per-k-mer walk (all their events sit in block 0's item)
      u32 q = s0;
      for (; q + 2 <= s0 + n; q += 2) {
        const Ev e0 = x.ev[q], e1 = x.ev[q + 1];
        const u32 p0 = __ldg(a.repart + e0.y), p1 = __ldg(a.repart + e1.y);     // Repartitor, PartiInfo.hpp:381
        x.ev[q].y = p0 | (atomicAdd(&x.hist[p0], 1u) << 16);
        x.ev[q + 1].y = p1 | (atomicAdd(&x.hist[p1], 1u) << 16);
        atomicAdd(&x.kc[p0], (e0.x >> 19) & 127u);
        atomicAdd(&x.kc[p1], (e1.x >> 19) & 127u);
      }
      if (q < s0 + n) {
        const Ev e0 = x.ev[q];
        const u32 p0 = __ldg(a.repart + e0.y);
        x.ev[q].y = p0 | (atomicAdd(&x.hist[p0], 1u) << 16);
        atomicAdd(&x.kc[p0], (e0.x >> 19) & 127u);
      }
    }
    __syncthreads();
    for (u32 p = tid; p < a.P; p += V5_THREADS) {
      const u32 cnt = x.hist[p];
      if (cnt) {
        x.gbase[p] = atomicAdd(&a.cursor[p], cnt);
        atomicAdd(&a.kcnt[p], (u64)x.kc[p]);
        x.hist[p] = 0; x.kc[p] = 0;
      }
    }
    __syncthreads();
    // ---- P4 pass 2: build and store the records
    uint4* out = reinterpret_cast<uint4*>(a.records);
    for (u32 q = tid; q < nev; q += V5_THREADS) {
      const Ev e = x.ev[q];
      const u32 p = e.y & 0xFFFFu;
      const u32 rd = e.x & 127u, iend = (e.x >> 7) & 4095u, nkr = (e.x >> 19) & 127u;
      const u32 nb = (u32)a.k + nkr - 1u;
      u32 v[4 * W];
      build_record<4 * W>(x.BE + rd * geo.LW, geo.nch, iend, nb, v);
      const u32 pos = x.gbase[p] + (e.y >> 16);
      if (pos < a.bcap[p]) {
        uint4* dst = out + (size_t)W * (a.boff[p] + pos);
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        if (W == 2) dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
      } else *a.overflow = 1u;
    }
    __syncthreads();
    first = end;
  }
}

// can this launch take the position-parallel kernel?  (events of one read must fit a flush round,
// the event word holds 7 bits of read index and 12 bits of base index, the arrays must leave room
// for >= 2 CTAs per SM)
bool s1_v5_usable(u32 max_len, int k, int m, u32 P, Geo* geo, size_t* smem)
{
  const bool off = kmx_env_flag("KMX_S1V5_OFF");
  if (off || max_len < (u32)k || max_len > 4000u) return false;
  u32 R = 32;
  if (const char* e = getenv("KMX_S1V5_R")) { int v = atoi(e); if (v >= 1 && v <= 128) R = (u32)v; }
  Geo g = make_geo(R, max_len, k, m);
  const size_t b = smem_bytes(g, P);
  const int max_nk = (k <= 32 ? KMX_REC1_MAXN : KMX_REC2_MAXN) - k + 1;
  if (b > 100 * 1024 || max_len - (u32)k + 1u > g.evcap || k - m + 1 > max_nk) return false;   // p3_count: a block's inner runs are single records
  *geo = g; *smem = b;
  return true;
}

cudaError_t launch_s1_v5(int W, const S1Args& a, const Geo& geo, size_t smem, cudaStream_t st, u64* launches)
{
  if (a.nseg == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((a.nseg + geo.R - 1) / geo.R);
  cudaError_t e;
  if (W == 1) {
    e = cudaFuncSetAttribute(s1_superk_v5<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    s1_superk_v5<1><<<grid, V5_THREADS, smem, st>>>(a, geo);
  } else {
    e = cudaFuncSetAttribute(s1_superk_v5<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    s1_superk_v5<2><<<grid, V5_THREADS, smem, st>>>(a, geo);
  }
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace kmx
