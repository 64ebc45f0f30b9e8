// s1_v5.cu -- stage 1, position-parallel kernel (phase bodies and their description: s1_v5.cuh).
// Used for launches whose longest read fits the per-CTA shared arrays (short reads); longer
// sequences keep the streaming kernel of s1_superk.cu.
#include "common.cuh"
#include "kmx_internal.h"
#include "s1_v5.cuh"
#include <cstring>
#include <cstdlib>

namespace kmx {

using namespace s1v5;

static constexpr int V5_THREADS = 320;       // 10 warps: one round of the pack phase for 32 reads of <= 160 bases
static constexpr int V5_WARPS = V5_THREADS / 32;

// byte position (in the coordinates of the mask array) of newline number 4 * S1_FUSED_R * c, for every CTA c of the
// self-indexing launch.  One warp per 16 KiB tile: 8 x 32 mask words, ranked by popcount + warp scan.
__global__ void __launch_bounds__(256)
fq_cta_pos(const u64* __restrict__ nlmask64, const u64* __restrict__ tile_prefix, u64 ntiles, u32* __restrict__ cta_pos, u64 ncta)
{
  const u32 lane = threadIdx.x & 31u;
  const u64 t = (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= ntiles) return;
  const u64 per = 4ull * S1_FUSED_R;
  u64 g = tile_prefix[t];
  u64 m[8];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = nlmask64[t * 256 + (u64)i * 32 + lane];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 cnt = (u32)__popcll(m[i]);
    u32 inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += y; }
    const u64 gg = g + inc - cnt;                         // number of this word's first newline
    const u64 c = (gg + per - 1) / per;                   // a word holds <= 64 newlines: at most one multiple of 128
    if (cnt && c * per < gg + cnt && c < ncta) {
      u64 mm = m[i];
      for (u32 j = (u32)(c * per - gg); j; j--) mm &= mm - 1;
      cta_pos[c] = (u32)((t * 256 + (u64)i * 32 + lane) * 64 + (u32)(__ffsll((long long)mm) - 1));
    }
    g += __shfl_sync(0xffffffffu, inc, 31);
  }
}

template <int W, bool FUSED>
__global__ void __launch_bounds__(V5_THREADS, 5)
s1_superk_v5(const S1Args a, const Geo geo, const S1Idx ix)
{
  extern __shared__ __align__(16) u32 smem5[];
  __shared__ u32 s_nev, s_pend;
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
  for (u32 t = tid; t < geo.R * geo.nblk; t += V5_THREADS) x.done[t] = 0;
  if (tid == 0) { s_nev = 0; s_pend = 0; }
  u32 nr = 0, bad = 0, vplus = '+', vat = '@';       // self-indexing launch: format checks, consumed after P0 (their loads overlap it)
  if (!FUSED) {
    for (u32 r = tid; r < R; r += V5_THREADS) {
      const u64 seg = seg0 + r;
      u32 len = 0, st = 0;
      if (seg < a.nseg) { len = a.seg_len[seg]; st = a.seg_start[seg]; }
      if (len < (u32)a.k) len = 0;                  // Sequence2SuperKmer.hpp:143-144
      x.len[r] = len; x.start[r] = st; x.inval[r] = 0;
    }
  } else {
    // The CTA's reads are records [32 c, 32 c + 32) of the strict 4-line FASTQ text: newline number 4 i ends the header of
    // record i, 4 i + 1 its sequence, 4 i + 3 the record.  Warp 0 ranks the newline masks of the count pass from the
    // position the host-launched table gives for the CTA's first newline (256 mask words = 16 KiB of text per round).
    u32* s_nl = x.U;                                   // [4 * S1_FUSED_R]; U is not written before P1
    nr = (u32)min((u64)S1_FUSED_R, a.nseg - seg0);
    const u32 need = 4 * nr;
    for (u32 i = tid; i < 4 * S1_FUSED_R; i += V5_THREADS) s_nl[i] = 0xFFFFFFFFu;      // "no such newline" (end of text)
    __syncthreads();
    {
      // every thread ranks one mask word of a 256-word (16 KiB) window per round: popcount, warp scan, warp totals through
      // shared memory; a window holds the 4 x 32 newlines of a CTA unless records are longer than 512 bytes
      u32* s_wsum = x.U + 4 * S1_FUSED_R;              // [2][V5_WARPS] (double-buffered: one barrier per round)
      const u64 nwords = ix.ntiles * (FQ_TILE_BYTES / 64);
      const u32 pos0 = ix.cta_pos[blockIdx.x];
      u64 w = pos0 >> 6;
      u32 found = 0, buf = 0;
      u64 keep = ~0ull << (pos0 & 63u);                // first word: newlines before the CTA's first one belong to the previous CTA
      while (found < need && w < nwords) {             // CTA-uniform
        const u64 idx = w + tid;
        u64 mm = idx < nwords ? ix.nlmask64[idx] : 0ull;
        if (tid == 0) mm &= keep;
        keep = ~0ull;
        const u32 cnt = (u32)__popcll(mm);
        u32 inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += y; }
        if (lane == 31) s_wsum[buf * V5_WARPS + wid] = inc;
        __syncthreads();
        u32 before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < V5_WARPS; i++) { const u32 v = s_wsum[buf * V5_WARPS + i]; if ((u32)i < wid) before += v; total += v; }
        u32 rank = found + before + inc - cnt;
        const u32 wbase = (u32)(idx * 64 - ix.lead);
        while (mm && rank < need) {
          s_nl[rank++] = wbase + (u32)(__ffsll((long long)mm) - 1);
          mm &= mm - 1;
        }
        found += total;
        w += V5_THREADS; buf ^= 1u;
      }
    }
    __syncthreads();
    for (u32 r = tid; r < R; r += V5_THREADS) {
      u32 len = 0, st = 0;
      if (r < nr) {
        const u64 nbytes = ix.tot - ix.lead;
        const u32 h = s_nl[4 * r], e = s_nl[4 * r + 1], q = s_nl[4 * r + 3];
        if (h == 0xFFFFFFFFu || e == 0xFFFFFFFFu) bad = 1;
        else {
          st = h + 1; len = e - st;
          // A trailing CR (dropped by kseq when the line is longer than 1, BankFasta.cpp:476-477) is kept as an invalid last
          // base: the k-mers that would cover it do not exist either way.  Only a read one longer than the geometry looks.
          if (len == ix.geo_maxlen + 1 && a.text[e - 1] == '\r') len--;
          vplus = ((u64)e + 1 < nbytes) ? a.text[e + 1] : (u32)'+';
          vat = (q != 0xFFFFFFFFu && (u64)q + 1 < nbytes) ? a.text[q + 1] : (u32)'@';
        }
        if (seg0 + r == 0 && a.text[0] != '@') bad = 1;
        if (len > ix.geo_maxlen) { atomicOr(&ix.flags[3], 1u); len = 0; }
        if (bad) len = 0;
      }
      if (len < (u32)a.k) len = 0;
      x.len[r] = len; x.start[r] = st; x.inval[r] = 0;
    }
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
  if (FUSED && tid < nr && (bad || vplus != '+' || vat != '@')) atomicOr(&ix.flags[0], 1u);   // not strict 4-line FASTQ: the host redoes the block
  __syncthreads();

  // ---- P1: lut values, one warp per read, lanes over the m-mers
  for (u32 r = wid; r < R; r += V5_WARPS) {
    const u32 len = x.len[r];
    if (!len) continue;
    p1_row(x, r, lane, len - (u32)a.m + 1u);
  }
  __syncthreads();

  // ---- P2: sliding minimum, lane = read (odd row pitch: conflict-free), warp = block
  if (a.wlen == 22) {                                // the default k = 31, m = 10: straight-line code
    for (u32 g = wid; g < geo.nblk; g += V5_WARPS)
      for (u32 r = lane; r < R; r += 32) p2_block<22>(x, r, g, x.len[r]);
  } else {
    for (u32 g = wid; g < geo.nblk; g += V5_WARPS)
      for (u32 r = lane; r < R; r += 32) p2_block<0>(x, r, g, x.len[r]);
  }
  __syncthreads();

  // ---- P3 + P4, in rounds (one round unless the CTA logs more events than its queue holds): every pending item
  // (read, block) completes its change mask, counts its records, takes that many queue slots with ONE shared-memory
  // atomic, logs its events and (P4 pass 1) looks up their partition, per-partition rank and k-mer totals
  const u32 ntask = R * geo.nblk;
  for (;;) {
    for (u32 t = tid; t < ntask; t += V5_THREADS) {
      if (x.done[t]) continue;
      const u32 r = item_read(geo, t), g = t - r * geo.nblk;
      const u32 n = p3_prepare(x, r, g, x.len[r]);
      if (!n) { x.done[t] = 1; continue; }
      const u32 s0 = atomicAdd(&s_nev, n);
      if (s0 + n > geo.evcap) {                      // does not fit this round: pad what was taken, try again after the flush
        for (u32 q = s0; q < geo.evcap; q++) { Ev z; z.x = 0; z.y = 0; x.ev[q] = z; }
        s_pend = 1;
        continue;
      }
      x.done[t] = 1;
      if (!x.inval[r]) p3_emit_item(x, r, g, s0);
      else p3_slow<true>(x, r, x.len[r], s0);       // reads with invalid bases: per-k-mer walk (all their events sit in block 0's item)
      u32 q = s0;
      for (; q + 2 <= s0 + n; q += 2) {
        const Ev e0 = x.ev[q], e1 = x.ev[q + 1];
        const u32 p0 = __ldg(a.repart + e0.y), p1 = __ldg(a.repart + e1.y);     // Repartitor, PartiInfo.hpp:381
        if (a.mload) {                                  // repartition estimate: k-mers per minimizer
          atomicAdd(a.mload + e0.y, (u64)((e0.x >> 19) & 127u)); atomicAdd(a.mload + e1.y, (u64)((e1.x >> 19) & 127u));
        }
        x.ev[q].y = p0 | (atomicAdd(&x.hist[p0], 1u) << 16);
        x.ev[q + 1].y = p1 | (atomicAdd(&x.hist[p1], 1u) << 16);
        atomicAdd(&x.kc[p0], (e0.x >> 19) & 127u);
        atomicAdd(&x.kc[p1], (e1.x >> 19) & 127u);
      }
      if (q < s0 + n) {
        const Ev e0 = x.ev[q];
        const u32 p0 = __ldg(a.repart + e0.y);
        if (a.mload) atomicAdd(a.mload + e0.y, (u64)((e0.x >> 19) & 127u));
        x.ev[q].y = p0 | (atomicAdd(&x.hist[p0], 1u) << 16);
        atomicAdd(&x.kc[p0], (e0.x >> 19) & 127u);
      }
    }
    __syncthreads();
    const u32 nev = min(s_nev, geo.evcap), pend = s_pend;
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
      if (!nkr) continue;                             // padding of an item that did not fit
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
    if (!pend) break;
    if (tid == 0) { s_nev = 0; s_pend = 0; }
    __syncthreads();
  }
}

// can this launch take the position-parallel kernel?  (the events of one read must fit the event queue,
// the event word holds 7 bits of read index and 12 bits of base index, the arrays must leave room
// for >= 2 CTAs per SM, and a block of w k-mers must fit one record: p3_prepare's closed-form count)
bool s1_v5_usable(u32 max_len, int k, int m, u32 P, Geo* geo, size_t* smem)
{
  const bool off = kmx_env_flag("KMX_S1V5_OFF");
  if (off || max_len < (u32)k || max_len > 4000u) return false;
  u32 R = 32;
  if (const char* e = getenv("KMX_S1V5_R")) { int v = atoi(e); if (v >= 1 && v <= 128) R = (u32)v; }
  Geo g = make_geo(R, max_len, k, m);
  const size_t b = smem_bytes(g, P);
  const int max_nk = (k <= 32 ? KMX_REC1_MAXN : KMX_REC2_MAXN) - k + 1;
  if (b > 100 * 1024 || max_len - (u32)k + 1u > g.evcap || k - m + 1 > max_nk) return false;
  *geo = g; *smem = b;
  return true;
}

cudaError_t launch_s1_v5(int W, const S1Args& a, const Geo& geo, size_t smem, const S1Idx* idx, cudaStream_t st, u64* launches)
{
  if (a.nseg == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((a.nseg + geo.R - 1) / geo.R);
  S1Idx ix; memset(&ix, 0, sizeof ix);
  if (idx) ix = *idx;
  cudaError_t e;
#define KMX_V5_LAUNCH(WW, FF)                                                                                              \
  do {                                                                                                                     \
    e = cudaFuncSetAttribute(s1_superk_v5<WW, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
    if (e != cudaSuccess) return e;                                                                                        \
    s1_superk_v5<WW, FF><<<grid, V5_THREADS, smem, st>>>(a, geo, ix);                                                      \
  } while (0)
  if (idx) { if (W == 1) KMX_V5_LAUNCH(1, true); else KMX_V5_LAUNCH(2, true); }
  else { if (W == 1) KMX_V5_LAUNCH(1, false); else KMX_V5_LAUNCH(2, false); }
#undef KMX_V5_LAUNCH
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_fq_cta_pos(const u64* nlmask64, const u64* tile_prefix, u64 ntiles, u32* cta_pos, u64 ncta, cudaStream_t st, u64* launches)
{
  if (!ntiles) return cudaSuccess;
  fq_cta_pos<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(nlmask64, tile_prefix, ntiles, cta_pos, ncta);
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace kmx
