// s1_superk.cu -- stage 1: FASTQ text -> 2-bit encode -> minimizer -> repartition scatter.
//
// Replaces (behaviour, not code) gatb Model::iterate / ModelMinimizer (Model.hpp:725-765,
// 1040-1139,1220-1287), Sequence2SuperKmer (Sequence2SuperKmer.hpp:90-158) and
// KmFillPartitions::processSuperkmer (include/kmtricks/gatb/fill_partitions.hpp:59-105).
//
// Kernels:
//   fq_count_newlines / fq_index_lines : strict 4-line FASTQ line index (sequence start/len
//       per record) built on the device with 128-bit loads, ballot-free popcount of '\n'.
//   s1_superk<W> : one thread per sequence segment streams its bases, keeps the rolling
//       forward / reverse-complement m-mer, computes lut(m-mer) arithmetically (canonical
//       m-mer + "AA" ban) instead of gathering an 8 MiB table, and maintains the sliding
//       minimum over the k-m+1 m-mers with a divergence-free block prefix/suffix-min ring
//       in shared memory.  A super-k-mer record is cut when the minimizer changes, the
//       k-mer is invalid, or the record is full.  Records are staged per CTA in shared
//       memory and flushed with ONE global atomic per (CTA flush, partition) into the
//       per-partition bucket slabs in HBM (128-bit stores).
#include "common.cuh"
#include "kmx_internal.h"

namespace kmx {

// ------------------------------------------------------------------------------------
// FASTQ line index
// ------------------------------------------------------------------------------------
static constexpr int FQ_THREADS = 256;
static constexpr int FQ_BYTES_PER_THREAD = 64;                 // 4 x uint4
static constexpr int FQ_TILE = FQ_THREADS * FQ_BYTES_PER_THREAD;

// 16-bit mask of the '\n' bytes of one 16-byte piece (bit b = byte b)
__device__ __forceinline__ u32 nl_mask16(const uint4 v)
{
  const u32 wv[4] = {v.x, v.y, v.z, v.w};
  u32 m16 = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const u32 x = wv[q] ^ 0x0A0A0A0Au;                           // zero byte <=> '\n'
    const u32 y = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;        // high bit set iff byte != 0 (exact, no borrow)
    const u32 m = (~y & 0x80808080u) >> 7;                       // bit 8b set <=> byte b is '\n'
    m16 |= (((m * 0x00204081u) >> 21) & 0xFu) << (4 * q);        // gather bits 0,8,16,24 -> 0..3
  }
  return m16;
}

// Pass 1 over the text: newline count per 16 KiB tile AND the newline bitmask itself (one u16 per
// 16-byte piece, bytes outside [lead, nbytes_total) masked off), so that the index pass below reads
// 1/8 of the bytes and does no SWAR work.  `base` is the 16B-aligned address at or below the text
// start, `lead` = text - base.
__global__ void __launch_bounds__(FQ_THREADS)
fq_count_newlines(const uint4* __restrict__ base, u64 lead, u64 nbytes_total /* lead + n */,
                  u32* __restrict__ tile_counts, uint16_t* __restrict__ nlmask)
{
  __shared__ u32 s_sum[FQ_THREADS / 32];
  u64 tile0 = (u64)blockIdx.x * FQ_TILE;
  u32 cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    u64 off = tile0 + ((u64)j * FQ_THREADS + threadIdx.x) * 16;
    u32 m16 = 0;
    if (off < nbytes_total) {
      m16 = nl_mask16(__ldg(base + off / 16));
      if (off < lead) { const u32 d = (u32)min((u64)16, lead - off); m16 = d >= 16 ? 0u : (m16 >> d) << d; }
      if (off + 16 > nbytes_total) m16 &= 0xFFFFu >> (16 - (u32)(nbytes_total - off));
    }
    nlmask[off / 16] = (uint16_t)m16;                            // the mask array covers whole tiles
    cnt += __popc(m16);
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int i = 0; i < FQ_THREADS / 32; i++) t += s_sum[i];
    tile_counts[blockIdx.x] = t;
  }
}

// single-CTA exclusive scan of tile counts (u32 -> u64 prefix); also total
__global__ void __launch_bounds__(1024) scan_u32_to_u64(const u32* __restrict__ in, u64* __restrict__ out, u64 n, u64* total)
{
  // every thread owns a contiguous range: one round of independent loads, one block scan, one round of stores
  __shared__ u64 s_warp[32];
  const u64 per = (n + 1023) / 1024;
  const u64 i0 = min(n, (u64)threadIdx.x * per), i1 = min(n, i0 + per);
  u64 sum = 0;
  for (u64 i = i0; i < i1; i += 8) {
    u32 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = i + q < i1 ? in[i + q] : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) sum += v[q];
  }
  u64 x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    const u64 w = s_warp[threadIdx.x]; u64 xw = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, xw, o); if (threadIdx.x >= (u32)o) xw += y; }
    s_warp[threadIdx.x] = xw - w;
  }
  __syncthreads();
  u64 run = s_warp[threadIdx.x >> 5] + x - sum;
  for (u64 i = i0; i < i1; i += 8) {
    u32 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = i + q < i1 ? in[i + q] : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) if (i + q < i1) { out[i + q] = run; run += v[q]; }
  }
  if (threadIdx.x == 1023 && total) *total = run;
}

// For newline number g (0-based) at text position pos:
//   g%4==0 -> the sequence of record g/4 starts at pos+1 and runs to the next newline
//   g%4==1 -> next char must be '+' ; g%4==3 -> next char must be '@' (or end)
// Each thread owns 64 contiguous bytes (blocked, so newline order == thread order) and folds its
// newline bytes into ONE 64-bit mask, so the per-newline loop runs max-newlines-per-lane times
// (2-3 for short reads) instead of once per 32-bit word.  The end of a sequence line is found
// in the CTA's shared mask array (the next set bit after the start), falling back to a byte scan
// only when the line crosses the 16 KiB tile; length, '\r' strip (kseq drops one trailing '\r'
// when the line is longer than 1, BankFasta.cpp:476-477) and the maximum length are done here.
__global__ void __launch_bounds__(FQ_THREADS)
fq_index_lines(const uint4* __restrict__ base, u64 lead, u64 nbytes_total,
               const u64* __restrict__ tile_prefix, const u64* __restrict__ nlmask64, u32* __restrict__ seq_start,
               u32* __restrict__ seq_len, u64 nrec, u32* __restrict__ flags /* [0]=format error, [1]=max len */)
{
  __shared__ u32 s_warp[FQ_THREADS / 32];
  __shared__ u64 s_mask[FQ_THREADS];
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(base);
  const u64 tile0 = (u64)blockIdx.x * FQ_TILE;
  const u64 off = tile0 + (u64)threadIdx.x * FQ_BYTES_PER_THREAD;
  const u64 mask = nlmask64[off / 64];                             // written by fq_count_newlines (4 x u16, whole tiles)
  const u32 cnt = (u32)__popcll(mask);
  u32 x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  s_mask[threadIdx.x] = mask;
  __syncthreads();
  u32 wbase = 0;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) wbase += s_warp[i];
  u64 g = tile_prefix[blockIdx.x] + wbase + x - cnt;
  u32 maxlen = 0, bad = 0;
  u64 msk = mask;
  // Only two of the four newlines of a record need work: #4r (the sequence starts behind it; its end is the
  // next newline, behind which a '+' must follow) and #4r+3 (an '@' must follow).  The bytes to look at are
  // independent of each other, so up to FQ_BATCH newlines are located first and their loads issued together.
  constexpr int FQ_BATCH = 3;
  while (msk) {
    u64 a1[FQ_BATCH], a2[FQ_BATCH], e_rec[FQ_BATCH]; u32 e_st[FQ_BATCH], e_len[FQ_BATCH], e_kind[FQ_BATCH];
#pragma unroll
    for (int e = 0; e < FQ_BATCH; e++) {          // static slot index: the batch stays in registers
      e_kind[e] = 9; a1[e] = 0; a2[e] = 0; e_rec[e] = 0; e_st[e] = 0; e_len[e] = 0;
      bool found = false; u64 pos = 0, rec = 0; u32 ph = 0;
      while (msk && !found) {
        const int bit = __ffsll((long long)msk) - 1; msk &= msk - 1;
        pos = off + bit; rec = g >> 2; ph = (u32)(g & 3);
        g++;
        found = rec < nrec && (ph == 0 || ph == 3);
      }
      if (!found) continue;
      if (ph == 0) {
        // end of the sequence line = next newline: own mask, then the following threads' masks
        u64 end;
        if (msk) end = off + (__ffsll((long long)msk) - 1);
        else {
          u32 t = threadIdx.x + 1;
          while (t < FQ_THREADS && s_mask[t] == 0) t++;
          if (t < FQ_THREADS) end = tile0 + (u64)t * FQ_BYTES_PER_THREAD + (__ffsll((long long)s_mask[t]) - 1);
          else {
            end = tile0 + FQ_TILE;
            while (end < nbytes_total && bytes[end] != '\n') end++;
          }
        }
        e_kind[e] = 0; e_rec[e] = rec; e_st[e] = (u32)(pos + 1 - lead); e_len[e] = (u32)(end - (pos + 1));
        a1[e] = end - 1;                        // a trailing '\r' is dropped when the line is longer than 1
        a2[e] = end + 1;                        // '+' line
      } else {
        e_kind[e] = 3; a1[e] = pos + 1;         // '@' of the next record (or end of text)
      }
    }
    u32 b1[FQ_BATCH], b2[FQ_BATCH];
#pragma unroll
    for (int e = 0; e < FQ_BATCH; e++) {
      b1[e] = 0; b2[e] = 0;
      if (e_kind[e] == 3) { b1[e] = a1[e] < nbytes_total ? bytes[a1[e]] : (u32)'@'; }
      else if (e_kind[e] == 0) {
        b1[e] = e_len[e] > 1 ? bytes[a1[e]] : 0u;
        b2[e] = a2[e] < nbytes_total ? bytes[a2[e]] : (u32)'+';
      }
    }
#pragma unroll
    for (int e = 0; e < FQ_BATCH; e++) {
      if (e_kind[e] == 3) bad |= (u32)(b1[e] != '@');
      else if (e_kind[e] == 0) {
        const u32 len = e_len[e] - (u32)(b1[e] == '\r');
        seq_start[e_rec[e]] = e_st[e]; seq_len[e_rec[e]] = len;
        maxlen = max(maxlen, len);
        bad |= (u32)(b2[e] != '+');
      }
    }
  }
  if (bad) atomicOr(&flags[0], 1u);
  if (blockIdx.x == 0 && threadIdx.x == 0 && bytes[lead] != '@') atomicOr(&flags[0], 1u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  if ((threadIdx.x & 31) == 0 && maxlen) atomicMax(&flags[1], maxlen);
}

// ------------------------------------------------------------------------------------
// main stage-1 kernel
// ------------------------------------------------------------------------------------
// One thread per sequence segment streams its bases 8 at a time:
//   * 2-bit code + validity from the character (code = (c>>1)&3; valid <=> "ACTG"[code] == c&0xDF)
//   * rolling forward / reverse-complement m-mer, lut value computed arithmetically
//   * sliding minimum over the k-m+1 m-mers: block prefix / suffix-min ring in shared memory
//     (all threads of the CTA are at the same base index, so the ring index is uniform)
//   * the bases are also packed 16 per word (big-endian) into shared memory
//   * a super-k-mer CUT only logs an 6-byte event (thread, start, #k-mers, partition) -- the
//     divergent path is ~8 instructions
// Every few rounds the CTA flushes cooperatively: per-partition counting in shared memory, ONE
// global atomic per (flush, partition), then every event is turned into a 16/32-byte record by
// funnel-shifting the packed bases (dense loop, no divergence) and stored to the bucket slab.
template <int W> struct RecT;
template <> struct RecT<1> { typedef uint4 type; };
template <> struct RecT<2> { struct __align__(16) type { uint4 a, b; }; };

static constexpr int S1_THREADS = 128;
static_assert(S1_THREADS * 4 == 512, "lds_next_row hard-codes the ring row pitch");

// bits [32*jw, 32*jw+32) of (X >> s), X = big-endian words S(0..nwords) of thread t's packed read
__device__ __forceinline__ u32 pack_word(const u32* __restrict__ s_pack, u32 t, int nwords, int top_word, int s, int jw)
{
  // X is the (top_word+1)-word big-endian number ending at word `top_word`; LSB word index 0 == top_word
  int q = s + 32 * jw;
  int wi = q >> 5, sh = q & 31;
  int a = top_word - wi, b = a - 1;               // b is the more significant neighbour
  u32 xa = (a >= 0 && a < nwords) ? s_pack[a * S1_THREADS + t] : 0u;
  u32 xb = (b >= 0 && b < nwords) ? s_pack[b * S1_THREADS + t] : 0u;
  return __funnelshift_r(xa, xb, sh);
}

// suffix-min rebuild of one thread's ring column when a block of wlen m-mers is complete.
// The loads of a batch of 8 rows are independent of the running minimum, so they are issued
// together and only the min/store chain is serial.
__device__ __noinline__ void ring_suffix_min(u32* __restrict__ col, int wlen)
{
  u32 accm = 0xFFFFFFFFu;
  int t2 = wlen - 1;
  for (; t2 >= 7; t2 -= 8) {
    u32 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = col[(t2 - q) * S1_THREADS];
#pragma unroll
    for (int q = 0; q < 8; q++) { accm = min(accm, v[q]); col[(t2 - q) * S1_THREADS] = accm; }
  }
  for (; t2 >= 0; t2--) {
    accm = min(accm, col[t2 * S1_THREADS]);
    col[t2 * S1_THREADS] = accm;
  }
}

static constexpr u32 S1_EVW = 640;        // cut events per warp queue
static constexpr u32 S1_EVTHR = S1_EVW - 32 * 17;  // a warp logs <= 32 x 16 (+32 terminators) events between two checks

// shared-memory accesses of the hot loop by 32-bit shared address (keeps the generic->shared window
// arithmetic out of the per-base instruction stream)
__device__ __forceinline__ u32 lds_next_row(u32 sa)            // [sa + one ring row (S1_THREADS words)]
{
  u32 v; asm volatile("ld.shared.u32 %0, [%1+512];" : "=r"(v) : "r"(sa) : "memory"); return v;
}
__device__ __forceinline__ void sts_u32(u32 sa, u32 v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v2(u32 sa, u32 x, u32 y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sa), "r"(x), "r"(y) : "memory"); }
// a kernel parameter pinned in a register (the compiler otherwise re-reads it from the constant bank per base)
__device__ __forceinline__ u32 pin_u32(u32 v) { u32 r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }

template <int W>
__global__ void __launch_bounds__(S1_THREADS)
s1_superk(const S1Args a)
{
  extern __shared__ __align__(16) unsigned char smem[];
  // layout: pack[pack_words*128] | ring[wlen*128] | hist[P] | gbase[P] | kc[P] | ev[4*S1_EVW] (uint2: event, partition | rank << 16)
  u32* s_pack = reinterpret_cast<u32*>(smem);
  u32* s_ring = s_pack + a.pack_words * S1_THREADS;
  u32* s_hist = s_ring + a.wlen * S1_THREADS;
  u32* s_gbase = s_hist + a.P;
  u32* s_kc = s_gbase + a.P;
  uint2* s_ev = reinterpret_cast<uint2*>(s_kc + a.P + (a.P & 1u));     // 8-byte aligned
  __shared__ u32 s_wcount[4];
  __shared__ u32 s_maxlen;

  const u32 tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const u32 ltmask = (1u << lane) - 1u;
  const int k = a.k, m = a.m, wlen = a.wlen;
  const u32 mmask = (1u << (2 * m)) - 1u;
  const u32 ban_mask = 0x55555555u & ((1u << (2 * (m - 2))) - 1u);
  const int rsh = 2 * (m - 1);
  const u32 max_nk = (u32)a.max_nk;

  for (u32 p = tid; p < a.P; p += S1_THREADS) { s_hist[p] = 0; s_kc[p] = 0; }
  for (int j = 0; j < wlen; j++) s_ring[j * S1_THREADS + tid] = 0xFFFFFFFFu;
  if (tid < 4) s_wcount[tid] = 0;
  if (tid == 0) s_maxlen = 0;
  __syncthreads();

  u64 seg = (u64)blockIdx.x * S1_THREADS + tid;
  u32 len = 0; u64 start = 0;
  if (seg < a.nseg) { len = a.seg_len[seg]; start = a.seg_start[seg]; }
  if (len < (u32)k) len = 0;                    // Sequence2SuperKmer.hpp:143-144
  {
    u32 ml = len;
    for (int o = 16; o > 0; o >>= 1) ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, o));
    if (lane == 0) atomicMax(&s_maxlen, ml);
  }
  __syncthreads();
  const u32 maxlen = s_maxlen;
  if (maxlen == 0) return;

  // character reader: aligned 32-bit words + funnel shift
  const uint8_t* addr = a.text + start;
  const u32* wp = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(addr) & ~(uintptr_t)3);
  const u32 csh = 8u * (u32)(reinterpret_cast<uintptr_t>(addr) & 3);
  const u32* wend = reinterpret_cast<const u32*>((reinterpret_cast<uintptr_t>(a.text) + a.text_bytes + 3) & ~(uintptr_t)3);
  // software pipeline: the word consumed in iteration t was requested in iteration t-2
  auto ldw = [&](const u32* q) -> u32 { return (len && q < wend) ? __ldg(q) : 0u; };
  u32 w0 = ldw(wp), w1 = ldw(wp + 1), w2 = ldw(wp + 2);
  wp += 3;

  u32 fm = 0, rm = 0;                 // rolling forward / revcomp m-mer
  u32 pk = 0;                         // packed bases of the current 16-base word
  u32 vrun = 0;                       // valid bases in a row ending here: the k-mer ending here is valid iff vrun >= k
  u32 nk = 0;                         // k-mers in the open record
  u32 cur_min = 0;
  u32 pre = 0xFFFFFFFFu;
  u32 roff = 0;                       // (m-mer index mod wlen) * 512 = byte offset of the ring row (uniform)
  const u32 ring_bytes = (u32)wlen * S1_THREADS * 4u;
  u32 wcnt = 0;                       // events in this warp's queue (same value in every lane)
  u32* ring_col = s_ring + tid;
  const u32 ring_sa = (u32)__cvta_generic_to_shared(ring_col);
  u32 ra = ring_sa;                   // shared address of this thread's word in the current ring row
  const u32 evq_sa = (u32)__cvta_generic_to_shared(s_ev + wid * S1_EVW);
  const u32 kk = pin_u32((u32)k);
  const uint16_t* __restrict__ repart = a.repart;
  const u32 mm1 = (u32)m - 1u;

  // one extra step (i == len) acts as an invalid terminator that closes the last record
  for (u32 i0 = 0; i0 <= maxlen; i0 += 4) {
    const u32 c4 = __funnelshift_r(w0, w1, csh);
    w0 = w1; w1 = w2;
    w2 = (i0 + 8 < len + 4) ? ldw(wp) : 0u;
    ++wp;
#pragma unroll
    for (u32 ii = 0; ii < 4; ii++) {
      const u32 i = i0 + ii;
      const bool active = i < len;
      const u32 ch = (c4 >> (8 * ii)) & 0xFFu;
      const u32 c = (ch >> 1) & 3u;
      const bool valid = active && (((0x47544341u >> (8 * c)) & 0xFFu) == (ch & 0xDFu));   // "ACTG"[c]
      fm = ((fm << 2) | c) & mmask;
      rm = (rm >> 2) | ((c ^ 2u) << rsh);
      if (active) pk = (pk << 2) | c;
      vrun = valid ? vrun + 1u : 0u;
      // lut value of the m-mer ending here (garbage before base m-1: it only feeds windows of
      // k-mers that are not valid yet)
      u32 canon = min(fm, rm);
      u32 t = ~(canon | (canon >> 2));
      t = ((t >> 1) & t) & ban_mask;
      const u32 lutv = (t || i < mm1) ? mmask : canon;
      u32 s = 0xFFFFFFFFu;
      if (roff + 512u < ring_bytes) s = lds_next_row(ra);
      sts_u32(ra, lutv);
      pre = (roff == 0) ? lutv : min(pre, lutv);
      const u32 wmin = min(s, pre);
      roff += 512u; ra += 512u;
      if (roff == ring_bytes) { roff = 0; ra = ring_sa; ring_suffix_min(ring_col, wlen); }
      // ---- cut decision (uniform code: ballot-allocated slot in the warp's event queue)
      const bool kvalid = vrun >= kk;
      const bool cut = nk && (!kvalid || wmin != cur_min || nk == max_nk);
      const u32 cmask = __ballot_sync(0xffffffffu, cut);
      if (cut) {
        const u32 slot = wcnt + __popc(cmask & ltmask);
        sts_v2(evq_sa + slot * 8u, tid | (i << 7) | (nk << 19), cur_min); // record = bases [i-(k+nk-1), i); partition looked up at the flush
        nk = 0;
      }
      wcnt += __popc(cmask);
      if (kvalid && nk == 0) cur_min = wmin;
      nk += kvalid ? 1u : 0u;
    }
    if ((i0 & 12u) == 12u && i0 + 3 < len) s_pack[(i0 >> 4) * S1_THREADS + tid] = pk;
    // every 16 bases: decide (CTA-uniformly) whether to flush the events
    if ((i0 & 12u) == 12u || i0 + 4 > maxlen) {
      const bool last = i0 + 4 > maxlen;
      const int need = __syncthreads_or((int)(wcnt > S1_EVTHR) | (int)last);
      if (need) {
        const u32 iend = min(i0 + 4u, len);            // bases [0, iend) of this thread are packed: publish the partial word
        if (iend && (iend & 15u)) s_pack[((iend - 1u) >> 4) * S1_THREADS + tid] = pk << (2u * (16u - (iend & 15u)));
        if (lane == 0) s_wcount[wid] = wcnt;
        wcnt = 0;
        __syncthreads();
        // pass 1: per-partition rank of every event, k-mer totals
        for (u32 w = 0; w < 4; w++) {
          const u32 n = s_wcount[w];
          for (u32 r = tid; r < n; r += S1_THREADS) {
            const u32 q = w * S1_EVW + r;
            const uint2 e = s_ev[q];
            const u32 p = __ldg(repart + e.y);              // e.y = minimizer of the record (Repartitor, PartiInfo.hpp:381)
            if (a.mload) atomicAdd(a.mload + e.y, (u64)((e.x >> 19) & 127u));   // repartition estimate: k-mers per minimizer
            s_ev[q].y = p | (atomicAdd(&s_hist[p], 1u) << 16);
            atomicAdd(&s_kc[p], (e.x >> 19) & 127u);
          }
        }
        __syncthreads();
        for (u32 p = tid; p < a.P; p += S1_THREADS) {
          u32 cnt = s_hist[p];
          if (cnt) {
            s_gbase[p] = atomicAdd(&a.cursor[p], cnt);
            atomicAdd(&a.kcnt[p], (u64)s_kc[p]);
            s_hist[p] = 0; s_kc[p] = 0;
          }
        }
        __syncthreads();
        // pass 2: build every record from the packed bases and store it
        uint4* out = reinterpret_cast<uint4*>(a.records);
        for (u32 w = 0; w < 4; w++) {
          const u32 n = s_wcount[w];
          for (u32 r = tid; r < n; r += S1_THREADS) {
            const u32 q = w * S1_EVW + r;
            const uint2 e = s_ev[q];
            const u32 ev = e.x, p = e.y & 0xFFFFu;
            const u32 t = ev & 127u, nkr = (ev >> 19) & 127u;
            const u32 nb = (u32)k + nkr - 1u;                     // bases in the record
            const u32 st = ((ev >> 7) & 4095u) - nb;              // the event carries the base index just past the record
            const u32 pos = s_gbase[p] + (e.y >> 16);
            const int top = (int)((st + nb - 1u) >> 4);
            const int s = 2 * (int)(15u - ((st + nb - 1u) & 15u));
            const u32 bits = 2u * nb;
            u32 v[4 * W];
#pragma unroll
            for (int jw = 0; jw < 4 * W; jw++) {
              u32 x = pack_word(s_pack, t, a.pack_words, top, s, jw);
              const int lo = 32 * jw;
              if ((int)bits <= lo) x = 0u;
              else if ((int)bits < lo + 32) x &= (1u << (bits - lo)) - 1u;
              v[jw] = x;
            }
            v[4 * W - 1] |= nb << 24;
            if (pos < a.bcap[p]) {
              uint4* dst = out + (size_t)W * (a.boff[p] + pos);
              dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
              if (W == 2) dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
            } else *a.overflow = 1u;
          }
        }
        __syncthreads();
        if (tid < 4) s_wcount[tid] = 0;
        __syncthreads();
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// host-side launchers (called from kmx_api.cu)
// ------------------------------------------------------------------------------------
size_t s1_smem_bytes(u32 pack_words, u32 stage_cap, int wlen, u32 P)
{
  (void)stage_cap;
  return (size_t)pack_words * S1_THREADS * 4 + (size_t)wlen * S1_THREADS * 4 + (size_t)P * 12 + 8 + (size_t)4 * S1_EVW * 8;
}

cudaError_t launch_fq_index(const uint8_t* text, u64 nbytes, u32* tile_counts, u64* tile_prefix, u64* nlmask,
                            u64* d_total, u32* seq_start, u32* seq_len, u64 nrec_cap, u32* flags,
                            int phase, cudaStream_t st, u64* launches)
{
  uintptr_t ta = reinterpret_cast<uintptr_t>(text);
  const uint4* base = reinterpret_cast<const uint4*>(ta & ~(uintptr_t)15);
  u64 lead = ta & 15, tot = lead + nbytes;
  u64 ntiles = (tot + FQ_TILE - 1) / FQ_TILE;
  if (phase == 0) {
    fq_count_newlines<<<(unsigned)ntiles, FQ_THREADS, 0, st>>>(base, lead, tot, tile_counts, (uint16_t*)nlmask);
    scan_u32_to_u64<<<1, 1024, 0, st>>>(tile_counts, tile_prefix, ntiles, d_total);
    *launches += 2;
  } else {
    fq_index_lines<<<(unsigned)ntiles, FQ_THREADS, 0, st>>>(base, lead, tot, tile_prefix, nlmask, seq_start, seq_len, nrec_cap, flags);
    *launches += 1;
  }
  return cudaGetLastError();
}

cudaError_t launch_scan_u32(const u32* in, u64* out, u64 n, u64* total, cudaStream_t st, u64* launches)
{
  scan_u32_to_u64<<<1, 1024, 0, st>>>(in, out, n, total);
  *launches += 1;
  return cudaGetLastError();
}

u64 fq_num_tiles(const uint8_t* text, u64 nbytes)
{
  u64 lead = reinterpret_cast<uintptr_t>(text) & 15;
  return (lead + nbytes + FQ_TILE - 1) / FQ_TILE;
}

cudaError_t launch_s1(int W, const S1Args& a, cudaStream_t st, u64* launches)
{
  if (a.nseg == 0) return cudaSuccess;
  size_t smem = s1_smem_bytes(a.pack_words, a.stage_cap, a.wlen, a.P);
  unsigned grid = (unsigned)((a.nseg + S1_THREADS - 1) / S1_THREADS);
  cudaError_t e;
  if (W == 1) {
    e = cudaFuncSetAttribute(s1_superk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    s1_superk<1><<<grid, S1_THREADS, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(s1_superk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    s1_superk<2><<<grid, S1_THREADS, smem, st>>>(a);
  }
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace kmx
