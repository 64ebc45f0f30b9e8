// s1_v5.cuh -- stage 1, position-parallel formulation (reads of <= a few hundred bases).
//
// Same function as s1_superk (s1_superk.cu): per read, every valid k-mer's minimizer
// (gatb Model.hpp:1254-1287: plain minimum of lut[m-mer] over the k-m+1 m-mers of the forward
// k-mer, lut = canonical m-mer or 4^m-1 when banned, Model.hpp:1040-1064,1220-1251), super-k-mer
// cuts (Sequence2SuperKmer.hpp:90-158: minimizer change, invalid k-mer, record full) and the
// scatter of the records to the partition of their minimizer (fill_partitions.hpp:59-63).
//
// Instead of one thread walking one read base by base, a CTA takes R reads through phases
// whose work items are independent (no rolling state, no per-base control flow):
//   P0  16 bases per item: characters -> 2-bit codes by SWAR, packed big-endian (forward
//       strand) and little-endian complemented (reverse strand), validity bits
//   P1  one m-mer per item: forward and reverse-complement m-mer by ONE funnel shift each
//       out of the two packed streams, lut value by arithmetic
//   P2  one block of w = k-m+1 k-mers per item: sliding-window minimum by block suffix /
//       running prefix minima (van Herk), minimizer-change bitmask of the block
//   P3  one block of one read per item: completes the block's change mask, counts its records (about
//       one per 11 k-mers), takes that many slots of the CTA's event queue with one shared-memory
//       atomic and logs one event per record; a read with invalid bases takes a per-k-mer walk by one thread
//   P4  flush (as in s1_superk): per-partition ranking in shared memory, ONE global atomic per
//       (CTA, partition), records built by funnel shifts from the packed forward stream
//
// The phase bodies are plain functions of (item index, shared arrays) and compile for the host
// too: tests/emul/s1v5_emul.cpp runs a CTA phase by phase on the CPU and checks the emitted records,
// so the index arithmetic is tested without a GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define KMX_HD __host__ __device__ __forceinline__
#else
#define KMX_HD inline
#endif

namespace kmx {
namespace s1v5 {

typedef unsigned int u32;
typedef unsigned long long u64;

static const u32 EVCAP_MAX = 1024;    // cut events per flush round (at most)
static const u32 INF = 0xFFFFFFFFu;

KMX_HD u32 fsr(u32 lo, u32 hi, u32 s)      // low 32 bits of (hi:lo) >> s, 0 <= s < 32
{
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
KMX_HD u32 fsl(u32 lo, u32 hi, u32 s)      // high 32 bits of (hi:lo) << s, 0 <= s < 32
{
#ifdef __CUDA_ARCH__
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
KMX_HD u32 umin(u32 a, u32 b) { return a < b ? a : b; }
KMX_HD u32 popc64(u64 v)
{
#ifdef __CUDA_ARCH__
  return (u32)__popcll(v);
#else
  return (u32)__builtin_popcountll(v);
#endif
}
KMX_HD u32 clz64(u64 v)                      // v != 0
{
#ifdef __CUDA_ARCH__
  return (u32)__clzll((long long)v);
#else
  return (u32)__builtin_clzll(v);
#endif
}
// two adjacent words at an 8-byte aligned address
KMX_HD void ld2(const u32* p, u32& a, u32& b)
{
#ifdef __CUDA_ARCH__
  const uint2 v = *reinterpret_cast<const uint2*>(p); a = v.x; b = v.y;
#else
  a = p[0]; b = p[1];
#endif
}
KMX_HD u32 ctz64(u64 v)
{
#ifdef __CUDA_ARCH__
  return (u32)(__ffsll((long long)v) - 1);
#else
  return (u32)__builtin_ctzll(v);
#endif
}

// geometry of one CTA's shared arrays (host-computed from the longest read of the launch)
struct Geo {
  u32 R;          // reads per CTA
  u32 nch;        // 16-base chunks per read = ceil(maxlen / 16)
  u32 LW;         // words per read in BE / LE = 2 nch: word pairs (chunk c, chunk c+1) so that P1 takes one aligned 64-bit load per strand
  u32 Lpad;       // words per read in U: >= maxlen - m + 1, odd (bank-conflict-free with lane = read)
  u32 Spad;       // words per read in S: >= nblk * w, odd
  u32 nblk;       // blocks of w k-mers per read = ceil((maxlen - k + 1) / w)
  u32 evcap;      // events per flush round: the queue reuses U (dead after P2)
  u32 inv_nblk;   // ceil(2^32 / nblk) (nblk >= 2): item -> read by one multiply
};
KMX_HD Geo make_geo(u32 R, u32 maxlen, int k, int m)
{
  Geo g; g.R = R; g.nch = (maxlen + 15) / 16; g.LW = 2 * g.nch;
  g.Lpad = (maxlen - (u32)m + 1) | 1u;
  const u32 w = (u32)(k - m + 1), nk = maxlen - (u32)k + 1;
  g.nblk = (nk + w - 1) / w;
  g.Spad = (g.nblk * w) | 1u;
  g.evcap = umin(EVCAP_MAX, (R * g.Lpad) / 2);
  g.inv_nblk = g.nblk >= 2 ? (u32)((0x100000000ull + g.nblk - 1) / g.nblk) : 0u;
  return g;
}
// 32-bit words of shared memory: BE | LE | CH | U (later the event queue) | S | NX | len | start | inval | done | hist | gbase | kc | VB (u16)
KMX_HD size_t smem_words(const Geo& g, u32 P)
{
  const size_t nt = (size_t)g.R * g.nblk;
  return (size_t)g.R * g.Lpad + (size_t)g.R * g.Spad + (size_t)2 * g.R * g.LW + ((size_t)g.R * g.nch + 1) / 2 + 3 * nt + (size_t)3 * g.R + nt + 1 + (size_t)3 * P;
}
KMX_HD size_t smem_bytes(const Geo& g, u32 P) { return smem_words(g, P) * 4 + 16; }

// read of item t = t / nblk (exact: t * nblk < 2^32)
KMX_HD u32 item_read(const Geo& g, u32 t)
{
  if (g.nblk < 2) return t;
#ifdef __CUDA_ARCH__
  return __umulhi(t, g.inv_nblk);
#else
  return (u32)(((u64)t * g.inv_nblk) >> 32);
#endif
}
struct Ev { u32 x, y; };   // x = read | (base index just past the record) << 7 | (#k-mers) << 19 ; y = minimizer, later partition | rank << 16

struct Cta {
  int k, m, w; u32 max_nk, mmask, ban_mask;
  Geo g;
  u32 *BE, *LE, *U, *S, *CH, *NX, *len, *start, *inval, *done, *hist, *gbase, *kc;
  uint16_t* VB;
  Ev* ev;
};
KMX_HD void carve(Cta& x, u32* base /* 8-byte aligned */, u32 P)
{
  const Geo& g = x.g;
  const size_t nt = (size_t)g.R * g.nblk;
  u32* p = base;
  x.BE = p; p += (size_t)g.R * g.LW;                         // even word offsets: 64-bit loads
  x.LE = p; p += (size_t)g.R * g.LW;
  x.CH = p; p += 2 * nt;
  x.U = p; x.ev = reinterpret_cast<Ev*>(p); p += (size_t)g.R * g.Lpad;
  x.S = p; p += (size_t)g.R * g.Spad;
  x.NX = p; p += nt;
  x.len = p; p += g.R; x.start = p; p += g.R; x.inval = p; p += g.R;
  x.done = p; p += nt + 1;
  x.hist = p; p += P; x.gbase = p; p += P; x.kc = p; p += P;
  x.VB = reinterpret_cast<uint16_t*>(p);
}

// ---- P0: bases [16c, 16c+16) of read r ---------------------------------------------------
// wend = first 4-byte-aligned address at or past the end of the text (words at or past it read as 0)
KMX_HD void p0_pack(const Cta& x, u32 r, u32 c, const uint8_t* rd /* first base of the read */, u32 len, const u32* wend)
{
  const u32 b0 = 16u * c;
  u32 be = 0, le = 0, vb = 0, act16 = 0;
  if (b0 < len) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(rd + b0);
    const u32* wp = reinterpret_cast<const u32*>(ad & ~(uintptr_t)3);
    const u32 sh = 8u * (u32)(ad & 3);
    u32 wv[5];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 5; j++) wv[j] = 0u;
    if (wp + 5 <= wend) {                          // all but the last few bytes of the text
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int j = 0; j < 5; j++) wv[j] = wp[j];
    } else {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int j = 0; j < 5; j++) if (wp + j < wend) wv[j] = wp[j];
    }
    u32 inv = 0;                                   // bit i: character i of the chunk is not a valid letter
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 4; j++) {
      const u32 c4 = fsr(wv[j], wv[j + 1], sh);
      const u32 cd = (c4 >> 1) & 0x03030303u;      // code = (c >> 1) & 3 (Data.hpp:179)
      const u32 b0m = cd & 0x01010101u, b1m = (cd >> 1) & 0x01010101u;
      const u32 e = 0x41414141u + 2u * b0m + 0x13u * b1m - 0x0Fu * (b0m & b1m);   // "ACTG"[code] per byte
      const u32 d = (c4 & 0xDFDFDFDFu) ^ e;                                          // zero byte <=> valid letter
      const u32 nz = ((((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) >> 7) & 0x01010101u;  // bit 8b <=> byte b != 0
      inv |= (((nz * 0x00204081u) >> 21) & 0xFu) << (4 * j);                         // gather bits 0,8,16,24
      be |= ((cd * 0x40100401u) >> 24) << (24 - 8 * j);        // byte = c0<<6 | c1<<4 | c2<<2 | c3
      le |= ((cd * 0x01041040u) >> 24) << (8 * j);             // byte = c0 | c1<<2 | c2<<4 | c3<<6
    }
    le ^= 0xAAAAAAAAu;                                            // complement
    const u32 rem = len - b0;                                     // active bases from here on (may exceed 16)
    act16 = rem >= 16 ? 0xFFFFu : ((1u << rem) - 1u);
    if (rem < 16) be &= ~(0xFFFFFFFFu >> (2 * rem));              // bases past the end read as 0
    vb = ~inv & act16;
  }
  // word pairs: pair c = (chunk c, chunk c+1)
  u32* bep = x.BE + r * x.g.LW + 2 * c;
  u32* lep = x.LE + r * x.g.LW + 2 * c;
  bep[0] = be; lep[0] = le;
  if (c) { bep[-1] = be; lep[-1] = le; }
  if (c + 1 == x.g.nch) { bep[1] = 0u; lep[1] = 0u; }
  x.VB[r * x.g.nch + c] = (uint16_t)vb;
  if (vb != act16) x.inval[r] = 1u;                // same value from every writer
}

// ---- P1: lut values of the m-mers a = lane, lane + 32, ... < nm of read r (one warp per read) ----
// a advances by 32 bases = 2 packed words, so the shift amounts are loop-invariant
KMX_HD void p1_row(const Cta& x, u32 r, u32 lane, u32 nm)
{
  const u32 q2 = 2u * (lane & 15u), sh = 32u - 2u * (u32)x.m;
  const u32 mmask = x.mmask, ban = x.ban_mask;
  const u32* be = x.BE + r * x.g.LW + 2 * (lane >> 4);
  const u32* le = x.LE + r * x.g.LW + 2 * (lane >> 4);
  u32* u = x.U + r * x.g.Lpad + lane;
  u32 a = lane;
#define KMX_P1_ONE(I)                                                                                   \
  {                                                                                                     \
    u32 b0w, b1w, l0w, l1w;                                                                             \
    ld2(be + 4 * (I), b0w, b1w); ld2(le + 4 * (I), l0w, l1w);                                           \
    const u32 fm = fsl(b1w, b0w, q2) >> sh;       /* bases a .. a+m-1, first base most significant */   \
    const u32 rm = fsr(l0w, l1w, q2) & mmask;     /* its reverse complement */                          \
    const u32 canon = umin(fm, rm);                                                                     \
    const u32 o2 = canon | (canon >> 2);          /* bit 2j (and 2j+1) clear <=> bases j and j+1 are both A ... */       \
    const u32 t = ~(o2 | (o2 >> 1)) & ban;        /* ... "AA" anywhere but at the two leading bases (Model.hpp:1220-1251) */ \
    u[32 * (I)] = t ? mmask : canon;                                                                    \
  }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (; a + 96 < nm; a += 128, be += 16, le += 16, u += 128) { KMX_P1_ONE(0) KMX_P1_ONE(1) KMX_P1_ONE(2) KMX_P1_ONE(3) }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (; a < nm; a += 32, be += 4, le += 4, u += 32) KMX_P1_ONE(0)
#undef KMX_P1_ONE
}

// ---- P2: block g of read r: minimizers of k-mers [g w, g w + w) and their change mask ----
// window of k-mer t = U[t .. t+w-1] = suffix of block g from t, then prefix of block g+1 up to t+w-1.
// Loads go in batches of 8 ahead of the serial min chain.  WC > 0: w is the compile-time constant WC (the default
// k = 31, m = 10 gives 22) and a full block is one straight line of code with constant offsets; WC = 0: any w.
template <int WC>
KMX_HD void p2_block(const Cta& x, u32 r, u32 g, u32 len)
{
  const u32 w = WC ? (u32)WC : (u32)x.w;
  const u32 nk = len >= (u32)x.k ? len - (u32)x.k + 1u : 0u;
  const u32 lo = g * w;
  u32* ch = x.CH + 2 * (r * x.g.nblk + g);
  if (lo >= nk) { ch[0] = 0; ch[1] = 0; return; }
  const u32* U = x.U + r * x.g.Lpad;
  u32* S = x.S + r * x.g.Spad;
  const u32 hi = lo + w;                              // k-mer lo exists, so m-mers lo .. lo+w-1 do
  if (WC > 0 && hi <= nk) {                           // all w k-mers of the block exist
    constexpr int WW = WC > 0 ? WC : 1;
    const u32* Ub = U + lo; u32* Sb = S + lo;
    u32 acc = INF;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = (WW - 1) / 8; b >= 0; b--) {         // suffix minima, batches [8b, 8b+8) from the top
      u32 v[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 8; q++) if (8 * b + q < WW) v[q] = Ub[8 * b + q];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 7; q >= 0; q--) if (8 * b + q < WW) { acc = umin(acc, v[q]); Sb[8 * b + q] = acc; }
    }
    u32 prev = acc, pre = INF, clo = 0, chi = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = 0; 4 * b + 1 < WW; b++) {            // k-mers lo + j, j = 1 + 4b + q: window = suffix from j, prefix of the next block up to j + w - 1
      u32 uu[4], ss[4];                               // (batches of 4: the kernel must stay within 32 registers for 5 CTAs per SM)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 4; q++) if (1 + 4 * b + q < WW) { uu[q] = Ub[WW + 4 * b + q]; ss[q] = Sb[1 + 4 * b + q]; }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 4; q++) if (1 + 4 * b + q < WW) {
        const int j = 1 + 4 * b + q;
        pre = umin(pre, uu[q]);
        const u32 mz = umin(ss[q], pre);
        Sb[j] = mz;
        if (mz != prev) { if (j < 32) clo |= 1u << (j & 31); else chi |= 1u << (j & 31); }
        prev = mz;
      }
    }
    ch[0] = clo; ch[1] = chi;
    return;
  }
  u32 acc = INF;
  {
    const u32* up = U + hi; u32* sp = S + hi;
    u32 n = w;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; n >= 8; n -= 8) {
      up -= 8; sp -= 8;
      u32 v[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 8; q++) v[q] = up[q];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 7; q >= 0; q--) { acc = umin(acc, v[q]); sp[q] = acc; }
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; n; n--) { --up; --sp; acc = umin(acc, *up); *sp = acc; }
  }
  const u32 tend = umin(hi, nk);
  u32 prev = acc, pre = INF;
  u64 chg = 0;
  {
    u32 t = lo + 1;
    const u32* up = U + t + w - 1; u32* sp = S + t;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; t + 8 <= tend; t += 8, up += 8, sp += 8) {
      u32 uu[8], ss[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 8; q++) { uu[q] = up[q]; ss[q] = sp[q]; }
      u32 cm = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = 0; q < 8; q++) {
        pre = umin(pre, uu[q]);
        const u32 mz = umin(ss[q], pre);
        sp[q] = mz;
        cm |= (u32)(mz != prev) << q;
        prev = mz;
      }
      chg |= (u64)cm << (t - lo);
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; t < tend; t++, up++, sp++) {
      pre = umin(pre, *up);
      const u32 mz = umin(*sp, pre);
      *sp = mz;
      chg |= (u64)(mz != prev) << (t - lo);
      prev = mz;
    }
  }
  ch[0] = (u32)chg; ch[1] = (u32)(chg >> 32);
}

// ---- P3 -------------------------------------------------------------------------------------
// P3a, item (r, g): completes the block's change mask (bit 0 needs the neighbour block's last minimizer),
// finds NX = the first cut after the block (looking ahead over blocks without a cut) and returns the number
// of records that START in the block.  Every run but the last one of the block ends inside the block
// (shorter than w <= max_nk: one record each); the last one runs to NX.
KMX_HD u32 p3_slow_count(const Cta& x, u32 r, u32 len);
KMX_HD u32 p3_prepare(const Cta& x, u32 r, u32 g, u32 len)
{
  const u32 k = (u32)x.k, w = (u32)x.w;
  if (len < k) return 0;
  const u32 nk = len - k + 1u, lo = g * w;
  if (lo >= nk) return 0;
  if (x.inval[r]) return g == 0 ? p3_slow_count(x, r, len) : 0u;
  const u32* M = x.S + r * x.g.Spad;
  u32* ch = x.CH + 2 * (r * x.g.nblk + g);
  u32 c0 = ch[0];
  if (g == 0 || M[lo] != M[lo - 1]) { c0 |= 1u; ch[0] = c0; }
  const u64 mask = (u64)c0 | ((u64)ch[1] << 32);
  if (!mask) return 0;
  u32 nxt = nk;
  for (u32 g2 = g + 1, lo2 = lo + w; lo2 < nk; g2++, lo2 += w) {
    // the owner of block g2 may or may not have set its bit 0 yet: recompute it
    const u64 m2 = (u64)(ch[2 * (g2 - g)] | (u32)(M[lo2] != M[lo2 - 1])) | ((u64)ch[2 * (g2 - g) + 1] << 32);
    if (m2) { nxt = lo2 + ctz64(m2); break; }
  }
  x.NX[r * x.g.nblk + g] = nxt;
  u32 last = nxt - (lo + 63u - clz64(mask));            // k-mers in the last run
  u32 n = popc64(mask);
  while (last > x.max_nk) { last -= x.max_nk; n++; }
  return n;
}

// events of the run of k-mers [a, b) with one minimizer: pieces of at most max_nk k-mers.
// returns their number; EMIT: also writes them at x.ev[slot...]
template <bool EMIT>
KMX_HD u32 p3_run(const Cta& x, u32 r, const u32* M, u32 a, u32 b, u32 slot)
{
  u32 n = 0;
  const u32 mz = M[a];
  for (u32 s = a; s < b; s += x.max_nk) {
    const u32 nkr = umin(x.max_nk, b - s);
    if (EMIT) { Ev e; e.x = r | ((s + nkr + (u32)x.k - 1u) << 7) | (nkr << 19); e.y = mz; x.ev[slot + n] = e; }
    n++;
  }
  return n;
}
// read with invalid bases, whole read by one thread: k-mer t is valid iff bases [t, t+k) are all valid letters
template <bool EMIT>
KMX_HD u32 p3_slow(const Cta& x, u32 r, u32 len, u32 slot)
{
  const u32 k = (u32)x.k, nk = len - k + 1u;
  const u32* M = x.S + r * x.g.Spad;
  const uint16_t* vb = x.VB + r * x.g.nch;
  u32 n = 0, vr = 0;
  for (u32 i = 0; i + 1 < k; i++) vr = ((vb[i >> 4] >> (i & 15u)) & 1u) ? vr + 1u : 0u;
  bool open = false; u32 s0 = 0, cm = 0, nn = 0;
  for (u32 t = 0; t < nk; t++) {
    const u32 i = t + k - 1u;
    vr = ((vb[i >> 4] >> (i & 15u)) & 1u) ? vr + 1u : 0u;
    if (vr >= k) {
      const u32 mz = M[t];
      if (!open) { open = true; s0 = t; cm = mz; nn = 1; }
      else if (mz != cm || nn == x.max_nk) { n += p3_run<EMIT>(x, r, M, s0, s0 + nn, slot + n); s0 = t; cm = mz; nn = 1; }
      else nn++;
    } else if (open) { n += p3_run<EMIT>(x, r, M, s0, s0 + nn, slot + n); open = false; }
  }
  if (open) n += p3_run<EMIT>(x, r, M, s0, s0 + nn, slot + n);
  return n;
}
KMX_HD u32 p3_slow_count(const Cta& x, u32 r, u32 len) { return p3_slow<false>(x, r, len, 0); }

// P3c, item (r, g): its events -> x.ev[slot ...] in k-mer order, y = minimizer.  (Reads with invalid
// bases are emitted by p3_slow<true>, one thread per read.)
KMX_HD void p3_put(const Cta& x, u32 slot, u32 r, u32 s, u32 e, u32 mz)
{
  Ev ev; ev.x = r | ((e + (u32)x.k - 1u) << 7) | ((e - s) << 19); ev.y = mz;
  x.ev[slot] = ev;
}
KMX_HD void p3_emit_item(const Cta& x, u32 r, u32 g, u32 slot)
{
  const u32 lo = g * (u32)x.w;
  const u32* M = x.S + r * x.g.Spad;
  const u32* ch = x.CH + 2 * (r * x.g.nblk + g);
  u64 mask = (u64)ch[0] | ((u64)ch[1] << 32);
  if (!mask) return;
  const u32 nxt = x.NX[r * x.g.nblk + g];
  u32 prev = lo + ctz64(mask); mask &= mask - 1;
  while (mask) {
    const u32 t = lo + ctz64(mask); mask &= mask - 1;
    p3_put(x, slot++, r, prev, t, M[prev]);
    prev = t;
  }
  const u32 mz = M[prev];
  for (u32 s0 = prev; s0 < nxt; s0 += x.max_nk) p3_put(x, slot++, r, s0, umin(s0 + x.max_nk, nxt), mz);
}

// ---- P4: the record of one event: nb = k + nk - 1 bases ending just before base `iend` ------
// v[0..NW) little-endian words of the big number (first base most significant), length in the top byte
template <int NW>
KMX_HD void build_record(const u32* be /* packed forward stream of the read (word pairs) */, u32 nwords /* chunks */, u32 iend, u32 nb, u32* v)
{
  const u32 last = iend - 1u;                      // last base of the record
  const int top = (int)(last >> 4);
  const int s = 2 * (int)(15u - (last & 15u));
  const int bits = 2 * (int)nb;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int jw = 0; jw < NW; jw++) {
    const int q = s + 32 * jw;
    const int wi = q >> 5, sh = q & 31;
    const int a = top - wi, b = a - 1;             // b is the more significant neighbour
    const u32 xa = (a >= 0 && a < (int)nwords) ? be[2 * a] : 0u;      // paired layout: chunk a is word 2a
    const u32 xb = (b >= 0 && b < (int)nwords) ? be[2 * b] : 0u;
    u32 xw = fsr(xa, xb, (u32)sh);
    const int lo = 32 * jw;
    if (bits <= lo) xw = 0u;
    else if (bits < lo + 32) xw &= (1u << (bits - lo)) - 1u;
    v[jw] = xw;
  }
  v[NW - 1] |= nb << 24;
}

}  // namespace s1v5
}  // namespace kmx
