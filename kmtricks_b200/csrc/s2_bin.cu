// s2_bin.cu -- stage 2, hash keys, k <= 32: hash-BINNED counting in shared memory.
//
// Replaces (behaviour, not code) ReadSuperkHash / HashSort / HashPartCounter::executeDump
// (include/kmtricks/gatb/sorting_count.hpp:387-470,525-528,971-990), KmXXHash (:346-363) and
// HashCountProcessor (include/kmtricks/gatb/count_processor.hpp:61-70): per (sample, partition) the
// ascending list of (hash key, count >= hard_min).
//
// The L2-histogram path (s2_hash.cu) pays one L2 atomic per k-mer (1.9e11 RED/s on B200 = 0.63 ms per
// 1.2e8 k-mers, measured by tools/ubench) and then streams the whole P x W histogram to find a few %
// of survivors.  Shared-memory atomics are ~7x faster (>= 1.3e12/s at full occupancy, same tool), so:
//
//   pass A  hash_bin_kernel       every k-mer is hashed ONCE (rolled forward/reverse k-mer, XXH64, Barrett
//                                 modulo) and its window slot is split into (bin = slot >> bs_log, 16-bit
//                                 offset).  A CTA takes a tile of 512 records (~5800 k-mers); every offset goes
//                                 straight to slot (bin, rank) of a shared-memory staging array, the rank from
//                                 one shared-memory atomic, and each bin's run is appended to the bin's region
//                                 of the lane's bin buffer with ONE global atomic per (tile, bin): 2 bytes per
//                                 k-mer, written in coalesced runs.
//   pass B  hash_bincount_kernel  one CTA per (window, bin): counts the bin's offsets into a 64 KB
//                                 shared-memory histogram (32 K slots x 16 bit, or 16 K x 32 bit), then
//                                 emits the survivors in slot order straight from shared memory; the
//                                 output offset of a bin comes from a decoupled look-back over the bins'
//                                 survivor counts (in-order tickets), so the lists are contiguous and
//                                 ascending without a separate scan or copy pass.
//
// No P x W histogram in HBM, no sweep.  A 16-bit counter that wraps is detected exactly (the fields of a
// bin must add up to the bin's entries) and the sample is redone with 32-bit counters, as before.
#include "common.cuh"
#include "kmx_internal.h"
#include "records.cuh"

#include <algorithm>
#include <mutex>

namespace kmx {

static constexpr int HB_THREADS = 256;
static constexpr int HB_WARPS = HB_THREADS / 32;
static constexpr int HB_TR = 512;                     // records per tile
static constexpr int HB_PER = HB_TR / HB_THREADS;
static constexpr int HB_GROUPS = HB_TR / 32;          // groups of 32 length-sorted records (one warp pass each)
static constexpr u32 HB_S = 12288;                    // staged offsets per round: NB bins x C slots, C = HB_S / NB
static constexpr u32 HB_TPAD = 7168;                  // k-mer slots hashed per round (a group counts 32 x its longest record): mean bin load <= 0.58 C

// ---- pass A ------------------------------------------------------------------------------------
// Hand-scheduled 32-bit arithmetic for the per-k-mer loop (the compiler's 64-bit expansion of XXH64 + modulo takes 55
// instructions, this takes 35; tools/ubench and the algebra check in tests/test_host_logic.py::test_hash_mod_halves).
__device__ __forceinline__ u64 hb_mulw(u32 a, u32 b) { u64 d; asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ u64 hb_madw(u32 a, u32 b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ u32 hb_mad(u32 a, u32 b, u32 c) { u32 d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// (l, h) <- (l, h) * C mod 2^64 : one wide multiply + two multiply-adds into the high word
#define HB_MUL64(l, h, C) do { const u64 t_ = hb_mulw(l, (u32)(C)); h = hb_mad(h, (u32)(C), hb_mad(l, (u32)((C) >> 32), (u32)(t_ >> 32))); l = (u32)t_; } while (0)
// XXH64 of the 8-byte word (xl, xh), seed 0, then x mod d (= fastmod64_d32(xxh64_8(x), fm)) for d < 2^30.
// p4: KMX_P4 held in a register pair (hb_opaque64), so that "* P1 + P4" is ONE wide multiply-add.
__device__ __forceinline__ u32 hb_hash_mod(u32 xl, u32 xh, u32 ml, u32 mh, u32 d, u64 p4)
{
  u32 l = xl, h = xh, a, b;
  HB_MUL64(l, h, KMX_P2);
  a = __funnelshift_l(h, l, 31); b = __funnelshift_l(l, h, 31); l = a; h = b;      // rotl 31
  HB_MUL64(l, h, KMX_P1);
  l ^= (u32)(KMX_P5 + 8ULL); h ^= (u32)((KMX_P5 + 8ULL) >> 32);
  a = __funnelshift_l(h, l, 27); b = __funnelshift_l(l, h, 27); l = a; h = b;      // rotl 27
  { const u64 t = hb_madw(l, (u32)KMX_P1, p4); h = hb_mad(h, (u32)KMX_P1, hb_mad(l, (u32)(KMX_P1 >> 32), (u32)(t >> 32))); l = (u32)t; }
  l ^= h >> 1;                                                                      // h ^= h >> 33
  HB_MUL64(l, h, KMX_P2);
  a = l ^ __funnelshift_r(l, h, 29); h ^= h >> 29; l = a;                          // h ^= h >> 29
  HB_MUL64(l, h, KMX_P3);
  l ^= h;                                                                           // h ^= h >> 32
  // Barrett with m64 = floor((2^64-1)/d): q = bits [64, 96) of x * m64 without the carry out of the lowest partial
  // product (xl * ml), so q is the true quotient or up to 2 less and r = x - q d < 3 d < 2^32: two conditional subtracts.
  const u64 s = hb_mulw(l, mh);
  const u64 t2 = hb_mulw(h, ml) + (u64)(u32)s;
  const u32 q = hb_mad(h, mh, (u32)(s >> 32)) + (u32)(t2 >> 32);
  u32 r = hb_mad(q, 0u - d, l);
  r = min(r, r - d);
  return min(r, r - d);
}
// a value the optimiser cannot rematerialise: it stays in a register across the loop
__device__ __forceinline__ u32 hb_opaque32(u32 v) { u32 r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ u64 hb_opaque64(u64 v) { u64 r; asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(v)); return r; }
// shared-memory accesses by 32-bit shared-window address (formed once: indexing a __shared__ array inside the loop
// makes the compiler rebuild the window base -- S2R SR_CgaCtaId, MOV, LEA -- at every use)
__device__ __forceinline__ u32 hb_saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 hb_atoms_inc(u32 addr) { u32 r; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(addr) : "memory"); return r; }
__device__ __forceinline__ void hb_sts16(u32 addr, u32 v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ u32 hb_lds16(u32 addr) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ u32 hb_lds32(u32 addr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// TMA (1-D bulk copy) + mbarrier: a tile's records are one contiguous 8 KB run of the bucket region ("bucket page"); an elected
// thread has the copy engine of the SM fetch the NEXT tile into the other half of a double buffer while the CTA hashes the
// current one, and the CTA picks it up by waiting on the buffer's mbarrier (SASS: UBLKCP / SYNCS).
__device__ __forceinline__ void hb_mbar_init(u32 mbar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void hb_mbar_expect_tx(u32 mbar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void hb_bulk_g2s(u32 dst, const void* src, u32 bytes, u32 mbar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void hb_mbar_wait(u32 mbar, u32 parity)
{
  asm volatile("{\n\t.reg .pred p;\n\tHB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra HB_DONE;\n\tbra HB_WAIT;\n\tHB_DONE:\n\t}" :: "r"(mbar), "r"(parity) : "memory");
}

// NBMAX: capacity of the per-bin arrays (static shared memory).  FAST: 28 <= k <= 32 -- the bases after a record's first
// k-mer fit one 64-bit word and the roll runs on 32-bit halves with the k-dependent shifts folded into constants.
template <bool FAST, int NBMAX>
__global__ void __launch_bounds__(HB_THREADS, 4)
hash_bin_kernel(HashBinArgs a, FastMod32 fm32)
{
  __shared__ __align__(16) uint16_t s_stage[HB_S];    // bin b: slots [b*C, b*C + C)
  __shared__ __align__(128) uint4 s_rec2[2][HB_TR];   // double buffer of record tiles, filled by bulk copies
  __shared__ __align__(8) u64 s_mbar[2];
  __shared__ u32 s_meta[2][4];                        // per buffer: ticket, window, first record of the tile, records in the tile
  __shared__ uint16_t s_perm[HB_TR];
  __shared__ u32 s_bcnt[NBMAX], s_gcnt[NBMAX], s_gdst[NBMAX];
  __shared__ u32 s_cnt[64], s_start[64];
  __shared__ u32 s_gmax[HB_GROUPS], s_ground[HB_GROUPS];
  __shared__ u32 s_nrounds;
  const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const u32 NB = a.NB, C = HB_S / NB;
  const int k = a.k;
  const u64 kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
  const int rcsh = 2 * (k - 1);
  const u32 bs_log = a.bs_log, bsmask = (1u << bs_log) - 1u;
  const u32 ml = (u32)fm32.m64, mh = (u32)(fm32.m64 >> 32), md = fm32.d;
  const u32 khmask = (u32)(kmask >> 32);              // FAST: 2k > 32
  const u32 rs = (u32)(rcsh - 32) & 31u, rc2 = 2u << rs;
  const u32 a_bcnt = hb_opaque32(hb_saddr(s_bcnt)), a_stage = hb_opaque32(hb_saddr(s_stage));
  const u32 a_gcnt = hb_opaque32(hb_saddr(s_gcnt)), a_gdst = hb_opaque32(hb_saddr(s_gdst));
  const u64 p4 = hb_opaque64(KMX_P4);
  const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(a.records);
  // persistent, in-order tickets over (window, tile) items numbered window-major (tile_pref = prefix of tiles per window)
  const u32 total = __ldg(a.tile_pref + a.nwin);
  u32 ywin = 0;                                        // thread 0: window of the last ticket (tickets only grow)
  // thread 0: describe ticket `item` in s_meta[buf] and start the bulk copy of its records into s_rec2[buf]
  auto fetch = [&](u32 item, u32 buf) {
    s_meta[buf][0] = item;
    if (item >= total) return;
    while (__ldg(a.tile_pref + ywin + 1) <= item) ywin++;
    const u32 n = __ldg(a.bcnt + ywin);
    const u32 tile0 = (item - __ldg(a.tile_pref + ywin)) * HB_TR;
    const u32 nt = min((u32)HB_TR, n - tile0);
    s_meta[buf][1] = ywin; s_meta[buf][2] = tile0; s_meta[buf][3] = nt;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer's earlier (generic) reads are ordered before the copy's writes
    const u32 mb = hb_saddr(&s_mbar[buf]);
    hb_mbar_expect_tx(mb, nt * 16u);
    hb_bulk_g2s(hb_saddr(&s_rec2[buf][0]), recs + __ldg(a.boff + ywin) + tile0, nt * 16u, mb);
  };
  if (tid == 0) {
    hb_mbar_init(hb_saddr(&s_mbar[0]), 1); hb_mbar_init(hb_saddr(&s_mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fetch(atomicAdd(a.tickets, 1u), 0);
  }
  for (u32 b = tid; b < NB; b += HB_THREADS) s_bcnt[b] = 0;
  u32 nxt = 0;
  for (u32 it = 0;; it++) {
    const u32 buf = it & 1u;
    __syncthreads();                               // previous tile fully consumed, this tile's s_meta visible
    if (s_meta[buf][0] >= total) break;
    const u32 y = s_meta[buf][1];
    const u32 nt = s_meta[buf][3];
    const uint4* __restrict__ s_rec = s_rec2[buf];
    if (tid < 64) s_cnt[tid] = 0;
    if (tid == 0) nxt = atomicAdd(a.tickets, 1u);   // next ticket: its latency hides behind the counting sort
    __syncthreads();
    hb_mbar_wait(hb_saddr(&s_mbar[buf]), (it >> 1) & 1u);      // the tile's records have landed
    // ---- counting sort by k-mers per record
    u32 nkr[HB_PER], rank[HB_PER];
#pragma unroll
    for (int i = 0; i < HB_PER; i++) {
      const u32 r = tid + HB_THREADS * i;
      nkr[i] = 0; rank[i] = 0;
      if (r < nt) {
        nkr[i] = ((s_rec[r].w >> 24) - (u32)k + 1u) & 63u;
        rank[i] = atomicAdd(&s_cnt[nkr[i]], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {
      const u32 c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
      u32 x = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, x, o); if (tid >= (u32)o) x += t; }
      s_start[2 * tid] = x - c0 - c1; s_start[2 * tid + 1] = x - c1;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HB_PER; i++) {
      const u32 r = tid + HB_THREADS * i;
      if (r < nt) s_perm[s_start[nkr[i]] + rank[i]] = (uint16_t)r;
    }
    const u32 ngroups = (nt + 31u) >> 5;
    __syncthreads();
    if (tid == 0) {                                 // groups are packed greedily into rounds of <= HB_TPAD hashed slots
      u32 r = 0, used = 0;
      for (u32 g = 0; g < ngroups; g++) {
        const uint4 v = s_rec[s_perm[min(g * 32u + 31u, nt - 1u)]];     // the group's longest record (ascending order)
        const u32 gm = ((v.w >> 24) - (u32)k + 1u) & 63u;
        s_gmax[g] = gm;
        const u32 sz = 32u * gm;
        if (used + sz > HB_TPAD) { r++; used = 0; }
        s_ground[g] = r; used += sz;
      }
      s_nrounds = r + 1;
      fetch(nxt, buf ^ 1u);                         // the next tile's records travel while this tile is hashed
    }
    __syncthreads();
    const u32 nrounds = s_nrounds;
    const u64 wbase = __ldg(a.win_base + y);
    const u32 wcap = __ldg(a.win_cap + y);
    uint16_t* __restrict__ wp = a.binbuf + wbase;
    u32* __restrict__ wcur = a.bin_cursor + (u64)y * NB;
    for (u32 rd = 0; rd < nrounds; rd++) {
      // ---- hash: lane = record, forward k-mer and reverse complement rolled base by base; the offset goes straight
      //      to slot (bin, rank) of the staging array, rank from ONE shared-memory atomic
      for (u32 g = w; g < ngroups; g += HB_WARPS) {
        if (s_ground[g] != rd) continue;
        const u32 idx = g * 32u + lane;
        u32 nk = 0; u64 f = 0, rc = 0, tlo = 0, thi = 0;
        if (idx < nt) {
          const uint4 v = s_rec[s_perm[idx]];
          const u64 lo = (u64)v.x | ((u64)v.y << 32);
          const u64 hh = (u64)v.z | ((u64)v.w << 32);
          const int nb = (int)(hh >> 56);
          const u64 hi = hh & 0x00FFFFFFFFFFFFFFULL;
          nk = (u32)(nb - k + 1);
          f = rec1_kmer(lo, hi, nb, k, 0);
          rc = revcomp64(f, k);
          const int tb = 2 * (nb - k);                     // the nb-k bases after the first k-mer, aligned to the top of (thi:tlo)
          if (FAST) { thi = tb ? (lo << (64 - tb)) : 0ULL; }
          else {
            const int sh = 128 - tb;                       // 128 >= sh > 0
            if (sh >= 128) { thi = 0; tlo = 0; }
            else if (sh >= 64) { thi = lo << (sh - 64); tlo = 0; }
            else { thi = (hi << sh) | (lo >> (64 - sh)); tlo = lo << sh; }
          }
        }
        const u32 gmax = s_gmax[g];
        if (FAST) {
          u32 fl = (u32)f, fh = (u32)(f >> 32), rl = (u32)rc, rh = (u32)(rc >> 32), tl = (u32)thi, th = (u32)(thi >> 32);
          for (u32 j = 0; j < gmax; j++) {
            const bool lt = (((u64)fh << 32) | fl) < (((u64)rh << 32) | rl);
            const u32 key = hb_hash_mod(lt ? fl : rl, lt ? fh : rh, ml, mh, md, p4);
            if (j < nk) {
              const u32 bin = key >> bs_log;
              const u32 r = hb_atoms_inc(a_bcnt + bin * 4u);
              if (r < C) hb_sts16(a_stage + (bin * C + r) * 2u, key & bsmask);
              else {                                       // a crowded bin (repeats inside one tile): straight to the bin's region
                const u32 g1 = atomicAdd(wcur + bin, 1u);
                if (g1 < wcap) wp[(u64)bin * wcap + g1] = (uint16_t)(key & bsmask);
                else a.flags[2] = 1u;
              }
            }
            const u32 b = th >> 30;
            th = __funnelshift_l(tl, th, 2); tl <<= 2;
            fh = __funnelshift_l(fl, fh, 2) & khmask; fl = (fl << 2) | b;
            rl = __funnelshift_r(rl, rh, 2); rh = (rh >> 2) | ((b << rs) ^ rc2);
          }
        } else {
          for (u32 j = 0; j < gmax; j++) {
            const u64 c = f < rc ? f : rc;
            const u32 key = hb_hash_mod((u32)c, (u32)(c >> 32), ml, mh, md, p4);
            if (j < nk) {
              const u32 bin = key >> bs_log;
              const u32 r = hb_atoms_inc(a_bcnt + bin * 4u);
              if (r < C) hb_sts16(a_stage + (bin * C + r) * 2u, key & bsmask);
              else {
                const u32 g1 = atomicAdd(wcur + bin, 1u);
                if (g1 < wcap) wp[(u64)bin * wcap + g1] = (uint16_t)(key & bsmask);
                else a.flags[2] = 1u;
              }
            }
            const u64 b = thi >> 62;
            thi = (thi << 2) | (tlo >> 62); tlo <<= 2;
            f = ((f << 2) | b) & kmask;
            rc = (rc >> 2) | ((b ^ 2ULL) << rcsh);
          }
        }
      }
      __syncthreads();
      // ---- one global reservation per (tile, bin); the counters are cleared for the next round / tile
      for (u32 b = tid; b < NB; b += HB_THREADS) {
        u32 c = min(s_bcnt[b], C);
        if (c) {
          const u32 g1 = atomicAdd(wcur + b, c);
          const u32 room = g1 < wcap ? wcap - g1 : 0u;
          if (c > room) { c = room; a.flags[2] = 1u; }       // the cursor keeps counting; the host retries with more room
          s_gdst[b] = b * wcap + g1;
        }
        s_gcnt[b] = c; s_bcnt[b] = 0;
      }
      __syncthreads();
      // ---- append every bin's run to its region (coalesced 2-byte runs)
      for (u32 b = w; b < NB; b += HB_WARPS) {
        const u32 c = hb_lds32(a_gcnt + b * 4u);
        const u32 d = hb_lds32(a_gdst + b * 4u);
        const u32 sa = a_stage + b * C * 2u;
#pragma unroll 1
        for (u32 j = lane; j < c; j += 32u) wp[d + j] = (uint16_t)hb_lds16(sa + j * 2u);
      }
      // (no barrier: the next round's / tile's staging writes are ordered behind the barriers that follow)
      if (rd + 1 < nrounds) __syncthreads();
    }
  }
}

// ---- pass B ------------------------------------------------------------------------------------
static constexpr int HC2_THREADS = 256;
static constexpr int HC2_WARPS = HC2_THREADS / 32;
static constexpr u32 HC2_WORDS = 16384;                         // 64 KB histogram per CTA ...
static constexpr u32 HC2_CHUNK = HC2_WORDS / HC2_THREADS;       // ... scanned in per-thread chunks of 64 consecutive words,
static constexpr u32 HC2_WORDS_P = HC2_WORDS + HC2_WORDS / 64;  // ... stored with one pad word per 64 so those chunk reads are conflict-free
static constexpr int HC2_PRE = 10;                              // 16-byte loads in flight per thread before the histogram is zeroed
static constexpr u64 LB_AGG = 1ULL << 62, LB_INCL = 2ULL << 62, LB_VAL = (1ULL << 62) - 1ULL;

template <bool H16>
__device__ __forceinline__ void hc2_add8(u32* __restrict__ s_h, const uint4& q, u32 base, u32 n)
{
  const u32 wv[4] = {q.x, q.y, q.z, q.w};
  if (base + 8u <= n) {
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const u32 off = (e & 1) ? (wv[e >> 1] >> 16) : (wv[e >> 1] & 0xFFFFu);
      if (H16) { const u32 wd = off >> 1; atomicAdd(&s_h[wd + (wd >> 6)], 1u << (16u * (off & 1u))); }
      else atomicAdd(&s_h[off + (off >> 6)], 1u);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; e++) {
      if (base + e < n) {
        const u32 off = (e & 1) ? (wv[e >> 1] >> 16) : (wv[e >> 1] & 0xFFFFu);
        if (H16) { const u32 wd = off >> 1; atomicAdd(&s_h[wd + (wd >> 6)], 1u << (16u * (off & 1u))); }
        else atomicAdd(&s_h[off + (off >> 6)], 1u);
      }
    }
  }
}

template <bool H16>
__global__ void __launch_bounds__(HC2_THREADS, 3)
hash_bincount_kernel(HashBinArgs a)
{
  extern __shared__ __align__(16) u32 s_h[];                    // [HC2_WORDS_P]
  __shared__ u32 s_wtot[HC2_WARPS], s_wfs[HC2_WARPS];
  __shared__ u64 s_excl;
  const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  // CTAs are dispatched in block-index order, so every predecessor a CTA looks back at is running or done
  const u32 item = blockIdx.x;
  const u32 v = item / a.NB, b = item - v * a.NB;
  const u32 cap = __ldg(a.win_cap + v);
  const u32 n = min(a.bin_cursor[item], cap);
  const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(a.binbuf + __ldg(a.win_base + v) + (u64)b * cap);
  uint4 q[HC2_PRE];
#pragma unroll
  for (int i = 0; i < HC2_PRE; i++) {                            // the bin's offsets are on their way while the histogram is zeroed
    const u32 g = tid + HC2_THREADS * i;
    q[i] = make_uint4(0, 0, 0, 0);
    if (g * 8u < n) q[i] = __ldcs(src4 + g);
  }
  {
    uint4* h4 = reinterpret_cast<uint4*>(s_h);
    for (u32 i = tid; i < HC2_WORDS_P / 4; i += HC2_THREADS) h4[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < HC2_PRE; i++) {
    const u32 g = tid + HC2_THREADS * i;
    if (g * 8u < n) hc2_add8<H16>(s_h, q[i], g * 8u, n);
  }
  for (u32 g = tid + HC2_THREADS * HC2_PRE; g * 8u < n; g += HC2_THREADS) hc2_add8<H16>(s_h, __ldcs(src4 + g), g * 8u, n);
  __syncthreads();
  // ---- survivors: thread t owns words [64 t, 64 t + 64) = slots in ascending order; their positions stay in registers
  const u32 hmin = a.hard_min;
  const u32* __restrict__ hw = s_h + tid * (HC2_CHUNK + 1u);
  constexpr int NM = H16 ? 4 : 2;                                // 32-slot survivor masks per thread
  u32 m[NM];
  u32 fsum = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) m[i] = 0;
#pragma unroll
  for (u32 j = 0; j < HC2_CHUNK; j++) {
    const u32 x = hw[j];
    if (H16) {
      const u32 f0 = x & 0xFFFFu, f1 = x >> 16;
      m[j >> 4] |= ((u32)(f0 >= hmin) | ((u32)(f1 >= hmin) << 1)) << (2u * (j & 15u));
      fsum = __dp2a_lo(x, 0x0101u, fsum);                        // + f0 + f1
    } else m[j >> 5] |= (u32)(x >= hmin) << (j & 31u);
  }
  u32 cnt = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) cnt += __popc(m[i]);
  u32 incl = cnt;                                                // inclusive scan over the warp's lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
  if (lane == 31) { s_wtot[w] = incl; s_wfs[w] = fsum; }
  __syncthreads();
  if (w == 0) {
    u32 total = 0, fs = 0;
#pragma unroll
    for (int i = 0; i < HC2_WARPS; i++) { total += s_wtot[i]; fs += s_wfs[i]; }
    if (H16 && lane == 0 && fs != n) a.flags[1] = 1u;            // a 16-bit counter wrapped: the sample is redone with 32-bit counters
    // decoupled look-back over the survivor counts of the preceding (window, bin) items
    volatile u64* st = a.status;
    u64 excl = 0;
    if (item == 0) { if (lane == 0) st[0] = LB_INCL | (u64)total; }
    else {
      if (lane == 0) st[item] = LB_AGG | (u64)total;
      int look = (int)item - 1;
      for (;;) {
        const int idx = look - (int)lane;
        u64 s = LB_INCL;                                         // before item 0: inclusive prefix 0
        if (idx >= 0) { do { s = st[idx]; } while ((s >> 62) == 0ULL); }
        const u32 inc = __ballot_sync(0xffffffffu, (s >> 62) == 2ULL);
        u64 val = s & LB_VAL;
        if (inc && lane > (u32)(__ffs(inc) - 1)) val = 0;       // stop at the nearest inclusive prefix
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (inc) break;
        look -= 32;
      }
      if (lane == 0) st[item] = LB_INCL | (excl + (u64)total);
    }
    if (lane == 0) {
      s_excl = excl;
      if (b == 0) a.list_off[v] = excl;
      if (excl + total > a.meta[2]) a.flags[0] = 1u;             // output space ran out: the host retries with the exact size
      if (item == a.nwin * a.NB - 1u) a.meta[0] = excl + total;
    }
  }
  __syncthreads();
  // ---- emit (key, count) in slot order straight from shared memory
  if (cnt) {
    u64 r = s_excl + (incl - cnt);
    for (u32 i = 0; i < w; i++) r += s_wtot[i];
    const u64 ocap = a.meta[2];
    const u64 kb = (u64)(a.win_part ? a.win_part[v] : v) * a.Wbits + ((u64)b << a.bs_log) + (u64)tid * HC2_CHUNK * (H16 ? 2u : 1u);
#pragma unroll
    for (int i = 0; i < NM; i++) {
      u32 mm = m[i];
      while (mm) {
        const u32 slot = (u32)i * 32u + (u32)__ffs(mm) - 1u;
        mm &= mm - 1u;
        const u32 x = hw[H16 ? (slot >> 1) : slot];
        const u32 f = H16 ? ((slot & 1u) ? (x >> 16) : (x & 0xFFFFu)) : x;
        if (r < ocap) { a.out_keys[r] = kb + slot; a.out_counts[r] = f; }
        r++;
      }
    }
  }
}

// phase 0: pass A over all windows; phase 1: pass B.  The caller zeroes bin_cursor / tickets / flags / status first.
cudaError_t launch_hash_binned(const HashBinArgs& a, u32 total_tiles, int phase, bool h16, cudaStream_t st, u64* launches)
{
  if (phase == 0) {
    if (!total_tiles) return cudaSuccess;
    FastMod32 f32; f32.d = (u32)a.Wbits; f32.m64 = (~0ULL) / a.Wbits;
    const bool fast = 2 * (KMX_REC1_MAXN - a.k) <= 64 && a.k <= 32;     // 28 <= k <= 32
    const unsigned grid = (unsigned)std::min<u64>(total_tiles, (u64)148 * 4);
    if (a.NB <= 128) {
      if (fast) hash_bin_kernel<true, 128><<<grid, HB_THREADS, 0, st>>>(a, f32);
      else hash_bin_kernel<false, 128><<<grid, HB_THREADS, 0, st>>>(a, f32);
    } else {
      if (fast) hash_bin_kernel<true, 512><<<grid, HB_THREADS, 0, st>>>(a, f32);
      else hash_bin_kernel<false, 512><<<grid, HB_THREADS, 0, st>>>(a, f32);
    }
    *launches += 1;
  } else {
    const u64 items = (u64)a.nwin * a.NB;
    if (!items || items >= 0x7FFFFFF0ULL) return items ? cudaErrorInvalidValue : cudaSuccess;
    const size_t smem = (size_t)HC2_WORDS_P * 4;
    // the attribute is set once per device and process (not per launch: several lanes launch concurrently)
    static std::mutex attr_mu; static bool attr_done[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> g(attr_mu);
      if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(hash_bincount_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(hash_bincount_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
      }
    }
    if (h16) hash_bincount_kernel<true><<<(unsigned)items, HC2_THREADS, smem, st>>>(a);
    else hash_bincount_kernel<false><<<(unsigned)items, HC2_THREADS, smem, st>>>(a);
    *launches += 1;
  }
  return cudaGetLastError();
}

u32 hash_bin_tile_records() { return HB_TR; }
u32 hash_bin_max_bins() { return 512; }

}  // namespace kmx
