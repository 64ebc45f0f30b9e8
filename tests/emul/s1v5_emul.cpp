// s1v5_emul.cpp -- runs the phase functions of kmtricks_b200/csrc/s1_v5.cuh (the bodies of the
// position-parallel stage-1 kernel) on the CPU, one CTA at a time, phase by phase, and compares
// the (partition, canonical k-mer) multiset decoded from the emitted records with the oracle's
// orc_s1_seq (oracle/kmx_oracle.c).  Test infrastructure only (built and run by
// tests/test_s1v5_emul.py); checks the index arithmetic of the kernel without a GPU.
#include "../../kmtricks_b200/csrc/s1_v5.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <tuple>
#include <vector>

extern "C" {
void orc_minim_lut(int m, uint32_t* lut);
void orc_repart_static(int m, uint32_t P, uint16_t* table);
size_t orc_s1_seq(const char* seq, size_t len, int k, int m, const uint32_t* lut, const uint16_t* table,
                  uint16_t* part_out, uint64_t* canon_lo, uint64_t* canon_hi);
}

using namespace kmx::s1v5;
typedef std::tuple<uint32_t, uint64_t, uint64_t> Key;   // partition, canon hi, canon lo

static int base_of(const uint32_t* v, int NW, uint32_t nb, uint32_t i)   // code of base i of a record
{
  const uint32_t bit = 2 * (nb - 1 - i);
  (void)NW;
  return (v[bit >> 5] >> (bit & 31)) & 3;
}

static void canon_of(const int* codes, int k, uint64_t& lo, uint64_t& hi)
{
  unsigned __int128 f = 0, r = 0;
  for (int i = 0; i < k; i++) f = (f << 2) | (unsigned)codes[i];
  for (int i = k - 1; i >= 0; i--) r = (r << 2) | (unsigned)(codes[i] ^ 2);
  unsigned __int128 c = f < r ? f : r;
  lo = (uint64_t)c; hi = (uint64_t)(c >> 64);
}

int main(int argc, char** argv)
{
  if (argc < 7) { fprintf(stderr, "usage: k m P nreads maxlen seed [R]\n"); return 2; }
  const int k = atoi(argv[1]), m = atoi(argv[2]); const uint32_t P = (uint32_t)atoi(argv[3]);
  const int nreads = atoi(argv[4]); const uint32_t maxlen_req = (uint32_t)atoi(argv[5]); const unsigned seed = (unsigned)atoi(argv[6]);
  const uint32_t R = argc > 7 ? (uint32_t)atoi(argv[7]) : 32u;
  const int W = (k + 31) / 32, NW = 4 * W;
  const uint32_t max_nk = (uint32_t)((W == 1 ? 60 : 124) - k + 1);

  std::vector<uint32_t> lut((size_t)1 << (2 * m));
  std::vector<uint16_t> table((size_t)1 << (2 * m));
  orc_minim_lut(m, lut.data());
  orc_repart_static(m, P, table.data());

  // reads: pieces of a random genome (so minimizers repeat), some homopolymer / low-complexity stretches,
  // invalid characters, lower case, lengths around k and up to maxlen, arbitrary alignment in the text
  std::mt19937 rng(seed);
  std::string genome(5000, 'A');
  for (auto& ch : genome) ch = "ACGT"[rng() & 3];
  for (int j = 0; j < 20; j++) { size_t p = rng() % (genome.size() - 100); int l = 20 + rng() % 60; char c = "ACGT"[rng() & 3]; for (int i = 0; i < l; i++) genome[p + i] = (rng() % 10) ? c : "ACGT"[rng() & 3]; }
  std::string text;
  std::vector<uint32_t> st, ln;
  for (int i = 0; i < nreads; i++) {
    uint32_t len;
    switch (rng() % 8) {
      case 0: len = (uint32_t)(k - 2 + rng() % 5); break;           // around k (some shorter than k)
      case 1: len = maxlen_req; break;
      case 2: len = 1 + rng() % maxlen_req; break;
      default: len = maxlen_req > 10 ? maxlen_req - rng() % 10 : maxlen_req;
    }
    if (len > maxlen_req) len = maxlen_req;
    if (len < 1) len = 1;
    std::string rd = genome.substr(rng() % (genome.size() - len), len);
    for (auto& ch : rd) if (rng() % 50 == 0) ch = "ACGT"[rng() & 3];
    const int kind = rng() % 6;
    if (kind == 0) for (int j = 0; j < 1 + (int)(rng() % 3); j++) rd[rng() % len] = "NnRY.-@"[rng() % 7];
    if (kind == 1) for (auto& ch : rd) ch = (char)(ch | 0x20);
    if (kind == 2) for (auto& ch : rd) if (rng() & 1) ch = (char)(ch | 0x20);
    text.append("@r\n", 1 + rng() % 3);                               // 1..3 bytes of junk: shifts the alignment
    st.push_back((uint32_t)text.size()); ln.push_back(len);
    text += rd;
    text += "\n+\nIIII\n";
  }
  const size_t nbytes = text.size();
  text.append(16, '\0');
  const uint8_t* tb = reinterpret_cast<const uint8_t*>(text.data());
  const uint32_t* wend = reinterpret_cast<const uint32_t*>((reinterpret_cast<uintptr_t>(tb) + nbytes + 3) & ~(uintptr_t)3);
  uint32_t maxlen = 0;
  for (uint32_t l : ln) maxlen = std::max(maxlen, l);
  if (maxlen < (uint32_t)k) { printf("OK (nothing to do)\n"); return 0; }

  // ---- oracle
  std::vector<Key> want;
  {
    std::vector<uint16_t> po(maxlen); std::vector<uint64_t> cl(maxlen), chh(maxlen);
    for (size_t i = 0; i < st.size(); i++) {
      size_t n = orc_s1_seq(text.data() + st[i], ln[i], k, m, lut.data(), table.data(), po.data(), cl.data(), chh.data());
      for (size_t j = 0; j < n; j++) want.emplace_back(po[j], W == 2 ? chh[j] : 0, cl[j]);
    }
  }

  // ---- emulated CTAs
  Cta x; x.k = k; x.m = m; x.w = k - m + 1; x.max_nk = max_nk;
  x.mmask = (uint32_t)(((uint64_t)1 << (2 * m)) - 1); x.ban_mask = 0x55555555u & ((1u << (2 * (m - 2))) - 1u);
  x.g = make_geo(R, maxlen, k, m);
  if (maxlen - k + 1 > x.g.evcap) { fprintf(stderr, "maxlen too large for the event queue\n"); return 2; }
  std::vector<uint32_t> smem(smem_bytes(x.g, P) / 4 + 16, 0xDEADBEEFu);   // garbage like real shared memory
  carve(x, smem.data(), P);
  std::vector<Key> got;
  std::vector<uint64_t> kcnt(P, 0), cursor(P, 0);
  uint64_t nrec = 0, nev_rounds = 0, nfull = 0;
  const size_t nseg = st.size();
  for (size_t cta = 0; cta * R < nseg; cta++) {
    std::fill(smem.begin(), smem.end(), 0xDEADBEEFu);
    for (uint32_t p = 0; p < P; p++) { x.hist[p] = 0; x.kc[p] = 0; }
    for (uint32_t r = 0; r < R; r++) {
      size_t seg = cta * R + r;
      uint32_t len = seg < nseg ? ln[seg] : 0;
      if (len < (uint32_t)k) len = 0;
      x.len[r] = len; x.inval[r] = 0;
    }
    // P0
    for (uint32_t task = 0; task < R * x.g.nch; task++) {
      uint32_t r = task / x.g.nch, c = task % x.g.nch; size_t seg = cta * R + r;
      p0_pack(x, r, c, tb + (seg < nseg ? st[seg] : 0), x.len[r], wend);
    }
    // P1
    for (uint32_t r = 0; r < R; r++) { uint32_t len = x.len[r]; if (!len) continue; for (uint32_t lane = 0; lane < 32; lane++) p1_row(x, r, lane, len - m + 1); }
    // P2
    for (uint32_t g = 0; g < x.g.nblk; g++) for (uint32_t r = 0; r < R; r++) { if (x.w == 22) p2_block<22>(x, r, g, x.len[r]); else p2_block<0>(x, r, g, x.len[r]); }
    // P3 rounds: every pending item (r, g) prepares (mask bit 0, look-ahead, count), takes its slots with an atomic add on the
    // CTA's event counter and emits if they fit the queue; otherwise it pads what it took and waits for the next round. P4 follows.
    const uint32_t ntask = R * x.g.nblk;                       // item = r * nblk + g
    std::vector<uint32_t> cnt(ntask, 0);
    for (uint32_t t = 0; t < ntask; t++) x.done[t] = 0;
    uint32_t pending = ntask;
    while (pending) {
      uint32_t s_nev = 0, progressed = 0;
      for (uint32_t t = 0; t < ntask; t++) {
        if (x.done[t]) continue;
        const uint32_t r = item_read(x.g, t);
        if (r != t / x.g.nblk) { fprintf(stderr, "item_read(%u) = %u\n", t, r); return 1; }
        const uint32_t g = t % x.g.nblk;
        const uint32_t n = p3_prepare(x, r, g, x.len[r]);
        cnt[t] = n;
        const uint32_t s0 = s_nev; s_nev += n;                  // atomicAdd
        if (s0 + n <= x.g.evcap) {
          if (n) {
            if (!x.inval[r]) p3_emit_item(x, r, g, s0);
            else if (p3_slow<true>(x, r, x.len[r], s0) != n) { fprintf(stderr, "count/emit mismatch read %u\n", r); return 1; }
          }
          x.done[t] = 1; pending--; progressed++;
        } else {
          for (uint32_t q = s0; q < x.g.evcap; q++) { x.ev[q].x = 0; x.ev[q].y = 0; }   // null events: skipped by the flush
        }
      }
      if (!progressed) { fprintf(stderr, "no item fits the event queue\n"); return 1; }
      const uint32_t acc = std::min(s_nev, x.g.evcap);
      nev_rounds++;
      for (uint32_t q = 0; q < acc; q++) {
        const Ev e = x.ev[q];
        if (!(e.x >> 19)) continue;                              // padding of an item that did not fit
        const uint32_t rd = e.x & 127u, iend = (e.x >> 7) & 4095u, nkr = (e.x >> 19) & 127u;
        if (e.y > x.mmask) { fprintf(stderr, "minimizer out of range\n"); return 1; }
        const uint32_t p = table[e.y];
        const uint32_t nb = (uint32_t)k + nkr - 1u;
        if (nkr == 0 || nkr > max_nk || iend > x.len[rd] || iend < nb) { fprintf(stderr, "bad event rd=%u iend=%u nk=%u len=%u\n", rd, iend, nkr, x.len[rd]); return 1; }
        uint32_t v[8];
        if (W == 1) build_record<4>(x.BE + rd * x.g.LW, x.g.nch, iend, nb, v); else build_record<8>(x.BE + rd * x.g.LW, x.g.nch, iend, nb, v);
        if ((v[NW - 1] >> 24) != nb) { fprintf(stderr, "length byte\n"); return 1; }
        v[NW - 1] &= 0x00FFFFFFu;
        std::vector<int> codes(nb);
        for (uint32_t i = 0; i < nb; i++) codes[i] = base_of(v, NW, nb, i);
        // the record must be the read's bases [iend-nb, iend)
        size_t seg = cta * R + rd;
        for (uint32_t i = 0; i < nb; i++) {
          int want_c = (tb[st[seg] + iend - nb + i] >> 1) & 3;
          if (codes[i] != want_c) { fprintf(stderr, "record bases differ from the read (cta %zu read %u base %u)\n", cta, rd, i); return 1; }
        }
        for (uint32_t bitpos = 2 * nb; bitpos < (uint32_t)(32 * NW - 8); bitpos++) if ((v[bitpos >> 5] >> (bitpos & 31)) & 1) { fprintf(stderr, "stray bits above the record\n"); return 1; }
        for (uint32_t j = 0; j < nkr; j++) { uint64_t lo, hi; canon_of(codes.data() + j, k, lo, hi); got.emplace_back(p, hi, lo); }
        kcnt[p] += nkr; cursor[p]++; nrec++; if (nkr == max_nk) nfull++;
      }
    }
  }
  std::sort(want.begin(), want.end()); std::sort(got.begin(), got.end());
  if (want != got) {
    fprintf(stderr, "MISMATCH: oracle %zu k-mers, emulated kernel %zu\n", want.size(), got.size());
    return 1;
  }
  printf("OK k=%d m=%d P=%u reads=%d maxlen=%u kmers=%zu records=%llu (%.2f k-mers/record) flush rounds=%llu full records=%llu\n", k, m, P, nreads, maxlen, got.size(),
         (unsigned long long)nrec, nrec ? (double)got.size() / (double)nrec : 0.0, (unsigned long long)nev_rounds, (unsigned long long)nfull);
  return 0;
}
