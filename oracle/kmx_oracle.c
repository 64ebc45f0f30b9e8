/* kmx_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the kmtricks (v1.6.0 @9bccc774) repart -> superk -> count ->
 * merge (+ Bloom rows / bit transpose) hot path.  It is the *checker* for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The
 * product (kmtricks_b200/, libkmx_sm100.so) never links, imports or calls anything here.
 *
 * Parity pin: this restatement is validated (tests/test_oracle_*.py) against
 *   - the reference's own golden vectors (tests/task_main.cpp:59-507, tests/merge_test.cpp:5-78,
 *     tests/kmer_test.cpp:117-154, tests/repartition_test.cpp:7-18, tests/data fixtures) and
 *   - outputs of the unmodified reference binary built by oracle/build_ref.sh (tests/golden/).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference;
 * "gatb/" = thirdparty/gatb-core-stripped/src/gatb/).  Written from the behaviour, not
 * copied: scalar loops, no super-k-mers, no kx-mers, no radix bins.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* XXH64 -- third-party dependency thirdparty/xxHash v0.8.3 (xxhash.h:3454-3673); restated
 * from the published XXH64 specification (all input lengths).  Call sites on the path:
 * include/kmtricks/repartition.hpp:52 (len 4) and gatb/sorting_count.hpp:356 (len 8w).   */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl64(acc, 31); return acc * P1; }
static inline uint64_t xmerge(uint64_t h, uint64_t v) { v = xround(0, v); h ^= v; return h * P1 + P4; }

uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed)
{
  const uint8_t* p = (const uint8_t*)data;
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* lim = end - 32;
    do {
      v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8));
      v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
  while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; p++; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

/* ------------------------------------------------------------------------------------ */
/* Nucleotide code and validity: gatb/tools/misc/api/Data.hpp:179 ((c>>1)&3; A0 C1 T2 G3),
 * validity table Data.hpp:183-196 (only ACGTacgt valid).                                 */
static inline int nt_code(unsigned char c) { return (c >> 1) & 3; }
static inline int nt_valid(unsigned char c)
{
  switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; default: return 0; }
}

/* revcomp of an m-mer held in the low 2m bits; complement = code ^ 2
 * (gatb/tools/math/LargeInt1.pri:137-158; comp_NT table gatb/kmer/impl/Model.hpp:415).     */
static uint64_t revcomp_small(uint64_t x, int n)
{
  uint64_t r = 0;
  for (int i = 0; i < n; i++) { r = (r << 2) | ((x & 3) ^ 2); x >>= 2; }
  return r;
}

/* Minimizer LUT: gatb/kmer/impl/Model.hpp:1040-1064 (canonical m-mer, banned -> mask) with
 * is_allowed() Model.hpp:1220-1251 ("AA" anywhere except as the two leading bases).        */
void orc_minim_lut(int m, uint32_t* lut)
{
  uint64_t n = 1ULL << (2 * m), mask = n - 1;
  for (uint64_t x = 0; x < n; x++) {
    uint64_t c = revcomp_small(x, m);
    if (x < c) c = x;
    int banned = 0;
    for (int j = 0; j + 2 < m; j++)  /* pairs (nt_j, nt_{j+1}), nt_0 = last base, j in [0, m-3] */
      if (((c >> (2 * j)) & 0xF) == 0) { banned = 1; break; }
    lut[x] = (uint32_t)(banned ? mask : c);
  }
}

/* Static repartition table: include/kmtricks/repartition.hpp:45-56
 * table[x] = XXH64(&x as uint32 LE, 4, seed 0) % P.                                        */
void orc_repart_static(int m, uint32_t P, uint16_t* table)
{
  uint64_t n = 1ULL << (2 * m);
  for (uint64_t x = 0; x < n; x++) {
    uint32_t v = (uint32_t)x;
    table[x] = (uint16_t)(orc_xxh64(&v, 4, 0) % P);
  }
}

/* ------------------------------------------------------------------------------------ */
/* 128-bit k-mer helpers (k <= 64): (hi,lo) pair, first base most significant
 * (gatb/tools/math/LargeInt2.pri:30-160).                                                */
typedef struct { uint64_t lo, hi; } k128;
static inline k128 k_shl2_or(k128 v, uint64_t c, int k)
{
  k128 r;
  r.hi = (v.hi << 2) | (v.lo >> 62);
  r.lo = (v.lo << 2) | c;
  if (k < 32) { r.hi = 0; r.lo &= ((1ULL << (2 * k)) - 1); }
  else if (k == 32) { r.hi = 0; }
  else if (k < 64) { r.hi &= ((1ULL << (2 * (k - 32))) - 1); }
  return r;
}
static inline k128 k_shr2_orhigh(k128 v, uint64_t c, int k)
{ /* r = (v >> 2) | (c << 2(k-1)) */
  k128 r;
  r.lo = (v.lo >> 2) | (v.hi << 62);
  r.hi = v.hi >> 2;
  int sh = 2 * (k - 1);
  if (sh < 64) r.lo |= c << sh; else r.hi |= c << (sh - 64);
  return r;
}
static inline int k_less(k128 a, k128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

/* Stage 1 for one sequence (already joined, no newlines).
 * Follows gatb/kmer/impl/Model.hpp:725-765 (iterate + invalid window), :857-884 (rolling
 * fwd/rev, canonical), :1254-1287 (minimizer = plain min of lut over the k-m+1 m-mers of the
 * forward k-mer), gatb/kmer/impl/Sequence2SuperKmer.hpp:137-147 (skip len<k),
 * include/kmtricks/gatb/fill_partitions.hpp:59-63 (p = repart[minimizer]).
 * Super-k-mer cuts are ignored: a k-mer's partition is a pure function of the k-mer (F6).
 * Emits, for every VALID k-mer in order: partition, canonical (lo,hi).  Returns the count.  */
size_t orc_s1_seq(const char* seq, size_t len, int k, int m,
                  const uint32_t* lut, const uint16_t* table,
                  uint16_t* part_out, uint64_t* canon_lo, uint64_t* canon_hi)
{
  if (len < (size_t)k) return 0;
  k128 f = {0, 0}, r = {0, 0};
  uint64_t mmask = (1ULL << (2 * m)) - 1;
  int bad = 0;
  size_t n = 0;
  for (size_t i = 0; i < len; i++) {
    unsigned char ch = (unsigned char)seq[i];
    uint64_t c = (uint64_t)nt_code(ch);
    f = k_shl2_or(f, c, k);
    r = k_shr2_orhigh(r, c ^ 2, k);
    if (nt_valid(ch)) { if (bad > 0) bad--; } else bad = k;
    if (i + 1 >= (size_t)k && bad == 0) {
      k128 canon = k_less(f, r) ? f : r;
      /* plain minimum over the k-m+1 m-mers of the forward k-mer */
      uint32_t best = (uint32_t)mmask;
      k128 v = f;
      for (int j = 0; j <= k - m; j++) {
        uint32_t cand = lut[v.lo & mmask];
        if (cand < best) best = cand;
        v.lo = (v.lo >> 2) | (v.hi << 62); v.hi >>= 2;
      }
      part_out[n] = table[best];
      canon_lo[n] = canon.lo;
      if (canon_hi) canon_hi[n] = canon.hi;
      n++;
    }
  }
  return n;
}

/* Minimizer value of one k-mer given as forward (lo,hi); for the KAT in
 * tests/kmer_test.cpp:117-154.                                                             */
uint32_t orc_minimizer_of(uint64_t lo, uint64_t hi, int k, int m, const uint32_t* lut)
{
  uint64_t mmask = (1ULL << (2 * m)) - 1;
  uint32_t best = (uint32_t)mmask;
  for (int j = 0; j <= k - m; j++) {
    uint32_t cand = lut[lo & mmask];
    if (cand < best) best = cand;
    lo = (lo >> 2) | (hi << 62); hi >>= 2;
  }
  return best;
}

/* ------------------------------------------------------------------------------------ */
/* FASTA/FASTQ record reader, kseq-style: gatb/bank/impl/BankFasta.cpp:391-560.
 * Sequence lines are joined; a trailing '\r' is dropped from lines longer than 1; FASTQ
 * quality is skipped until it is at least as long as the sequence.
 * Writes joined sequences back to back into seq_out and record offsets into off_out
 * (off_out[i]..off_out[i+1]).  Returns number of records (<= max_rec).                     */
size_t orc_fastx_parse(const char* buf, size_t n, char* seq_out, uint64_t* off_out, size_t max_rec)
{
  size_t pos = 0, nrec = 0; uint64_t o = 0;
  int last_char = 0;
  off_out[0] = 0;
  while (nrec < max_rec) {
    if (last_char == 0) {
      while (pos < n && buf[pos] != '>' && buf[pos] != '@') pos++;
      if (pos >= n) break;
      last_char = buf[pos]; pos++;
    }
    /* header: rest of line */
    if (pos >= n) break;
    while (pos < n && buf[pos] != '\n') pos++;
    if (pos < n) pos++;
    /* sequence lines */
    int c = -1;
    uint64_t start = o;
    while (pos < n) {
      c = (unsigned char)buf[pos++];
      if (c == '>' || c == '+' || c == '@') break;
      if (c == '\n') { c = -1; continue; }
      seq_out[o++] = (char)c;
      size_t ls = pos;
      while (pos < n && buf[pos] != '\n') pos++;
      memcpy(seq_out + o, buf + ls, pos - ls); o += pos - ls;
      if (pos < n) pos++;
      /* buffered_gets(allow_spaces): drop trailing '\r' if accumulated length > 1 */
      if (o - start > 1 && seq_out[o - 1] == '\r') o--;
      c = -1;
    }
    if (c == '>' || c == '@') last_char = c;
    if (c == '+') {
      while (pos < n && buf[pos] != '\n') pos++;   /* rest of the '+' line */
      if (pos < n) pos++;
      uint64_t qlen = 0, slen = o - start;
      /* quality lines until qlen >= slen (at least one line is always consumed) */
      while (pos < n) {
        size_t ls = pos;
        while (pos < n && buf[pos] != '\n') pos++;
        uint64_t l = pos - ls;
        if (pos < n) pos++;
        qlen += l;
        if (qlen > 1 && l > 0 && buf[ls + l - 1] == '\r') qlen--;
        if (qlen >= slen) break;
      }
      last_char = 0;
    }
    nrec++;
    off_out[nrec] = o;
    if (pos >= n) break;
  }
  return nrec;
}

/* ------------------------------------------------------------------------------------ */
/* Hash-mode key: gatb/sorting_count.hpp:346-363  XXH64(canonical words, 8w, 0) % W + W*p,
 * for w=2 the bytes are [lo u64][hi u64] little-endian (gatb/tools/math/LargeInt2.pri:151-154) */
uint64_t orc_hash_key(uint64_t lo, uint64_t hi, int w, uint64_t W, uint64_t p)
{
  uint64_t words[2] = {lo, hi};
  return orc_xxh64(words, (size_t)(8 * w), 0) % W + W * p;
}

/* Stage 2: sort + run-length + hard-min + saturate.
 * include/kmtricks/gatb/sorting_count.hpp:637-650,694-884 (k-mer), :934-990 (hash);
 * include/kmtricks/gatb/count_processor.hpp:61-70,135-146 (count >= hard_min, saturate u32).
 * Keys are (hi,lo) pairs (hi may be NULL for w=1).  In place: sorted distinct survivors are
 * written to out_*; returns their number.                                                  */
typedef struct { uint64_t hi, lo; } keypair;
static int cmp_keypair(const void* a, const void* b)
{
  const keypair* x = (const keypair*)a; const keypair* y = (const keypair*)b;
  if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
  if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
  return 0;
}
static int cmp_u64(const void* a, const void* b)
{
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
size_t orc_s2_count(const uint64_t* lo, const uint64_t* hi, size_t n, uint32_t hard_min,
                    uint64_t* out_lo, uint64_t* out_hi, uint32_t* out_count)
{
  if (n == 0) return 0;
  size_t m = 0;
  if (!hi) {
    uint64_t* t = (uint64_t*)malloc(n * sizeof(uint64_t));
    memcpy(t, lo, n * sizeof(uint64_t));
    qsort(t, n, sizeof(uint64_t), cmp_u64);
    size_t i = 0;
    while (i < n) {
      size_t j = i; while (j < n && t[j] == t[i]) j++;
      uint64_t c = j - i;
      if (c >= hard_min) { out_lo[m] = t[i]; out_count[m] = c > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)c; m++; }
      i = j;
    }
    free(t);
  } else {
    keypair* t = (keypair*)malloc(n * sizeof(keypair));
    for (size_t i = 0; i < n; i++) { t[i].hi = hi[i]; t[i].lo = lo[i]; }
    qsort(t, n, sizeof(keypair), cmp_keypair);
    size_t i = 0;
    while (i < n) {
      size_t j = i; while (j < n && t[j].hi == t[i].hi && t[j].lo == t[i].lo) j++;
      uint64_t c = j - i;
      if (c >= hard_min) { out_lo[m] = t[i].lo; out_hi[m] = t[i].hi; out_count[m] = c > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)c; m++; }
      i = j;
    }
    free(t);
  }
  return m;
}

/* ------------------------------------------------------------------------------------ */
/* Stage 3: N-way merge with soft-min, share-min rescue, recurrence-min and the six
 * statistics vectors.  include/kmtricks/merge.hpp:183-260 (KmerMerger::next), :441-517
 * (HashMerger::next), :49-100 (MergeStatistics).
 * Inputs: N ascending lists concatenated; list s = [off[s], off[s+1]).
 * Outputs: every merged row (emit_all != 0, the plugin case F11) or only kept rows:
 *   row_lo/row_hi, row_counts (N per row), row_keep (1 byte per emitted row).
 * stats = 6 vectors of N uint64 in the order NON_SOLID, RESCUED, UNIQUE_WO_RESCUE,
 * UNIQUE_W_RESCUE, TOTAL_WO_RESCUE, TOTAL_W_RESCUE.  Returns number of emitted rows;
 * *n_union receives the number of distinct keys.  Row arrays may be NULL to only count.     */
size_t orc_s3_merge(int w, uint32_t N, const uint64_t* off,
                    const uint64_t* lo, const uint64_t* hi, const uint32_t* cnt,
                    const uint32_t* soft_min, uint32_t r_min, uint32_t save_if, int emit_all,
                    uint64_t* row_lo, uint64_t* row_hi, uint32_t* row_counts, uint8_t* row_keep,
                    uint64_t* stats, uint64_t* n_union)
{
  uint64_t* head = (uint64_t*)malloc(N * sizeof(uint64_t));
  uint32_t* c = (uint32_t*)malloc(N * sizeof(uint32_t));
  uint8_t* chk = (uint8_t*)malloc(N);
  for (uint32_t s = 0; s < N; s++) head[s] = off[s];
  memset(stats, 0, 6 * (size_t)N * sizeof(uint64_t));
  uint64_t *ns = stats, *rd = stats + N, *uwo = stats + 2 * N, *uw = stats + 3 * N, *two = stats + 4 * N, *tw = stats + 5 * N;
  size_t nrows = 0; uint64_t nu = 0;
  for (;;) {
    int any = 0; uint64_t cl = 0, ch = 0;
    for (uint32_t s = 0; s < N; s++) if (head[s] < off[s + 1]) {
      uint64_t l = lo[head[s]], h = (w > 1) ? hi[head[s]] : 0;
      if (!any || h < ch || (h == ch && l < cl)) { cl = l; ch = h; any = 1; }
    }
    if (!any) break;
    nu++;
    uint32_t solid_in = 0;
    for (uint32_t s = 0; s < N; s++) {
      chk[s] = 0; c[s] = 0;
      if (head[s] < off[s + 1] && lo[head[s]] == cl && ((w > 1) ? hi[head[s]] : 0) == ch) {
        c[s] = cnt[head[s]];
        if (c[s] >= soft_min[s]) { solid_in++; two[s] += c[s]; tw[s] += c[s]; uwo[s]++; uw[s]++; }
        else { ns[s]++; if (save_if) chk[s] = 1; else c[s] = 0; }
        head[s]++;
      }
    }
    for (uint32_t s = 0; s < N; s++) if (chk[s]) {
      if (!(solid_in >= save_if)) c[s] = 0;
      else { rd[s]++; uw[s]++; tw[s] += c[s]; }
    }
    int keep = solid_in >= r_min;
    if (keep || emit_all) {
      if (row_lo) {
        row_lo[nrows] = cl; if (row_hi) row_hi[nrows] = ch;
        memcpy(row_counts + nrows * (size_t)N, c, N * sizeof(uint32_t));
        if (row_keep) row_keep[nrows] = (uint8_t)keep;
      }
      nrows++;
    }
  }
  if (n_union) *n_union = nu;
  free(head); free(c); free(chk);
  return nrows;
}

/* Presence/absence bytes of a row: include/kmtricks/utils.hpp:104-116 (set_bit_vector,
 * bit (s%8) of byte s/8 <=> counts[s] != 0).                                               */
void orc_pa_row(const uint32_t* counts, uint32_t N, uint8_t* out /* (N+7)/8 bytes */)
{
  memset(out, 0, (N + 7) / 8);
  for (uint32_t s = 0; s < N; s++) if (counts[s]) out[s >> 3] |= (uint8_t)(1u << (s & 7));
}

/* Dense Bloom slab of a partition: include/kmtricks/merge.hpp:575-600 (write_as_bf):
 * for h in [lower, lower+W): PA bytes of the kept row with key h, else zeros.              */
void orc_bf_slab(const uint64_t* row_key, const uint32_t* row_counts, size_t nrows, uint32_t N,
                 uint64_t lower, uint64_t W, uint8_t* slab /* W * (N+7)/8, zeroed here */)
{
  size_t rb = (N + 7) / 8;
  memset(slab, 0, (size_t)W * rb);
  for (size_t i = 0; i < nrows; i++)
    orc_pa_row(row_counts + i * (size_t)N, N, slab + (size_t)(row_key[i] - lower) * rb);
}

/* Bit-matrix transpose, definitional form of include/kmtricks/bitmatrix.hpp:209-214,238-289:
 * in = nrows x (ncols/8) bytes, out = ncols x (nrows/8) bytes, both LSB-first;
 * out[c][r] = in[r][c].  nrows, ncols multiples of 8.                                      */
void orc_transpose_bits(const uint8_t* in, size_t nrows, size_t ncols, uint8_t* out)
{
  size_t ib = ncols / 8, ob = nrows / 8;
  memset(out, 0, ncols * ob);
  for (size_t r = 0; r < nrows; r++)
    for (size_t c = 0; c < ncols; c++)
      if ((in[r * ib + (c >> 3)] >> (c & 7)) & 1) out[c * ob + (r >> 3)] |= (uint8_t)(1u << (r & 7));
}

/* Per-sample bit vector ("count --mode vector"): include/kmtricks/gatb/count_processor.hpp:84-120
 * bit (h - lower), LSB-first, for every hash with count >= hard_min (already filtered).    */
void orc_hash_vector(const uint64_t* keys, size_t n, uint64_t lower, uint64_t W, uint8_t* out)
{
  memset(out, 0, (size_t)(W / 8));
  for (size_t i = 0; i < n; i++) { uint64_t b = keys[i] - lower; out[b >> 3] |= (uint8_t)(1u << (b & 7)); }
}
