// arith_emul.cpp -- the device arithmetic of stage 2 (kmtricks_b200/csrc/common.cuh, records.cuh: XXH64 short
// paths, exact modulo by multiplication, reverse complement, k-mer extraction from super-k-mer records) compiled
// for the host through tests/emul/shim/cuda_runtime.h and checked against the CPU restatement (oracle/kmx_oracle.c)
// and plain integer arithmetic.  The records are built by stage 1's own record builder (s1_v5.cuh build_record),
// so the writer and the reader of the bucket format are checked against each other.  Test infrastructure only.
#include "../../kmtricks_b200/csrc/records.cuh"
#include "../../kmtricks_b200/csrc/s1_v5.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

extern "C" {
uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed);
uint64_t orc_hash_key(uint64_t lo, uint64_t hi, int w, uint64_t W, uint64_t p);
}

using namespace kmx;

#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, __VA_ARGS__); fprintf(stderr, " (%s:%d)\n", __FILE__, __LINE__); return 1; } } while (0)

static unsigned __int128 naive_kmer(const std::vector<int>& codes, int j, int k)
{
  unsigned __int128 f = 0;
  for (int i = 0; i < k; i++) f = (f << 2) | (unsigned)codes[j + i];
  return f;
}
static unsigned __int128 naive_rc(const std::vector<int>& codes, int j, int k)
{
  unsigned __int128 r = 0;
  for (int i = k - 1; i >= 0; i--) r = (r << 2) | (unsigned)(codes[j + i] ^ 2);
  return r;
}

int main()
{
  std::mt19937_64 rng(20261017);
  // ---- XXH64, 8 and 16 bytes, seed 0 (sorting_count.hpp:346-363 hashes kmer.get_data(), 8 w bytes)
  for (int it = 0; it < 200000; it++) {
    const u64 a = rng() >> (rng() % 64), b = rng() >> (rng() % 64);
    u64 buf[2] = {a, b};
    CHECK(xxh64_8(a) == orc_xxh64(buf, 8, 0), "xxh64_8(%llx)", (unsigned long long)a);
    CHECK(xxh64_16(a, b) == orc_xxh64(buf, 16, 0), "xxh64_16(%llx, %llx)", (unsigned long long)a, (unsigned long long)b);
  }
  // ---- exact modulo: 128-bit-magic fastmod for any divisor, one Barrett step for divisors < 2^31
  {
    std::vector<u64> ds = {1, 2, 3, 64, 3125056, 25000000, (1ULL << 31) - 1, (1ULL << 31), (1ULL << 32) + 5, 0xFFFFFFFFFFFFFFFFULL, 6400000000ULL};
    for (int i = 0; i < 200; i++) ds.push_back((rng() >> (rng() % 63)) | 1ULL);
    for (u64 d : ds) {
      const unsigned __int128 M = (~(unsigned __int128)0) / d + 1;          // as the host does (fastmod_magic, kmx_api.cu)
      FastMod64 fm; fm.d = d; fm.mlo = (u64)M; fm.mhi = (u64)(M >> 64);
      FastMod32 f32; f32.d = (u32)d; f32.m64 = d >= 2 ? (~0ULL) / d : 0;
      for (int it = 0; it < 3000; it++) {
        u64 x = rng() >> (rng() % 64);
        if (it < 8) x = it < 4 ? (u64)it : ~0ULL - (u64)(it - 4);
        if (it >= 8 && it < 16) x = d * (u64)(it - 7) + (it & 1 ? 0 : d - 1);  // around multiples of d (may wrap: still a valid x)
        CHECK(fastmod64(x, fm) == x % d, "fastmod64(%llu, %llu)", (unsigned long long)x, (unsigned long long)d);
        if (d >= 2 && d < (1ULL << 31)) CHECK(fastmod64_d32(x, f32) == (u32)(x % d), "fastmod64_d32(%llu, %llu)", (unsigned long long)x, (unsigned long long)d);
      }
    }
  }
  // ---- records: stage 1's builder -> stage 2's extraction, canonical k-mers, hash keys
  for (int k : {8, 15, 21, 31, 32, 33, 40, 47, 63}) {
    const int W = (k + 31) / 32, maxn = W == 1 ? KMX_REC1_MAXN : KMX_REC2_MAXN;
    for (int it = 0; it < 400; it++) {
      const int L = k + (int)(rng() % (300 - k));                                   // read length
      std::vector<int> codes(L);
      for (auto& c : codes) c = (int)(rng() & 3);
      if (it % 5 == 0) for (auto& c : codes) c = (it % 10 == 0) ? 0 : 3;            // poly-A / poly-G
      // packed forward stream in stage 1's paired layout: chunk c at word 2c
      const u32 nch = (u32)(L + 15) / 16;
      std::vector<u32> be(2 * nch + 2, 0);
      for (int i = 0; i < L; i++) be[2 * (i / 16)] |= (u32)codes[i] << (2 * (15 - i % 16));
      const int nb = k + (int)(rng() % (u64)(std::min(maxn, L) - k + 1));           // bases in the record
      const int st = (int)(rng() % (u64)(L - nb + 1));
      u32 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (W == 1) s1v5::build_record<4>(be.data(), nch, (u32)(st + nb), (u32)nb, v); else s1v5::build_record<8>(be.data(), nch, (u32)(st + nb), (u32)nb, v);
      uint4 rec[2] = {make_uint4(v[0], v[1], v[2], v[3]), make_uint4(v[4], v[5], v[6], v[7])};
      for (int j = 0; j + k <= nb; j++) {
        const unsigned __int128 f = naive_kmer(codes, st + j, k), rc = naive_rc(codes, st + j, k), cn = f < rc ? f : rc;
        u64 clo, chi = 0;
        if (W == 1) {
          const Rec1 r = load_rec1(rec, 0);
          CHECK(r.n == nb, "rec1 length");
          CHECK(rec1_kmer(r.lo, r.hi, r.n, k, j) == (u64)f, "rec1_kmer k=%d nb=%d j=%d", k, nb, j);
          CHECK(revcomp64((u64)f, k) == (u64)rc, "revcomp64 k=%d", k);
          canon1(r, k, j, clo);
        } else {
          const Rec2 r = load_rec2(rec, 0);
          CHECK(r.n == nb, "rec2 length");
          u64 flo, fhi, rlo, rhi;
          rec2_kmer(r.v0, r.v1, r.v2, r.v3, r.n, k, j, flo, fhi);
          CHECK(flo == (u64)f && fhi == (u64)(f >> 64), "rec2_kmer k=%d nb=%d j=%d", k, nb, j);
          revcomp128(flo, fhi, k, rlo, rhi);
          CHECK(rlo == (u64)rc && rhi == (u64)(rc >> 64), "revcomp128 k=%d", k);
          canon2(r, k, j, clo, chi);
        }
        CHECK(clo == (u64)cn && chi == (u64)(cn >> 64), "canonical k=%d", k);
        // hash key of the window (KmXXHash: XXH64 % W + W p) the way the fill kernels compute it
        const u64 Wb = 3125056, p = 5;
        const unsigned __int128 M = (~(unsigned __int128)0) / Wb + 1;
        FastMod64 fm; fm.d = Wb; fm.mlo = (u64)M; fm.mhi = (u64)(M >> 64);
        const u64 h = W == 1 ? xxh64_8(clo) : xxh64_16(clo, chi);
        CHECK(fastmod64(h, fm) + Wb * p == orc_hash_key(clo, chi, W, Wb, p), "hash key k=%d", k);
      }
    }
  }
  printf("OK\n");
  return 0;
}
