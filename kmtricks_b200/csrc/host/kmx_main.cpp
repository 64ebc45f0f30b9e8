// kmx -- C++ host for the kmtricks hot path on B200: keeps the `kmtricks pipeline` command line
// (reference src/cli.cpp:117-382), the run-directory layout (include/kmtricks/kmdir.hpp:195-236),
// the file headers (include/kmtricks/io/*.hpp) and the IMergePlugin surface
// (include/kmtricks/plugin.hpp, plugin_manager.hpp), and calls libkmx_sm100.so through the C ABI
// (include/kmx.h) for every compute stage.  No compute happens on the host.
#include <kmx.h>

#include <dlfcn.h>
#include <sys/stat.h>
#include <zlib.h>

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "kmx_plugin_abi.hpp"

namespace {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------- options
struct Options {
  std::string fof, dir, mode = "kmer:count:bin", repart_from, until = "all", plugin, plugin_config;
  uint32_t k = 31, m = 10, P = 0, hard_min = 2, soft_min = 1, rec_min = 1, share_min = 0, threads = 4;
  uint64_t bloom = 10000000;
  bool keep_tmp = false, static_repart = true, balanced_repart = false, hist = false;
  size_t repart_sample_mib = 64;     // --balanced-repart: text sampled from the head of each of up to 16 samples
  int device = 0;
  std::vector<int> devices;          // --devices a-b | a,b,c : partitions sharded over these GPUs (one in-process rank per GPU)
  size_t block_mib = 256;            // FASTQ is streamed to the device in blocks of this many MiB (whole records)
  std::string key_kind, what, fmt;
};

[[noreturn]] void usage(const char* why)
{
  if (why) std::cerr << "kmx: " << why << "\n";
  std::cerr << "usage: kmx pipeline --file <fof> --run-dir <dir> --nb-partitions <P> [--kmer-size 31]\n"
               "           [--mode <kmer|hash>:<count|pa|bf|bft>:bin] [--hard-min 2] [--soft-min 1] [--recurrence-min 1]\n"
               "           [--share-min 0] [--minimizer-size 10] [--bloom-size 10000000] [--static-repart | --repart-from <run-dir> | --balanced-repart]\n"
               "           [--until all|superk|count|merge] [--keep-tmp] [--hist] [--threads 4] [--plugin lib.so [--plugin-config s]] [--device 0 | --devices 0-7] [--block-mib 256]\n";
  std::exit(why ? EXIT_FAILURE : EXIT_SUCCESS);
}

Options parse(int argc, char** argv)
{
  if (argc < 2 || std::string(argv[1]) != "pipeline") usage(argc < 2 ? nullptr : "only the `pipeline` command is on the hot path");
  Options o;
  auto need = [&](int& i) -> std::string { if (i + 1 >= argc) usage((std::string(argv[i]) + " needs a value").c_str()); return argv[++i]; };
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--file") o.fof = need(i);
    else if (a == "--run-dir") o.dir = need(i);
    else if (a == "--kmer-size") o.k = std::stoul(need(i));
    else if (a == "--mode") o.mode = need(i);
    else if (a == "--hard-min") o.hard_min = std::stoul(need(i));
    else if (a == "--soft-min") o.soft_min = std::stoul(need(i));
    else if (a == "--recurrence-min") o.rec_min = std::stoul(need(i));
    else if (a == "--share-min") o.share_min = std::stoul(need(i));
    else if (a == "--nb-partitions") o.P = std::stoul(need(i));
    else if (a == "--minimizer-size") o.m = std::stoul(need(i));
    else if (a == "--bloom-size") o.bloom = std::stoull(need(i));
    else if (a == "--static-repart") o.static_repart = true;
    else if (a == "--balanced-repart") { o.balanced_repart = true; o.static_repart = false; }
    else if (a == "--repart-sample-mib") o.repart_sample_mib = std::stoull(need(i));
    else if (a == "--repart-from") { o.repart_from = need(i); o.static_repart = false; }
    else if (a == "--until") o.until = need(i);
    else if (a == "--keep-tmp") o.keep_tmp = true;
    else if (a == "--hist") o.hist = true;
    else if (a == "--threads" || a == "-t") o.threads = std::stoul(need(i));
    else if (a == "--plugin") o.plugin = need(i);
    else if (a == "--plugin-config") o.plugin_config = need(i);
    else if (a == "--device") o.device = std::stoi(need(i));
    else if (a == "--devices") {
      std::string v = need(i); size_t d = v.find('-');
      if (d != std::string::npos) { for (int g = std::stoi(v.substr(0, d)); g <= std::stoi(v.substr(d + 1)); g++) o.devices.push_back(g); }
      else { std::stringstream ds(v); for (std::string t; std::getline(ds, t, ',');) o.devices.push_back(std::stoi(t)); }
    }
    else if (a == "--block-mib") o.block_mib = std::stoull(need(i));
    else if (a == "--help" || a == "-h") usage(nullptr);
    else usage(("unknown option " + a).c_str());
  }
  if (o.fof.empty() || o.dir.empty()) usage("--file and --run-dir are required");
  if (o.P == 0) usage("--nb-partitions is required (the repartition map must be fixed, SURVEY F9)");
  std::stringstream ss(o.mode); std::getline(ss, o.key_kind, ':'); std::getline(ss, o.what, ':'); std::getline(ss, o.fmt, ':');
  if ((o.key_kind != "kmer" && o.key_kind != "hash") || (o.what != "count" && o.what != "pa" && o.what != "bf" && o.what != "bft") || (o.fmt != "bin" && o.fmt != "text" && !o.fmt.empty()))
    usage("--mode must be <kmer|hash>:<count|pa|bf|bft>:<bin|text>");
  if (o.fmt == "text" && o.what != "count" && o.what != "pa") usage("text output exists for count and pa matrices (merge.hpp:288-316,531-573)");
  if ((o.what == "bf" || o.what == "bft") && o.key_kind != "hash") usage("bf/bft need hash keys");
  if (o.until != "all" && o.until != "superk" && o.until != "count" && o.until != "merge") usage("--until must be all|superk|count|merge");
  if (o.devices.empty()) o.devices.push_back(o.device);
  if (o.block_mib < 1 || o.block_mib > 3072) usage("--block-mib must be 1..3072 (a block must stay below 4 GiB)");
  return o;
}

// ---------------------------------------------------------------------------- small helpers
void mkdirs(const std::string& p)
{
  std::string cur;
  for (size_t i = 0; i <= p.size(); i++) {
    if (i == p.size() || p[i] == '/') { if (!cur.empty() && mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) throw Error("cannot create " + cur); }
    if (i < p.size()) cur += p[i];
  }
}
std::string trim(const std::string& s)
{
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
template <class T> void put(std::string& b, T v) { b.append(reinterpret_cast<const char*>(&v), sizeof v); }
void write_file(const std::string& path, const std::string& head, const void* body, size_t n)
{
  std::ofstream out(path, std::ios::binary);
  if (!out) throw Error("Unable to write at " + path);
  out.write(head.data(), head.size());
  if (n) out.write(reinterpret_cast<const char*>(body), n);
}
std::string km_header() { std::string h; put<uint64_t>(h, 0x736b636972746d6bULL); put<uint32_t>(h, 0); put<uint8_t>(h, 0); return h; }

// XXH64 of one little-endian uint32 (len 4, seed 0): the --static-repart map (repartition.hpp:45-56)
uint64_t xxh64_u32(uint32_t x)
{
  const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P5 = 0x27D4EB2F165667C5ULL;
  uint64_t h = P5 + 4;
  h ^= (uint64_t)x * P1;
  h = ((h << 23) | (h >> 41)) * P2 + P3;
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

uint64_t window_bits(uint64_t bloom, uint32_t P)   // hash.hpp:31-38 (through doubles, as there)
{
  uint64_t w = (uint64_t)std::ceil((double)bloom / (double)P);
  return (w + 63) / 64 * 64;
}

struct Sample { std::string id; std::vector<std::string> files; uint32_t hard_min = 0; };

std::vector<Sample> read_fof(const std::string& path)   // grammar "ID : f1 ; f2 ! n" (io/fof.hpp:39-40,115-147)
{
  std::ifstream in(path);
  if (!in) throw Error("Unable to read at " + path);
  std::vector<Sample> v; std::map<std::string, int> seen;
  for (std::string line; std::getline(in, line);) {
    line = trim(line);
    if (line.empty()) continue;
    size_t c = line.find(':');
    if (c == std::string::npos) throw Error("Invalid fof format.");
    Sample s; s.id = trim(line.substr(0, c));
    std::string rest = line.substr(c + 1);
    size_t e = rest.find('!');
    if (e != std::string::npos) { s.hard_min = std::stoul(trim(rest.substr(e + 1))); rest = rest.substr(0, e); }
    std::stringstream ss(rest);
    for (std::string f; std::getline(ss, f, ';');) { f = trim(f); if (!f.empty()) s.files.push_back(f); }
    if (s.id.empty() || s.files.empty()) throw Error("Invalid fof format.");
    if (seen[s.id]++) throw Error(s.id + " -> sample identifiers must be unique.");
    v.push_back(s);
  }
  return v;
}

std::string read_all(const std::string& path)   // plain or .gz (zlib reads both)
{
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw Error("Unable to open " + path);
  gzbuffer(f, 1 << 20);
  std::string out; std::vector<char> buf(1 << 22);
  for (int n; (n = gzread(f, buf.data(), (unsigned)buf.size())) > 0;) out.append(buf.data(), n);
  gzclose(f);
  return out;
}

// kseq-style FASTA/FASTQ reader (gatb BankFasta.cpp:391-560 semantics): joined sequences + offsets
void parse_fastx(const std::string& b, std::string& seqs, std::vector<uint64_t>& off)
{
  const size_t n = b.size(); size_t pos = 0; int last = 0;
  if (off.empty()) off.push_back(0);
  for (;;) {
    if (last == 0) { while (pos < n && b[pos] != '>' && b[pos] != '@') pos++; if (pos >= n) break; last = b[pos++]; }
    if (pos >= n) break;
    while (pos < n && b[pos] != '\n') pos++;
    if (pos < n) pos++;
    const size_t s0 = seqs.size(); int c = -1;
    while (pos < n) {
      c = (unsigned char)b[pos++];
      if (c == '>' || c == '+' || c == '@') break;
      if (c == '\n') { c = -1; continue; }
      seqs.push_back((char)c);
      size_t ls = pos; while (pos < n && b[pos] != '\n') pos++;
      seqs.append(b, ls, pos - ls);
      if (pos < n) pos++;
      if (seqs.size() - s0 > 1 && seqs.back() == '\r') seqs.pop_back();
      c = -1;
    }
    if (c == '>' || c == '@') last = c;
    if (c == '+') {
      while (pos < n && b[pos] != '\n') pos++;
      if (pos < n) pos++;
      uint64_t qlen = 0; const uint64_t slen = seqs.size() - s0;
      while (pos < n) {
        size_t ls = pos; while (pos < n && b[pos] != '\n') pos++;
        uint64_t l = pos - ls; if (pos < n) pos++;
        qlen += l; if (qlen > 1 && l > 0 && b[ls + l - 1] == '\r') qlen--;
        if (qlen >= slen) break;
      }
      last = 0;
    }
    off.push_back(seqs.size());
    if (pos >= n) break;
  }
}

#define KX(call) do { int rc_ = (call); if (rc_) throw Error(std::string(#call) + ": " + kmx_last_error(ctx)); } while (0)

// ---------------------------------------------------------------------------- streaming FASTQ reader
// Hands out a (possibly gzipped) strict 4-line FASTQ file in blocks of whole records: a block ends after the last newline
// whose line number is a multiple of 4; the rest is carried into the next block.  Nothing but one block is ever resident.
struct FastqBlocks {
  gzFile f = nullptr; std::string carry; bool eof = false;
  explicit FastqBlocks(const std::string& path) { f = gzopen(path.c_str(), "rb"); if (!f) throw Error("Unable to open " + path); gzbuffer(f, 1 << 20); }
  ~FastqBlocks() { if (f) gzclose(f); }
  // fills buf (cap bytes); returns the bytes of the block, 0 at the end of the file; throws when one record exceeds cap
  size_t next(char* buf, size_t cap)
  {
    if (eof && carry.empty()) return 0;
    size_t n = carry.size();
    if (n > cap) throw Error("FASTQ record longer than a block (--block-mib)");
    memcpy(buf, carry.data(), n); carry.clear();
    while (!eof && n < cap) {
      int r = gzread(f, buf + n, (unsigned)std::min<size_t>(cap - n, (size_t)1 << 30));
      if (r < 0) throw Error("read error (gz)");
      if (r == 0) eof = true; else n += (size_t)r;
    }
    if (eof) return n;                                   // the last block takes everything (a final line without '\n' included)
    size_t lines = 0, cut = 0;
    for (const char* p = buf; (p = (const char*)memchr(p, '\n', buf + n - p)) != nullptr; ) { p++; if (++lines % 4 == 0) cut = (size_t)(p - buf); }
    if (cut == 0) throw Error("FASTQ record longer than a block (--block-mib)");
    carry.assign(buf + cut, n - cut);
    return cut;
  }
};

// ---------------------------------------------------------------------------- writer thread (matrices are written while the next partition merges)
struct WriteJob { std::string path, head; std::vector<uint8_t> body; };
struct Writer {
  std::mutex mu; std::condition_variable cv; std::deque<WriteJob> q; bool stop = false; std::exception_ptr err; std::thread th;
  Writer() : th([this] { run(); }) {}
  void run()
  {
    for (;;) {
      WriteJob j;
      { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return stop || !q.empty(); }); if (q.empty()) return; j = std::move(q.front()); q.pop_front(); }
      cv.notify_all();
      try { write_file(j.path, j.head, j.body.data(), j.body.size()); } catch (...) { std::lock_guard<std::mutex> l(mu); if (!err) err = std::current_exception(); }
    }
  }
  void push(WriteJob&& j) { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return q.size() < 2; }); q.push_back(std::move(j)); cv.notify_all(); }
  void finish() { { std::lock_guard<std::mutex> l(mu); stop = true; } cv.notify_all(); if (th.joinable()) th.join(); if (err) std::rethrow_exception(err); }
  ~Writer() { try { finish(); } catch (...) {} }
};

// ---------------------------------------------------------------------------- plugin host
struct PluginHost {
  void* handle = nullptr;
  km::IMergePlugin* (*create)() = nullptr;
  void (*destroy)(km::IMergePlugin*) = nullptr;
  std::string name, config;
  void load(const std::string& path, const std::string& cfg, uint32_t k)
  {
    config = cfg;
    handle = dlopen(path.c_str(), RTLD_LAZY);
    if (!handle) throw Error(std::string("Unable to load shared lib. dlerror: ") + dlerror());
    auto use_t = reinterpret_cast<int (*)()>(dlsym(handle, "use_template"));
    if (!use_t) throw Error("Unable to load symbol use_template");
    std::string sym = "create" + std::to_string(use_t() ? (k < 32 ? 32 : 64) : 0);   // MAX_K of the run (KMER_LIST 32 64)
    create = reinterpret_cast<km::IMergePlugin* (*)()>(dlsym(handle, sym.c_str()));
    destroy = reinterpret_cast<void (*)(km::IMergePlugin*)>(dlsym(handle, "destroy"));
    auto pname = reinterpret_cast<std::string (*)()>(dlsym(handle, "plugin_name"));
    if (!create || !destroy || !pname) throw Error("Unable to load symbol " + sym + " / destroy / plugin_name");
    name = pname();
  }
  ~PluginHost() { if (handle) dlclose(handle); }
};

// HowDeSBT Bloom filter file header of a simple uncompressed one-vector filter, as howde_utils.hpp:57-90 fills it.  The struct
// (bloom_filter_file.h of HowDeSBT) is NOT vendored in the reference: field order and magic numbers below are restated from
// HowDeSBT's published header -- parity UNPINNED (the reference cannot build one at this commit); the payload placement is pinned.
static const size_t HOWDE_HEADER_BYTES = 112;           // round_up_16(sizeof(bffileheader) with one bfvectorinfo) = 0x50 + 0x20
std::string howde_header(uint32_t k, uint64_t bloom_bits)
{
  std::string h(HOWDE_HEADER_BYTES, '\0');
  auto w32 = [&](size_t o, uint32_t v) { memcpy(&h[o], &v, 4); };
  auto w64 = [&](size_t o, uint64_t v) { memcpy(&h[o], &v, 8); };
  w64(0x00, 0xD532006662544253ULL);                      // bffileheaderMagic ("SBTbf")
  w32(0x08, (uint32_t)HOWDE_HEADER_BYTES); w32(0x0C, 1); // headerSize, version
  w32(0x10, 1);                                          // bfKind = bfkind_simple
  w32(0x18, k); w32(0x1C, 1);                            // smerSize, numHashes
  w64(0x20, 0); w64(0x28, 0);                            // hashSeed1, hashSeed2
  w64(0x30, bloom_bits); w64(0x38, bloom_bits);          // hashModulus, numBits
  w32(0x40, 1); w32(0x44, 0); w64(0x48, 0);              // numVectors, setSizeKnown, setSize
  w32(0x50, 1); w32(0x54, 0);                            // info[0].compressor = bvcomp_uncompressed, name
  w64(0x58, HOWDE_HEADER_BYTES);                         // info[0].offset
  w64(0x60, bloom_bits / 8 + 8); w64(0x68, 0);           // info[0].numBytes (sdsl bit_vector: u64 bit count + words), filterInfo
  return h;
}

// ---------------------------------------------------------------------------- balanced repartition (RepartTask, task.hpp:170-222)
// The reference samples the banks on the CPU and spreads the minimizers over the partitions by estimated load
// (gatb RepartitionAlgorithm.cpp:395-492, PartiInfo.cpp:48-106).  Here the head of each of up to 16 samples goes through stage 1 on
// the device with the minimizer-load counters on (kmx_minimizer_load_*), and the host assigns the minimizers heaviest first to
// the lightest partition (longest-processing-time rule); minimizers the sample never showed keep their static place.  The result
// is written as repartition_gatb/repartition.minimRepart, which the reference accepts through --repart-from.
std::vector<uint16_t> balanced_table(const Options& o, const std::vector<Sample>& samples, const std::vector<uint16_t>& static_table)
{
  const size_t tn = static_table.size();
  kmx_params prm{};
  prm.kmer_size = o.k; prm.minim_size = o.m; prm.nb_partitions = o.P; prm.key_kind = KMX_KEY_KMER; prm.window_bits = 0;
  prm.repart_table = static_table.data(); prm.nb_samples = 1;
  kmx_ctx* ctx = nullptr;
  if (int rc = kmx_create(o.devices[0], &prm, &ctx)) { std::string e = ctx ? kmx_last_error(ctx) : "failed"; if (ctx) kmx_destroy(ctx); throw Error("kmx_create (repartition estimate): " + e + " (code " + std::to_string(rc) + ")"); }
  std::vector<uint64_t> load(tn, 0);
  try {
    KX(kmx_minimizer_load_enable(ctx, 1));
    std::vector<char> buf(o.repart_sample_mib << 20);
    const size_t step = std::max<size_t>(1, samples.size() / 16);
    for (size_t s = 0; s < samples.size(); s += step) {
      KX(kmx_superk_begin(ctx));
      FastqBlocks fb(samples[s].files[0]);
      const size_t n = fb.next(buf.data(), buf.size());
      int rc = (n && buf[0] == '@') ? kmx_superk_push_fastq(ctx, buf.data(), n, 0) : KMX_ERR_FORMAT;
      if (rc == KMX_ERR_FORMAT) {                          // FASTA / multi-line: the head of the file through the kseq-style parser
        std::string text(buf.data(), n), seqs; std::vector<uint64_t> off;
        parse_fastx(text, seqs, off);
        if (off.size() > 2) { off.pop_back(); KX(kmx_superk_push_reads(ctx, seqs.data(), off.data(), off.size() - 1)); }   // the last record may be cut
      } else if (rc) throw Error(std::string("kmx_superk_push_fastq: ") + kmx_last_error(ctx));
      KX(kmx_superk_end(ctx, nullptr));
    }
    KX(kmx_minimizer_load_get(ctx, load.data()));
  } catch (...) { kmx_destroy(ctx); throw; }
  kmx_destroy(ctx);
  std::vector<uint32_t> order;
  for (size_t x = 0; x < tn; x++) if (load[x]) order.push_back((uint32_t)x);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return load[a] > load[b]; });
  std::vector<uint16_t> table = static_table;
  std::vector<std::pair<uint64_t, uint32_t>> heap;        // (load, partition), min-heap
  for (uint32_t p = 0; p < o.P; p++) heap.push_back({0, p});
  auto cmp = [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a > b; };
  std::make_heap(heap.begin(), heap.end(), cmp);
  for (uint32_t x : order) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    table[x] = (uint16_t)heap.back().second; heap.back().first += load[x];
    std::push_heap(heap.begin(), heap.end(), cmp);
  }
  return table;
}

// ---------------------------------------------------------------------------- one GPU (one rank)
struct Rank {
  int device = 0; kmx_ctx* ctx = nullptr;
  uint32_t first_part = 0, last_part = 0;                // owned partitions [first, last)
};

struct Run {
  Options o; std::vector<Sample> samples; uint32_t N = 0, P = 0, w = 1; bool hash = false; uint64_t W = 0;
  std::vector<std::vector<uint64_t>> pinfo;
  std::mutex err_mu; std::exception_ptr err; std::atomic<bool> failed{false};
  void fail(std::exception_ptr e) { std::lock_guard<std::mutex> l(err_mu); if (!err) err = e; failed = true; }
  uint32_t hard_min(uint32_t s) const { return samples[s].hard_min ? samples[s].hard_min : o.hard_min; }
};

// kseq-parsed whole file(s): FASTA, multi-line records, anything that is not strict 4-line FASTQ
void push_parsed(kmx_ctx* ctx, uint32_t lane, const std::string& path)
{
  std::string text = read_all(path);
  int rc = (!text.empty() && text[0] == '@' && text.size() < 0xFFFFFFF0ULL) ? kmx_lane_superk_push_fastq(ctx, lane, text.data(), text.size(), 0) : KMX_ERR_FORMAT;
  if (rc == KMX_ERR_FORMAT) {
    std::string seqs; std::vector<uint64_t> off; parse_fastx(text, seqs, off);
    text.clear(); text.shrink_to_fit();
    KX(kmx_lane_superk_push_reads(ctx, lane, seqs.data(), off.data(), off.size() - 1));
  } else if (rc) throw Error(std::string("kmx_lane_superk_push_fastq: ") + kmx_last_error(ctx));
}

// stage 1 + 2 of one sample on one lane, its files streamed block by block through the lane's pinned buffer
void process_sample(Run& R, kmx_ctx* ctx, uint32_t lane, uint32_t s, char* pin, size_t pin_cap)
{
  const Sample& smp = R.samples[s];
  bool streamed = true;
  KX(kmx_lane_superk_begin(ctx, lane));
  for (const std::string& f : smp.files) {
    FastqBlocks fb(f);
    bool first = true;
    for (size_t n; streamed && (n = fb.next(pin, pin_cap)) != 0; first = false) {
      if (first && pin[0] != '@') { streamed = false; break; }
      int rc = kmx_lane_superk_push_fastq(ctx, lane, pin, n, 0);
      if (rc == KMX_ERR_FORMAT) streamed = false;          // multi-line FASTQ, ...: the host parses it kseq-style
      else if (rc) throw Error(std::string("kmx_lane_superk_push_fastq: ") + kmx_last_error(ctx));
    }
    if (!streamed) break;
  }
  if (!streamed) {                                         // restart the sample (begin clears what the blocks had put into the buckets)
    KX(kmx_lane_superk_begin(ctx, lane));
    for (const std::string& f : smp.files) push_parsed(ctx, lane, f);
  }
  KX(kmx_lane_superk_end(ctx, lane, R.pinfo[s].data()));
  if (R.o.until == "superk") return;
  if (!R.o.hist) { KX(kmx_lane_count_sample(ctx, lane, s, R.hard_min(s))); return; }
  // --hist: histograms/<id>.hist (HistWriter, io/hist_file.hpp:31-130; KHist(i, k, 1, 255), task_scheduler.hpp:103)
  const uint64_t lower = 1, upper = 255, nb = upper - lower + 1;
  std::vector<uint64_t> hv(6 + 2 * nb);
  KX(kmx_lane_count_sample_hist(ctx, lane, s, R.hard_min(s), (uint32_t)lower, (uint32_t)upper, hv.data()));
  std::string h = km_header();
  put<uint64_t>(h, 0x747369686bULL); put<uint32_t>(h, R.o.k); put<uint32_t>(h, s); put<uint64_t>(h, lower); put<uint64_t>(h, upper);
  put<uint64_t>(h, hv[0]); put<uint64_t>(h, hv[1]);                       // uniq, total
  put<uint64_t>(h, hv[3]); put<uint64_t>(h, hv[2]); put<uint64_t>(h, hv[5]); put<uint64_t>(h, hv[4]);   // oob_ln, oob_lu, oob_un, oob_uu (the order serialize() writes them in)
  write_file(R.o.dir + "/histograms/" + smp.id + ".hist", h, hv.data() + 6, 2 * nb * 8);
}

std::string matrix_header(const Run& R, uint32_t p)
{
  const Options& o = R.o; const uint32_t N = R.N, w = R.w, nbytes = (N + 7) / 8;
  std::string h = km_header();
  if (!R.hash && o.what == "count") { put<uint64_t>(h, 0x6b5f78697274616dULL); put<uint32_t>(h, o.k); put<uint32_t>(h, w); put<uint32_t>(h, 1); put<uint32_t>(h, N); put<uint32_t>(h, 0); put<uint32_t>(h, 0); }
  else if (!R.hash) { put<uint64_t>(h, 0x6b5f74616d6170ULL); put<uint32_t>(h, o.k); put<uint32_t>(h, w); put<uint32_t>(h, N); put<uint32_t>(h, nbytes); put<uint32_t>(h, 0); put<uint32_t>(h, 0); }
  else if (o.what == "count") { put<uint64_t>(h, 0x685f78697274616dULL); put<uint32_t>(h, 4); put<uint32_t>(h, N); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
  else if (o.what == "pa") { put<uint64_t>(h, 0x685f74616d6170ULL); put<uint32_t>(h, N); put<uint32_t>(h, nbytes); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
  else { put<uint64_t>(h, 0x74616d746962ULL); put<uint32_t>(h, N); put<uint64_t>(h, R.W * p); put<uint64_t>(h, R.W); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
  return h;
}

// counts/partition_P/<id>.kmer|.hash of the partitions this rank holds (kmer_file.hpp:31-108, hash_file.hpp:31-131)
void write_counts(const Run& R, const Rank& rk)
{
  kmx_ctx* ctx = rk.ctx; const Options& o = R.o;
  for (uint32_t s = 0; s < R.N; s++) for (uint32_t p = rk.first_part; p < rk.last_part; p++) {
    uint64_t n = 0; KX(kmx_counts_size(ctx, s, p, &n));
    const uint32_t kw = R.hash ? 1 : R.w;
    std::vector<uint64_t> keys(n * kw); std::vector<uint32_t> cnt(n);
    KX(kmx_counts_get(ctx, s, p, keys.data(), cnt.data()));
    std::string path = o.dir + "/counts/partition_" + std::to_string(p) + "/" + R.samples[s].id + (R.hash ? ".hash" : ".kmer");
    std::string h = km_header(), body;
    if (!R.hash) {
      put<uint64_t>(h, 0x72656d6bULL); put<uint32_t>(h, o.k); put<uint32_t>(h, R.w); put<uint32_t>(h, 4); put<uint32_t>(h, s); put<uint32_t>(h, p);
      body.reserve(n * (8 * R.w + 4));
      for (uint64_t i = 0; i < n; i++) { body.append((const char*)&keys[i * R.w], 8 * R.w); body.append((const char*)&cnt[i], 4); }
    } else {
      put<uint64_t>(h, 0x68736168ULL); put<uint32_t>(h, 4); put<uint32_t>(h, s); put<uint32_t>(h, p);
      for (uint64_t i = 0; i < n; i += 4096) {
        uint64_t b = std::min<uint64_t>(4096, n - i);
        put<uint64_t>(body, b); body.append((const char*)&keys[i], 8 * b); body.append((const char*)&cnt[i], 4 * b);
      }
    }
    write_file(path, h, body.data(), body.size());
  }
}

// merge of the partitions this rank owns (KmerMergeTask / HashMergeTask, task.hpp:690-870); files go through the writer thread
void merge_partitions(Run& R, const Rank& rk, PluginHost& plug, const std::vector<int>& bf_fds)
{
  kmx_ctx* ctx = rk.ctx; const Options& o = R.o; const uint32_t N = R.N, w = R.w; const bool hash = R.hash; const uint64_t W = R.W;
  std::vector<uint32_t> soft(N, o.soft_min);
  const char* ext = !hash ? (o.what == "count" ? "count" : "pa") : (o.what == "count" ? "count_hash" : o.what == "pa" ? "pa_hash" : "cmbf");
  const uint32_t nbytes = (N + 7) / 8, kw = hash ? 1 : w;
  const uint32_t fmt = o.what == "count" ? KMX_FMT_COUNT : o.what == "pa" ? KMX_FMT_PA : o.what == "bf" ? KMX_FMT_BF : KMX_FMT_BFT;
  Writer writer;
  for (uint32_t p = rk.first_part; p < rk.last_part && !R.failed; p++) {
    // with a plugin every merged row goes through it, whatever the output mode (merge.hpp:249-257,509-514): the library
    // returns all rows with their counts, the plugin decides / edits, and the host encodes the mode's row from its counts
    kmx_merge_params mp{soft.data(), o.rec_min, o.share_min, plug.handle ? (uint32_t)KMX_FMT_COUNT : fmt, (uint32_t)(plug.handle ? 1 : 0)};
    kmx_merge_result r{};
    KX(kmx_merge_partition(ctx, p, &mp, &r));
    WriteJob job; job.path = o.dir + "/matrices/matrix_" + std::to_string(p) + "." + ext; job.head = matrix_header(R, p);
    std::vector<uint8_t> body(r.n_rows * r.row_bytes), keep(plug.handle ? r.n_rows : 0);
    std::vector<uint64_t> stats((size_t)6 * N);
    KX(kmx_merge_get(ctx, body.data(), stats.data(), plug.handle ? keep.data() : nullptr));
    if (!plug.handle) job.body = std::move(body);
    else {
      // one plugin instance per merge task (task.hpp:701-712), called once per merged row in key order;
      // its return value replaces the keep decision and it may edit the counts
      km::IMergePlugin* pl = plug.create();
      pl->configure(plug.config);
      pl->set_out_dir(o.dir + "/plugin_output");
      pl->set_kmer_size(hash ? 0 : o.k);
      pl->set_partition(p);
      std::vector<uint32_t> c(N);
      std::vector<uint8_t>& out = job.body;
      const bool dense = fmt == KMX_FMT_BF || fmt == KMX_FMT_BFT;
      if (dense) out.assign((size_t)W * nbytes, 0);
      for (uint64_t i = 0; i < r.n_rows; i++) {
        const uint8_t* row = &body[i * r.row_bytes];
        memcpy(c.data(), row + 8 * kw, 4 * N);
        uint64_t kbuf[2]; memcpy(kbuf, row, 8 * kw);
        const bool kp = hash ? pl->process_hash(kbuf[0], c) : pl->process_kmer(kbuf, c);
        if (!kp) continue;
        if (fmt == KMX_FMT_COUNT) { out.insert(out.end(), (const uint8_t*)kbuf, (const uint8_t*)kbuf + 8 * kw); out.insert(out.end(), (const uint8_t*)c.data(), (const uint8_t*)c.data() + 4 * N); }
        else {
          uint8_t* bits;
          if (dense) bits = &out[(size_t)(kbuf[0] - W * p) * nbytes];
          else { const size_t at = out.size(); out.resize(at + 8 * kw + nbytes, 0); memcpy(&out[at], kbuf, 8 * kw); bits = &out[at + 8 * kw]; }
          for (uint32_t s2 = 0; s2 < N; s2++) if (c[s2]) bits[s2 >> 3] |= (uint8_t)(1u << (s2 & 7));
        }
      }
      plug.destroy(pl);
      if (fmt == KMX_FMT_BFT) {                            // W x 8*nbytes bits -> 8*nbytes x W bits (BitMatrix::transpose) on the device
        std::vector<uint8_t> t(out.size());
        KX(kmx_transpose_bits(ctx, out.data(), W, (uint64_t)nbytes * 8, t.data()));
        out.swap(t);
      }
    }
    if (o.fmt == "text") {                                 // write_as_text / write_as_pa_text (merge.hpp:288-316,531-573): no header, one row per line
      const size_t rb = 8 * kw + (fmt == KMX_FMT_COUNT ? 4 * (size_t)N : nbytes);
      const size_t nrows = rb ? job.body.size() / rb : 0;
      std::string txt; txt.reserve(nrows * (o.k + 2 * (size_t)N + 2));
      std::string km(o.k, 'A');
      for (size_t i = 0; i < nrows; i++) {
        const uint8_t* row = &job.body[i * rb];
        uint64_t kbuf[2] = {0, 0}; memcpy(kbuf, row, 8 * kw);
        if (hash) txt += std::to_string(kbuf[0]);
        else {                                             // Kmer::to_string (kmer.hpp:541-550): first base most significant, "ACTG"[code]
          for (uint32_t b = 0; b < o.k; b++) km[o.k - 1 - b] = "ACTG"[(kbuf[b >> 5] >> (2 * (b & 31))) & 3];
          txt += km;
        }
        if (fmt == KMX_FMT_COUNT) { uint32_t c; for (uint32_t s2 = 0; s2 < N; s2++) { memcpy(&c, row + 8 * kw + 4 * s2, 4); txt += ' '; txt += std::to_string(c); } }
        else for (uint32_t s2 = 0; s2 < N; s2++) { txt += ' '; txt += ((row[8 * kw + (s2 >> 3)] >> (s2 & 7)) & 1) ? '1' : '0'; }
        txt += '\n';
      }
      job.path += ".txt"; job.head.clear();
      job.body.assign(txt.begin(), txt.end());
    }
    if (fmt == KMX_FMT_BFT && !bf_fds.empty()) {           // per-sample filter: row s of this partition at its place in filters/<id>.bf
      const size_t rb = W / 8;
      for (uint32_t s2 = 0; s2 < N; s2++)
        if (pwrite(bf_fds[s2], &job.body[(size_t)s2 * rb], rb, (off_t)(HOWDE_HEADER_BYTES + 8 + (uint64_t)p * rb)) != (ssize_t)rb) throw Error("Unable to write filters/" + R.samples[s2].id + ".bf");
    }
    writer.push(std::move(job));
    {   // merge_infos/partitionP.merge_info (merge.hpp:72-83)
      static const char* names[6] = {"NON_SOLID", "RESCUED", "UNIQUE_WO_RESCUE", "UNIQUE_W_RESCUE", "TOTAL_WO_RESCUE", "TOTAL_W_RESCUE"};
      std::ofstream mi(o.dir + "/merge_infos/partition" + std::to_string(p) + ".merge_info");
      for (int q = 0; q < 6; q++) { mi << names[q] << '\t'; for (uint32_t s2 = 0; s2 < N; s2++) mi << stats[(size_t)q * N + s2] << '\t'; mi << "\n"; }
    }
    if (o.what == "bf") {   // fpr/partition_P.txt (task.hpp:849-860; utils.hpp:239-243) -- the only floating point on the path
      std::ofstream fp(o.dir + "/fpr/partition_" + std::to_string(p) + ".txt");
      for (uint32_t s2 = 0; s2 < N; s2++) {
        double nn = (double)stats[(size_t)3 * N + s2];
        double fpr = std::pow(1.0 - std::pow(std::exp(1.0), (-(1.0 * nn)) / (double)W), 1.0);
        fp << std::fixed << fpr << "\n";
      }
    }
  }
  writer.finish();
}

}  // namespace

int main(int argc, char** argv)
{
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 1; i < argc; i++) if (std::string(argv[i]) == "--devices")
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);      // multi-GPU: one hardware queue per stream, so a lane's NCCL kernel never sits in front of another lane's
  std::vector<Rank> ranks;
  try {
    if (argc == 3 && std::string(argv[1]) == "fof") {     // host-side check, no device: prints the parsed input list
      for (const Sample& s : read_fof(argv[2])) {
        std::cout << s.id << "\t" << s.hard_min << "\t";
        for (size_t i = 0; i < s.files.size(); i++) std::cout << (i ? ";" : "") << s.files[i];
        std::cout << "\n";
      }
      return EXIT_SUCCESS;
    }
    if (argc == 4 && std::string(argv[1]) == "blocks") {  // host-side check, no device: how a FASTQ file is cut into blocks of whole records
      FastqBlocks fb(argv[2]);
      std::vector<char> buf(std::stoull(argv[3]));
      uint64_t h = 1469598103934665603ULL, total = 0;      // FNV-1a of the concatenated blocks
      for (size_t n; (n = fb.next(buf.data(), buf.size())) != 0;) {
        size_t lines = 0; for (size_t i = 0; i < n; i++) lines += buf[i] == '\n';
        for (size_t i = 0; i < n; i++) { h ^= (unsigned char)buf[i]; h *= 1099511628211ULL; }
        total += n;
        std::cout << n << "\t" << lines << "\t" << (int)(buf[0] == '@') << "\n";
      }
      std::cout << "total\t" << total << "\t" << h << "\n";
      return EXIT_SUCCESS;
    }
    Run R;
    R.o = parse(argc, argv);
    const Options& o = R.o;
    R.samples = read_fof(o.fof);
    const uint32_t N = R.N = (uint32_t)R.samples.size(), P = R.P = o.P; R.w = (o.k + 31) / 32;
    const bool hash = R.hash = o.key_kind == "hash";
    const uint64_t W = R.W = window_bits(o.bloom, P);
    const int G = (int)o.devices.size();
    if (G > 1 && (uint32_t)G > P) throw Error("--devices: more GPUs than partitions");
    R.pinfo.assign(N, std::vector<uint64_t>(P, 0));
    // ---- run directory (kmdir.hpp:195-236)
    for (const char* d : {"", "/config_gatb", "/repartition_gatb", "/superkmers", "/counts", "/matrices", "/merge_infos",
                          "/partition_infos", "/fpr", "/plugin_output", "/histograms", "/minimizers", "/filters", "/howde_index"})
      mkdirs(o.dir + d);
    for (uint32_t p = 0; p < P; p++) mkdirs(o.dir + "/counts/partition_" + std::to_string(p));
    { std::ifstream src(o.fof, std::ios::binary); std::ofstream dst(o.dir + "/kmtricks.fof", std::ios::binary); dst << src.rdbuf(); }
    {
      std::ofstream op(o.dir + "/options.txt");
      op << "Options: dir=" << o.dir << ", nb_threads=" << o.threads << ", fof=" << o.fof << ", kmer_size=" << o.k << ", c_ab_min=" << o.hard_min
         << ", m_ab_min=" << o.soft_min << ", r_min=" << o.rec_min << ", save_if=" << o.share_min << ", minim_size=" << o.m << ", nb_parts=" << P
         << ", bloom_size=" << o.bloom << ", keep_tmp=" << o.keep_tmp << ", static_repart=" << o.static_repart << ", use_plugin=" << !o.plugin.empty()
         << ", plugin=" << o.plugin << ", plugin_config=" << o.plugin_config << ", mode=" << o.what << ", format=" << (o.fmt == "text" ? "text" : "bin") << ", count_format=" << o.key_kind
         << ", until=" << o.until << ", engine=kmx_sm100, gpus=" << G << "\n";
    }
    { std::string h; put<uint64_t>(h, W * P); put<uint64_t>(h, P); put<uint64_t>(h, W); put<uint64_t>(h, W / 8); put<uint32_t>(h, o.m); write_file(o.dir + "/hash.info", h, nullptr, 0); }
    // ---- repartition table (RepartTask, task.hpp:170-222)
    const size_t tn = (size_t)1 << (2 * o.m);
    std::vector<uint16_t> table(tn);
    if (o.static_repart || o.balanced_repart) for (size_t x = 0; x < tn; x++) table[x] = (uint16_t)(xxh64_u32((uint32_t)x) % P);
    if (o.balanced_repart) table = balanced_table(o, R.samples, table);
    else if (!o.static_repart) {
      std::ifstream in(o.repart_from + "/repartition_gatb/repartition.minimRepart", std::ios::binary);
      if (!in) throw Error("Unable to read at " + o.repart_from + "/repartition_gatb/repartition.minimRepart");
      uint16_t fp, npass; uint64_t fn;
      in.read((char*)&fp, 2); in.read((char*)&fn, 8); in.read((char*)&npass, 2);
      if (fp != P || fn != tn) throw Error("--repart-from: incompatible repartition (partitions / minimizer size)");
      in.read((char*)table.data(), tn * 2);
    }
    { std::string h; put<uint16_t>(h, (uint16_t)P); put<uint64_t>(h, tn); put<uint16_t>(h, 1);
      std::string tail; put<uint8_t>(tail, 0); put<uint32_t>(tail, 0x12345678);
      std::ofstream out(o.dir + "/repartition_gatb/repartition.minimRepart", std::ios::binary);
      out.write(h.data(), h.size()); out.write((const char*)table.data(), tn * 2); out.write(tail.data(), tail.size()); }

    if (o.m <= 12) {   // minimizers/minimizers.P (RepartTask::postprocess, task.hpp:160-168; Repartition::write_minimizers, repartition.hpp:116-124)
      std::vector<std::string> buf(P);
      std::string mm(o.m, 'A');
      for (size_t x = 0; x < tn; x++) {
        for (uint32_t b = 0; b < o.m; b++) mm[o.m - 1 - b] = "ACTG"[(x >> (2 * b)) & 3];      // Mmer::to_string (kmer.hpp:115-128)
        buf[table[x]] += mm; buf[table[x]] += '\n';
      }
      for (uint32_t p = 0; p < P; p++) write_file(o.dir + "/minimizers/minimizers." + std::to_string(p), buf[p], nullptr, 0);
    }
    {   // config_gatb/gatb.config: the fields Configuration::save writes, in its order (gatb Configuration.cpp:145-178).  The reference reads
        // k, m and the partition count back from it when this directory is given to --repart-from (check_repart_compatibility,
        // task.hpp:135-147); the estimates (sequence counts, volumes) are GATB planning figures nothing downstream reads: zero here.
      std::string c;
      put<uint64_t>(c, o.k); put<uint64_t>(c, o.m); put<uint64_t>(c, 0); put<uint64_t>(c, 0);   // _kmerSize, _minim_size, _repartitionType, _minimizerType
      put<uint64_t>(c, 0); put<uint32_t>(c, 8000);                                              // _max_disk_space, _max_memory
      put<uint64_t>(c, 1); put<uint64_t>(c, 1); put<uint64_t>(c, 1); put<uint64_t>(c, 1);       // _nbCores, _nb_partitions_in_parallel, _abundanceUserNb, _nbCores_per_partition
      for (int q = 0; q < 6; q++) put<uint64_t>(c, 0);                                          // _estimateSeqNb/TotalSize/MaxSize, _available_space, _volume, _kmersNb
      put<uint32_t>(c, 1); put<uint32_t>(c, P); put<uint16_t>(c, (uint16_t)(64 * ((o.k + 31) / 32)));
      size_t nfiles = 0; for (const Sample& sm : R.samples) nfiles += sm.files.size();
      put<uint16_t>(c, (uint16_t)nfiles); put<uint32_t>(c, 0);   // _nb_passes, _nb_partitions, _nb_bits_per_kmer, _nb_banks, _nb_cached_items_per_core_per_part
      write_file(o.dir + "/config_gatb/gatb.config", c, nullptr, 0);
    }
    {
      std::ofstream bi(o.dir + "/build_infos.txt");
      bi << "kmx (kmtricks hot path on sm_100a; run-directory layout of kmtricks v1.6.0)\n\n- BUILD -\nkmer: 32,64\nmax_c: 4294967295\nengine: libkmx_sm100\n";
    }

    kmx_params prm{};
    prm.kmer_size = o.k; prm.minim_size = o.m; prm.nb_partitions = P; prm.key_kind = hash ? KMX_KEY_HASH : KMX_KEY_KMER;
    prm.window_bits = hash ? W : 0; prm.repart_table = table.data(); prm.nb_samples = N;
    ranks.resize(G);
    for (int g = 0; g < G; g++) {
      ranks[g].device = o.devices[g];
      ranks[g].first_part = (uint32_t)(((uint64_t)g * P) / G); ranks[g].last_part = (uint32_t)(((uint64_t)(g + 1) * P) / G);
      if (int rc = kmx_create(o.devices[g], &prm, &ranks[g].ctx))
        throw Error(std::string("kmx_create: ") + (ranks[g].ctx ? kmx_last_error(ranks[g].ctx) : "failed") + " (code " + std::to_string(rc) + ")");
    }
    const uint32_t lanes = std::max(1u, std::min(o.threads, 8u));        // --threads = samples in flight per GPU

    // ---- superk + count (TaskScheduler::exec_superk_count, task_scheduler.hpp:251-348)
    if (G == 1) {
      // one host thread per lane: it reads / inflates a block of its sample into its pinned buffer and pushes it, while the
      // other lanes' blocks are on the device -- no input is ever resident as a whole, a block is < 4 GiB by construction
      kmx_ctx* ctx = ranks[0].ctx;
      const uint32_t T = std::min<uint32_t>(lanes, std::max(1u, N));
      KX(kmx_lanes(ctx, T));
      std::atomic<uint32_t> next{0};
      auto work = [&](uint32_t t) {
        void* pin = nullptr; const size_t cap = o.block_mib << 20;
        try {
          if (kmx_host_alloc(cap, &pin)) throw Error("pinned block buffer: out of memory");
          for (uint32_t s; !R.failed && (s = next.fetch_add(1)) < N;) process_sample(R, ctx, t, s, (char*)pin, cap);
        } catch (...) { R.fail(std::current_exception()); }
        if (pin) kmx_host_free(pin);
      };
      std::vector<std::thread> th;
      for (uint32_t t = 0; t < T; t++) th.emplace_back(work, t);
      for (auto& x : th) x.join();
      if (R.err) std::rethrow_exception(R.err);
      KX(kmx_sync(ctx));
    } else {
      // one in-process rank per GPU: rank g parses the samples [g nl, (g+1) nl), the buckets travel to the partitions' owners
      // (one NCCL all-to-all-v per sample), every rank counts and later merges its own partitions.  Batches of a few samples.
      if (o.until == "superk") throw Error("--until superk needs a single --device");
      if (o.hist) throw Error("--hist needs a single --device");
      const uint32_t nl = (N + G - 1) / G, B = std::max(lanes, 4u);
      std::vector<uint8_t> ids((size_t)128 * lanes);
      for (uint32_t t = 0; t < lanes; t++) if (kmx_dist_unique_id(&ids[(size_t)128 * t])) throw Error("kmx_dist_unique_id failed (libnccl.so.2 not loadable?)");
      auto work = [&](int g) {
        kmx_ctx* ctx = ranks[g].ctx;
        try {
          KX(kmx_dist_init(ctx, g, G, lanes, ids.data()));
          for (uint32_t b0 = 0; b0 < nl; b0 += B) {
            const uint32_t nb = std::min(B, nl - b0);
            std::vector<std::string> texts(nb); std::vector<const char*> ptr(nb); std::vector<size_t> sz(nb); std::vector<uint32_t> hm(nb, o.hard_min);
            std::vector<uint64_t> flat((size_t)nb * P);
            for (uint32_t i = 0; i < nb; i++) {
              const uint64_t s = (uint64_t)g * nl + b0 + i;
              if (s < N) {
                if (R.samples[s].files.size() != 1) throw Error("--devices: one strict 4-line FASTQ file per sample");
                texts[i] = read_all(R.samples[s].files[0]); hm[i] = R.hard_min((uint32_t)s);
                if (texts[i].size() >= 0xFFFFFFF0ULL) throw Error("--devices: sample files must be < 4 GiB uncompressed");
              }
              ptr[i] = texts[i].data(); sz[i] = texts[i].size();
            }
            // every rank makes the same sequence of calls, also when one of them has failed: the library reports a failed peer
            int rc = kmx_dist_run_batch(ctx, nb, ptr.data(), sz.data(), 0, hm.data(), b0, nl, flat.data());
            if (rc) throw Error(std::string("kmx_dist_run_batch: ") + kmx_last_error(ctx));
            for (uint32_t i = 0; i < nb; i++) { const uint64_t s = (uint64_t)g * nl + b0 + i; if (s < N) std::copy(flat.begin() + (size_t)i * P, flat.begin() + (size_t)(i + 1) * P, R.pinfo[s].begin()); }
          }
          KX(kmx_sync(ctx));
        } catch (...) { R.fail(std::current_exception()); }
      };
      std::vector<std::thread> th;
      for (int g = 0; g < G; g++) th.emplace_back(work, g);
      for (auto& x : th) x.join();
      if (R.err) std::rethrow_exception(R.err);
    }
    for (uint32_t s = 0; s < N; s++) {        // partition_infos/<id>.pinfo (gatb_utils.hpp:46-51)
      std::ofstream pi(o.dir + "/partition_infos/" + R.samples[s].id + ".pinfo");
      for (uint32_t p = 0; p < P; p++) pi << R.pinfo[s][p] << "\n";
    }
    // ---- counts/ files (kept with --keep-tmp or when stopping at count)
    if (o.until != "superk" && (o.keep_tmp || o.until == "count")) for (const Rank& rk : ranks) write_counts(R, rk);
    // ---- merge: every rank its own partitions, concurrently
    if (o.until == "all" || o.until == "merge") {
      PluginHost plug;
      if (!o.plugin.empty()) plug.load(o.plugin, o.plugin_config, o.k);
      std::vector<int> bf_fds;
      if (o.what == "bft") {                               // filters/<id>.bf: HowDeSBT header, u64 number of bits, then the partitions' rows
        const std::string hd = howde_header(o.k, W * P);
        const uint64_t bits = W * P;
        for (uint32_t s = 0; s < N; s++) {
          const std::string path = o.dir + "/filters/" + R.samples[s].id + ".bf";
          int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
          if (fd < 0) throw Error("Unable to write at " + path);
          if (pwrite(fd, hd.data(), hd.size(), 0) != (ssize_t)hd.size() || pwrite(fd, &bits, 8, HOWDE_HEADER_BYTES) != 8) throw Error("Unable to write at " + path);
          bf_fds.push_back(fd);
        }
      }
      std::vector<std::thread> th;
      for (int g = 0; g < G; g++) th.emplace_back([&, g] { try { merge_partitions(R, ranks[g], plug, bf_fds); } catch (...) { R.fail(std::current_exception()); } });
      for (auto& x : th) x.join();
      for (int fd : bf_fds) close(fd);
      if (R.err) std::rethrow_exception(R.err);
    }
    {
      double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      uint64_t launches = 0; for (const Rank& rk : ranks) launches += kmx_launch_count(rk.ctx);
      std::ofstream ri(o.dir + "/run_infos.txt");
      ri << "Time: " << (long)sec << " seconds\n" << "GPU kernels launched: " << launches << "\n";
    }
    for (Rank& rk : ranks) kmx_destroy(rk.ctx);
    return EXIT_SUCCESS;
  } catch (const std::exception& e) {
    std::cerr << "[error] " << e.what() << "\n";      // reference: spdlog::error + exit(EXIT_FAILURE), src/kmtricks.cpp:109-123
    for (Rank& rk : ranks) if (rk.ctx) kmx_destroy(rk.ctx);
    return EXIT_FAILURE;
  }
}
