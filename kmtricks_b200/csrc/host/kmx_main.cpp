// kmx -- C++ host for the kmtricks hot path on B200: keeps the `kmtricks pipeline` command line
// (reference src/cli.cpp:117-382), the run-directory layout (include/kmtricks/kmdir.hpp:195-236),
// the file headers (include/kmtricks/io/*.hpp) and the IMergePlugin surface
// (include/kmtricks/plugin.hpp, plugin_manager.hpp), and calls libkmx_sm100.so through the C ABI
// (include/kmx.h) for every compute stage.  No compute happens on the host.
#include <kmx.h>

#include <dlfcn.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
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
  bool keep_tmp = false, static_repart = true;
  int device = 0;
  std::string key_kind, what, fmt;
};

[[noreturn]] void usage(const char* why)
{
  if (why) std::cerr << "kmx: " << why << "\n";
  std::cerr << "usage: kmx pipeline --file <fof> --run-dir <dir> --nb-partitions <P> [--kmer-size 31]\n"
               "           [--mode <kmer|hash>:<count|pa|bf|bft>:bin] [--hard-min 2] [--soft-min 1] [--recurrence-min 1]\n"
               "           [--share-min 0] [--minimizer-size 10] [--bloom-size 10000000] [--static-repart | --repart-from <run-dir>]\n"
               "           [--until all|superk|count|merge] [--keep-tmp] [--threads 4] [--plugin lib.so [--plugin-config s]] [--device 0]\n";
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
    else if (a == "--repart-from") { o.repart_from = need(i); o.static_repart = false; }
    else if (a == "--until") o.until = need(i);
    else if (a == "--keep-tmp") o.keep_tmp = true;
    else if (a == "--threads" || a == "-t") o.threads = std::stoul(need(i));
    else if (a == "--plugin") o.plugin = need(i);
    else if (a == "--plugin-config") o.plugin_config = need(i);
    else if (a == "--device") o.device = std::stoi(need(i));
    else if (a == "--help" || a == "-h") usage(nullptr);
    else usage(("unknown option " + a).c_str());
  }
  if (o.fof.empty() || o.dir.empty()) usage("--file and --run-dir are required");
  if (o.P == 0) usage("--nb-partitions is required (the repartition map must be fixed, SURVEY F9)");
  std::stringstream ss(o.mode); std::getline(ss, o.key_kind, ':'); std::getline(ss, o.what, ':'); std::getline(ss, o.fmt, ':');
  if ((o.key_kind != "kmer" && o.key_kind != "hash") || (o.what != "count" && o.what != "pa" && o.what != "bf" && o.what != "bft") || (o.fmt != "bin" && !o.fmt.empty()))
    usage("--mode must be <kmer|hash>:<count|pa|bf|bft>:bin");
  if ((o.what == "bf" || o.what == "bft") && o.key_kind != "hash") usage("bf/bft need hash keys");
  if (o.until != "all" && o.until != "superk" && o.until != "count" && o.until != "merge") usage("--until must be all|superk|count|merge");
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

}  // namespace

int main(int argc, char** argv)
{
  const auto t0 = std::chrono::steady_clock::now();
  kmx_ctx* ctx = nullptr;
  try {
    if (argc == 3 && std::string(argv[1]) == "fof") {     // host-side check, no device: prints the parsed input list
      for (const Sample& s : read_fof(argv[2])) {
        std::cout << s.id << "\t" << s.hard_min << "\t";
        for (size_t i = 0; i < s.files.size(); i++) std::cout << (i ? ";" : "") << s.files[i];
        std::cout << "\n";
      }
      return EXIT_SUCCESS;
    }
    Options o = parse(argc, argv);
    std::vector<Sample> samples = read_fof(o.fof);
    const uint32_t N = (uint32_t)samples.size(), P = o.P, w = (o.k + 31) / 32;
    const bool hash = o.key_kind == "hash";
    const uint64_t W = window_bits(o.bloom, P);
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
         << ", plugin=" << o.plugin << ", plugin_config=" << o.plugin_config << ", mode=" << o.what << ", format=bin, count_format=" << o.key_kind
         << ", until=" << o.until << ", engine=kmx_sm100\n";
    }
    { std::string h; put<uint64_t>(h, W * P); put<uint64_t>(h, P); put<uint64_t>(h, W); put<uint64_t>(h, W / 8); put<uint32_t>(h, o.m); write_file(o.dir + "/hash.info", h, nullptr, 0); }
    // ---- repartition table (RepartTask, task.hpp:170-222)
    const size_t tn = (size_t)1 << (2 * o.m);
    std::vector<uint16_t> table(tn);
    if (o.static_repart) for (size_t x = 0; x < tn; x++) table[x] = (uint16_t)(xxh64_u32((uint32_t)x) % P);
    else {
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

    kmx_params prm{};
    prm.kmer_size = o.k; prm.minim_size = o.m; prm.nb_partitions = P; prm.key_kind = hash ? KMX_KEY_HASH : KMX_KEY_KMER;
    prm.window_bits = hash ? W : 0; prm.repart_table = table.data(); prm.nb_samples = N;
    if (int rc = kmx_create(o.device, &prm, &ctx)) throw Error(std::string("kmx_create: ") + (ctx ? kmx_last_error(ctx) : "failed") + " (code " + std::to_string(rc) + ")");

    // ---- superk + count (TaskScheduler::exec_superk_count, task_scheduler.hpp:251-348)
    // fast path: every sample is one plain strict FASTQ file -> kmx_run_samples over pinned buffers
    std::vector<std::vector<uint64_t>> pinfo(N, std::vector<uint64_t>(P, 0));
    std::vector<std::string> texts(N);
    bool all_fastq = true;
    for (uint32_t s = 0; s < N; s++) {
      if (samples[s].files.size() != 1) { all_fastq = false; break; }
      texts[s] = read_all(samples[s].files[0]);
      if (texts[s].empty() || texts[s][0] != '@') { all_fastq = false; break; }
    }
    bool done = false;
    if (all_fastq && o.until != "superk") {
      std::vector<const char*> ptr(N); std::vector<size_t> nb(N); std::vector<uint32_t> hm(N); std::vector<uint64_t> flat((size_t)N * P);
      for (uint32_t s = 0; s < N; s++) { ptr[s] = texts[s].data(); nb[s] = texts[s].size(); hm[s] = samples[s].hard_min ? samples[s].hard_min : o.hard_min; }
      int rc = kmx_run_samples(ctx, N, ptr.data(), nb.data(), 0, nullptr, hm.data(), std::max(1u, std::min(o.threads, 8u)), flat.data());
      if (rc == KMX_OK) { for (uint32_t s = 0; s < N; s++) std::copy(flat.begin() + (size_t)s * P, flat.begin() + (size_t)(s + 1) * P, pinfo[s].begin()); done = true; }
      else if (rc != KMX_ERR_FORMAT) throw Error(std::string("kmx_run_samples: ") + kmx_last_error(ctx));
      else KX(kmx_reset(ctx));
    }
    if (!done) {
      for (uint32_t s = 0; s < N; s++) {
        KX(kmx_superk_begin(ctx));
        for (const std::string& f : samples[s].files) {
          std::string text = read_all(f);
          int rc = (!text.empty() && text[0] == '@') ? kmx_superk_push_fastq(ctx, text.data(), text.size(), 0) : KMX_ERR_FORMAT;
          if (rc == KMX_ERR_FORMAT) {
            std::string seqs; std::vector<uint64_t> off; parse_fastx(text, seqs, off);
            KX(kmx_superk_push_reads(ctx, seqs.data(), off.data(), off.size() - 1));
          } else if (rc) throw Error(std::string("kmx_superk_push_fastq: ") + kmx_last_error(ctx));
        }
        KX(kmx_superk_end(ctx, pinfo[s].data()));
        if (o.until != "superk") KX(kmx_count_sample(ctx, s, samples[s].hard_min ? samples[s].hard_min : o.hard_min));
      }
    }
    for (uint32_t s = 0; s < N; s++) {        // partition_infos/<id>.pinfo (gatb_utils.hpp:46-51)
      std::ofstream pi(o.dir + "/partition_infos/" + samples[s].id + ".pinfo");
      for (uint32_t p = 0; p < P; p++) pi << pinfo[s][p] << "\n";
    }
    // ---- counts/ files (kept with --keep-tmp or when stopping at count: kmer_file.hpp:31-108, hash_file.hpp:31-131)
    if (o.until != "superk" && (o.keep_tmp || o.until == "count")) {
      for (uint32_t s = 0; s < N; s++) for (uint32_t p = 0; p < P; p++) {
        uint64_t n = 0; KX(kmx_counts_size(ctx, s, p, &n));
        const uint32_t kw = hash ? 1 : w;
        std::vector<uint64_t> keys(n * kw); std::vector<uint32_t> cnt(n);
        KX(kmx_counts_get(ctx, s, p, keys.data(), cnt.data()));
        std::string path = o.dir + "/counts/partition_" + std::to_string(p) + "/" + samples[s].id + (hash ? ".hash" : ".kmer");
        std::string h = km_header();
        std::string body;
        if (!hash) {
          put<uint64_t>(h, 0x72656d6bULL); put<uint32_t>(h, o.k); put<uint32_t>(h, w); put<uint32_t>(h, 4); put<uint32_t>(h, s); put<uint32_t>(h, p);
          body.reserve(n * (8 * w + 4));
          for (uint64_t i = 0; i < n; i++) { body.append((const char*)&keys[i * w], 8 * w); body.append((const char*)&cnt[i], 4); }
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
    // ---- merge (KmerMergeTask / HashMergeTask, task.hpp:690-870)
    if (o.until == "all" || o.until == "merge") {
      PluginHost plug;
      if (!o.plugin.empty()) { if (o.what != "count") throw Error("--plugin needs a count matrix mode"); plug.load(o.plugin, o.plugin_config, o.k); }
      std::vector<uint32_t> soft(N, o.soft_min);
      const char* ext = !hash ? (o.what == "count" ? "count" : "pa") : (o.what == "count" ? "count_hash" : o.what == "pa" ? "pa_hash" : "cmbf");
      const uint32_t nbytes = (N + 7) / 8;
      for (uint32_t p = 0; p < P; p++) {
        kmx_merge_params mp{soft.data(), o.rec_min, o.share_min,
                            (uint32_t)(o.what == "count" ? KMX_FMT_COUNT : o.what == "pa" ? KMX_FMT_PA : o.what == "bf" ? KMX_FMT_BF : KMX_FMT_BFT),
                            (uint32_t)(plug.handle ? 1 : 0)};
        kmx_merge_result r{};
        KX(kmx_merge_partition(ctx, p, &mp, &r));
        std::vector<uint8_t> body(r.n_rows * r.row_bytes), keep(plug.handle ? r.n_rows : 0);
        std::vector<uint64_t> stats((size_t)6 * N);
        KX(kmx_merge_get(ctx, body.data(), stats.data(), plug.handle ? keep.data() : nullptr));
        std::string h = km_header();
        if (!hash && o.what == "count") { put<uint64_t>(h, 0x6b5f78697274616dULL); put<uint32_t>(h, o.k); put<uint32_t>(h, w); put<uint32_t>(h, 1); put<uint32_t>(h, N); put<uint32_t>(h, 0); put<uint32_t>(h, 0); }
        else if (!hash) { put<uint64_t>(h, 0x6b5f74616d6170ULL); put<uint32_t>(h, o.k); put<uint32_t>(h, w); put<uint32_t>(h, N); put<uint32_t>(h, nbytes); put<uint32_t>(h, 0); put<uint32_t>(h, 0); }
        else if (o.what == "count") { put<uint64_t>(h, 0x685f78697274616dULL); put<uint32_t>(h, 4); put<uint32_t>(h, N); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
        else if (o.what == "pa") { put<uint64_t>(h, 0x685f74616d6170ULL); put<uint32_t>(h, N); put<uint32_t>(h, nbytes); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
        else { put<uint64_t>(h, 0x74616d746962ULL); put<uint32_t>(h, N); put<uint64_t>(h, W * p); put<uint64_t>(h, W); put<uint32_t>(h, 0); put<uint32_t>(h, p); }
        const std::string mpath = o.dir + "/matrices/matrix_" + std::to_string(p) + "." + ext;
        if (!plug.handle) write_file(mpath, h, body.data(), body.size());
        else {
          // one plugin instance per merge task (task.hpp:701-712), called once per merged row in key order;
          // its return value replaces the keep decision and it may edit the counts (merge.hpp:249-257)
          km::IMergePlugin* pl = plug.create();
          pl->configure(plug.config);
          pl->set_out_dir(o.dir + "/plugin_output");
          pl->set_kmer_size(hash ? 0 : o.k);
          pl->set_partition(p);
          const uint32_t kw = hash ? 1 : w;
          std::string out; std::vector<uint32_t> c(N);
          for (uint64_t i = 0; i < r.n_rows; i++) {
            const uint8_t* row = &body[i * r.row_bytes];
            memcpy(c.data(), row + 8 * kw, 4 * N);
            const uint64_t* key = reinterpret_cast<const uint64_t*>(row);
            uint64_t kbuf[2]; memcpy(kbuf, key, 8 * kw);
            bool kp = hash ? pl->process_hash(kbuf[0], c) : pl->process_kmer(kbuf, c);
            if (kp) { out.append((const char*)kbuf, 8 * kw); out.append((const char*)c.data(), 4 * N); }
          }
          plug.destroy(pl);
          write_file(mpath, h, out.data(), out.size());
        }
        {   // merge_infos/partitionP.merge_info (merge.hpp:72-83)
          static const char* names[6] = {"NON_SOLID", "RESCUED", "UNIQUE_WO_RESCUE", "UNIQUE_W_RESCUE", "TOTAL_WO_RESCUE", "TOTAL_W_RESCUE"};
          std::ofstream mi(o.dir + "/merge_infos/partition" + std::to_string(p) + ".merge_info");
          for (int q = 0; q < 6; q++) { mi << names[q] << '\t'; for (uint32_t s = 0; s < N; s++) mi << stats[(size_t)q * N + s] << '\t'; mi << "\n"; }
        }
        if (o.what == "bf") {   // fpr/partition_P.txt (task.hpp:849-860; utils.hpp:239-243) -- the only floating point on the path
          std::ofstream fp(o.dir + "/fpr/partition_" + std::to_string(p) + ".txt");
          for (uint32_t s = 0; s < N; s++) {
            double nn = (double)stats[(size_t)3 * N + s];
            double fpr = std::pow(1.0 - std::pow(std::exp(1.0), (-(1.0 * nn)) / (double)W), 1.0);
            fp << std::fixed << fpr << "\n";
          }
        }
      }
    }
    {
      double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::ofstream ri(o.dir + "/run_infos.txt");
      ri << "Time: " << (long)sec << " seconds\n" << "GPU kernels launched: " << kmx_launch_count(ctx) << "\n";
    }
    kmx_destroy(ctx);
    return EXIT_SUCCESS;
  } catch (const std::exception& e) {
    std::cerr << "[error] " << e.what() << "\n";      // reference: spdlog::error + exit(EXIT_FAILURE), src/kmtricks.cpp:109-123
    if (ctx) kmx_destroy(ctx);
    return EXIT_FAILURE;
  }
}
