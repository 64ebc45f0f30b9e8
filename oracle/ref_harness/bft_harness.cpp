// TEST INFRASTRUCTURE ONLY -- drives the reference's own HashMerger::write_as_bft / write_as_bf
// (include/kmtricks/merge.hpp:575-644), which the CLI cannot reach at this commit.
// Compiled against the headers under /root/reference by oracle/build_ref.sh into oracle/_ref/bin.
// usage: bft_harness <bf|bft> <out> <lower> <upper> <soft_min> <r_min> <save_if> <a.hash> [b.hash ...]
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#define WITH_PLUGIN
#include <kmtricks/merge.hpp>

int main(int argc, char** argv)
{
  if (argc < 9) { std::cerr << "usage: bft_harness <bf|bft> out lower upper soft_min r_min save_if files...\n"; return 2; }
  std::string what = argv[1], out = argv[2];
  uint64_t lower = std::strtoull(argv[3], nullptr, 10), upper = std::strtoull(argv[4], nullptr, 10);
  uint32_t soft = std::atoi(argv[5]), rmin = std::atoi(argv[6]), save_if = std::atoi(argv[7]);
  std::vector<std::string> paths(argv + 8, argv + argc);
  std::vector<uint32_t> amin(paths.size(), soft);
  km::HashMerger<DMAX_C, 32768, km::HashReader<DMAX_C, 32768>> m(paths, amin, rmin, save_if);
  if (what == "bft") m.write_as_bft(out, lower, upper, false);
  else m.write_as_bf(out, lower, upper, false);
  return 0;
}
