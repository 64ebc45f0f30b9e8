// kmx_plugin_abi.hpp -- the in-process operator interface of the merge stage, binary compatible
// with kmtricks' km::IMergePlugin (reference include/kmtricks/plugin.hpp:12-30) so that plugin
// shared objects built against kmtricks load unchanged.  What makes up the ABI:
//   * vtable order: ~IMergePlugin (complete + deleting), set_out_dir, set_partition,
//     set_kmer_size, configure, process_kmer, process_hash
//   * data members, in order: std::string m_output_directory; size_t m_kmer_size; size_t m_partition
//   * count vector element type selectC<DMAX_C>::type -- uint32_t for the default DMAX_C
//   * C-linkage factory symbols (plugin_manager.hpp:38-90): int use_template();
//     km::IMergePlugin* create0() | create32() / create64() ...; void destroy(km::IMergePlugin*);
//     std::string plugin_name()
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace km {

class IMergePlugin
{
public:
  IMergePlugin() = default;
  virtual ~IMergePlugin() {}
  virtual void set_out_dir(const std::string& s) final { m_output_directory = s; }
  virtual void set_partition(size_t p) final { m_partition = p; }
  virtual void set_kmer_size(const size_t kmer_size) { m_kmer_size = kmer_size; }
  virtual void configure(const std::string&) {}
  virtual bool process_kmer(const uint64_t*, std::vector<uint32_t>&) { return true; }
  virtual bool process_hash(uint64_t, std::vector<uint32_t>&) { return true; }

protected:
  std::string m_output_directory;
  size_t m_kmer_size;
  size_t m_partition;
};

}  // namespace km
