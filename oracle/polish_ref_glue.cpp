// TEST INFRASTRUCTURE ONLY.  C entry point around the REFERENCE's own fill_bfs
// (subprojects/goldpolish/src/utils.cpp:96-123, compiled unmodified from /root/reference against
// the stand-in btllib headers in shim_polish/), set up the way serve_batch sets its filters up
// (goldpolish_targeted_bfs.cpp:68-77) and called read by read in the given order (:131-135).
// Built into oracle/_ref/libgoldpolish_ref.so by oracle/Makefile when /root/reference is present.
#include "utils.hpp"

#include <cstring>

extern "C" int
grbp_ref_fill(const char* seqs, const uint64_t* off, const uint32_t* thresholds, uint64_t n_reads,
              unsigned hash_num, const unsigned* k_values, unsigned n_k, size_t cbf_bytes, size_t bf_bytes,
              uint8_t* out_bfs)
{
  std::vector<unsigned> ks(k_values, k_values + n_k);
  std::vector<std::unique_ptr<btllib::KmerCountingBloomFilter8>> cbfs;
  std::vector<std::unique_ptr<btllib::KmerBloomFilter>> bfs;
  for (const auto k : ks) {
    cbfs.push_back(std::unique_ptr<btllib::KmerCountingBloomFilter8>(
      new btllib::KmerCountingBloomFilter8(cbf_bytes, hash_num, k)));
    bfs.push_back(
      std::unique_ptr<btllib::KmerBloomFilter>(new btllib::KmerBloomFilter(bf_bytes, hash_num, k)));
  }
  for (uint64_t r = 0; r < n_reads; ++r) {
    fill_bfs(seqs + off[r], off[r + 1] - off[r], hash_num, ks, thresholds[r], cbfs, bfs);
  }
  for (unsigned i = 0; i < n_k; ++i) {
    memcpy(out_bfs + (size_t)i * bf_bytes, bfs[i]->raw().data(), bf_bytes);
  }
  return 0;
}

// ---- the input side of the builder, by the reference's own classes ----
// goldpolish-index (goldpolish_index.cpp:13-14)
#include "mappings.hpp"
#include "seqindex.hpp"

extern "C" int
grbp_ref_index_build(const char* seqs_path, const char* index_path)
{
  SeqIndex index{ std::string(seqs_path) };
  index.save(index_path);
  return 0;
}

// the reference's serve_batch (goldpolish_targeted_bfs.cpp:53-149, compiled unmodified, its main()
// renamed on the command line); declared here because the file has no header
void
serve_batch(const SeqIndex& target_seqs_index, const SeqIndex& mapped_seqs_index, const AllMappings& all_mappings,
            const size_t cbf_bytes, const size_t bf_bytes, const std::string& batch_name,
            const std::string& target_ids_input_pipe, const std::string& bfs_ready_pipe,
            const std::vector<std::string>& bf_names, const unsigned hash_num,
            const std::vector<unsigned>& k_values, const double subsample_max_mapped_seqs_per_target_10kbp);

// What goldpolish-targeted-bfs does for a list of batches, with regular files where it has named
// pipes: loads both indexes and the mappings as its main() does (:271-281, thresholds :35-36), then
// for every batch b hands serve_batch the file `<work>/b<b>-target_ids_input` (target ids, then
// "x").  serve_batch writes `<work>/b<b>-k<k>.bf` through the stand-in KmerBloomFilter::save (raw
// bit array).
extern "C" int
grbp_ref_serve_batches(const char* target_seqs, const char* target_index, const char* mappings,
                       const char* mapped_seqs, const char* mapped_index, double mx_max_mapped_seqs_per_target_10kbp,
                       double subsample_max_mapped_seqs_per_target_10kbp, unsigned hash_num,
                       const unsigned* k_values, unsigned n_k, size_t cbf_bytes, size_t bf_bytes,
                       const char* work_dir, unsigned n_batches)
{
  std::vector<unsigned> ks(k_values, k_values + n_k);
  std::vector<std::string> bf_names;
  for (const auto k : ks) {
    bf_names.push_back("k" + std::to_string(k) + ".bf"); // :208-211
  }
  SeqIndex target_seqs_index(target_index, target_seqs);
  SeqIndex mapped_seqs_index(mapped_index, mapped_seqs);
  AllMappings all_mappings(mappings, target_seqs_index, 1, 30, mx_max_mapped_seqs_per_target_10kbp);
  for (unsigned b = 0; b < n_batches; ++b) {
    const std::string batch_name = std::string(work_dir) + "/b" + std::to_string(b);
    serve_batch(target_seqs_index, mapped_seqs_index, all_mappings, cbf_bytes, bf_bytes, batch_name,
                batch_name + "-target_ids_input", batch_name + "-bfs_ready", bf_names, hash_num, ks,
                subsample_max_mapped_seqs_per_target_10kbp);
  }
  return 0;
}

// AllMappings as main() builds it (:275-281): the ids kept for one target, '\n' after each, into buf
extern "C" long
grbp_ref_mappings(const char* target_seqs, const char* target_index, const char* mappings,
                  double mx_max_mapped_seqs_per_target_10kbp, const char* target_id, char* buf, size_t cap)
{
  SeqIndex target_seqs_index(target_index, target_seqs);
  AllMappings all_mappings(mappings, target_seqs_index, 1, 30, mx_max_mapped_seqs_per_target_10kbp);
  const auto& ids = all_mappings.get_mappings(target_id);
  size_t at = 0;
  for (const auto& id : ids) {
    if (at + id.size() + 2 <= cap) {
      memcpy(buf + at, id.data(), id.size());
      buf[at + id.size()] = '\n';
      buf[at + id.size() + 1] = 0;
    }
    at += id.size() + 1;
  }
  return (long)ids.size();
}
