// (f4) GoldPolish targeted Bloom filters, the builder's input side: the sequence index
// (subprojects/goldpolish/src/seqindex.cpp), the mappings of reads to targets (mappings.cpp: ntLink
// triples, SAM, PAF) and what serve_batch does with them for the target ids of a batch
// (goldpolish_targeted_bfs.cpp:84-136) up to the point where k-mers are hashed -- that part is
// grb_polish_fill_batches on the GPU.  Host C++; the named pipes and the .bf files stay with the caller.
#include "goldrush_b200.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

int
fail(char* err, size_t cap, int code, const std::string& msg)
{
  if (err && cap) {
    snprintf(err, cap, "%s", msg.c_str());
  }
  return code;
}

bool
read_file(const char* path, std::string* out)
{
  FILE* f = fopen(path, "rb");
  if (!f) {
    return false;
  }
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) {
    out->append(buf, n);
  }
  fclose(f);
  return true;
}

// operator>> of a stream into a std::string: runs of non-space characters
struct Tokens
{
  const std::string& s;
  size_t at = 0;
  explicit Tokens(const std::string& str)
    : s(str)
  {
  }
  bool next(std::string* tok)
  {
    while (at < s.size() && isspace((unsigned char)s[at])) {
      ++at;
    }
    if (at >= s.size()) {
      return false;
    }
    const size_t b = at;
    while (at < s.size() && !isspace((unsigned char)s[at])) {
      ++at;
    }
    tok->assign(s, b, at - b);
    return true;
  }
};

bool
to_u64(const std::string& t, uint64_t* v) // std::stoul: a number, then anything
{
  errno = 0;
  char* end = nullptr;
  *v = strtoull(t.c_str(), &end, 10);
  return end != t.c_str() && errno == 0;
}

bool
to_double(const std::string& t, double* v) // std::stod
{
  errno = 0;
  char* end = nullptr;
  *v = strtod(t.c_str(), &end);
  return end != t.c_str();
}

struct SeqEntry
{
  uint64_t start, len;
  double phred_avg;
};

struct Index
{
  std::unordered_map<std::string, SeqEntry> seqs;
  // SeqIndex::SeqIndex(index_filepath, seqs_filepath), seqindex.cpp:86-123: whitespace-separated
  // quadruples; a repeated id keeps its first entry (emplace)
  int load(const char* path, char* err, size_t cap)
  {
    std::string text;
    if (!read_file(path, &text)) {
      return fail(err, cap, GRB_ERR_ARG, std::string("cannot read index ") + path);
    }
    Tokens tk(text);
    std::string tok, id;
    SeqEntry e{ 0, 0, 0.0 };
    for (uint64_t i = 0; tk.next(&tok); ++i) {
      bool ok = true;
      switch (i % 4) {
        case 0:
          id = tok;
          break;
        case 1:
          ok = to_u64(tok, &e.start);
          break;
        case 2:
          ok = to_u64(tok, &e.len);
          break;
        default:
          ok = to_double(tok, &e.phred_avg);
          seqs.emplace(id, e);
      }
      if (!ok) {
        return fail(err, cap, GRB_ERR_ARG, std::string(path) + ": not a number: " + tok);
      }
    }
    return GRB_OK;
  }
};

struct Mappings
{
  std::unordered_map<std::string, std::vector<std::string>> by_target;
  std::unordered_map<std::string, std::unordered_set<std::string>> inserted;
  std::unordered_map<std::string, std::vector<unsigned>> mx;

  // AllMappings::load_mapping, mappings.cpp:33-69: targets outside the index are dropped, a read is
  // listed once per target (its first line wins)
  void add(const std::string& mapped, const std::string& target, const Index& targets, unsigned m)
  {
    if (targets.seqs.find(target) == targets.seqs.end()) {
      return;
    }
    if (inserted[target].insert(mapped).second) {
      by_target[target].push_back(mapped);
      mx[target].push_back(m);
    }
  }

  // load_ntlink, mappings.cpp:71-107
  int load_ntlink(const std::string& text, const Index& targets, unsigned mx_min, char* err, size_t cap)
  {
    Tokens tk(text);
    std::string tok, mapped, target;
    for (uint64_t i = 0; tk.next(&tok); ++i) {
      if (i % 3 == 0) {
        mapped = tok;
      } else if (i % 3 == 1) {
        target = tok;
      } else {
        uint64_t m = 0;
        if (!to_u64(tok, &m)) {
          return fail(err, cap, GRB_ERR_ARG, "mappings: not a minimizer count: " + tok);
        }
        if (m >= mx_min) {
          add(mapped, target, targets, (unsigned)m);
        }
      }
    }
    return GRB_OK;
  }

  // load_sam (columns 1 and 3) and load_paf (columns 1 and 6), mappings.cpp:109-165,167-224: header
  // lines start with '@'; a line too short to hold the column keeps the id of the line before it
  void load_columns(const std::string& text, const Index& targets, unsigned target_col)
  {
    std::string mapped, target, tok;
    size_t b = 0;
    while (b < text.size()) {
      size_t e = text.find('\n', b);
      e = e == std::string::npos ? text.size() : e + 1;
      if (text[b] != '@') {
        const std::string line(text, b, e - b);
        Tokens tk(line);
        for (unsigned col = 1; tk.next(&tok); ++col) {
          if (col == 1) {
            mapped = tok;
          } else if (col == target_col) {
            target = tok;
          }
        }
        add(mapped, target, targets, 0);
      }
      b = e;
    }
  }

  // AllMappings::filter, mappings.cpp:226-320: per target, the smallest minimizer threshold in
  // [mx_min, mx_max] that leaves at most ceil(len * per_10kbp / 10000) reads
  int filter(double per_10kbp, unsigned mx_min, unsigned mx_max, const Index& targets, char* err, size_t cap)
  {
    if (per_10kbp <= 0) {
      return fail(err, cap, GRB_ERR_ARG, "filter: max_mapped_seqs_per_target_10kbp is not positive.");
    }
    for (auto& tm : by_target) {
      std::vector<std::string>& maps = tm.second;
      if (maps.empty()) {
        continue;
      }
      const std::vector<unsigned>& m = mx.at(tm.first);
      const auto it = targets.seqs.find(tm.first);
      if (it == targets.seqs.end()) {
        continue;
      }
      const int max_mapped = (int)std::ceil(double(it->second.len) * per_10kbp / 10000.0);
      if (max_mapped <= 0) {
        return fail(err, cap, GRB_ERR_ARG, "filter: max_mapped_seqs <= 0.");
      }
      auto count_at = [&](int thr) {
        int n = 0;
        for (unsigned v : m) {
          n += v >= (unsigned)thr;
        }
        return n;
      };
      int lo = (int)mx_min, hi = (int)mx_max, thr;
      if ((int)maps.size() <= max_mapped) {
        thr = lo;
      } else if (count_at(hi) > max_mapped) {
        thr = hi;
      } else {
        while (hi - lo > 1) {
          const int mid = (hi + lo) / 2;
          if (count_at(mid) > max_mapped) {
            lo = mid;
          } else {
            hi = mid;
          }
        }
        thr = hi;
      }
      std::vector<std::string> kept;
      for (size_t i = 0; i < maps.size(); ++i) {
        if ((int)m[i] >= thr) {
          kept.push_back(maps[i]);
        }
      }
      maps.swap(kept);
    }
    return GRB_OK;
  }
};

bool
ends_with(const std::string& s, const char* suffix)
{
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

} // namespace

struct grb_polish_inputs
{
  Index targets, mapped;
  Mappings maps;
  std::string mapped_seqs_path;
  int fd = -1;
};

namespace {

// What serve_batch prepares for its batch before fill_bfs runs (goldpolish_targeted_bfs.cpp:84-136):
// per target id the mapped reads sorted by (Phred average as size_t descending, id ascending), cut to
// min(n, size_t(len * subsample / 10000)), the k-mer threshold from their summed lengths; the reads
// of all targets of a batch in that order.
struct Work
{
  std::vector<uint64_t> batch_first; // reads
  std::vector<uint64_t> seq_off;
  std::vector<uint32_t> thresholds;
  std::vector<uint64_t> file_start;
  std::vector<char> seqs;
};

int
gather(const grb_polish_inputs* in, double subsample, uint32_t n_batches, const uint64_t* batch_first,
       const char* const* target_ids, Work* w, char* err, size_t cap)
{
  static const uint64_t max_seqlen = 20ull * 1024 * 1024; // seqindex.hpp:70,88-90
  w->batch_first.assign(1, 0);
  w->seq_off.assign(1, 0);
  for (uint32_t b = 0; b < n_batches; ++b) {
    for (uint64_t t = batch_first[b]; t < batch_first[b + 1]; ++t) {
      const std::string tid = target_ids[t];
      const auto ti = in->targets.seqs.find(tid);
      if (ti == in->targets.seqs.end()) { // the reference: std::out_of_range from .at(), :87
        return fail(err, cap, GRB_ERR_ARG, "target id not in the target index: " + tid);
      }
      const auto mi = in->maps.by_target.find(tid);
      if (mi == in->maps.by_target.end() || mi->second.empty()) {
        continue; // :89-92
      }
      const std::vector<std::string>& ids = mi->second;
      const uint32_t n = (uint32_t)ids.size();
      std::vector<const char*> id_ptr(n);
      std::vector<double> phred(n);
      std::vector<uint64_t> lens(n), starts(n);
      for (uint32_t i = 0; i < n; ++i) {
        const auto si = in->mapped.seqs.find(ids[i]);
        if (si == in->mapped.seqs.end()) { // the reference: .at() throws, :103
          return fail(err, cap, GRB_ERR_ARG, "mapped read not in the mapped-sequence index: " + ids[i]);
        }
        id_ptr[i] = ids[i].c_str();
        phred[i] = si->second.phred_avg;
        lens[i] = si->second.len;
        starts[i] = si->second.start;
      }
      std::vector<uint32_t> order(n);
      uint32_t n_used = 0;
      int32_t thr = 0;
      if (grb_polish_plan_target(ti->second.len, subsample, n, id_ptr.data(), phred.data(), lens.data(),
                                 order.data(), &n_used, &thr) != GRB_OK) {
        return fail(err, cap, GRB_ERR_ARG, "serve_batch: k-mer threshold must be >0."); // :126-127
      }
      for (uint32_t i = 0; i < n_used; ++i) {
        const uint32_t j = order[i];
        if (lens[j] >= max_seqlen) {
          return fail(err, cap, GRB_ERR_ARG, "get_seq: Seq size over buffer size: " + ids[j]);
        }
        w->file_start.push_back(starts[j]);
        w->seq_off.push_back(w->seq_off.back() + lens[j]);
        w->thresholds.push_back((uint32_t)thr);
      }
    }
    w->batch_first.push_back(w->thresholds.size());
  }
  // SeqIndex::get_seq, seqindex.hpp:92-99: seq_len bytes at seq_start of the sequence file
  w->seqs.resize(std::max<uint64_t>(w->seq_off.back(), 1));
  const uint64_t n_reads = w->thresholds.size();
  std::atomic<uint64_t> next{ 0 };
  std::atomic<int> bad{ 0 };
  auto worker = [&]() {
    for (;;) {
      const uint64_t r = next.fetch_add(1);
      if (r >= n_reads) {
        return;
      }
      uint64_t done = 0;
      const uint64_t len = w->seq_off[r + 1] - w->seq_off[r];
      while (done < len) {
        const ssize_t got = pread(in->fd, w->seqs.data() + w->seq_off[r] + done, len - done,
                                  (off_t)(w->file_start[r] + done));
        if (got <= 0) {
          bad.store(1);
          return;
        }
        done += (uint64_t)got;
      }
    }
  };
  const unsigned n_thr = (unsigned)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()),
                                                      std::max<uint64_t>(1, n_reads / 64));
  std::vector<std::thread> pool;
  for (unsigned i = 1; i < n_thr; ++i) {
    pool.emplace_back(worker);
  }
  worker();
  for (auto& th : pool) {
    th.join();
  }
  if (bad.load()) {
    return fail(err, cap, GRB_ERR_ARG, "get_seq: read did not read all bytes: " + in->mapped_seqs_path);
  }
  return GRB_OK;
}

} // namespace

extern "C" {

int
grb_polish_index_build(const char* seqs_path, const char* index_path, char* err, size_t err_cap)
{
  std::string text;
  if (!read_file(seqs_path, &text)) {
    return fail(err, err_cap, GRB_ERR_ARG, std::string("cannot read ") + seqs_path);
  }
  FILE* out = fopen(index_path, "wb");
  if (!out) {
    return fail(err, err_cap, GRB_ERR_ARG, std::string("cannot write ") + index_path);
  }
  const bool fastq = !text.empty() && text[0] == '@'; // seqindex.cpp:20
  std::unordered_set<std::string> seen;
  std::string id;
  uint64_t id_end = 0, seq_start = 0, seq_len = 0, byte = 0;
  int rc = GRB_OK;
  for (uint64_t i = 0; byte < text.size() && rc == GRB_OK; ++i) { // std::getline, :26
    size_t nl = text.find('\n', byte);
    nl = nl == std::string::npos ? text.size() : nl;
    const uint64_t end = nl;
    const char* line = text.data() + byte;
    const uint64_t len = end - byte;
    const unsigned phase = fastq ? i % 4 : (i % 2) * 1;
    if (phase == 0) { // :29-33,53-55: the id is the header up to the first blank (FASTQ: or tab)
      id_end = end;
      if (len == 0) {
        rc = fail(err, err_cap, GRB_ERR_ARG, std::string(seqs_path) + ": empty header line");
        break;
      }
      uint64_t n = 1;
      while (n < len && line[n] != ' ' && !(fastq && line[n] == '\t')) {
        ++n;
      }
      id.assign(line + 1, n - 1);
    } else if (phase == 1) { // :35-37,57-58
      seq_start = id_end + 1;
      seq_len = end - id_end - 1;
      if (!fastq && seen.insert(id).second) {
        fprintf(out, "%s\t%llu\t%llu\t%g\n", id.c_str(), (unsigned long long)seq_start,
                (unsigned long long)seq_len, 0.0);
      }
    } else if (fastq && phase == 3) { // :39-47: mean of (Q - 33) over all but the last character
      if (len == 0) {
        rc = fail(err, err_cap, GRB_ERR_ARG, std::string(seqs_path) + ": empty quality line");
        break;
      }
      const uint64_t n = len == 1 ? 1 : len - 1;
      uint64_t sum = 0;
      for (uint64_t q = 0; q < n; ++q) {
        sum += (uint64_t)(unsigned char)line[q] - 33;
      }
      if (seen.insert(id).second) {
        fprintf(out, "%s\t%llu\t%llu\t%g\n", id.c_str(), (unsigned long long)seq_start,
                (unsigned long long)seq_len, (double)sum / (double)n);
      }
    }
    byte = end + 1;
  }
  fclose(out);
  return rc;
}

int
grb_polish_inputs_open(const char* target_index_path, const char* mappings_path, const char* mapped_seqs_path,
                       const char* mapped_index_path, double mx_max_mapped_seqs_per_target_10kbp,
                       grb_polish_inputs** out, char* err, size_t err_cap)
{
  static const unsigned MX_THRESHOLD_MIN = 1, MX_THRESHOLD_MAX = 30; // goldpolish_targeted_bfs.cpp:35-36
  *out = nullptr;
  grb_polish_inputs* in = new grb_polish_inputs;
  int rc = in->targets.load(target_index_path, err, err_cap);
  if (rc == GRB_OK) {
    rc = in->mapped.load(mapped_index_path, err, err_cap);
  }
  std::string text;
  const std::string mp = mappings_path;
  if (rc == GRB_OK && ends_with(mp, ".bam")) {
    rc = fail(err, err_cap, GRB_ERR_ARG, "BAM mappings are not read here: pass `samtools view -h` output as .sam");
  }
  if (rc == GRB_OK && !read_file(mappings_path, &text)) {
    rc = fail(err, err_cap, GRB_ERR_ARG, "cannot read mappings " + mp);
  }
  if (rc == GRB_OK) { // AllMappings::AllMappings, mappings.cpp:12-31
    if (ends_with(mp, ".sam")) {
      in->maps.load_columns(text, in->targets, 3);
    } else if (ends_with(mp, ".paf")) {
      in->maps.load_columns(text, in->targets, 6);
    } else {
      rc = in->maps.load_ntlink(text, in->targets, MX_THRESHOLD_MIN, err, err_cap);
      if (rc == GRB_OK) {
        rc = in->maps.filter(mx_max_mapped_seqs_per_target_10kbp, MX_THRESHOLD_MIN, MX_THRESHOLD_MAX,
                             in->targets, err, err_cap);
      }
    }
    in->maps.inserted.clear();
    in->maps.mx.clear();
  }
  if (rc == GRB_OK) {
    in->mapped_seqs_path = mapped_seqs_path;
    in->fd = open(mapped_seqs_path, O_RDONLY);
    if (in->fd < 0) {
      rc = fail(err, err_cap, GRB_ERR_ARG, std::string("cannot open ") + mapped_seqs_path);
    } else {
      posix_fadvise(in->fd, 0, 0, POSIX_FADV_RANDOM); // seqindex.hpp:74-76
    }
  }
  if (rc != GRB_OK) {
    grb_polish_inputs_close(in);
    return rc;
  }
  *out = in;
  return GRB_OK;
}

void
grb_polish_inputs_close(grb_polish_inputs* in)
{
  if (!in) {
    return;
  }
  if (in->fd >= 0) {
    close(in->fd);
  }
  delete in;
}

int64_t
grb_polish_inputs_mappings(const grb_polish_inputs* in, const char* target_id, char* buf, size_t cap)
{
  const auto it = in->maps.by_target.find(target_id);
  if (it == in->maps.by_target.end()) {
    return 0; // AllMappings::get_mappings: EMPTY_MAPPINGS, mappings.cpp:322-329
  }
  size_t at = 0;
  for (const std::string& id : it->second) {
    if (buf && at + id.size() + 1 <= cap) {
      memcpy(buf + at, id.data(), id.size());
      buf[at + id.size()] = '\n';
    }
    at += id.size() + 1;
  }
  if (buf && cap) {
    buf[std::min(at, cap - 1)] = 0;
  }
  return (int64_t)it->second.size();
}

int
grb_polish_serve_batches(grb_ctx* ctx, const grb_polish_inputs* in, const grb_polish_params* p,
                         double subsample_max_mapped_seqs_per_target_10kbp, uint32_t n_batches,
                         const uint64_t* batch_first, const char* const* target_ids, uint8_t* out_bfs, char* err,
                         size_t err_cap)
{
  Work w;
  const int rc = gather(in, subsample_max_mapped_seqs_per_target_10kbp, n_batches, batch_first, target_ids, &w,
                        err, err_cap);
  if (rc != GRB_OK) {
    return rc;
  }
  const int rf = grb_polish_fill_batches(ctx, p, n_batches, w.batch_first.data(), w.seqs.data(), w.seq_off.data(),
                                         w.thresholds.data(), out_bfs);
  return rf == GRB_OK ? GRB_OK : fail(err, err_cap, rf, grb_last_error(ctx));
}

int
grb_test_polish_serve_batches_host(const grb_polish_inputs* in, const grb_polish_params* p,
                                   double subsample_max_mapped_seqs_per_target_10kbp, uint32_t n_batches,
                                   const uint64_t* batch_first, const char* const* target_ids, uint8_t* out_bfs,
                                   uint64_t* n_reads, uint64_t* n_bases, char* err, size_t err_cap)
{
  Work w;
  const int rc = gather(in, subsample_max_mapped_seqs_per_target_10kbp, n_batches, batch_first, target_ids, &w,
                        err, err_cap);
  if (rc != GRB_OK) {
    return rc;
  }
  if (n_reads) {
    *n_reads = w.thresholds.size();
  }
  if (n_bases) {
    *n_bases = w.seq_off.back();
  }
  return grb_test_polish_fill_host(p, n_batches, w.batch_first.data(), w.seqs.data(), w.seq_off.data(),
                                   w.thresholds.data(), out_bfs);
}

} // extern "C"
