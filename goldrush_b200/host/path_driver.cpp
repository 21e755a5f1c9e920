// grb_run_path — the GoldRush-Path stage on top of the engine's C ABI: what the reference's main()
// does between option parsing and exit (goldrush_path/goldrush_path.cpp:1096-1275), with the three
// file passes of the reference collapsed into one ingest:
//   ingest (K1) -> [ntcard (K5)] -> Phred median -> filters -> bit vector (K2+K4a) -> rank (K4b)
//   -> ordered selection (K2+K3+K4c) -> silver-path FASTQ / golden-path FASTA writers.
// Everything data-parallel runs on the device; this file only holds per-read bookkeeping
// (O(#reads)), glibc-exact scalar maths, name handling and output.
#include "goldrush_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <future>
#include <memory>
#include <thread>
#include <unistd.h>
#include <unordered_set>
#include <vector>

namespace {

struct Log
{
  bool on;
  void operator()(const char* fmt, ...) const
  {
    if (!on) {
      return;
    }
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
  }
};

struct Fnv
{
  uint64_t h = 1469598103934665603ull;
  void add(const char* p, size_t n)
  {
    for (size_t i = 0; i < n; ++i) {
      h ^= (unsigned char)p[i];
      h *= 1099511628211ull;
    }
  }
};

// Hash of one output record (grb_run_result.out_digest folds these in output order): FNV-1a over
// the record's bytes taken as little-endian 8-byte words (the last word zero-padded), then the
// length.  One multiply per 8 bytes: the byte-wise form hashed under 1 GB/s per thread, which made
// the record assembly -- not the GPU -- the slowest stage of a human-scale run on 4 cores per rank.
uint64_t
record_hash(const char* p, size_t n)
{
  uint64_t h = 1469598103934665603ull;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 1099511628211ull;
  }
  if (i < n) {
    uint64_t w = 0;
    memcpy(&w, p + i, n - i);
    h = (h ^ w) * 1099511628211ull;
  }
  return (h ^ (uint64_t)n) * 1099511628211ull;
}

struct OutFile
{
  FILE* f = nullptr;
  bool write = false;
  Fnv* digest = nullptr;
  void open(const std::string& path)
  {
    close();
    if (write) {
      f = fopen(path.c_str(), "wb");
    }
  }
  void close()
  {
    if (f) {
      fclose(f);
    }
    f = nullptr;
  }
};

// first byte >= from at which a FASTQ record starts: a line that begins with '@' and whose line
// after next begins with '+' (a quality line may begin with '@' too, but then the line after
// next is a sequence).  n if there is none.
size_t
next_record_start(const char* d, size_t n, size_t from)
{
  if (from == 0) {
    return 0;
  }
  size_t line = from;
  if (d[from - 1] != '\n') {
    const char* nl = (const char*)memchr(d + from, '\n', n - from);
    if (!nl) {
      return n;
    }
    line = (size_t)(nl - d) + 1;
  }
  while (line < n) {
    const char* e1 = (const char*)memchr(d + line, '\n', n - line);
    if (!e1) {
      return n;
    }
    const size_t l1 = (size_t)(e1 - d) + 1;
    if (d[line] == '@' && l1 < n) {
      const char* e2 = (const char*)memchr(d + l1, '\n', n - l1);
      if (e2 && (size_t)(e2 - d) + 1 < n && e2[1] == '+') {
        return line;
      }
    }
    line = l1;
  }
  return n;
}

// fn(i) for i in [lo, hi) on `threads` plain std::threads, work handed out `grain` items at a time.
// Used for the record assembly that runs beside pass 2 instead of an OpenMP region: libgomp's
// workers spin at the end of every region (GOMP_SPINCOUNT), and on a host with four cores per rank
// that spinning took the core of the thread that launches the kernels.
template<class F>
void
parallel_for(int threads, int64_t lo, int64_t hi, int64_t grain, F fn)
{
  if (threads <= 1 || hi - lo <= grain) {
    for (int64_t i = lo; i < hi; ++i) {
      fn(i);
    }
    return;
  }
  std::atomic<int64_t> next(lo);
  auto work = [&] {
    while (true) {
      const int64_t b = next.fetch_add(grain);
      if (b >= hi) {
        return;
      }
      for (int64_t i = b; i < std::min(hi, b + grain); ++i) {
        fn(i);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) {
    pool.emplace_back(work);
  }
  work();
  for (std::thread& t : pool) {
    t.join();
  }
}

// Two-stage call in slice mode: the joined silver paths are a sequence of parts (path q, rank r) --
// q in the order of `order` (the shell glob's), r in rank order; part (q, r) holds bytes[r * n_paths
// + q] bytes on rank r.  Each golden-stage rank gets a run of whole, consecutive parts of about
// total / ranks bytes: to[i] = receiving rank of the i-th part of that sequence (non-decreasing).
// Returns false if some rank would get nothing (the caller then gathers everything everywhere).
bool
plan_silver_parts(const std::vector<uint64_t>& bytes, size_t n_paths, int ranks,
                  const std::vector<size_t>& order, std::vector<int>* to, std::vector<uint64_t>* got)
{
  uint64_t total = 0;
  for (uint64_t v : bytes) {
    total += v;
  }
  to->clear();
  got->assign((size_t)ranks, 0);
  if (total == 0) {
    return false;
  }
  uint64_t start = 0;
  int prev = 0;
  for (size_t q : order) {
    for (int r = 0; r < ranks; ++r) {
      const uint64_t b = bytes[(size_t)r * n_paths + q];
      int t = (int)std::min<uint64_t>((uint64_t)ranks - 1,
                                      (uint64_t)((unsigned __int128)(start + b / 2) * (unsigned)ranks / total));
      t = std::max(t, prev);
      prev = t;
      to->push_back(t);
      (*got)[(size_t)t] += b;
      start += b;
    }
  }
  for (uint64_t v : *got) {
    if (v == 0) {
      return false;
    }
  }
  return true;
}

double
now_ms()
{
  return std::chrono::duration<double, std::milli>(
           std::chrono::steady_clock::now().time_since_epoch())
    .count();
}

int
set_err(char* err, size_t cap, const std::string& msg, int code)
{
  if (err && cap) {
    snprintf(err, cap, "%s", msg.c_str());
  }
  return code;
}

// calc_phred_average.cpp:45-58 over a byte range (host side: only for the few selected reads'
// "Average Phred" log line)
double
sum_phred_host(const char* q, size_t n, const double* tab)
{
  double s = 0;
  for (size_t i = 0; i < n; ++i) {
    s += tab[(unsigned char)q[i]];
  }
  return s;
}

void
log_path_stat(const Log& log, uint64_t curr_path, const grb_path_stats& s, double phred_sum)
{
  // goldrush_path.cpp:126-154
  log("Visited %llu reads to generate %llu silver paths\n", (unsigned long long)s.valid_reads,
      (unsigned long long)curr_path);
  log("Saw: %llu tiles to generate %llu silver paths\n", (unsigned long long)s.total_tiles,
      (unsigned long long)curr_path);
  log("Assigned: %llu tiles to generate %llu silver paths\n", (unsigned long long)s.assigned_tiles,
      (unsigned long long)curr_path);
  log("Unassigned: %llu tiles to generate %llu silver paths\n",
      (unsigned long long)s.unassigned_tiles, (unsigned long long)curr_path);
  log("Total queries: %llu to generate %llu silver paths\n", (unsigned long long)s.queries,
      (unsigned long long)curr_path);
  log("Total hits: %llu to generate %llu silver paths\n", (unsigned long long)s.hits,
      (unsigned long long)curr_path);
  log("Total misses: %llu to generate %llu silver paths\n", (unsigned long long)s.misses,
      (unsigned long long)curr_path);
  log("Num reads: %llu in silver path %llu\n", (unsigned long long)s.num_reads_in_path,
      (unsigned long long)curr_path);
  const uint32_t avg = (uint32_t)(-10 * log10(phred_sum / s.inserted_bases));
  log("Average Phred: %u in silver path %llu\n", avg, (unsigned long long)curr_path);
}

} // namespace

// What grb_run_two_stage takes over from the silver stage: the bytes of every output record, one
// buffer per output file (per_path[n - 1] = what <p>_n.fq holds; slice mode: this rank's records
// only), and at the end of the stage `joined`: what `cat <p>_*.fq` gives, on every rank.
struct Capture
{
  std::vector<std::vector<char>> per_path;
  std::unique_ptr<char[]> joined;
  size_t joined_len = 0;
  // several ranks in slice mode: `joined` is this rank's consecutive share of the joined paths
  // (bytes [slice_offset, slice_offset + joined_len) of slice_total), cut between records
  uint64_t slice_offset = 0, slice_total = 0;
};

static int
run_path_impl(const grb_run_options* o, const char* fastq, size_t fastq_len, grb_run_result* res,
              char* err, size_t err_cap, Capture* cap_out)
{
  const double t_wall0 = now_ms();
  const bool timing = getenv("GRB_TIMING") != nullptr; // host-side phase clock on stderr
  double t_mark = t_wall0;
  auto mark = [&](const char* what) {
    if (timing) {
      const double t = now_ms();
      fprintf(stderr, "[grb timing] %-12s %9.1f ms\n", what, t - t_mark);
      t_mark = t;
    }
  };
  grb_run_result R{};
  std::vector<std::vector<char>>* capture = cap_out ? &cap_out->per_path : nullptr;
  const Log log{ !o->quiet };
  grb_params p = o->params;

  // ---- input ----
  const char* data = fastq;
  size_t n = fastq_len;
  void* map = nullptr;
  size_t map_len = 0;
  if (!data) {
    const int fd = open(o->input_path ? o->input_path : "", O_RDONLY);
    if (fd < 0) {
      return set_err(err, err_cap, std::string("cannot open ") + (o->input_path ? o->input_path : ""),
                     GRB_ERR_ARG);
    }
    struct stat sb;
    fstat(fd, &sb);
    map_len = (size_t)sb.st_size;
    if (map_len) {
      map = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
      if (map == MAP_FAILED) {
        close(fd);
        return set_err(err, err_cap, "mmap failed", GRB_ERR_NOMEM);
      }
      madvise(map, map_len, MADV_SEQUENTIAL);
    }
    close(fd);
    data = (const char*)map;
    n = map_len;
  }
  struct Unmap
  {
    void* m;
    size_t l;
    ~Unmap()
    {
      if (m) {
        munmap(m, l);
      }
    }
  } unmap{ map, map_len };

  // ---- seeds (spaced_seeds.cpp) ----
  const unsigned k = (unsigned)p.kmer_size, h = (unsigned)p.hash_num;
  std::vector<std::string> seed_store(h, std::string(k + h + 2, '\0'));
  std::vector<char*> seed_ptr(h);
  for (unsigned i = 0; i < h; ++i) {
    seed_ptr[i] = &seed_store[i][0];
  }
  const bool preset = o->seed_preset && o->seed_preset[0];
  if (preset) {
    log("Using preset spaced seed\nwith:\n\tspan: %zu\n\tweight: %ld\n", strlen(o->seed_preset),
        (long)std::count(o->seed_preset, o->seed_preset + strlen(o->seed_preset), '1'));
  } else {
    log("Designing base symmetrical spaced seed\nUsing:\nspan: %u\nweight: %u\n", k,
        (unsigned)p.weight);
  }
  if (grb_make_seed_pattern(o->seed_preset, k, (unsigned)p.weight, h, seed_ptr.data()) != GRB_OK) {
    return set_err(err, err_cap, "cannot design a spaced seed for this k / w", GRB_ERR_ARG);
  }
  std::vector<const char*> seed_c(seed_ptr.begin(), seed_ptr.end());
  p.seeds = seed_c.data();

  grb_ctx* ctx = nullptr;
  int rc = grb_create(&p, &ctx);
  if (rc != GRB_OK) {
    return set_err(err, err_cap, grb_last_error(nullptr), rc);
  }
  struct Guard
  {
    grb_ctx* c;
    ~Guard() { grb_destroy(c); }
  } guard{ ctx };
  auto fail = [&](int code) { return set_err(err, err_cap, grb_last_error(ctx), code); };

  int c_world_now = 1;
  bool shard_ingest = false;
  const bool slice_mode = o->fastq_total != 0;
  if (slice_mode && o->write_outputs) {
    return set_err(err, err_cap, "slice mode (fastq_total != 0) cannot write output files: no rank "
                                 "holds the whole input", GRB_ERR_ARG);
  }
  // The reference opens (truncates) its first output file after sizing and before pass 1
  // (goldrush_path.cpp:1109-1123,1174-1179): a run that stops later -- input not FASTQ, a read shorter
  // than the seed span, no read passing the filters -- leaves that file behind, empty; a run that dies
  // inside --ntcard sizing does not.
  auto touch_first_output = [&]() {
    if (o->write_outputs) {
      const std::string prefix0 = o->prefix ? o->prefix : "goldrush_out";
      if (FILE* f0 = fopen((p.silver_path ? prefix0 + "_1.fq" : prefix0 + ".fa").c_str(), "wb")) {
        fclose(f0);
      }
    }
  };
  // byte `off` of the whole input, as this process can address it (slice mode: own reads only)
  const uint64_t data_origin = slice_mode ? o->fastq_offset : 0;
  bool early = false; // pass 1 already done chunk by chunk during the ingest
  uint64_t early_bits = 0;
  double early_ms = 0;
  std::vector<uint8_t> early_flags;

  // goldrush_path.cpp:1138-1161
  auto log_parameters = [&]() {
    log("Calculating %s\nUsing:\n\ttile length: %llu\n\tblock size: %llu\n\tseed patterns: %llu\n"
        "\tthreshold: %llu\n\tbase seed pattern: %s\n\tminimum unassigned tiles: %llu\n"
        "\tmaximum assigned tiles: %llu\n\texpected hash space: %llu\n"
        "\tminimum average phred quality score: %u\n"
        "\tmaximum average phred delta between first and second half of read: %u\n"
        "\toccupancy: %g\n\tjobs: %d\n",
        p.silver_path ? (std::to_string(p.max_paths) + " silver path(s)").c_str() : "the golden path",
        (unsigned long long)p.tile_length, (unsigned long long)p.block_size,
        (unsigned long long)p.hash_num, (unsigned long long)p.threshold, seed_c[0],
        (unsigned long long)p.unassigned_min, (unsigned long long)p.assigned_max,
        (unsigned long long)p.hash_universe, p.phred_min, p.phred_delta, p.occupancy, o->jobs);
  };

  // everything the reference has logged when fill_bit_vector starts reading (:1138-1199), for runs
  // that end there before this driver's own logging has reached that point
  auto log_up_to_pass1 = [&]() {
    if (p.hash_universe == 0) {
      p.hash_universe = grb_default_hash_universe(p.weight, p.genome_size, p.hash_num);
    }
    log_parameters();
    if (o->filter_file && o->filter_file[0]) {
      log("Using only reads not found in: %s\n", o->filter_file);
    }
    log("allocating bit vector\nm_filterSize: %llu\nfinished allocating bit vector\n"
        "opening: %s\ninserting bit vector\n",
        (unsigned long long)grb_calc_optimal_size(p.hash_universe, 1, p.occupancy),
        o->input_path ? o->input_path : "(memory)");
  };

  // ---- K1: one pass over the file ----
  if (n == 0 || data[0] != '@') {
    // The reference meets its format check inside fill_bit_vector, after it has logged its parameters
    // and sized the filter (goldrush_path.cpp:1109-1199): the same lines first, where they do not
    // depend on reading the input (-H or -g sizing, -P given).
    if (!(p.hash_universe == 0 && o->ntcard) && p.phred_min != 0) {
      log_up_to_pass1();
    }
    touch_first_output();
    log("Gold Path requires fastq format\n"); // goldrush_path.cpp:247-250
    return set_err(err, err_cap, "Gold Path requires fastq format", GRB_ERR_FORMAT);
  }
  {
    int c_rank = 0, c_world = 1; // several GPUs: each rank ingests its own share of the input
    grb_comm_info(ctx, &c_rank, &c_world);
    c_world_now = c_world;
    // the share of this rank: the whole buffer (one GPU, or slice mode: the caller cut the input),
    // else bytes [cut(r), cut(r + 1)) with the cuts moved forward to record boundaries
    size_t lo = 0, hi = n;
    {
      const char* e = getenv("GRB_SHARD_INGEST");
      shard_ingest = c_world > 1 && !(e && strcmp(e, "0") == 0);
    }
    if (slice_mode) {
      shard_ingest = c_world > 1;
      if ((rc = grb_reads_set_origin(ctx, o->fastq_offset)) != GRB_OK) {
        return fail(rc);
      }
    } else if (shard_ingest) {
      lo = next_record_start(data, n, (size_t)((unsigned __int128)n * (unsigned)c_rank / (unsigned)c_world));
      hi = c_rank + 1 == c_world
             ? n
             : next_record_start(data, n, (size_t)((unsigned __int128)n * (unsigned)(c_rank + 1) / (unsigned)c_world));
      if ((rc = grb_reads_set_origin(ctx, lo)) != GRB_OK) {
        return fail(rc);
      }
    }
    if ((rc = grb_reads_reserve(ctx, hi - lo)) != GRB_OK) {
      return fail(rc);
    }
    const size_t kChunk = (size_t)1 << 30;
    size_t off = lo;
    if ((rc = grb_reads_readahead(ctx, data + lo, hi - lo)) != GRB_OK) { // next chunk's copy under this one's decode
      return fail(rc);
    }
    // Pass 1 under the ingest: when the filter size and the Phred threshold do not depend on the
    // whole file (-P given, no --ntcard), a read's pass-1 verdict (goldrush_path.cpp:261-301) only
    // needs its own record, so each chunk's reads are hashed into the bit vector while the next
    // chunk is still crossing PCIe.  The regular filter stage below re-derives the same flags.
    {
      const char* e = getenv("GRB_EARLY_PASS1");
      early = p.phred_min != 0 && (p.hash_universe != 0 || !o->ntcard) && !(e && strcmp(e, "0") == 0);
    }
    if (early) {
      const uint64_t hu = p.hash_universe
                            ? p.hash_universe
                            : grb_default_hash_universe(p.weight, p.genome_size, p.hash_num);
      early_bits = grb_calc_optimal_size(hu, 1, p.occupancy);
      if ((rc = grb_filter_alloc(ctx, early_bits)) != GRB_OK) {
        return fail(rc);
      }
    }
    uint64_t early_done = 0;
    std::vector<grb_read_meta> cm;
    while (off < hi) {
      const size_t len = std::min(kChunk, hi - off);
      const int final = off + len == hi;
      size_t used = 0;
      const double t_c0 = now_ms();
      if ((rc = grb_reads_ingest_fastq(ctx, data + off, len, final, &used)) != GRB_OK) {
        return fail(rc);
      }
      const double t_c1 = now_ms();
      R.ms_ingest += grb_last_device_ms(ctx);
      if (used == 0) {
        if (final) {
          break;
        }
        return set_err(err, err_cap, "FASTQ record larger than the 1 GiB ingest chunk", GRB_ERR_ARG);
      }
      off += used;
      if (early) {
        const uint64_t now = grb_reads_count(ctx), cnt = now - early_done;
        if (cnt) {
          cm.resize(cnt);
          if ((rc = grb_reads_get_meta(ctx, early_done, cnt, cm.data())) != GRB_OK) {
            return fail(rc);
          }
          early_flags.resize(now, 0);
          for (uint64_t i = 0; i < cnt; ++i) {
            uint32_t a = 0, d = 0;
            grb_phred_finalize(cm[i].phred_first_half_sum, cm[i].phred_total_sum, cm[i].qual_len, &a, &d);
            early_flags[early_done + i] = (cm[i].len >= p.min_length && a >= p.phred_min &&
                                           d < p.phred_delta && !cm[i].non_acgt)
                                            ? GRB_READ_PASS1
                                            : 0;
          }
          // own slice ingested: all of the chunk's reads are this rank's; whole input on every
          // rank (GRB_SHARD_INGEST=0): an equal share of each chunk
          const uint64_t my_lo = shard_ingest ? early_done : early_done + cnt * (uint64_t)c_rank / (uint64_t)c_world;
          const uint64_t my_hi = shard_ingest ? now : early_done + cnt * (uint64_t)(c_rank + 1) / (uint64_t)c_world;
          if ((rc = grb_reads_set_flags(ctx, early_done, cnt, early_flags.data() + early_done)) != GRB_OK ||
              (rc = grb_build_bitvector_range(ctx, my_lo, my_hi - my_lo)) != GRB_OK) {
            log_up_to_pass1(); // a read shorter than the seed span: the reference dies inside pass 1
            touch_first_output();
            log("%s\n", grb_last_error(ctx));
            return fail(rc);
          }
          early_ms += grb_last_device_ms(ctx);
          early_done = now;
        }
      }
      if (timing && getenv("GRB_TIMING")[0] == '2') {
        fprintf(stderr, "[grb chunk] ingest %.1f ms (device %.1f)  early pass 1 %.1f ms\n", t_c1 - t_c0,
                grb_last_device_ms(ctx), now_ms() - t_c1);
      }
    }
    grb_reads_readahead(ctx, nullptr, 0);
    if (shard_ingest) { // every rank's store now holds all reads, in file order (NVLink)
      if ((rc = grb_reads_allgather(ctx)) != GRB_OK) {
        return fail(rc);
      }
      R.ms_ingest += grb_last_device_ms(ctx);
    }
  }
  mark("create+ingest");
  uint64_t own_first = 0, own_count = 0;
  grb_reads_own_range(ctx, &own_first, &own_count);
  const uint64_t nreads = grb_reads_count(ctx);
  std::vector<grb_read_meta> meta(nreads);
  if (nreads && (rc = grb_reads_get_meta(ctx, 0, nreads, meta.data())) != GRB_OK) {
    return fail(rc);
  }
  R.num_reads = nreads;

  // ---- filter sizing (goldrush_path.cpp:1109-1123) ----
  if (p.hash_universe == 0) {
    if (o->ntcard) {
      log("Calculating expected entries\n");
      std::vector<uint64_t> per(h);
      uint64_t total = 0;
      if ((rc = grb_estimate_cardinality(ctx, n, per.data(), &total)) != GRB_OK) {
        return fail(rc);
      }
      for (unsigned i = 0; i < h; ++i) {
        log("Expected entries for seed pattern %s : %llu\n", seed_c[i], (unsigned long long)per[i]);
      }
      log("Total expected entries for seed patterns: %llu\n", (unsigned long long)total);
      p.hash_universe = total;
    } else {
      p.hash_universe = grb_default_hash_universe(p.weight, p.genome_size, p.hash_num);
    }
  }
  touch_first_output();

  // host threads for the bookkeeping loops: -j if given, else the machine's cores divided by the
  // ranks sharing it (one process per GPU).  Set explicitly on every region: launchers such as
  // torchrun export OMP_NUM_THREADS=1.
  int n_threads = o->jobs > 0 ? o->jobs : (int)std::max(1u, std::thread::hardware_concurrency());
  {
    int c_rank = 0, c_world = 1;
    grb_comm_info(ctx, &c_rank, &c_world);
    if (o->jobs <= 0 && c_world > 1) {
      n_threads = std::max(1, n_threads / c_world);
    }
  }
  n_threads = std::min(n_threads, 64);
  // record assembly runs beside pass 2: with several ranks sharing the host, the thread that
  // launches the kernels and NCCL's proxy thread must keep a core each, or every collective of the
  // lock-stepped ranks waits for whichever rank was descheduled last
  const int emit_threads = c_world_now > 1 ? std::max(1, n_threads - 2) : n_threads;

  // ---- per-read Phred statistics from the device sums ----
  std::vector<uint32_t> avg(nreads), delta(nreads);
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (int64_t i = 0; i < (int64_t)nreads; ++i) {
    grb_phred_finalize(meta[i].phred_first_half_sum, meta[i].phred_total_sum, meta[i].qual_len,
                       &avg[i], &delta[i]);
  }
  // calc_min_phred_threshold (goldrush_path.cpp:79-107): median over the first 50000 reads that
  // are long enough, in file order (the reference's sample is "first to arrive" among its threads)
  if (p.phred_min == 0) {
    log("Calculating minimum phred score via median\n");
    const size_t cap = 50000;
    std::vector<uint32_t> scores(cap, 0);
    size_t cnt = 0;
    for (uint64_t i = 0; i < nreads; ++i) {
      if (meta[i].len < p.min_length) {
        continue;
      }
      if (cnt >= cap) {
        ++cnt;
        break;
      }
      scores[cnt++] = avg[i];
    }
    std::sort(scores.begin(), scores.end(), std::greater<uint32_t>());
    p.phred_min = std::max<uint32_t>(10u, scores[cnt / 2]);
    if (o->verbose) {
      log("Minimum phred score calculated with median: %u\n", p.phred_min);
    }
  }
  R.phred_min = p.phred_min;

  log_parameters();

  // ---- name filter (-f, goldrush_path.cpp:1163-1172) ----
  std::unordered_set<std::string> filter_out;
  if (o->filter_file && o->filter_file[0]) {
    log("Using only reads not found in: %s\n", o->filter_file);
    FILE* ff = fopen(o->filter_file, "r");
    if (ff) {
      char name[4096];
      while (fscanf(ff, "%4095s", name) == 1) {
        filter_out.insert(name);
      }
      fclose(ff);
    }
  }
  // ---- pass-1 filters (goldrush_path.cpp:261-301) ----
  // The reference inserts the NAME of every read dropped here into filter_out_reads and pass 2
  // skips every read whose name is in that set (:907-932), i.e. the dropped read itself and any
  // other read carrying the same name.  Same rule, without 10^5 string inserts: the dropped names
  // (and the -f names) are kept as sorted 64-bit hashes, a surviving read whose hash matches is
  // confirmed by comparing the strings.
  std::vector<uint8_t> flags(nreads, 0);
  uint64_t by_len = 0, by_phred = 0, by_delta = 0, by_bases = 0, passed = 0;
  auto have_bytes = [&](uint64_t i) { return !slice_mode || (i >= own_first && i < own_first + own_count); };
  auto id_span = [&](uint64_t i, const char** s0) {
    const char* hdr = data + (meta[i].hdr_off - data_origin);
    size_t l = 0;
    while (l < meta[i].hdr_len && hdr[l] != ' ' && hdr[l] != '\t') {
      ++l;
    }
    *s0 = hdr;
    return l;
  };
  auto hash_bytes = [](const char* s0, size_t l) {
    Fnv f;
    f.add(s0, l);
    return f.h;
  };
  std::vector<uint8_t> dropped(nreads, 0);
  for (uint64_t i = 0; i < nreads; ++i) {
    if (meta[i].len < p.min_length) {
      ++by_len;
      continue;
    }
    if (avg[i] < p.phred_min || delta[i] >= p.phred_delta) {
      by_phred += avg[i] < p.phred_min;
      by_delta += delta[i] >= p.phred_delta;
      dropped[i] = 1;
      continue;
    }
    if (meta[i].non_acgt) {
      ++by_bases;
      dropped[i] = 1;
      continue;
    }
    ++passed;
    R.bases_pass1 += meta[i].len;
    flags[i] = GRB_READ_PASS1;
  }
  std::vector<uint64_t> name_hash(nreads); // FNV-1a of the record id, computed by K1 (k_records)
  for (uint64_t i = 0; i < nreads; ++i) {
    name_hash[i] = meta[i].name_hash;
  }
  struct Dropped
  {
    uint64_t hash;
    int64_t read; // -1: a name of the -f list
  };
  std::vector<Dropped> drop_list;
  for (uint64_t i = 0; i < nreads; ++i) {
    if (dropped[i]) {
      drop_list.push_back(Dropped{ name_hash[i], (int64_t)i });
    }
  }
  for (const std::string& nm : filter_out) {
    drop_list.push_back(Dropped{ hash_bytes(nm.data(), nm.size()), -1 });
  }
  std::sort(drop_list.begin(), drop_list.end(),
            [](const Dropped& x, const Dropped& y) { return x.hash < y.hash; });
  auto name_is_dropped = [&](uint64_t i) {
    auto it = std::lower_bound(drop_list.begin(), drop_list.end(), name_hash[i],
                               [](const Dropped& x, uint64_t hsh) { return x.hash < hsh; });
    if (it == drop_list.end() || it->hash != name_hash[i]) {
      return false;
    }
    if (!have_bytes(i)) {
      return true; // slice mode, another rank's read: the 64-bit hash of the id stands for the id
    }
    const char* s0;
    const size_t l = id_span(i, &s0);
    for (; it != drop_list.end() && it->hash == name_hash[i]; ++it) {
      if (it->read < 0) {
        if (filter_out.count(std::string(s0, l))) {
          return true;
        }
        continue;
      }
      if (!have_bytes((uint64_t)it->read)) {
        return true;
      }
      const char* t0;
      const size_t tl = id_span((uint64_t)it->read, &t0);
      if (tl == l && memcmp(s0, t0, l) == 0) {
        return true;
      }
    }
    return false;
  };
  // pass-2 eligibility (goldrush_path.cpp:907-932): long enough and name not filtered out
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (int64_t i = 0; i < (int64_t)nreads; ++i) {
    if (meta[i].len >= p.min_length && !dropped[i] &&
        (drop_list.empty() || !name_is_dropped((uint64_t)i))) {
      flags[i] |= GRB_READ_PASS2;
    }
  }
  R.num_passed_reads = passed;

  mark("filters");
  log("allocating bit vector\n");
  const uint64_t filter_bits = grb_calc_optimal_size(p.hash_universe, 1, p.occupancy);
  log("m_filterSize: %llu\n", (unsigned long long)filter_bits);
  if (early) { // the early pass must have used this very size and these very flags
    const uint64_t e0 = shard_ingest ? own_first : 0;
    early = filter_bits == early_bits && early_flags.size() == (shard_ingest ? own_count : nreads);
    for (uint64_t i = 0; early && i < early_flags.size(); ++i) {
      early = (flags[e0 + i] & GRB_READ_PASS1) == early_flags[i];
    }
  }
  if (!early && (rc = grb_filter_alloc(ctx, filter_bits)) != GRB_OK) {
    return fail(rc);
  }
  R.filter_bits = filter_bits;
  log("finished allocating bit vector\n");
  log("opening: %s\n", o->input_path ? o->input_path : "(memory)");
  log("inserting bit vector\n");
  if (nreads && (rc = grb_reads_set_flags(ctx, 0, nreads, flags.data())) != GRB_OK) {
    return fail(rc);
  }
  if (early) {
    if ((rc = grb_bitvector_or_reduce(ctx)) != GRB_OK) { // no-op on one GPU
      return fail(rc);
    }
    R.ms_pass1 = early_ms + (c_world_now > 1 ? grb_last_device_ms(ctx) : 0.0);
  } else {
    if ((rc = grb_build_bitvector(ctx)) != GRB_OK) {
      log("%s\n", grb_last_error(ctx));
      return fail(rc);
    }
    R.ms_pass1 = grb_last_device_ms(ctx);
  }
  if (o->verbose) {
    log("num_passed_reads: %llu\nnum_reads: %llu\nnum_reads - num_passed_reads: %llu\n"
        "num_reads - num_passed_reads / num_reads: %.0f\nnum_reads_skipped_by_phred: %llu\n"
        "num_reads_skipped_by_delta: %llu\nnum_reads_skipped_by_length: %llu\n"
        "num_reads_skipped_by_invalid_bases: %llu\nTotal reads skipped: %llu\n",
        (unsigned long long)passed, (unsigned long long)nreads,
        (unsigned long long)(nreads - passed), floor((double)(nreads - passed) / nreads),
        (unsigned long long)by_phred, (unsigned long long)by_delta, (unsigned long long)by_len,
        (unsigned long long)by_bases, (unsigned long long)(by_phred + by_delta + by_len + by_bases));
  }
  if (passed == 0) { // goldrush_path.cpp:327-334
    log("Error: no reads passed the Phred score and min length requirements\n"
        "Try again with a lower Phred threshold or lower min length\n");
    return set_err(err, err_cap, "no reads passed the Phred score and min length requirements",
                   GRB_ERR_ARG);
  }
  log("finished inserting bit vector\nin %.4f\n", R.ms_pass1 / 1e3);

  uint64_t pop = 0;
  if ((rc = grb_finalize_bitvector(ctx, &pop)) != GRB_OK) {
    return fail(rc);
  }
  R.ms_rank = grb_last_device_ms(ctx);
  R.pop = pop;

  mark("pass1+rank");
  // ---- pass 2 ----
  log("assigning tiles\n");
  // Pass 2 runs in slices of reads (grb_select_reads keeps its loop state between calls); while the
  // GPU works on slice i + 1 a host task turns slice i's decisions into output records, so the
  // writers (goldrush_path.cpp:973-976,996-1002,1055-1070,1174-1179,182-184) cost no wall time.
  std::vector<grb_decision> dec(nreads);
  std::vector<grb_path_stats> snaps(p.max_paths + 2);
  uint32_t n_snaps = 0;
  int finished = 0;

  // ---- output stage state (touched only by the emit task, one slice at a time, in order) ----
  // Pass A (serial, O(#reads)): which records go where.  Pass B (parallel over records, chunked):
  // assemble the record bytes.  Pass C (serial, in record order): write, digest, path log lines.
  double tab[256];
  for (int b = 0; b < 256; ++b) {
    tab[b] = pow(10.0, -(int)((char)b - 33) / 10.0);
  }
  struct Rec
  {
    uint64_t read;
    size_t s0, sl, ql;
    size_t at, bytes;   // position inside the chunk buffer (or inside the capture buffer of its path)
    uint32_t id_len;
    bool trimmed, closes_path;
    double phred;
    uint64_t hash;
  };
  std::vector<Rec> recs;
  // slice mode: this rank assembles (and hashes) only the records of its own reads; the digest and
  // the path log lines are replayed from the exchanged per-record {hash, Phred sum} pairs afterwards
  struct Sel
  {
    uint64_t read;
    bool closes_path;
  };
  struct HashPhred
  {
    uint64_t hash;
    double phred;
  };
  std::vector<Sel> sel_order;
  std::vector<HashPhred> own_pairs;
  const uint64_t T = p.tile_length;
  uint64_t visited_reads = 0;
  bool past_end = false; // a GRB_NOT_VISITED read was seen: nothing after it was reached
  uint32_t snap_a = 0;   // pass A's cursor into snaps
  Fnv digest;
  OutFile out;
  out.write = o->write_outputs != 0;
  out.digest = &digest;
  const std::string prefix = o->prefix ? o->prefix : "goldrush_out";
  out.open(p.silver_path ? prefix + "_1.fq" : prefix + ".fa");
  const char first_char = p.silver_path ? '@' : '>';
  uint32_t path_now = 1, snap_i = 0;
  double phred_sum = 0;
  const bool want_phred = !o->quiet && o->verbose;
  const bool want_bytes = out.write || capture != nullptr;
  std::vector<char> buf;
  const size_t kChunkBytes = (size_t)512 << 20;

  // snaps_seen = snapshots known when the slice was handed over (the array itself is shared)
  auto emit = [&](uint64_t first, uint64_t count, uint32_t snaps_seen) {
    recs.clear();
    for (uint64_t i = first; i < first + count && !past_end; ++i) {
      const grb_decision& d = dec[i];
      if (d.verdict == GRB_NOT_VISITED) {
        past_end = true;
        break;
      }
      ++visited_reads;
      if (d.verdict != GRB_SKIPPED) {
        ++R.reads_visited;
        R.bases_pass2 += (uint64_t)d.num_tiles * T;
      }
      if (d.verdict != GRB_UNTRIMMED && d.verdict != GRB_TRIMMED) {
        continue;
      }
      Rec r{};
      r.read = i;
      r.trimmed = d.verdict == GRB_TRIMMED;
      r.s0 = 0;
      r.sl = meta[i].len;
      if (r.trimmed) {
        r.s0 = (size_t)d.trim_start * T;
        r.sl = (d.trim_end == d.num_tiles - 1) ? meta[i].len - r.s0
                                               : (size_t)(d.trim_end - d.trim_start + 1) * T;
      }
      r.ql = std::min<size_t>(r.sl, meta[i].qual_len > r.s0 ? meta[i].qual_len - r.s0 : 0);
      // silver_path_check (goldrush_path.cpp:156-187): a snapshot was taken right after this read
      r.closes_path = snap_a < snaps_seen && snaps[snap_a].rollover_read == i;
      if (r.closes_path) {
        ++snap_a;
      }
      ++R.reads_selected;
      if (slice_mode) {
        sel_order.push_back(Sel{ i, r.closes_path });
        if (!have_bytes(i)) {
          R.bases_selected += r.sl;
          continue;
        }
      }
      const char* hdr = data + (meta[i].hdr_off - data_origin);
      uint32_t l = 0;
      while (l < meta[i].hdr_len && hdr[l] != ' ' && hdr[l] != '\t') {
        ++l;
      }
      r.id_len = l;
      r.bytes = 1 + l + (r.trimmed ? 9 : 11) + r.sl + 1 + (p.silver_path ? 2 + r.ql + 1 : 0);
      recs.push_back(r);
      R.bases_selected += r.sl;
    }
    size_t r0 = 0;
    while (r0 < recs.size()) {
      size_t r1 = r0, bytes = 0;
      while (r1 < recs.size() && (r1 == r0 || bytes + recs[r1].bytes <= kChunkBytes)) {
        recs[r1].at = bytes;
        bytes += recs[r1].bytes;
        ++r1;
      }
      if (capture) {
        // two-stage call: the records are assembled straight into the buffer of their path (no copy
        // afterwards); the buffers grow once per chunk of records
        std::vector<size_t> grow;
        for (size_t ri = r0; ri < r1; ++ri) {
          const uint32_t pth = dec[recs[ri].read].path;
          if (capture->size() < pth) {
            capture->resize(pth);
          }
          if (grow.size() < pth) {
            grow.resize(pth, 0);
          }
          recs[ri].at = (*capture)[pth - 1].size() + grow[pth - 1];
          grow[pth - 1] += recs[ri].bytes;
        }
        for (size_t q = 0; q < grow.size(); ++q) {
          std::vector<char>& cp = (*capture)[q];
          if (grow[q]) {
            if (cp.capacity() < cp.size() + grow[q]) { // a path ends near ratio * genome bases: about 2.1 bytes each
              cp.reserve(std::max<size_t>(cp.size() + grow[q],
                                          std::max<size_t>(2 * cp.capacity(), (size_t)(2.1 * p.ratio * p.genome_size) /
                                                                                (size_t)std::max(1, slice_mode ? c_world_now : 1))));
            }
            cp.resize(cp.size() + grow[q]);
          }
        }
      } else if (want_bytes && buf.size() < bytes) {
        buf.resize(bytes);
      }
      {
        parallel_for(emit_threads, (int64_t)r0, (int64_t)r1, 4, [&](int64_t ri) {
          thread_local std::vector<char> local; // record scratch when nothing is written (digest only)
          Rec& r = recs[ri];
          const grb_read_meta& m = meta[r.read];
          char* dst;
          if (capture) {
            dst = (*capture)[dec[r.read].path - 1].data() + r.at;
          } else if (want_bytes) {
            dst = buf.data() + r.at;
          } else {
            if (local.size() < r.bytes) {
              local.resize(r.bytes);
            }
            dst = local.data();
          }
          char* w = dst;
          *w++ = first_char;
          memcpy(w, data + (m.hdr_off - data_origin), r.id_len);
          w += r.id_len;
          if (r.trimmed) {
            memcpy(w, "_trimmed\n", 9);
            w += 9;
          } else {
            memcpy(w, "_untrimmed\n", 11);
            w += 11;
          }
          const char* sq = data + (m.seq_off - data_origin) + r.s0;
#pragma omp simd
          for (size_t j = 0; j < r.sl; ++j) { // SeqReader folds the sequence to upper case
            const unsigned char ch = (unsigned char)sq[j];
            w[j] = (char)(ch - (((unsigned)(ch - 'a') < 26u) << 5));
          }
          w += r.sl;
          *w++ = '\n';
          const char* ql = data + (m.qual_off - data_origin) + r.s0;
          if (p.silver_path) {
            *w++ = '+';
            *w++ = '\n';
            memcpy(w, ql, r.ql);
            w += r.ql;
            *w++ = '\n';
          }
          r.hash = record_hash(dst, r.bytes);
          // the reference adds sum_phred of the written quality string (goldrush_path.cpp:1005-1008);
          // for a whole read that is the running sum the device already holds
          r.phred = 0;
          if (want_phred) {
            r.phred = (!r.trimmed && r.ql == m.qual_len) ? m.phred_total_sum : sum_phred_host(ql, r.ql, tab);
          }
        });
      }
      for (size_t ri = r0; ri < r1; ++ri) {
        const Rec& r = recs[ri];
        if (slice_mode) {
          own_pairs.push_back(HashPhred{ r.hash, r.phred });
          continue;
        }
        if (out.f) {
          const char* src = capture ? (*capture)[dec[r.read].path - 1].data() + r.at : buf.data() + r.at;
          fwrite(src, 1, r.bytes, out.f);
        }
        digest.add((const char*)&r.hash, 8);
        phred_sum += r.phred;
        path_now = dec[r.read].path;
        if (r.closes_path) {
          if (o->verbose) {
            log_path_stat(log, path_now, snaps[snap_i], phred_sum);
          }
          ++snap_i;
          phred_sum = 0;
          if (path_now + 1 <= p.max_paths) {
            out.open(prefix + "_" + std::to_string(path_now + 1) + ".fq");
          }
        }
      }
      r0 = r1;
    }
  };

  // reads per grb_select_reads call: long enough that the calls' fixed costs (a handful of host
  // round trips each) stay small on inputs of millions of reads, short enough that the record
  // assembly of one slice has the next slice's device time to hide under
  uint64_t slice = std::min<uint64_t>(262144, std::max<uint64_t>(16384, nreads / 64));
  if (const char* e = getenv("GRB_SLICE_READS")) { // 0 = one call for the whole store
    const long long v = atoll(e);
    slice = v <= 0 ? nreads : (uint64_t)v;
  }
  slice = std::max<uint64_t>(slice, 1);
  std::future<void> pending;
  int sel_rc = GRB_OK;
  for (uint64_t first = 0; first < nreads; first += slice) {
    const uint64_t count = std::min<uint64_t>(slice, nreads - first);
    if (!finished) {
      uint32_t got = 0;
      sel_rc = grb_select_reads(ctx, first, count, dec.data() + first, snaps.data() + n_snaps,
                                (uint32_t)snaps.size() - n_snaps, &got, &finished);
      if (sel_rc != GRB_OK) {
        break;
      }
      n_snaps += got;
      R.ms_pass2 += grb_last_device_ms(ctx);
    } // after exit(0) in the reference nothing else is visited: dec stays GRB_NOT_VISITED
    if (pending.valid()) {
      pending.get();
    }
    const uint32_t seen = n_snaps;
    pending = std::async(std::launch::async, emit, first, count, seen);
    if (finished) {
      break; // the remaining reads were never reached
    }
  }
  if (pending.valid()) {
    pending.get();
  }
  if (sel_rc != GRB_OK) {
    return fail(sel_rc);
  }
  if (slice_mode) {
    // records of rank r follow those of rank r - 1 (file order): the pairs of all ranks, back to
    // back, line up with sel_order
    std::vector<HashPhred> all(sel_order.size());
    std::vector<uint64_t> sizes((size_t)std::max(1, c_world_now));
    if ((rc = grb_comm_allgather_host(ctx, own_pairs.data(), own_pairs.size() * sizeof(HashPhred),
                                      all.data(), all.size() * sizeof(HashPhred), sizes.data())) != GRB_OK) {
      return fail(rc);
    }
    uint64_t got = 0;
    for (uint64_t v : sizes) {
      got += v;
    }
    if (got != all.size() * sizeof(HashPhred)) {
      return set_err(err, err_cap, "slice mode: the ranks' record lists do not add up to the selection",
                     GRB_ERR_STATE);
    }
    for (size_t j = 0; j < all.size(); ++j) {
      digest.add((const char*)&all[j].hash, 8);
      phred_sum += all[j].phred;
      path_now = dec[sel_order[j].read].path;
      if (sel_order[j].closes_path) {
        if (o->verbose) {
          log_path_stat(log, path_now, snaps[snap_i], phred_sum);
        }
        ++snap_i;
        phred_sum = 0;
      }
    }
  }
  grb_path_stats cur{};
  uint64_t curr_path = 1;
  if ((rc = grb_select_state(ctx, &cur, &curr_path, nullptr)) != GRB_OK) {
    return fail(rc);
  }
  mark("pass2+outputs");
  for (uint64_t done = 10000; done <= visited_reads + 1; done += 10000) {
    log("processed %llu reads\n", (unsigned long long)done);
  }
  out.close();
  R.paths = (uint32_t)curr_path;
  if (!finished) {
    if (p.silver_path && p.max_paths > curr_path) { // goldrush_path.cpp:1257-1264
      log("WARNING: Expected %llu silver paths, but only %llu generated.\n"
          "Possible reasons include:\n\t- Input reads sorted by chromosome/position\n"
          "\t- Genome size set too large\n",
          (unsigned long long)p.max_paths, (unsigned long long)curr_path);
    }
    if (o->verbose) {
      log_path_stat(log, curr_path, cur, phred_sum);
    }
    log("assigned\nin %.4f\n", R.ms_pass2 / 1e3);
  }
  if (cap_out) {
    // The golden run's input: `cat $(p1)_*.fq` (bin/goldrush:250-251).  The shell expands the glob in
    // lexicographic order of the file names (_1, _10, _11, _2, ...), and the selection depends on
    // read order, so the paths are joined in that order, not numerically.  Slice mode: every rank
    // holds the records of its own reads only, and each path is put together from the ranks' parts
    // (rank order = read order) on every rank, straight into its place in the joined buffer.
    const bool spread = slice_mode && c_world_now > 1;
    const uint32_t n_paths =
      spread ? (uint32_t)std::min<uint64_t>(curr_path, std::max<uint64_t>(1, p.max_paths))
             : (uint32_t)capture->size();
    capture->resize(n_paths);
    std::vector<size_t> order(n_paths);
    for (size_t i = 0; i < order.size(); ++i) {
      order[i] = i;
    }
    std::sort(order.begin(), order.end(), [](size_t a, size_t b) {
      return std::to_string(a + 1) + ".fq" < std::to_string(b + 1) + ".fq";
    });
    std::vector<uint64_t> mine(n_paths), all((size_t)n_paths * c_world_now), sizes((size_t)c_world_now);
    for (uint32_t q = 0; q < n_paths; ++q) {
      mine[q] = (*capture)[q].size();
    }
    if (spread) {
      if ((rc = grb_comm_allgather_host(ctx, mine.data(), mine.size() * 8, all.data(), all.size() * 8,
                                        sizes.data())) != GRB_OK) {
        return fail(rc);
      }
    } else {
      all = mine;
    }
    const int ranks = spread ? c_world_now : 1;
    size_t total = 0;
    for (uint64_t v : all) {
      total += v;
    }
    // Slice mode: the joined stream is a sequence of parts (path q, rank r), q in glob order, r in
    // rank order; the golden stage gives rank g a run of whole parts of about total / W bytes (the
    // ranks' shares of a path are about equal: reads are in random order), sent point to point
    // over NVLink.  No rank ever holds the whole silver output.
    bool redistributed = false;
    if (spread && total > 0) {
      struct Part
      {
        size_t q;
        int r;
        uint64_t start, bytes;
        int to;
      };
      std::vector<Part> parts;
      std::vector<int> to;
      std::vector<uint64_t> got;
      const bool all_fed = plan_silver_parts(all, n_paths, ranks, order, &to, &got);
      uint64_t cum = 0;
      for (size_t q : order) {
        for (int r = 0; r < ranks; ++r) {
          const uint64_t b = all[(size_t)r * n_paths + q];
          parts.push_back(Part{ q, r, cum, b, to.empty() ? 0 : to[parts.size()] });
          cum += b;
        }
      }
      if (all_fed) {
        int me = 0, w = 1;
        grb_comm_info(ctx, &me, &w);
        cap_out->joined.reset(new char[std::max<uint64_t>(got[(size_t)me], 1)]);
        cap_out->joined_len = got[(size_t)me];
        cap_out->slice_total = total;
        std::vector<grb_host_msg> sends, recvs;
        uint64_t at_local = 0;
        bool first = true;
        for (const Part& pt : parts) {
          if (pt.r == me && pt.bytes) {
            sends.push_back(grb_host_msg{ pt.to, 0, (*capture)[pt.q].data(), pt.bytes });
          }
          if (pt.to == me) {
            if (first) {
              cap_out->slice_offset = pt.start;
              first = false;
            }
            if (pt.bytes) {
              recvs.push_back(grb_host_msg{ pt.r, 0, cap_out->joined.get() + at_local, pt.bytes });
            }
            at_local += pt.bytes;
          }
        }
        if ((rc = grb_comm_exchange_host(ctx, sends.data(), (uint32_t)sends.size(), recvs.data(),
                                         (uint32_t)recvs.size())) != GRB_OK) {
          return fail(rc);
        }
        for (auto& v : *capture) {
          std::vector<char>().swap(v);
        }
        redistributed = true;
      }
    }
    if (!redistributed) {
      cap_out->joined.reset(new char[std::max<size_t>(total, 1)]);
      cap_out->joined_len = total;
    }
    size_t at = 0;
    for (size_t q : order) {
      if (redistributed) {
        break;
      }
      size_t path_bytes = 0;
      for (int r = 0; r < ranks; ++r) {
        path_bytes += all[(size_t)r * n_paths + q];
      }
      std::vector<char>& part = (*capture)[q];
      if (spread) {
        if ((rc = grb_comm_allgather_host(ctx, part.data(), part.size(), cap_out->joined.get() + at,
                                          path_bytes, sizes.data())) != GRB_OK) {
          return fail(rc);
        }
      } else if (path_bytes) {
        memcpy(cap_out->joined.get() + at, part.data(), path_bytes);
      }
      std::vector<char>().swap(part);
      at += path_bytes;
    }
    mark("join paths");
  }
  R.launches = grb_launch_count(ctx);
  R.out_digest = digest.h;
  R.ms_wall = now_ms() - t_wall0;
  if (res) {
    *res = R;
  }
  return GRB_OK;
}

extern "C" int
grb_test_plan_silver_parts(const uint64_t* bytes, uint32_t n_paths, int32_t ranks, int32_t* to,
                           uint64_t* got)
{
  std::vector<uint64_t> b(bytes, bytes + (size_t)n_paths * ranks), g;
  std::vector<size_t> order(n_paths);
  for (size_t i = 0; i < order.size(); ++i) {
    order[i] = i;
  }
  std::sort(order.begin(), order.end(), [](size_t x, size_t y) {
    return std::to_string(x + 1) + ".fq" < std::to_string(y + 1) + ".fq";
  });
  std::vector<int> t;
  const bool ok = plan_silver_parts(b, n_paths, ranks, order, &t, &g);
  for (size_t i = 0; i < t.size(); ++i) {
    to[i] = t[i];
  }
  for (size_t i = 0; i < g.size(); ++i) {
    got[i] = g[i];
  }
  return ok ? 1 : 0;
}

extern "C" size_t
grb_test_next_record_start(const char* fastq, size_t n, size_t from)
{
  return next_record_start(fastq, n, from);
}

extern "C" int
grb_run_path(const grb_run_options* o, const char* fastq, size_t fastq_len, grb_run_result* res,
             char* err, size_t err_cap)
{
  return run_path_impl(o, fastq, fastq_len, res, err, err_cap, nullptr);
}

// The two goldrush-path launches of one assembly (bin/goldrush:240-260) in one call: the silver
// run, then the golden run on the concatenated silver paths, which stay in host memory instead of
// travelling through <p>_N.fq files and a second process.  Both stages write exactly the files the
// two separate runs write (when write_outputs is set).
extern "C" int
grb_run_two_stage(const grb_run_options* silver, const grb_run_options* golden, const char* fastq,
                  size_t fastq_len, grb_run_result* res_silver, grb_run_result* res_golden,
                  char* err, size_t err_cap)
{
  if (!silver || !golden || !silver->params.silver_path || golden->params.silver_path) {
    return set_err(err, err_cap, "grb_run_two_stage: first stage must be --silver_path, second not",
                   GRB_ERR_ARG);
  }
  Capture cap;
  int rc = run_path_impl(silver, fastq, fastq_len, res_silver, err, err_cap, &cap);
  if (rc != GRB_OK) {
    return rc;
  }
  if (cap.joined_len == 0) { // the golden run would stop on an empty file (goldrush_path.cpp:247-250)
    return set_err(err, err_cap, "grb_run_two_stage: the silver stage selected no read", GRB_ERR_FORMAT);
  }
  grb_run_options g = *golden;
  g.fastq_offset = cap.slice_offset; // slice mode: each rank holds its share of the joined paths
  g.fastq_total = cap.slice_total;   // (0: the joined silver paths are whole on every rank)
  if (!g.input_path) {
    g.input_path = "(silver paths in memory)";
  }
  return run_path_impl(&g, cap.joined.get(), cap.joined_len, res_golden, err, err_cap, nullptr);
}
