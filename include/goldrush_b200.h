/*
 * goldrush_b200.h — C ABI of libgoldrush_b200.so, the B200 (sm_100a) engine behind GoldRush-Path's
 * read-selection loop.  Plain pointers and sizes only; every device allocation lives behind the
 * opaque grb_ctx.  All citations are file:line under the reference tree (bcgsc/goldrush v1.2.2).
 *
 * The reference has no FFI; the seams below are the function boundaries inside
 * goldrush_path/goldrush_path.cpp that a maintainer would re-point at this library
 * (see INTEGRATION.md for the binding on the reference side).
 *
 * Conventions: every call returns 0 on success and a negative grb_status otherwise;
 * grb_last_error(ctx) gives the message.  A context is single-threaded (one host thread drives it)
 * and owns one CUDA device.  Nothing here falls back to a CPU implementation: without a usable
 * CUDA device grb_create fails with GRB_ERR_CUDA.
 */
#ifndef GOLDRUSH_B200_H
#define GOLDRUSH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grb_ctx grb_ctx;

typedef enum grb_status
{
  GRB_OK = 0,
  GRB_ERR_ARG = -1,    /* invalid argument / option (reference: opt.cpp:176-215 -> exit(1)) */
  GRB_ERR_CUDA = -2,   /* CUDA runtime error or no device */
  GRB_ERR_STATE = -3,  /* call out of order (e.g. query before grb_finalize_bitvector) */
  GRB_ERR_FORMAT = -4, /* input is not FASTQ (reference: goldrush_path.cpp:247-250) */
  GRB_ERR_NOMEM = -5
} grb_status;

/* ---- options: one field per opt:: global (goldrush_path/opt.hpp:9-38, defaults opt.cpp:5-32) ---- */
typedef struct grb_params
{
  uint64_t assigned_max;   /* -a */
  uint64_t unassigned_min; /* -u */
  uint64_t tile_length;    /* -t */
  uint64_t block_size;     /* -b */
  uint64_t hash_universe;  /* -H (0 = derive, goldrush_path.cpp:1109-1123) */
  uint64_t genome_size;    /* -g */
  uint64_t kmer_size;      /* -k */
  uint64_t weight;         /* -w */
  uint64_t min_length;     /* -m */
  uint64_t hash_num;       /* -h */
  double occupancy;        /* -o */
  double ratio;            /* -r */
  uint64_t max_paths;      /* -M */
  uint64_t threshold;      /* -x */
  uint32_t phred_min;      /* -P (already resolved: never 0 when pass 1 starts) */
  uint32_t phred_delta;    /* -d */
  int32_t silver_path;     /* --silver_path */
  int32_t device;          /* CUDA device ordinal */
  const char* const* seeds; /* hash_num NUL-terminated strings from grb_make_seed_pattern */
} grb_params;

/* Fills the reference's defaults (opt.cpp:5-32); seeds = NULL, device = 0. */
void grb_params_default(grb_params* p);

/* ---- host-side scalar helpers (no device) ---- */

/* spaced_seeds.cpp:7-68 make_seed_pattern.  out[i] must hold k + h bytes.  preset may be "" / NULL
 * (random design, srand(123) + rand()%2 exactly as the reference). */
int grb_make_seed_pattern(const char* preset, unsigned k, unsigned weight, unsigned h, char** out);
/* MIBloomFilter.hpp:94-101 calcOptimalSize */
uint64_t grb_calc_optimal_size(uint64_t entries, unsigned hash_num, double occupancy);
/* goldrush_path.cpp:1114-1121 default hash universe */
uint64_t grb_default_hash_universe(uint64_t weight, uint64_t genome_size, uint64_t hash_num);
/* calc_phred_average.cpp:32-42: final log10 / casts from the two running sums the device returns */
void grb_phred_finalize(double first_half_sum, double total_sum, uint64_t n, uint32_t* avg,
                        uint32_t* delta);

/* ---- context ---- */
int grb_create(const grb_params* p, grb_ctx** out);
void grb_destroy(grb_ctx* ctx);
const char* grb_last_error(const grb_ctx* ctx); /* ctx may be NULL: message of a failed grb_create */
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
uint64_t grb_launch_count(const grb_ctx* ctx);
/* Device allocations of a destroyed context are kept in a process-wide cache and handed to the next
 * context of the same shape (the reference re-allocates its filter per process,
 * MIBFConstructSupport.hpp:66-84 — a resident service or the bench loop need not).  These two
 * calls report and empty that cache; GRB_POOL=0 in the environment disables it. */
uint64_t grb_cached_memory_bytes(void);
void grb_release_cached_memory(void);
/* Page-locks a host range for the ingest copies (cudaHostRegister in pieces of 1 GiB, so that a
 * range larger than the system lets one call lock is pinned as far as possible: the copies of the
 * rest go through the driver's staging buffer, slower but correct).  *pinned_bytes = leading bytes
 * that are now page-locked.  grb_host_unpin releases what grb_host_pin locked. */
int grb_host_pin(void* p, size_t n, size_t* pinned_bytes);
void grb_host_unpin(void* p, size_t pinned_bytes);

/* Size hint before the first grb_reads_ingest_fastq: total FASTQ bytes to come, so that the read
 * store is allocated once instead of grown chunk by chunk. */
int grb_reads_reserve(grb_ctx* ctx, uint64_t fastq_bytes);

/* ---- K1: FASTQ decode + Phred sums into the device read store ----
 * replaces btllib::SeqReader (goldrush_path.cpp:87,246; read_hashing.cpp:89-90; ntcard.hpp:200)
 * and the summation loop of calc_phred_average (calc_phred_average.cpp:15-30). */
typedef struct grb_read_meta
{
  uint64_t hdr_off;  /* byte offsets into the concatenation of everything ingested so far */
  uint64_t seq_off;
  uint64_t qual_off;
  uint32_t hdr_len;  /* header line length without '@' and newline */
  uint32_t len;      /* bases */
  double phred_first_half_sum; /* running sum captured at i == n/2 - 1 (calc_phred_average.cpp:26-28) */
  double phred_total_sum;
  uint32_t non_acgt; /* 1 if the sequence holds a byte outside ACGTacgt (goldrush_path.cpp:293) */
  uint32_t qual_len; /* bytes of the quality line (the n of calc_phred_average) */
  uint64_t name_hash; /* FNV-1a of the record id = header up to the first blank (btllib Record::id,
                         the key of filter_out_reads, goldrush_path.cpp:266-299,919-932) */
} grb_read_meta;

/* grb_phred_finalize for n reads at once, from the metadata K1 returned (host threads) */
void grb_phred_finalize_batch(const grb_read_meta* meta, uint64_t n, uint32_t* avg, uint32_t* delta);

/* Decodes every complete 4-line record in bytes[0, n) and appends it to the read store.
 * *consumed = bytes used; re-send the tail with the next chunk (final != 0: a last record without
 * trailing newline is accepted).  bytes is HOST memory (pageable or pinned). */
int grb_reads_ingest_fastq(grb_ctx* ctx, const char* bytes, size_t n, int final, size_t* consumed);
/* Read-ahead hint: the following grb_reads_ingest_fastq calls take consecutive chunks of the HOST
 * range [base, base + total) (a whole file, pinned for full effect).  While one chunk is decoded
 * the next one's bytes are already copied on a second stream.  base = NULL switches it off; the
 * range must stay valid until the last chunk has been ingested. */
int grb_reads_readahead(grb_ctx* ctx, const char* base, size_t total);
/* Byte offset of the first ingested byte within the whole input (default 0): the offsets in
 * grb_read_meta are relative to the whole file.  Before the first grb_reads_ingest_fastq only.
 * Used when a rank ingests its own slice of the file (several GPUs). */
int grb_reads_set_origin(grb_ctx* ctx, uint64_t byte_offset);
/* Several GPUs: every rank has ingested its own consecutive slice of the input (rank order = file
 * order); afterwards every rank's store holds ALL reads in file order -- packed bases, masks and
 * grb_read_meta travel over NVLink (NCCL broadcasts), 0.28 bytes per base instead of 2 bytes per base
 * and rank over PCIe.  No-op for a context without communicator.  Collective: every rank calls it. */
int grb_reads_allgather(grb_ctx* ctx);
/* reads [*first, *first + *count) of the store are the ones this rank ingested itself */
int grb_reads_own_range(const grb_ctx* ctx, uint64_t* first, uint64_t* count);
uint64_t grb_reads_count(const grb_ctx* ctx);
int grb_reads_get_meta(grb_ctx* ctx, uint64_t first, uint64_t count, grb_read_meta* out);
/* per-read flags decided by the host from grb_read_meta (length / Phred / delta / ACGT / -f list) */
#define GRB_READ_PASS1 1u /* hash whole read into the bit vector  (goldrush_path.cpp:302-305) */
#define GRB_READ_PASS2 2u /* visit in the ordered selection loop  (goldrush_path.cpp:907-932) */
int grb_reads_set_flags(grb_ctx* ctx, uint64_t first, uint64_t count, const uint8_t* flags);
void grb_reads_clear(grb_ctx* ctx);
/* one quality string: the two sums of calc_phred_average.cpp:15-30, computed on the device */
int grb_phred_sums(grb_ctx* ctx, const char* qual, size_t n, double* first_half_sum,
                   double* total_sum);

/* ---- K5: ntCard-style estimate (ntcard.hpp:248-274 calc_ntcard_genome_size) ----
 * Hashes every read in the store.  input_bytes = size of the FASTQ file (ntcard.hpp:180-183 picks
 * sBits from it).  per_pattern may be NULL. */
int grb_estimate_cardinality(grb_ctx* ctx, uint64_t input_bytes, uint64_t* per_pattern,
                             uint64_t* total);

/* ---- K2 + K4a/K4b: the bit vector (MIBFConstructSupport.hpp:66-84,134-147,165-181) ---- */
int grb_filter_alloc(grb_ctx* ctx, uint64_t filter_bits);
/* hashes every GRB_READ_PASS1 read whole and sets hash % filter_bits for each pattern */
int grb_build_bitvector(grb_ctx* ctx);
/* same for reads [first, first+count) only (multi-GPU sharding of pass 1) */
int grb_build_bitvector_range(grb_ctx* ctx, uint64_t first, uint64_t count);
/* rank build + allocation of the ID / count slots; *pop = number of set bits (= m_data length) */
int grb_finalize_bitvector(grb_ctx* ctx, uint64_t* pop);
/* silver-path rollover: reset_counts + reset_ID_vector (MIBFConstructSupport.hpp:183-186,
 * MIBloomFilter.hpp:679-682) */
int grb_reset_ids(grb_ctx* ctx);

/* ---- K2 + K3 + K4c: the ordered selection loop (goldrush_path.cpp:892-1094, 1229-1256) ---- */
typedef enum grb_verdict
{
  GRB_NOT_VISITED = 0, /* after exit(0) at path M+1, or never reached */
  GRB_SKIPPED = 1,     /* too short / filtered (goldrush_path.cpp:907-932) */
  GRB_UNTRIMMED = 2,   /* inserted whole, "_untrimmed" (:978-1011) */
  GRB_TRIMMED = 3,     /* inserted tiles [trim_start, trim_end], "_trimmed" (:1035-1079) */
  GRB_ASSIGNED = 4     /* dropped: fully assigned or bad flanks (:1013-1023, :1083-1088) */
} grb_verdict;

typedef struct grb_decision
{
  uint8_t verdict;     /* grb_verdict */
  uint8_t pad[3];
  uint32_t path;       /* 1-based silver path the record is written to (1 in golden mode) */
  uint32_t trim_start; /* tiles, inclusive; valid for GRB_TRIMMED */
  uint32_t trim_end;
  uint32_t num_tiles;
  uint32_t num_assigned;
} grb_decision;

/* counters of log_info_struct (goldrush_path.cpp:41-51), snapshot at each path rollover */
typedef struct grb_path_stats
{
  uint64_t valid_reads;
  uint64_t total_tiles;
  uint64_t assigned_tiles;
  uint64_t unassigned_tiles;
  uint64_t queries;
  uint64_t hits;
  uint64_t misses;
  uint64_t num_reads_in_path;
  uint64_t inserted_bases;
  uint64_t rollover_read; /* store index of the read whose insertion closed this path */
  double phred_sum_in_path; /* filled by grb_run_path (host side); 0 from grb_select_reads */
} grb_path_stats;

/* Runs the selection loop over reads [first, first+count) of the store IN ORDER, continuing from
 * the state left by the previous call.  decisions[count].  *finished != 0 once the reference would
 * have called exit(0) (goldrush_path.cpp:174-176).  stats: up to stats_cap snapshots appended at
 * each rollover during this call (*n_stats). */
int grb_select_reads(grb_ctx* ctx, uint64_t first, uint64_t count, grb_decision* decisions,
                     grb_path_stats* stats, uint32_t stats_cap, uint32_t* n_stats, int* finished);
/* running counters of the current (unfinished) path + loop state */
int grb_select_state(grb_ctx* ctx, grb_path_stats* current, uint64_t* curr_path,
                     uint32_t* ids_inserted);

/* ---- parity / debug exports (used by tests; each names the reference routine it mirrors) ---- */
/* multiLensfrHashIterator over one sequence (multiLensfrHashIterator.hpp:29-68):
 * out[frame * hash_num + pattern], frames = n - k + 1, stale-tail semantics included. */
int grb_hash_sequence(grb_ctx* ctx, const char* seq, size_t n, uint64_t* out);
/* plain LSB-first bit vector, (filter_bits + 63) / 64 words (sdsl::bit_vector layout) */
int grb_copy_bitvector(grb_ctx* ctx, uint64_t* words);
int grb_load_bitvector(grb_ctx* ctx, const uint64_t* words); /* before grb_finalize_bitvector */
/* rank_support_il<1>(pos): set bits in [0, pos); also the bit itself (MIBloomFilter.hpp:465-491) */
int grb_rank(grb_ctx* ctx, const uint64_t* pos, size_t n, uint64_t* rank, uint8_t* bit);
int grb_get_ids(grb_ctx* ctx, const uint64_t* rank, size_t n, uint32_t* ids, uint32_t* counts);
int grb_set_ids(grb_ctx* ctx, const uint64_t* rank, size_t n, const uint32_t* ids,
                const uint32_t* counts);
/* per-tile vote of calc_num_assigned_tiles (goldrush_path.cpp:544-626) for one stored read:
 * best_id/best_count = arg-max (ties -> smallest id); candidates = every (id,count) with count > 2,
 * at most cand_cap per tile written to cand_ids/cand_counts[tile * cand_cap + j], n_cand[tile] =
 * true number.  counters[3] += {queries, hits, misses}. */
int grb_query_read(grb_ctx* ctx, uint64_t read_idx, uint32_t* best_id, uint32_t* best_count,
                   uint32_t* n_cand, uint32_t* cand_ids, uint32_t* cand_counts, uint32_t cand_cap,
                   uint64_t* counters);
/* insertMIBF(miBF, hashes, start, end, id) for one stored read (MIBFConstructSupport.hpp:247-283) */
int grb_insert_tiles(grb_ctx* ctx, uint64_t read_idx, uint32_t tile_start, uint32_t tile_end,
                     uint32_t id);

/* ---- multi-GPU: one process per GPU, the filter replicated on every rank ----
 * A process-wide NCCL communicator, bound at run time (dlopen of libnccl.so.2; no link dependency).
 * Rank 0 calls grb_comm_unique_id and ships the 128 bytes to the other ranks by any means
 * (torch.distributed broadcast, MPI, a file); every rank then calls grb_comm_init on its device.
 * Contexts created afterwards on that device shard what shards:
 *   grb_reads_ingest_fastq + grb_reads_allgather
 *                        - each rank copies and decodes its own slice of the FASTQ, the packed
 *                          read store travels over NVLink (grb_run_path does this by itself)
 *   grb_build_bitvector  - pass 1 over this rank's share of the reads + OR-reduce of the vectors
 *                          (replaces the OpenMP team of fill_bit_vector, goldrush_path.cpp:252-311)
 *   grb_select_reads     - with GRB_SHARD_QUERY=1 the speculative query of each batch over this
 *                          rank's share of its tiles + all-gather of the per-tile results (off by
 *                          default: the exchange costs what it saves).  The ordered commit is
 *                          replicated either way, so every rank returns the same decisions
 * and grb_run_path inherits all of it.  GRB_COMM=0 in the environment keeps contexts single-GPU.
 * Errors of grb_comm_unique_id / grb_comm_init are reported by grb_last_error(NULL). */
int grb_comm_unique_id(uint8_t* out128);
int grb_comm_init(const uint8_t* id128, int rank, int world, int device);
void grb_comm_destroy(void);
/* OR-reduce of the ranks' bit vectors, for callers that shard pass 1 themselves with
 * grb_build_bitvector_range (every rank must call it, before grb_finalize_bitvector); no-op for a
 * context without communicator */
int grb_bitvector_or_reduce(grb_ctx* ctx);
/* ctx == NULL: the process-wide communicator; else what this context uses (0 / 1 if unsharded) */
int grb_comm_info(const grb_ctx* ctx, int* rank, int* world);
/* 1 if this context shards each batch's speculative query over the ranks, else 0 */
int grb_query_sharded(const grb_ctx* ctx);
/* Host-side all-gather of variable-length byte strings through the context's communicator (staged
 * through device memory): rank r contributes send[0, n); out (caller-allocated, capacity out_cap)
 * receives the strings of ranks 0..W-1 back to back, sizes[r] = bytes of rank r.  A context
 * without communicator copies send to out.  Collective. */
int grb_comm_allgather_host(grb_ctx* ctx, const void* send, uint64_t n, void* out, uint64_t out_cap,
                            uint64_t* sizes);
/* Host-side point-to-point exchange through the communicator (staged through device memory, NCCL
 * send / recv in one group): message i of `sends` goes to rank peer[i]; `recvs` lists what arrives,
 * with the sizes the senders use.  Messages between one pair of ranks are matched in list order.
 * A message to oneself is a host copy.  Collective: every rank calls it (possibly with empty lists). */
typedef struct grb_host_msg
{
  int32_t peer;
  int32_t pad;
  void* ptr;
  uint64_t bytes;
} grb_host_msg;
int grb_comm_exchange_host(grb_ctx* ctx, const grb_host_msg* sends, uint32_t n_sends,
                           const grb_host_msg* recvs, uint32_t n_recvs);

/* ---- plumbing for callers that issue their own collectives (e.g. torch.distributed) on the raw
 * device pointers ---- */
int grb_bitvector_device(grb_ctx* ctx, void** dev_ptr, uint64_t* bytes);
/* dst |= src on this context's stream, n 8-byte words, both DEVICE pointers */
int grb_or_words(grb_ctx* ctx, void* dst, const void* src, uint64_t n_words);
int grb_sync(grb_ctx* ctx);
/* device timing of the last grb_build_bitvector / grb_select_reads / ingest call, milliseconds */
double grb_last_device_ms(const grb_ctx* ctx);
/* the cudaStream_t every kernel of this context is launched on (for the caller's own CUDA events) */
int grb_stream(grb_ctx* ctx, void** cuda_stream);
/* per-kernel-class device time: with profiling on, a CUDA event pair brackets every launch group of
 * the classes below on the context's stream; grb_kernel_time returns the sums since the last
 * grb_profile_enable (which also resets them).  Used by bench.py for the roofline. */
typedef enum grb_kernel_class
{
  GRB_K_FILL = 0,   /* K2+K4a  k_fill_part + k_fill_apply (or k_fill_bits, the direct form) */
  GRB_K_RANK = 1,   /* K4b     k_rank_partial + k_scan_u32 + k_rank_write */
  GRB_K_QUERY = 2,  /* K2+K3   k2_query (batch engine) / k_query (serial engine) */
  GRB_K_DECIDE = 3, /*         k_decide (serial engine only) */
  GRB_K_INSERT = 4, /* K4c     k3_bulk: reservoir inserts of the private ranks + write-back of the
                               shared ones (serial engine: k_insert_collect + k_insert_apply) */
  GRB_K_SMOOTH = 5, /*         k2_cmat: count matrix + smoothing + plan on the speculative votes */
  GRB_K_DEDUPE = 6, /*         the batch index: k3_mark ... k3_frames (which ranks do two probes of
                               the batch share, member lists, conflict-frame records) */
  GRB_K_COMMIT = 7, /* K4c     k3_fix: the ordered commit as a parallel fixed point */
  GRB_K_GATHER = 8, /*         multi-GPU exchange: NCCL all-gathers + k_or_gathered (0 on one GPU) */
  GRB_K_COUNT = 9
} grb_kernel_class;
int grb_profile_enable(grb_ctx* ctx, int on);
/* Phase clocks of the ordered commit kernel (k3_fix) since the loop state was created, as counted
 * by its CTA 0, each including the grid barrier in front of it: out[0] = SM cycles walking the
 * shared ranks' histories, out[2] = re-validating its share of the reads, out[3] = waiting for the
 * slowest CTA's reads + the in-order scan, out[4] = decisions and counters of the committed reads;
 * out[1] = scans that found a plan contradicted, out[5] = conflict frames, out[6] = reads committed,
 * out[7] = fixed-point iterations, out[8] = reads that inserted, out[9] = batches. */
int grb_commit_profile(grb_ctx* ctx, uint64_t* out10);
int grb_kernel_time(grb_ctx* ctx, int kclass, double* ms, uint64_t* n_launches);

/* ---- miBF probe microbenchmark (BASELINE.json configs[4]) ----
 * Builds a filter of filter_bits bits on this context (any earlier filter is dropped), sets a
 * fraction `fill` of them from synthetic keys (splitmix64), gives half of the ID slots an ID, then
 * times the product path's probe sequences over n_probes probes (n_probes / h keys drawn from the
 * inserted ones, so every probe meets a set bit): query = h block probes (bit + rank from one
 * 32-byte block) followed by the h slot reads; insert = the same block probes followed by the
 * reservoir read-modify-write of each slot (MIBFConstructSupport.hpp:271-282).  Best of `reps`
 * timed launches each, CUDA events.  Needs about filter_bits / 6 + 16 * fill * filter_bits bytes. */
typedef struct grb_probe_bench_result
{
  double query_ms;
  double insert_ms;
  uint64_t probes;
  uint64_t pop;
  uint64_t filter_bits;
  uint64_t footprint_bytes; /* filter blocks + ID slots */
  uint64_t checksum;        /* sum of the IDs read by the query launches */
  uint64_t keys_filled;
  uint64_t probes_missed;   /* keys of the query launches that met a clear bit: must be 0 */
  double line_query_ms;     /* the same probes as ONE random 128-byte line each (two dependent 16-byte
                               reads inside it) over line_bytes of memory: what a line-local filter
                               layout could reach */
  uint64_t line_bytes;
} grb_probe_bench_result;
int grb_probe_bench(grb_ctx* ctx, uint64_t filter_bits, double fill, uint32_t h, uint64_t n_probes,
                    uint64_t seed, int reps, grb_probe_bench_result* out);

/* ---- whole stage: what goldrush_path.cpp main() does between option parsing and exit ----
 * (goldrush_path.cpp:1096-1275).  fastq = the whole input file in host memory.  Writes
 * <prefix>_N.fq / <prefix>.fa exactly as the reference.  log may be NULL (else a FILE*-like
 * callback receives the stderr text). */
typedef struct grb_run_options
{
  grb_params params;       /* seeds may be NULL: derived from seed_preset */
  const char* seed_preset; /* -s */
  const char* prefix;      /* -p */
  const char* filter_file; /* -f, may be NULL */
  const char* input_path;  /* -i, used for messages and (if fastq == NULL) read from disk */
  int32_t ntcard;          /* --ntcard */
  int32_t verbose;         /* --verbose */
  int32_t debug;           /* --debug */
  int32_t write_outputs;   /* 0: decide only (bench) */
  int32_t quiet;           /* 1: no stderr text at all */
  int32_t jobs;            /* -j: host threads for the host-side bookkeeping (0 = default) */
  /* Several GPUs, slice mode: `fastq` holds only bytes [fastq_offset, fastq_offset + fastq_len) of an
   * input of fastq_total bytes, starting and ending at record boundaries, rank r's slice following
   * rank r-1's (fastq_total = 0: `fastq` is the whole input and each rank cuts its own share out of
   * it).  A rank then touches no byte of another rank's reads: the record digest is assembled from
   * per-record hashes exchanged between the ranks, and write_outputs must be 0. */
  uint64_t fastq_offset;
  uint64_t fastq_total;
} grb_run_options;

typedef struct grb_run_result
{
  uint64_t num_reads;
  uint64_t num_passed_reads;  /* pass 1 */
  uint64_t bases_pass1;       /* bases hashed into the bit vector */
  uint64_t reads_visited;     /* pass 2: reads that reached the query */
  uint64_t bases_pass2;       /* sum of num_tiles * tile_length over visited reads */
  uint64_t reads_selected;    /* untrimmed + trimmed */
  uint64_t bases_selected;
  uint64_t filter_bits;
  uint64_t pop;
  uint32_t phred_min;
  uint32_t paths;             /* silver paths completed or in progress */
  double ms_ingest;           /* device time per phase (CUDA events) */
  double ms_pass1;
  double ms_rank;
  double ms_pass2;
  double ms_wall;             /* host wall clock of the whole call */
  uint64_t launches;
  uint64_t out_digest;        /* order-sensitive digest of the output records (also when nothing is
                                 written): FNV-1a over the 8-byte hashes of the records in output
                                 order, a record's hash being FNV-1a over its bytes taken as
                                 little-endian 8-byte words (last word zero-padded), then its length.
                                 oracle/grb_digest.cpp computes the same from output files. */
} grb_run_result;

int grb_run_path(const grb_run_options* opt, const char* fastq, size_t fastq_len,
                 grb_run_result* result, char* err, size_t err_cap);

/* Both goldrush-path launches of one assembly (bin/goldrush:240-260: the --silver_path run, `cat` of
 * its <p>_N.fq files, the golden run on that file) in one call.  The silver paths go to the second
 * stage through host memory; each stage still writes the files its own launch would write when
 * its write_outputs is set, byte for byte.  silver->params.silver_path must be set, golden's not. */
int grb_run_two_stage(const grb_run_options* silver, const grb_run_options* golden,
                      const char* fastq, size_t fastq_len, grb_run_result* res_silver,
                      grb_run_result* res_golden, char* err, size_t err_cap);

/* ---- (f4) GoldPolish targeted Bloom filters ----
 * The computational core of goldpolish-targeted-bfs (subprojects/goldpolish/src/
 * goldpolish_targeted_bfs.cpp:53-149 serve_batch, utils.cpp:96-123 fill_bfs): per batch of target
 * sequences and per k value, every k-mer of the mapped reads goes, in order, through a counting
 * Bloom filter and into a plain Bloom filter once its count reaches the target's threshold.  The
 * named-pipe protocol, the sequence / mapping index files and the .bf file header stay on the
 * caller's side (control plane).  hash_num = 4, cbf_bytes = 10 MiB, bf_bytes = 512 KiB in the
 * reference (:261-263). */
typedef struct grb_polish_params
{
  uint32_t hash_num;
  uint32_t n_k;
  const uint32_t* k_values; /* one Bloom filter per k per batch ("k<k>.bf", :217-221) */
  uint64_t cbf_bytes;
  uint64_t bf_bytes;
} grb_polish_params;
/* goldpolish_targeted_bfs.cpp:43-51 mappings_bases_to_kmer_threshold */
int grb_polish_kmer_threshold(uint64_t mappings_bases);
/* What serve_batch derives for one target (:88-127): its mapped reads sorted by (Phred average as
 * size_t descending, id ascending), the first min(n, target_len * subsample_max_per_10kbp / 10000)
 * of them used, the k-mer threshold from their total length.  order_out[n_mappings] = indices into
 * the inputs in sorted order, *n_used = how many of them are inserted.  Host only. */
int grb_polish_plan_target(uint64_t target_len, double subsample_max_per_10kbp, uint32_t n_mappings,
                           const char* const* ids, const double* phred_avg, const uint64_t* lens,
                           uint32_t* order_out, uint32_t* n_used, int32_t* kmer_threshold);
/* fill_bfs over every batch: batch b = reads [batch_first[b], batch_first[b + 1]) in serve_batch's
 * order, read r = seqs[seq_off[r], seq_off[r + 1]) (HOST memory) inserted with thresholds[r] (its
 * target's k-mer threshold, >= 4).  out_bfs[(b * n_k + i) * bf_bytes ...] = the Bloom filter of
 * batch b and k_values[i], raw bit array (bit of hash h = h % (8 * bf_bytes), LSB first). */
int grb_polish_fill_batches(grb_ctx* ctx, const grb_polish_params* p, uint32_t n_batches,
                            const uint64_t* batch_first, const char* seqs, const uint64_t* seq_off,
                            const uint32_t* thresholds, uint8_t* out_bfs);
/* test hook: the same jobs (csrc/polish_core.h, one code for host and device) run on the host, for
 * CPU-side checks against the oracle.  Never called by the product path. */
int grb_test_polish_fill_host(const grb_polish_params* p, uint32_t n_batches, const uint64_t* batch_first,
                              const char* seqs, const uint64_t* seq_off, const uint32_t* thresholds,
                              uint8_t* out_bfs);
/* test hook: the warp kernel's algorithm lane by lane on the host (packed segments, table-driven
 * hashing, groups of 32 k-mers applied at once when no counter is shared between two lanes, in lane
 * order otherwise); also reports how many groups there were and how many needed lane order. */
int grb_test_polish_fill_host_grouped(const grb_polish_params* p, uint32_t n_batches,
                                      const uint64_t* batch_first, const char* seqs, const uint64_t* seq_off,
                                      const uint32_t* thresholds, uint8_t* out_bfs, uint64_t* groups_total,
                                      uint64_t* groups_in_lane_order);

/* ---- (f4) the builder's inputs: sequence indexes and mappings (host only) ---- */
/* goldpolish-index (subprojects/goldpolish/src/goldpolish_index.cpp:13-14; SeqIndex::SeqIndex(seqs)
 * seqindex.cpp:12-66 and save :68-84): one line `id \t seq_start \t seq_len \t phred_avg` per record of
 * a FASTQ (4-line records) or FASTA (2-line records) file, phred_avg printed as the reference's stream
 * does (%g).  Lines come in file order (the reference: in the order of its hash table); a repeated id
 * keeps its first record. */
int grb_polish_index_build(const char* seqs_path, const char* index_path, char* err, size_t err_cap);
/* What goldpolish-targeted-bfs loads before it serves batches (goldpolish_targeted_bfs.cpp:271-281):
 * the index of the target sequences and of the mapped sequences (seqindex.cpp:86-123) and the
 * mappings -- SAM (columns 1, 3) for `.sam`, PAF (columns 1, 6) for `.paf`, otherwise ntLink's
 * `read target minimizers` triples, filtered per target to at most ceil(len * mx_max / 10000) reads
 * by raising the minimizer threshold within [1, 30] (mappings.cpp:12-31,71-107,226-320).  Mappings to
 * targets outside the index are dropped, a read counts once per target.  `.bam` is refused (the
 * reference pipes it through samtools). */
typedef struct grb_polish_inputs grb_polish_inputs;
int grb_polish_inputs_open(const char* target_index_path, const char* mappings_path,
                           const char* mapped_seqs_path, const char* mapped_index_path,
                           double mx_max_mapped_seqs_per_target_10kbp, grb_polish_inputs** out, char* err,
                           size_t err_cap);
void grb_polish_inputs_close(grb_polish_inputs* in);
/* AllMappings::get_mappings (mappings.cpp:322-329): how many reads are kept for the target; their ids,
 * each followed by '\n', into buf as far as cap allows (buf may be NULL). */
int64_t grb_polish_inputs_mappings(const grb_polish_inputs* in, const char* target_id, char* buf, size_t cap);
/* serve_batch (goldpolish_targeted_bfs.cpp:84-136) for n_batches batches in one call: batch b = the
 * target ids target_ids[batch_first[b] .. batch_first[b + 1]) in the order the caller would write them
 * to the batch's pipe.  Per target: grb_polish_plan_target over its mapped reads, the chosen reads
 * fetched from the mapped-sequence file (SeqIndex::get_seq, seqindex.hpp:64-103), then
 * grb_polish_fill_batches; out_bfs as there.  A target id missing from the target index, or a mapped
 * read missing from the mapped-sequence index, is an error (the reference: uncaught std::out_of_range). */
int grb_polish_serve_batches(grb_ctx* ctx, const grb_polish_inputs* in, const grb_polish_params* p,
                             double subsample_max_mapped_seqs_per_target_10kbp, uint32_t n_batches,
                             const uint64_t* batch_first, const char* const* target_ids, uint8_t* out_bfs,
                             char* err, size_t err_cap);
/* test hook: the same gathering, then the host-compiled job code instead of the GPU; also reports how
 * many reads and bases the batches hold.  Never called by the product path. */
int grb_test_polish_serve_batches_host(const grb_polish_inputs* in, const grb_polish_params* p,
                                       double subsample_max_mapped_seqs_per_target_10kbp, uint32_t n_batches,
                                       const uint64_t* batch_first, const char* const* target_ids,
                                       uint8_t* out_bfs, uint64_t* n_reads, uint64_t* n_bases, char* err,
                                       size_t err_cap);

/* ---- synthetic reads; host only, used by bench.py and the tests ---- */
typedef struct grb_synth_params
{
  uint64_t genome_len;
  uint64_t seed;
  double coverage;
  uint32_t read_len; /* 0 = log-normal lengths with N50 = n50 */
  uint32_t n50;
  double sub_rate, ins_rate, del_rate;
  uint32_t qmin, qmax; /* per-read base quality ~ UniformInt[qmin, qmax) */
} grb_synth_params;

uint64_t grb_synth_num_reads(const grb_synth_params* p);
char* grb_synth_fastq(const grb_synth_params* p, uint64_t first, uint64_t count,
                      uint64_t* out_len);
void grb_free_host(void* p);

/* ---- test hook: the device decision code (csrc/decide.cuh) compiled for the host, for CPU-side
 * fuzzing against the oracle.  out_plan[9] = verdict, trim_start, trim_end, first_id, id_bump,
 * n_blocks, n_assigned, out_bases lo/hi.  Never called by the product path. ---- */
int grb_test_decide_host(uint32_t n_tiles, const uint32_t* best_id, const uint32_t* best_count,
                         const uint32_t* n_cand, const uint32_t* cand_id, const uint32_t* cand_cnt,
                         uint32_t cand_cap, uint64_t threshold, uint64_t read_len,
                         uint64_t tile_length, uint64_t block_size, uint64_t unassigned_min,
                         uint64_t assigned_max, uint32_t* ids_inserted, uint32_t* out_ids,
                         uint8_t* out_assigned, uint32_t* out_plan);

/* ---- test hook: sizeof of every struct that crosses this ABI, in declaration order (grb_params,
 * grb_read_meta, grb_decision, grb_path_stats, grb_probe_bench_result, grb_run_options,
 * grb_run_result, grb_synth_params, grb_host_msg), so that a binding can check its mirror. ---- */
void grb_abi_sizes(uint64_t* out9);

/* ---- test hook: the record-boundary search grb_run_path uses to cut a FASTQ buffer into the ranks'
 * shares (several GPUs): first byte >= from at which a record starts (a line beginning with '@'
 * whose line after next begins with '+'), n if there is none.  Host only. ---- */
size_t grb_test_next_record_start(const char* fastq, size_t n, size_t from);
/* ---- test hook: grb_run_two_stage in slice mode hands the silver records to the golden stage's
 * ranks in whole parts (path q of rank r = bytes[r * n_paths + q]), paths in shell-glob order, ranks
 * in rank order: to[n_paths * ranks] = receiving rank of each part of that sequence, got[ranks] =
 * bytes per receiving rank; returns 0 if some rank would get nothing.  Host only. ---- */
int grb_test_plan_silver_parts(const uint64_t* bytes, uint32_t n_paths, int32_t ranks, int32_t* to,
                               uint64_t* got);

/* ---- test hook: the grouped half-hash code of csrc/nthash.cuh (what the query and fill kernels
 * inline) compiled for the host; out[frame * h + pattern], frames = n - k + 1, ACGT only.  Checked
 * against the oracle's SeedNtHash restatement by the CPU test suite.  Never called by the product. */
int grb_test_group_hash_host(const char* const* seeds, uint32_t h, const char* seq, size_t n,
                             uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* GOLDRUSH_B200_H */
