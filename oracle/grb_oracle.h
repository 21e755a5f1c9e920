/*
 * TEST INFRASTRUCTURE ONLY — the CPU oracle of GoldRush-Path's read-selection loop.
 *
 * A flat-array restatement of the reference algorithm (bcgsc/goldrush v1.2.2, goldrush_path/),
 * each routine citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this; the product (libgoldrush_b200.so, goldrush-path)
 * never links or calls it.
 *
 * Pinning: checked byte-for-byte against oracle/_ref/goldrush-path-ref (the reference's own
 * sources compiled unmodified against the stand-in headers in oracle/shim/) by
 * tests/test_oracle_vs_ref.py, and against the fixtures under tests/golden/ that were generated
 * from that binary.  PARITY UNPINNED at the btllib boundary (SeedNtHash values, SeqReader
 * parsing): btllib is not vendored in the reference tree and cannot be installed offline; the
 * reference ships no golden vectors for this path.
 */
#ifndef GRB_ORACLE_H
#define GRB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grbo_filter grbo_filter;

/* spaced_seeds.cpp:7-68 */
int grbo_make_seed_pattern(const char* preset, unsigned k, unsigned w, unsigned h, char** out);
/* MIBloomFilter.hpp:94-101 */
uint64_t grbo_calc_optimal_size(uint64_t entries, unsigned hash_num, double occupancy);
/* goldrush_path.cpp:1114-1121 */
uint64_t grbo_default_hash_universe(uint64_t weight, uint64_t genome_size, uint64_t hash_num);
/* calc_phred_average.cpp:8-43; sums[0] = first-half running sum, sums[1] = total */
void grbo_calc_phred_average(const char* qual, size_t n, uint32_t* avg, uint32_t* delta,
                             double* sums);
/* calc_phred_average.cpp:45-58 */
double grbo_sum_phred(const char* qual, size_t n);
/* multiLensfrHashIterator.hpp:29-68 over btllib::SeedNtHash: out[frame*h + p]; returns frames
 * (0 if the sequence is shorter than the longest seed) */
size_t grbo_hash_sequence(const char* seq, size_t n, const char* const* seeds, unsigned h,
                          uint64_t* out, size_t out_cap);

/* MIBFConstructSupport / MIBloomFilter pair */
grbo_filter* grbo_filter_new(uint64_t filter_bits, unsigned h);
void grbo_filter_free(grbo_filter* f);
/* insertBV, MIBFConstructSupport.hpp:134-147 */
void grbo_filter_insert_bv(grbo_filter* f, const uint64_t* hashes, size_t n);
/* setup + getEmptyMIBF, MIBFConstructSupport.hpp:165-181; returns pop (MIBloomFilter.hpp:538-546) */
uint64_t grbo_filter_setup(grbo_filter* f);
const uint64_t* grbo_filter_words(const grbo_filter* f, uint64_t* n_words);
/* rank_support_il<1>: set bits in [0,pos) */
uint64_t grbo_filter_rank(const grbo_filter* f, uint64_t pos, int* bit);
uint32_t grbo_filter_get_id(const grbo_filter* f, uint64_t rank);
uint32_t grbo_filter_get_count(const grbo_filter* f, uint64_t rank);
void grbo_filter_set(grbo_filter* f, uint64_t rank, uint32_t id, uint32_t count);
void grbo_filter_reset_ids(grbo_filter* f);
/* per-tile vote of calc_num_assigned_tiles, goldrush_path.cpp:544-626.  hashes[frames*h] of ONE
 * tile.  cand arrays sized cand_cap; returns number of candidates (count > 2). counters[3] +=
 * queries, hits, misses. */
uint32_t grbo_query_tile(const grbo_filter* f, const uint64_t* hashes, size_t frames,
                         uint32_t* best_id, uint32_t* best_count, uint32_t* cand_ids,
                         uint32_t* cand_counts, uint32_t cand_cap, uint64_t* counters);
/* insertMIBF(miBF, hash_vec, start, end, id), MIBFConstructSupport.hpp:247-283, on the
 * concatenated hashes of tiles [start,end) */
void grbo_insert_mibf(grbo_filter* f, const uint64_t* hashes, size_t n, uint32_t id);
/* threshold + smoothing part of calc_num_assigned_tiles, goldrush_path.cpp:628-889.
 * cand_off[num_tiles+1] indexes cand_ids/cand_counts.  ids/assigned are in/out. Returns number of
 * assigned tiles. */
size_t grbo_smooth_tiles(size_t num_tiles, uint32_t* ids, uint8_t* assigned,
                         const uint32_t* cand_off, const uint32_t* cand_ids,
                         const uint32_t* cand_counts, uint64_t threshold);
/* find_longest_stretch goldrush_path.cpp:195-233 + eval_flanks :341-527 */
void grbo_find_longest_stretch(const uint8_t* assigned, size_t n, int64_t* start, int64_t* end);
int grbo_eval_flanks(int64_t ls, int64_t le, const uint32_t* ids, size_t n, uint64_t* trim_start,
                     uint64_t* trim_end);
/* ntcard.hpp:248-274 on a FASTQ file; per_pattern[h] may be NULL */
uint64_t grbo_ntcard(const char* fastq_path, const char* const* seeds, unsigned h,
                     uint64_t* per_pattern);
/* the same with the input size given (0 = the file's own): ntcard.hpp:180-183 picks sBits = 11 from
 * 50 GB up, which no test file reaches */
uint64_t grbo_ntcard_sized(const char* fastq_path, const char* const* seeds, unsigned h,
                           uint64_t file_bytes, uint64_t* per_pattern);

/* the whole stage with the reference's command line (goldrush_path.cpp:1096-1275); returns the
 * process exit code instead of calling exit() */
int grbo_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif
