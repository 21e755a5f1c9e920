// TEST INFRASTRUCTURE ONLY.  grb-digest: the order-sensitive digest of a goldrush-path output, as
// grb_run_result.out_digest defines it (include/goldrush_b200.h): FNV-1a over the 8-byte FNV-1a
// hashes of the output records, in output order; a record's hash is FNV-1a over its bytes taken as
// little-endian 8-byte words (last word zero-padded), then its length.  A record is 4 lines of a
// silver-path FASTQ
// (goldrush_path.cpp:996-1002) or 2 lines of the golden-path FASTA (:1063-1070).
//   grb-digest <p>_1.fq <p>_2.fq ...      (files in path order)   |   grb-digest <p>.fa
// prints: <digest as decimal> <records> <bytes>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static inline uint64_t
fnv(uint64_t h, const unsigned char* p, size_t n)
{
  for (size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

int
main(int argc, char** argv)
{
  const uint64_t kInit = 1469598103934665603ull;
  uint64_t digest = kInit, records = 0, bytes = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = fopen(argv[a], "rb");
    if (!f) {
      perror(argv[a]);
      return 1;
    }
    const size_t nl = strlen(argv[a]);
    const int lines_per_record = (nl > 3 && strcmp(argv[a] + nl - 3, ".fa") == 0) ? 2 : 4;
    std::vector<unsigned char> buf((size_t)64 << 20), rec;
    int lines = 0;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
      bytes += got;
      size_t i = 0;
      while (i < got) {
        const unsigned char* nlp = (const unsigned char*)memchr(buf.data() + i, '\n', got - i);
        const size_t end = nlp ? (size_t)(nlp - buf.data()) + 1 : got;
        rec.insert(rec.end(), buf.data() + i, buf.data() + end);
        i = end;
        if (nlp && ++lines == lines_per_record) {
          uint64_t h = kInit;
          size_t q = 0;
          for (; q + 8 <= rec.size(); q += 8) {
            uint64_t w;
            memcpy(&w, rec.data() + q, 8);
            h = (h ^ w) * 1099511628211ull;
          }
          if (q < rec.size()) {
            uint64_t w = 0;
            memcpy(&w, rec.data() + q, rec.size() - q);
            h = (h ^ w) * 1099511628211ull;
          }
          h = (h ^ (uint64_t)rec.size()) * 1099511628211ull;
          digest = fnv(digest, (const unsigned char*)&h, 8);
          ++records;
          rec.clear();
          lines = 0;
        }
      }
    }
    fclose(f);
    if (lines != 0) {
      fprintf(stderr, "%s: truncated record\n", argv[a]);
      return 1;
    }
  }
  printf("%llu %llu %llu\n", (unsigned long long)digest, (unsigned long long)records,
         (unsigned long long)bytes);
  return 0;
}
