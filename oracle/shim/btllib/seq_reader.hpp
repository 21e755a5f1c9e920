// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for btllib's <btllib/seq_reader.hpp>, so the
// reference sources compile unmodified into oracle/_ref/. Surface used by the reference:
//   goldrush_path/goldrush_path.cpp:87-89,246-258 (ctor, get_format, range-for from an OpenMP team),
//   goldrush_path/read_hashing.cpp:15-26,88-90 (read_block, get_block_size),
//   goldrush_path/goldrush_path.cpp:1210-1212 (LONG_MODE_BUFFER_SIZE / LONG_MODE_BLOCK_SIZE),
//   goldrush_path/ntcard.hpp:200-203.
// Behaviour restated from btllib's documented SeqReader: four-line FASTQ / two-line-or-wrapped
// FASTA, Record::id = header up to the first blank, Record::comment = the rest, sequence
// folded to upper case (the default, no NO_FOLD_CASE flag is passed by the reference),
// Record::num = 0-based record index, records handed out once each to whichever thread asks.
#ifndef GRB_SHIM_BTLLIB_SEQ_READER_HPP
#define GRB_SHIM_BTLLIB_SEQ_READER_HPP

#include "order_queue.hpp"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <mutex>
#include <string>

namespace btllib {

class SeqReader
{
public:
  struct Flag
  {
    static const unsigned NO_FOLD_CASE = 1;
    static const unsigned NO_TRIM_MASKED = 2;
    static const unsigned SHORT_MODE = 4;
    static const unsigned LONG_MODE = 8;
  };

  enum class Format
  {
    UNDETERMINED,
    FASTA,
    FASTQ,
    SAM,
    INVALID
  };

  struct Record
  {
    size_t num = std::numeric_limits<size_t>::max();
    std::string id;
    std::string comment;
    std::string seq;
    std::string qual;

    operator bool() const { return !seq.empty(); }
  };

  static const size_t SHORT_MODE_BUFFER_SIZE = 32;
  static const size_t SHORT_MODE_BLOCK_SIZE = 32;
  static const size_t LONG_MODE_BUFFER_SIZE = 4;
  static const size_t LONG_MODE_BLOCK_SIZE = 1;

  SeqReader(const std::string& path, unsigned flags, unsigned threads = 3)
    : m_fold_case(!(flags & Flag::NO_FOLD_CASE))
    , m_block_size((flags & Flag::LONG_MODE) ? LONG_MODE_BLOCK_SIZE : SHORT_MODE_BLOCK_SIZE)
  {
    (void)threads;
    m_buf.resize(1 << 22);
    m_in.rdbuf()->pubsetbuf(&m_buf[0], (std::streamsize)m_buf.size());
    m_in.open(path, std::ios::in | std::ios::binary);
    if (!m_in) {
      std::cerr << "SeqReader: cannot open " << path << std::endl;
      std::exit(EXIT_FAILURE);
    }
    const int c = m_in.peek();
    if (c == '@') {
      m_format = Format::FASTQ;
    } else if (c == '>') {
      m_format = Format::FASTA;
    } else {
      m_format = Format::INVALID;
    }
  }

  Format get_format() const { return m_format; }
  size_t get_block_size() const { return m_block_size; }

  // Thread-safe: the next record in file order, or an empty record at end of input.
  Record read()
  {
    Record rec;
    std::lock_guard<std::mutex> lock(m_mutex);
    read_locked(rec);
    return rec;
  }

  OrderQueueMPMC<Record>::Block read_block()
  {
    OrderQueueMPMC<Record>::Block block(m_block_size);
    std::lock_guard<std::mutex> lock(m_mutex);
    block.num = m_block_num;
    while (block.count < m_block_size) {
      Record rec;
      if (!read_locked(rec)) {
        break;
      }
      block.data[block.count++] = std::move(rec);
    }
    if (block.count > 0) {
      ++m_block_num;
    }
    return block;
  }

  class RecordIterator
  {
  public:
    void operator++() { m_record = m_reader.read(); }
    bool operator!=(const RecordIterator& other) const
    {
      return bool(m_record) || !other.m_end;
    }
    Record operator*() { return std::move(m_record); }

  private:
    friend class SeqReader;
    RecordIterator(SeqReader& reader, bool end)
      : m_reader(reader)
      , m_end(end)
    {
      if (!end) {
        ++(*this);
      }
    }
    SeqReader& m_reader;
    Record m_record;
    bool m_end;
  };

  RecordIterator begin() { return RecordIterator(*this, false); }
  RecordIterator end() { return RecordIterator(*this, true); }

private:
  static void rtrim(std::string& s)
  {
    while (!s.empty() && std::isspace((unsigned char)s.back())) {
      s.pop_back();
    }
  }

  bool read_locked(Record& rec)
  {
    std::string header;
    while (std::getline(m_in, header)) {
      rtrim(header);
      if (!header.empty()) {
        break;
      }
    }
    if (header.empty()) {
      return false;
    }
    size_t ws = 1;
    while (ws < header.size() && header[ws] != ' ' && header[ws] != '\t') {
      ++ws;
    }
    rec.id = header.substr(1, ws - 1);
    size_t cs = ws;
    while (cs < header.size() && (header[cs] == ' ' || header[cs] == '\t')) {
      ++cs;
    }
    rec.comment = header.substr(cs);
    if (m_format == Format::FASTQ) {
      std::string plus;
      if (!std::getline(m_in, rec.seq) || !std::getline(m_in, plus) ||
          !std::getline(m_in, rec.qual)) {
        return false;
      }
      rtrim(rec.seq);
      rtrim(rec.qual);
    } else {
      // FASTA: concatenate lines up to the next header
      std::string line;
      while (m_in.peek() != '>' && std::getline(m_in, line)) {
        rtrim(line);
        rec.seq += line;
      }
    }
    if (m_fold_case) {
      for (auto& c : rec.seq) {
        c = (char)std::toupper((unsigned char)c);
      }
    }
    rec.num = m_record_num++;
    return !rec.seq.empty();
  }

  std::ifstream m_in;
  std::string m_buf;
  bool m_fold_case;
  size_t m_block_size;
  Format m_format = Format::UNDETERMINED;
  std::mutex m_mutex;
  size_t m_record_num = 0;
  size_t m_block_num = 0;
};

} // namespace btllib

#endif
