// fastx.h — FASTA/FASTQ record reader of the host: what the reference takes from the `fxread`
// crate (count.rs:24,64,87; offsetter.rs:172-173,190,195): gzip iff the path ends in ".gz"
// (all members of a multi-member file), format sniffed from the first byte ('>' = 2-line
// FASTA, '@' = 4-line FASTQ), id = header line without the marker, seq = raw bytes.
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace sgh {

struct FastxError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Buffered line source over a plain or gzip file.
class LineSource {
 public:
  explicit LineSource(const std::string& path);
  ~LineSource();
  LineSource(const LineSource&) = delete;
  LineSource& operator=(const LineSource&) = delete;
  // Next line without its terminator ('\n' or "\r\n" are both stripped of '\n' only, like
  // BufRead::read_until + trim of the newline).  Returns false at end of input.  The view is
  // valid until the next call.
  bool next(const char*& begin, size_t& len);

 private:
  bool refill();
  FILE* fp_ = nullptr;
  bool gz_ = false;
  z_stream zs_{};
  bool z_init_ = false, z_eof_ = false;
  std::vector<unsigned char> in_;
  std::vector<char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
};

class FastxReader {
 public:
  explicit FastxReader(const std::string& path);
  // Next record; views valid until the next call.  Throws FastxError on a truncated record.
  bool next(const char*& id, size_t& id_len, const char*& seq, size_t& seq_len);

 private:
  LineSource src_;
  int lines_per_record_ = 0;  // 2 FASTA, 4 FASTQ; 0 = not sniffed yet
  std::string id_;
};

}  // namespace sgh
