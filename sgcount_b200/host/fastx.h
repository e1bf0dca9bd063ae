// fastx.h — FASTA/FASTQ record reader of the host: what the reference takes from the `fxread`
// crate (count.rs:24,64,87; offsetter.rs:172-173,190,195): gzip iff the path ends in ".gz"
// (all members of a multi-member file), format sniffed from the first byte ('>' = 2-line
// FASTA, '@' = 4-line FASTQ), id = header line without the marker, seq = raw bytes.
//
// Ingest is the end-to-end bottleneck, so a gzip file that consists of
// several members (bgzip, pigz -i, or any writer that restarts the stream) is inflated by a
// pool of threads, one member each, and handed to the parser in order; a single-member file
// falls back to one inflate thread.
#pragma once

#include <zlib.h>

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace sgh {

struct FastxError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Source of decompressed (or plain) bytes, in file order.
class ByteSource {
 public:
  virtual ~ByteSource() = default;
  // Appends more bytes to `out` (at least one unless the input has ended).  Returns false at
  // the end of the input.
  virtual bool read_more(std::vector<char>& out) = 0;
};

std::unique_ptr<ByteSource> open_byte_source(const std::string& path, unsigned inflate_threads);

// Buffered line source.
class LineSource {
 public:
  LineSource(const std::string& path, unsigned inflate_threads);
  // Next line without its '\n'.  Returns false at end of input.  The view is valid until the
  // next call.
  bool next(const char*& begin, size_t& len);

 private:
  std::unique_ptr<ByteSource> src_;
  std::vector<char> buf_;
  size_t pos_ = 0;
  bool eof_ = false;
};

class FastxReader {
 public:
  explicit FastxReader(const std::string& path, unsigned inflate_threads = 1);
  // Next record; views valid until the next call.  Throws FastxError on a truncated record.
  bool next(const char*& id, size_t& id_len, const char*& seq, size_t& seq_len);
  // Sequence line only (the hot loop of count_sample): a view into the read buffer.
  bool next_seq(const char*& seq, size_t& seq_len);

 private:
  void sniff(const char* line, size_t len);
  LineSource src_;
  bool pending_skip_ = false;
  int lines_per_record_ = 0;  // 2 FASTA, 4 FASTQ; 0 = not sniffed yet
  std::string id_, seq_;
};

}  // namespace sgh
