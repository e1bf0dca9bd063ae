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
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
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

// Allocator of the big ingest buffers (a member's decompressed bytes, its sequence lines):
// large blocks come from mmap with MADV_HUGEPAGE, so that sixteen inflate threads writing fresh
// buffers do not serialise on 4 KiB page faults, and elements are default-initialised (resize()
// does not zero memory that inflate is about to overwrite).
template <typename T>
struct BigAlloc {
  using value_type = T;
  BigAlloc() = default;
  template <typename U>
  BigAlloc(const BigAlloc<U>&) {}
  T* allocate(size_t n);
  void deallocate(T* p, size_t n);
  template <typename U, typename... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0)
      ::new (static_cast<void*>(p)) U;
    else
      ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
  template <typename U>
  bool operator==(const BigAlloc<U>&) const { return true; }
  template <typename U>
  bool operator!=(const BigAlloc<U>&) const { return false; }
};
void* big_alloc_bytes(size_t bytes);
void big_free_bytes(void* p, size_t bytes);
template <typename T>
T* BigAlloc<T>::allocate(size_t n) { return static_cast<T*>(big_alloc_bytes(n * sizeof(T))); }
template <typename T>
void BigAlloc<T>::deallocate(T* p, size_t n) { big_free_bytes(p, n * sizeof(T)); }
using Bytes = std::vector<char, BigAlloc<char>>;

// Span framing (include/sgcount_cuda.h sgc_span_geometry): when every read of a sample has the
// length of its first one, the inflate threads keep only bytes [start, start + len) of every
// sequence, in records of `stride` bytes — all the count kernels ever look at, a third of the
// bytes to copy into pinned memory and over PCIe.
struct SpanSpec {
  uint32_t read_len = 0;  // the sample's read length (that of its first record)
  uint32_t start = 0, len = 0, stride = 0;
};

// Sequence lines of a run of whole records, packed the way the count kernels take them: every
// sequence followed by '\n' — or, in span mode, fixed-stride span records.
struct SeqBlock {
  Bytes lines;
  std::vector<uint32_t> len;  // length of every sequence (without the '\n'); empty in span mode
  uint64_t n = 0;
  bool uniform = true;  // every sequence is as long as the first
  uint32_t first_len = 0;
  bool spans = false;   // records are spans: `first_len` = span length, `stride` = record size
  uint32_t stride = 0;  // bytes per record when uniform (first_len + 1 for whole lines)
  void clear() {
    lines.clear();
    len.clear();
    n = 0;
    uniform = true;
    first_len = 0;
    spans = false;
    stride = 0;
  }
  void push(const char* seq, size_t l);
};

// Incremental record framing: bytes in (cut anywhere), sequence lines out.  Same framing as
// FastxReader: '>' = 2-line FASTA, '@' = 4-line FASTQ, sniffed from the first byte.
class SeqParser {
 public:
  struct State {
    int lines_per_record = 0;  // 0 = not sniffed yet
    int phase = 0;             // line of the record the next line is
    std::string carry;         // bytes of a line whose '\n' has not arrived
    bool clean() const { return phase == 0 && carry.empty(); }
  };
  State st;
  // span framing of uniform-length reads; feed() returns false at the first sequence of another
  // length (the block is then incomplete: frame the bytes again without a spec)
  const SpanSpec* spans = nullptr;
  bool feed(const char* data, size_t len, SeqBlock& out);
  // end of input: a last line without '\n' counts; throws FastxError on a truncated record
  void finish(SeqBlock& out);

 private:
  bool line(const char* p, size_t len, SeqBlock& out);
  const char* fastq_records(const char* p, const char* end, SeqBlock& out, bool& ok);
};

// The hot ingest path of count_sample (count.rs:15-45 hands `Counter::new` a record iterator;
// here the kernels want packed sequence lines): blocks of sequence lines in file order.  For a
// multi-member gzip file every inflate thread also frames the records of its member (of its run
// of 64 blocks, for BGZF) from the first record start it can recognise in the text
// (fastx_first_record_start); the consumer runs the few bytes before that start through the real
// framing state, takes the thread's block as it is when that state comes out clean, and re-frames
// the member's bytes itself when it does not.
class SeqBlockReader {
 public:
  // spans: frame span records instead of whole lines wherever a gzip member's reads all have
  // spans->read_len bytes (blocks say which they hold); nullptr = whole lines only
  explicit SeqBlockReader(const std::string& path, unsigned inflate_threads = 1, const SpanSpec* spans = nullptr);
  ~SeqBlockReader();
  // Next block (possibly of zero records); false at the end of the input.
  bool next(SeqBlock& out);
  // gzip members (runs of BGZF blocks) whose framing by an inflate thread was taken as it was /
  // whose bytes this thread had to frame again
  uint64_t members_adopted() const;
  uint64_t members_reframed() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

// Offset of the first record start that can be recognised in text[0, n), a piece of a FASTA
// (lines_per_record 2) or FASTQ (4) file that begins anywhere in a record, or SIZE_MAX.  FASTQ: the
// first line (offset 0 counts as a line start) that begins with '@', whose third line begins with
// '+' and whose second and fourth lines have the same length (see fastq_last_record_end for why a
// quality line cannot pass); FASTA: the first line that begins with '>'.  A guess that is wrong
// (the text began inside a header that contains the marker) is caught by the caller's own state.
size_t fastx_first_record_start(const char* text, size_t n, int lines_per_record);

// Start offsets of the blocks of a BGZF file (gzip members with a 'BC' extra field that holds the
// member's size) plus the file size, and every block's ISIZE; false if the bytes are anything else.
// A file of `min_parallel_bytes` or more is walked by up to `threads` threads side by side (same
// result; `in_parts` says whether it was).
bool bgzf_index(const uint8_t* d, size_t n, std::vector<uint64_t>& begin, std::vector<uint32_t>& isize, unsigned threads = 1,
                size_t min_parallel_bytes = 64u << 20, bool* in_parts = nullptr);

// Offset just past the last complete FASTQ record that can be recognised in text[0, n) without
// knowing where the text starts in its file (a BGZF block begins anywhere in a record), or
// SIZE_MAX.  A record start is a line that begins with '@', whose third line begins with '+' and
// whose second and fourth lines have the same length; a quality line that begins with '@' cannot
// pass, because two lines further comes a sequence line, which never begins with '+'.
size_t fastq_last_record_end(const char* text, size_t n);

class FastxReader {
 public:
  explicit FastxReader(const std::string& path, unsigned inflate_threads = 1);
  // Next record; views valid until the next call.  Throws FastxError on a truncated record.
  bool next(const char*& id, size_t& id_len, const char*& seq, size_t& seq_len);
  // Sequence line only (the hot loop of count_sample): a view into the read buffer.
  bool next_seq(const char*& seq, size_t& seq_len);
  // '@' records of four lines (known once a record has been read)
  bool is_fastq() const { return lines_per_record_ == 4; }

 private:
  void sniff(const char* line, size_t len);
  LineSource src_;
  bool pending_skip_ = false;
  int lines_per_record_ = 0;  // 2 FASTA, 4 FASTQ; 0 = not sniffed yet
  std::string id_, seq_;
};

}  // namespace sgh
