#include "fastx.h"

#include "inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

namespace sgh {

// ---- big buffers ----------------------------------------------------------------------------
namespace {
constexpr size_t kBigThreshold = 4u << 20, kHugePage = 2u << 20;
}
void* big_alloc_bytes(size_t bytes) {
  if (bytes < kBigThreshold) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p) throw std::bad_alloc();
    return p;
  }
  const size_t len = (bytes + kHugePage - 1) / kHugePage * kHugePage;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
  madvise(p, len, MADV_HUGEPAGE);
#endif
  return p;
}
void big_free_bytes(void* p, size_t bytes) {
  if (bytes < kBigThreshold)
    free(p);
  else
    munmap(p, (bytes + kHugePage - 1) / kHugePage * kHugePage);
}

namespace {

bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

// ---- plain file ---------------------------------------------------------------------------
class PlainSource : public ByteSource {
 public:
  explicit PlainSource(const std::string& path) : fp_(fopen(path.c_str(), "rb")) {
    if (!fp_) throw FastxError("cannot open " + path);
  }
  ~PlainSource() override { fclose(fp_); }
  bool read_more(std::vector<char>& out) override {
    const size_t chunk = 8u << 20, at = out.size();
    out.resize(at + chunk);
    const size_t got = fread(out.data() + at, 1, chunk, fp_);
    out.resize(at + got);
    return got > 0;
  }

 private:
  FILE* fp_;
};

// ---- gzip, possibly multi-member, inflated member-parallel ------------------------------------
// The compressed file is mapped; every position that looks like the start of a gzip member
// (1f 8b 08, reserved flag bits clear) is a CANDIDATE.  Workers inflate candidates
// speculatively, each until its stream ends, and record where it ended.  The consumer walks
// the chain from byte 0: the member at `pos` is delivered, `pos` moves to its end, which must
// again be a candidate (or the end of the file).  A candidate that is not on the chain was a
// coincidence inside compressed data and is dropped; if the chain breaks, the rest of the file
// is inflated sequentially from `pos`.  Either way the output is exactly what a sequential
// multi-member decoder (flate2's MultiGzDecoder in the reference) produces.
class GzSource : public ByteSource {
 public:
  // frame_records: the workers also frame the records of their member (SeqBlockReader)
  GzSource(const std::string& path, unsigned threads, bool frame_records = false, const SpanSpec* spans = nullptr)
      : spans_(spans) {
    fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw FastxError("cannot open " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0) throw FastxError("cannot stat " + path);
    size_ = (size_t)st.st_size;
    if (size_ > 0) {
      void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (p == MAP_FAILED) throw FastxError("cannot map " + path);
      data_ = static_cast<const unsigned char*>(p);
      madvise(p, size_, MADV_SEQUENTIAL);
    }
    threads_ = std::max(1u, threads);
    if (threads_ > 1 && size_ > (1u << 16) && !group_bgzf_blocks()) find_candidates();
    if (cand_.size() <= 1) threads_ = 1;  // a single member: nothing to run in parallel
    if (threads_ > 1 && frame_records) frame_lpr_ = peek_lines_per_record();
    if (threads_ > 1) {
      jobs_.resize(cand_.size());
      // members in flight: one per thread plus a few finished ones waiting for the consumer, which
      // only copies and is never the slow side (each holds ~1.5x the member's decompressed size)
      window_ = threads_ + 2;
      for (unsigned t = 0; t < threads_; ++t) pool_.emplace_back([this] { worker(); });
    }
  }
  ~GzSource() override {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : pool_) t.join();
    if (seq_init_) inflateEnd(&seq_);
    if (data_) munmap(const_cast<unsigned char*>(data_), size_);
    if (fd_ >= 0) close(fd_);
  }

  bool read_more(std::vector<char>& out) override {
    Bytes data;
    bool framed;
    size_t head;
    SeqBlock block;
    SeqParser::State state;
    if (!next_member(data, framed, head, block, state)) return false;
    if (!data.empty()) {
      const size_t at = out.size();
      out.resize(at + data.size());
      memcpy(out.data() + at, data.data(), data.size());
    }
    recycle(std::move(data), SeqBlock());
    if (out.empty()) return read_more(out);  // an empty member
    return true;
  }

  // Hands buffers back to the inflate threads (their pages stay mapped and faulted).
  void recycle(Bytes&& raw, SeqBlock&& block) {
    std::lock_guard<std::mutex> lk(mu_);
    if (raw.capacity() && free_raw_.size() < window_ + 2) {
      raw.clear();
      free_raw_.push_back(std::move(raw));
    }
    if (block.lines.capacity() && free_blocks_.size() < window_ + 2) {
      block.clear();
      free_blocks_.push_back(std::move(block));
    }
  }

  // The next member (member-parallel mode) or the next chunk (sequential mode) of decompressed
  // bytes, possibly none.  `framed`: the worker framed the records of raw[head, size) from a clean
  // state into `block`, ending in `end_state`; `head` is where it saw the first record start.
  // Returns false at the end of the input.
  bool next_member(Bytes& raw, bool& framed, size_t& head, SeqBlock& block, SeqParser::State& end_state) {
    raw.clear();  // keeps its pages: the sequential path refills it, the parallel path recycles it
    framed = false;
    head = 0;
    if (threads_ > 1 && !sequential_) {
      if (pos_ >= size_) return false;
      // the candidate that starts exactly at pos_
      auto it = std::lower_bound(cand_.begin(), cand_.end(), pos_);
      if (it != cand_.end() && *it == pos_) {
        const size_t j = (size_t)(it - cand_.begin());
        std::unique_lock<std::mutex> lk(mu_);
        // candidates the chain stepped over were coincidences: whatever they inflated to is dropped
        for (size_t skipped = consumer_at_; skipped < j; ++skipped) {
          Bytes().swap(jobs_[skipped].out);
          jobs_[skipped].block = SeqBlock();
        }
        consumer_at_ = j;
        cv_work_.notify_all();
        cv_done_.wait(lk, [&] { return jobs_[j].done; });
        Job& job = jobs_[j];
        if (job.ok) {
          if (raw.capacity() && free_raw_.size() < window_ + 2) free_raw_.push_back(std::move(raw));
          raw = std::move(job.out);
          framed = job.framed;
          if (framed) {
            head = job.head;
            block = std::move(job.block);
            end_state = std::move(job.end_state);
          }
          pos_ = job.end;
          return true;
        }
      }
      // the chain is broken (not a member boundary, or a corrupt member): go on sequentially
      sequential_ = true;
      {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
      }
      cv_work_.notify_all();
    }
    return read_sequential(raw);
  }

 private:
  struct Job {
    Bytes out;
    size_t end = 0;
    bool ok = false, done = false, taken = false;
    // record framing of the member from its first recognisable record start, out[head]
    bool framed = false;
    size_t head = 0;
    SeqBlock block;
    SeqParser::State end_state;
  };

  // '>' -> 2, '@' -> 4, anything else (or an empty first member) -> 0: no framing by the workers,
  // the consumer's own parser reports what is wrong
  int peek_lines_per_record() {
    z_stream zs{};
    if (inflateInit2(&zs, 15 + 16) != Z_OK) return 0;
    unsigned char first = 0;
    zs.next_in = const_cast<unsigned char*>(data_);
    zs.avail_in = (uInt)std::min<size_t>(size_, 1u << 16);
    zs.next_out = &first;
    zs.avail_out = 1;
    inflate(&zs, Z_NO_FLUSH);
    const bool got = zs.avail_out == 0;
    inflateEnd(&zs);
    return !got ? 0 : (first == '>' ? 2 : (first == '@' ? 4 : 0));
  }


  // positions in [from, to) that look like the start of a member
  void scan_candidates(size_t from, size_t to, std::vector<size_t>& found) const {
    const unsigned char* p = data_ + from;
    const unsigned char* const last = data_ + to;
    while (p < last) {
      p = static_cast<const unsigned char*>(memchr(p, 0x1f, (size_t)(last - p)));
      if (!p) break;
      if (p[1] == 0x8b && p[2] == 0x08 && (p[3] & 0xE0) == 0) found.push_back((size_t)(p - data_));
      ++p;
    }
  }
  // BGZF: the blocks are 64 KB of text each and say how long they are, so the file is cut into
  // runs of blocks without looking at a byte of compressed data; one job inflates one run (a few
  // MB of text, like a member of the files synth writes).  False if the file is anything else.
  bool group_bgzf_blocks() {
    std::vector<uint64_t> begin;
    std::vector<uint32_t> isize;
    if (!bgzf_index(data_, size_, begin, isize, threads_)) return false;
    const size_t n_blocks = isize.size();
    const size_t run = std::min<size_t>(64, std::max<size_t>(1, n_blocks / (4 * (size_t)threads_)));
    for (size_t b = 0; b < n_blocks; b += run) cand_.push_back((size_t)begin[b]);
    run_text_ = run * 65536 + 1024;
    grouped_ = true;
    return true;
  }
  void find_candidates() {
    const size_t last = size_ - 18;  // smallest member: 10 header + 8 trailer
    // the scan touches every page of the file once: worth spreading over the threads that are
    // about to inflate it
    const unsigned parts = size_ < (64u << 20) ? 1u : std::min(threads_, 16u);
    std::vector<std::vector<size_t>> found(parts);
    std::vector<std::thread> pool;
    const size_t step = last / parts + 1;
    for (unsigned i = 1; i < parts; ++i)
      pool.emplace_back([&, i] { scan_candidates(step * i, std::min(last, step * (i + 1)), found[i]); });
    scan_candidates(0, std::min(last, step), found[0]);
    for (auto& t : pool) t.join();
    for (auto& f : found) cand_.insert(cand_.end(), f.begin(), f.end());
    if (cand_.empty() || cand_[0] != 0) cand_.clear();  // not a gzip file: let zlib report it
  }

  // inflate ONE member starting at `from`; false if the bytes there are not a complete member
  bool inflate_member(size_t from, Bytes& out, size_t& end) {
    // the whole-member decoder first (inflate.h); zlib decides whenever it declines
    if (use_fast_inflate_) {
      size_t consumed = 0;
      if (gunzip_member(data_ + from, size_ - from, out, consumed)) {
        end = from + consumed;
        return true;
      }
    }
    z_stream zs{};
    if (inflateInit2(&zs, 15 + 16) != Z_OK) return false;
    zs.next_in = const_cast<unsigned char*>(data_ + from);
    size_t in_left = size_ - from;
    size_t produced = 0;
    out.resize(std::max<size_t>(out.capacity(), 1u << 20));
    int rc = Z_OK;
    for (;;) {
      const uInt give = (uInt)std::min<size_t>(in_left, 1u << 30);
      zs.avail_in = give;
      if (produced == out.size()) out.resize(out.size() * 2);
      zs.next_out = reinterpret_cast<Bytef*>(out.data() + produced);
      const uInt room = (uInt)std::min<size_t>(out.size() - produced, 1u << 30);
      zs.avail_out = room;
      rc = inflate(&zs, Z_NO_FLUSH);
      produced += room - zs.avail_out;
      in_left -= give - zs.avail_in;
      if (rc == Z_STREAM_END) break;
      if (rc != Z_OK && rc != Z_BUF_ERROR) break;
      if (rc == Z_BUF_ERROR && zs.avail_in == 0 && in_left == 0) break;  // truncated
    }
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    out.resize(produced);
    end = size_ - in_left;
    return true;
  }

  void worker() {
    for (;;) {
      size_t j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return stop_ || (next_job_ < jobs_.size() && next_job_ < consumer_at_ + window_); });
        if (stop_) return;
        j = next_job_++;
      }
      Bytes out;
      SeqBlock block;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!free_raw_.empty()) {
          out = std::move(free_raw_.back());
          free_raw_.pop_back();
        }
        if (frame_lpr_ && !free_blocks_.empty()) {
          block = std::move(free_blocks_.back());
          free_blocks_.pop_back();
        }
      }
      // size hint: ISIZE of a member that ends where the next candidate starts.  If this or the
      // next candidate is a coincidence inside compressed data those four bytes are noise: the hint
      // is capped at 64x the compressed size (DEFLATE of sequence data stays far below that)
      const size_t nxt = j + 1 < cand_.size() ? cand_[j + 1] : size_;
      if (grouped_) {
        out.reserve(run_text_);
      } else if (nxt >= cand_[j] + 18) {
        uint32_t isize;
        memcpy(&isize, data_ + nxt - 4, 4);
        const size_t cap = 64 * (nxt - cand_[j]);
        if (isize < (1u << 30)) out.reserve(std::min<size_t>(isize, cap) + 1024);
      }
      size_t end = 0, head = 0;
      bool ok = false, framed = false;
      SeqParser framer;
      try {  // nothing may escape a worker: a failed job sends the consumer down the sequential path
        if (grouped_) {  // every block of the run, one after the other
          size_t at = cand_[j];
          Bytes one;
          ok = true;
          while (ok && at < nxt) {
            ok = inflate_member(at, one, end);
            if (ok) {
              out.insert(out.end(), one.begin(), one.end());
              at = end;
            }
          }
          ok = ok && at == nxt;
          end = at;
        } else {
          ok = inflate_member(cand_[j], out, end);
        }
        // the member starts anywhere in a record (BGZF cuts every 64 KB of text): frame from the
        // first record start on, the consumer finishes the record before it
        if (ok && frame_lpr_) head = fastx_first_record_start(out.data(), out.size(), frame_lpr_);
        if (ok && frame_lpr_ && head != SIZE_MAX) {
          framer.st.lines_per_record = frame_lpr_;
          block.lines.reserve(out.size() / 2);
          const char* const text = out.data() + head;
          const size_t text_len = out.size() - head;
          try {
            // span records while every read of the sample has had the first one's length
            framer.spans = spans_ && !spans_abandoned_.load(std::memory_order_relaxed) ? spans_ : nullptr;
            if (!framer.feed(text, text_len, block)) {
              spans_abandoned_.store(true, std::memory_order_relaxed);
              framer = SeqParser();
              framer.st.lines_per_record = frame_lpr_;
              block.clear();
              framer.feed(text, text_len, block);
            }
            framed = true;
          } catch (const std::exception&) {  // the consumer frames these bytes itself and reports
          }
        }
      } catch (const std::exception&) {
        ok = false;
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        // a failed candidate, or one the consumer has already stepped over, keeps no buffer
        if (j < consumer_at_) ok = framed = false;
        if (ok) jobs_[j].out = std::move(out);
        if (framed) {
          jobs_[j].framed = true;
          jobs_[j].head = head;
          jobs_[j].block = std::move(block);
          jobs_[j].end_state = std::move(framer.st);
        }
        jobs_[j].end = end;
        jobs_[j].ok = ok;
        jobs_[j].done = true;
      }
      cv_done_.notify_all();
    }
  }

  bool read_sequential(Bytes& out) {
    if (!seq_init_) {
      if (inflateInit2(&seq_, 15 + 16) != Z_OK) throw FastxError("inflateInit2 failed");
      seq_init_ = true;
    }
    const size_t chunk = 8u << 20;
    size_t got = 0;
    while (got == 0) {
      if (pos_ >= size_) return false;
      const size_t at = out.size();
      out.resize(at + chunk);
      const uInt give = (uInt)std::min<size_t>(size_ - pos_, 1u << 30);
      seq_.next_in = const_cast<unsigned char*>(data_ + pos_);
      seq_.avail_in = give;
      seq_.next_out = reinterpret_cast<Bytef*>(out.data() + at);
      seq_.avail_out = (uInt)chunk;
      const int rc = inflate(&seq_, Z_NO_FLUSH);
      got = chunk - seq_.avail_out;
      pos_ += give - seq_.avail_in;
      out.resize(at + got);
      if (rc == Z_STREAM_END) {
        if (pos_ < size_ && inflateReset(&seq_) != Z_OK) throw FastxError("inflateReset failed");
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        throw FastxError(std::string("gzip stream is corrupt: ") + (seq_.msg ? seq_.msg : "inflate error"));
      } else if (rc == Z_BUF_ERROR && got == 0 && seq_.avail_in == 0 && pos_ >= size_) {
        throw FastxError("gzip stream is truncated");
      }
    }
    return true;
  }

  int fd_ = -1;
  const unsigned char* data_ = nullptr;
  size_t size_ = 0, pos_ = 0;
  unsigned threads_ = 1;
  std::vector<size_t> cand_;
  std::vector<Job> jobs_;
  std::vector<std::thread> pool_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  size_t next_job_ = 0, consumer_at_ = 0, window_ = 2;
  bool stop_ = false, sequential_ = false;
  bool grouped_ = false;  // BGZF: a job is the run of blocks between two candidates
  size_t run_text_ = 0;   // most text a run inflates to
  int frame_lpr_ = 0;  // lines per record the workers frame with; 0 = they do not
  const SpanSpec* spans_ = nullptr;
  std::atomic<bool> spans_abandoned_{false};  // a read of another length was seen: whole lines from here on
  const bool use_fast_inflate_ = []() {
    const char* v = getenv("SGC_INFLATE");  // "zlib": every member through zlib (A/B and tests)
    return !(v && !strcmp(v, "zlib"));
  }();
  std::vector<Bytes> free_raw_;        // recycled buffers (under mu_)
  std::vector<SeqBlock> free_blocks_;
  z_stream seq_{};
  bool seq_init_ = false;
};

}  // namespace

std::unique_ptr<ByteSource> open_byte_source(const std::string& path, unsigned inflate_threads) {
  if (ends_with(path, ".gz")) return std::unique_ptr<ByteSource>(new GzSource(path, inflate_threads));
  return std::unique_ptr<ByteSource>(new PlainSource(path));
}

LineSource::LineSource(const std::string& path, unsigned inflate_threads)
    : src_(open_byte_source(path, inflate_threads)) {}

bool LineSource::next(const char*& begin, size_t& len) {
  for (;;) {
    const char* p = buf_.data() + pos_;
    const char* nl = buf_.size() > pos_ ? static_cast<const char*>(memchr(p, '\n', buf_.size() - pos_)) : nullptr;
    if (nl) {
      begin = p;
      len = (size_t)(nl - p);
      pos_ += len + 1;
      return true;
    }
    if (!eof_) {
      // keep the unread tail, append more
      if (pos_ > 0) {
        buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)pos_);
        pos_ = 0;
      }
      if (!src_->read_more(buf_)) eof_ = true;
      continue;
    }
    if (pos_ >= buf_.size()) return false;
    begin = buf_.data() + pos_;  // last line without a newline
    len = buf_.size() - pos_;
    pos_ = buf_.size();
    return true;
  }
}

// ---- packed sequence lines ---------------------------------------------------------------------
void SeqBlock::push(const char* seq, size_t l) {
  if (l >= 0xFFFFFFFFull) throw FastxError("a sequence line of 4 GiB or more");
  if (n == 0)
    first_len = (uint32_t)l;
  else
    uniform &= l == first_len;
  lines.insert(lines.end(), seq, seq + l);
  lines.push_back('\n');
  len.push_back((uint32_t)l);
  if (n == 0) stride = (uint32_t)l + 1;
  ++n;
}

bool SeqParser::line(const char* p, size_t len, SeqBlock& out) {
  if (st.lines_per_record == 0) {
    if (len == 0) throw FastxError("empty first line: not FASTA/FASTQ");
    if (p[0] == '>')
      st.lines_per_record = 2;
    else if (p[0] == '@')
      st.lines_per_record = 4;
    else
      throw FastxError("first byte is neither '>' nor '@'");
  }
  if (st.phase == 1) {
    if (spans) {
      if (len != spans->read_len) return false;
      if (out.n == 0) {
        out.spans = true;
        out.first_len = spans->len;
        out.stride = spans->stride;
      }
      const size_t at = out.lines.size();
      out.lines.resize(at + spans->stride);  // BigAlloc default-initialises: the pad bytes are never looked at
      memcpy(out.lines.data() + at, p + spans->start, spans->len);
      ++out.n;
    } else {
      out.push(p, len);
    }
  }
  if (++st.phase == st.lines_per_record) st.phase = 0;
  return true;
}

// Whole FASTQ records at a record start: as many as lie complete in [p, end), framed without the
// per-line state machine and written through a pointer into space reserved once (the same four
// newline searches per record, the same bytes out as line() produces; the record that is cut by
// `end`, FASTA, and the first line of a file are left to the line-by-line code).  Returns where it
// stopped; `ok` = false at a sequence of another length in span mode (see feed()).
const char* SeqParser::fastq_records(const char* p, const char* end, SeqBlock& out, bool& ok) {
  ok = true;
  const size_t room = (size_t)(end - p);
  const size_t at = out.lines.size();
  // an upper bound of what the records in [p, end) can add
  const size_t most = spans ? (room / ((size_t)spans->read_len + 4) + 1) * spans->stride : room + 1;
  out.lines.resize(at + most);  // (BigAlloc default-initialises: no byte is touched here)
  char* w = out.lines.data() + at;
  auto newline = [end](const char* from) { return static_cast<const char*>(memchr(from, '\n', (size_t)(end - from))); };
  while (p < end) {
    const char* h = newline(p);
    if (!h) break;
    const char* seq = h + 1;
    if (seq >= end) break;
    const char* s_end = newline(seq);
    if (!s_end || s_end + 1 >= end) break;
    // (the usual third line is "+" alone: its newline is found without a search)
    const char* plus_end = s_end + 2 < end && s_end[1] == '+' && s_end[2] == '\n' ? s_end + 2 : newline(s_end + 1);
    if (!plus_end || plus_end + 1 >= end) break;
    const char* q_end = newline(plus_end + 1);
    if (!q_end) break;
    const size_t l = (size_t)(s_end - seq);
    if (spans) {
      if (l != spans->read_len) {
        ok = false;
        break;
      }
      if (out.n == 0) {
        out.spans = true;
        out.first_len = spans->len;
        out.stride = spans->stride;
      }
      memcpy(w, seq + spans->start, spans->len);
      w += spans->stride;
    } else {
      if (l >= 0xFFFFFFFFull) throw FastxError("a sequence line of 4 GiB or more");
      if (out.n == 0) {
        out.first_len = (uint32_t)l;
        out.stride = (uint32_t)l + 1;
      } else {
        out.uniform &= l == out.first_len;
      }
      memcpy(w, seq, l + 1);  // the sequence and its newline
      w += l + 1;
      out.len.push_back((uint32_t)l);
    }
    ++out.n;
    p = q_end + 1;
  }
  out.lines.resize((size_t)(w - out.lines.data()));
  return p;
}

bool SeqParser::feed(const char* data, size_t len, SeqBlock& out) {
  const char* p = data;
  const char* const end = data + len;
  if (!st.carry.empty()) {  // finish the line the previous chunk left open
    const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
    if (!nl) {
      st.carry.append(p, end);
      return true;
    }
    st.carry.append(p, nl);
    if (!line(st.carry.data(), st.carry.size(), out)) return false;
    st.carry.clear();
    p = nl + 1;
  }
  while (p < end) {
    if (st.lines_per_record == 4 && st.phase == 0) {  // at a FASTQ record start: whole records at once
      bool ok;
      p = fastq_records(p, end, out, ok);
      if (!ok) return false;
      if (p >= end) break;
    }
    const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
    if (!nl) {
      st.carry.assign(p, end);
      return true;
    }
    if (!line(p, (size_t)(nl - p), out)) return false;
    p = nl + 1;
  }
  return true;
}

void SeqParser::finish(SeqBlock& out) {
  if (!st.carry.empty()) {  // last line without a newline
    line(st.carry.data(), st.carry.size(), out);
    st.carry.clear();
  }
  if (st.phase == 1) throw FastxError("truncated record: header without a sequence line");
  if (st.phase != 0) throw FastxError("truncated FASTQ record");
}

struct SeqBlockReader::Impl {
  std::unique_ptr<GzSource> gz;
  std::unique_ptr<ByteSource> plain;
  SeqParser parser;
  Bytes raw;                  // gzip: the member / chunk being framed
  std::vector<char> raw_plain;
  bool eof = false;
  SpanSpec spans;  // what the inflate threads frame with (they hold a pointer to it)
  bool have_spans = false;
  SeqBlock pending;  // an inflate thread's block, due after the record that was finished before it
  bool have_pending = false;
  uint64_t adopted = 0, reframed = 0;
};

SeqBlockReader::SeqBlockReader(const std::string& path, unsigned inflate_threads, const SpanSpec* spans)
    : impl_(new Impl()) {
  if (spans) impl_->spans = *spans;
  impl_->have_spans = spans != nullptr;
  if (ends_with(path, ".gz"))
    impl_->gz.reset(new GzSource(path, inflate_threads, /*frame_records=*/true, spans ? &impl_->spans : nullptr));
  else
    impl_->plain.reset(new PlainSource(path));
}
SeqBlockReader::~SeqBlockReader() = default;

bool SeqBlockReader::next(SeqBlock& out) {
  Impl& m = *impl_;
  out.clear();
  if (m.eof) return false;
  if (m.have_pending) {
    std::swap(out, m.pending);  // the caller's previous block goes back to the inflate threads
    m.have_pending = false;
    m.gz->recycle(Bytes(), std::move(m.pending));
    m.pending = SeqBlock();
    return true;
  }
  bool more;
  if (m.gz) {
    bool framed = false;
    size_t head = 0;
    SeqBlock block;
    SeqParser::State end_state;
    more = m.gz->next_member(m.raw, framed, head, block, end_state);
    if (more) {
      // The worker framed raw[head, size) from a clean state: valid iff that is the state the bytes
      // before `head` (the end of the record the previous member left open) bring this parser to.
      bool adopt = false;
      if (framed) {
        SeqParser probe;
        probe.st = m.parser.st;
        if (head > 0) {
          probe.spans = block.spans && m.have_spans ? &m.spans : nullptr;  // the same kind of block
          if (!probe.feed(m.raw.data(), head, out)) {
            out.clear();
            probe.st = m.parser.st;
            probe.spans = nullptr;
            probe.feed(m.raw.data(), head, out);
          }
        }
        const SeqParser::State& st = probe.st;
        adopt = st.clean() && (st.lines_per_record == 0 || st.lines_per_record == end_state.lines_per_record);
      }
      if (adopt) {
        if (out.n == 0) {
          std::swap(out, block);
        } else {  // this call returns the record that was finished, the next one the worker's block
          m.pending = std::move(block);
          block = SeqBlock();
          m.have_pending = true;
        }
        m.parser.st = std::move(end_state);
        ++m.adopted;
      } else {
        out.clear();
        m.parser.feed(m.raw.data(), m.raw.size(), out);
        ++m.reframed;
      }
      m.gz->recycle(Bytes(), std::move(block));
    }
  } else {
    m.raw_plain.clear();
    more = m.plain->read_more(m.raw_plain);
    if (more) m.parser.feed(m.raw_plain.data(), m.raw_plain.size(), out);
  }
  if (!more) {
    m.parser.finish(out);
    m.eof = true;
  }
  return true;
}

uint64_t SeqBlockReader::members_adopted() const { return impl_->adopted; }
uint64_t SeqBlockReader::members_reframed() const { return impl_->reframed; }

size_t fastx_first_record_start(const char* text, size_t n, int lines_per_record) {
  if (lines_per_record == 2) {
    if (n && text[0] == '>') return 0;
    for (const char* p = text; (p = static_cast<const char*>(memchr(p, '\n', (size_t)(text + n - p)))) != nullptr; ++p)
      if (p + 1 < text + n && p[1] == '>') return (size_t)(p + 1 - text);
    return SIZE_MAX;
  }
  if (lines_per_record != 4) return SIZE_MAX;
  // starts of the first lines of the text: a record start is among the first four, a few more
  // for quality lines that happen to look like headers
  size_t start[13];
  int lines = 0;
  start[lines++] = 0;
  for (const char* p = text; lines < 13 && (p = static_cast<const char*>(memchr(p, '\n', (size_t)(text + n - p)))) != nullptr; ++p)
    start[lines++] = (size_t)(p + 1 - text);
  for (int i = 0; i + 4 < lines; ++i) {
    if (start[i + 2] >= n) break;
    if (text[start[i]] == '@' && text[start[i + 2]] == '+' && start[i + 2] - start[i + 1] == start[i + 4] - start[i + 3])
      return start[i];
  }
  return SIZE_MAX;
}

namespace {

// Size of the BGZF block whose header starts at d[pos], or 0 if the bytes there are not one.
inline size_t bgzf_block_size(const uint8_t* d, size_t n, size_t pos) {
  if (n - pos < 28 || d[pos] != 0x1f || d[pos + 1] != 0x8b || d[pos + 2] != 8 || !(d[pos + 3] & 4)) return 0;
  const size_t xlen = d[pos + 10] | ((size_t)d[pos + 11] << 8);
  if (pos + 12 + xlen > n) return 0;
  size_t bsize = 0;
  for (size_t at = pos + 12; at + 4 <= pos + 12 + xlen;) {
    const size_t slen = d[at + 2] | ((size_t)d[at + 3] << 8);
    if (d[at] == 'B' && d[at + 1] == 'C' && slen == 2 && at + 6 <= pos + 12 + xlen) bsize = (d[at + 4] | ((size_t)d[at + 5] << 8)) + 1;
    at += 4 + slen;
  }
  if (bsize < 28 || pos + bsize > n) return 0;
  return bsize;
}

// Walks the blocks from `pos` until a block starts at or past `stop`; returns where it stopped, or
// SIZE_MAX at bytes that are not a block.
size_t bgzf_walk(const uint8_t* d, size_t n, size_t pos, size_t stop, std::vector<uint64_t>& begin, std::vector<uint32_t>& isize) {
  while (pos < stop) {
    const size_t bsize = bgzf_block_size(d, n, pos);
    if (!bsize) return SIZE_MAX;
    begin.push_back(pos);
    uint32_t sz;
    memcpy(&sz, d + pos + bsize - 4, 4);
    isize.push_back(sz);
    pos += bsize;
  }
  return pos;
}

}  // namespace

// The walk reads 18 bytes of every block, one block per page or two: it is the page faults of a
// first pass over the mapping that cost (30 ms per GB).  A large file is therefore walked in parts
// side by side: every part but the first begins at a GUESS — the first offset past its cut where
// three block headers follow each other — and the guess is confirmed when the part before it
// arrives at exactly that offset.  A guess that is not confirmed (compressed bytes that look like
// three chained headers) sends the whole file through the sequential walk.
bool bgzf_index(const uint8_t* d, size_t n, std::vector<uint64_t>& begin, std::vector<uint32_t>& isize, unsigned threads,
                size_t min_parallel_bytes, bool* in_parts) {
  begin.clear();
  isize.clear();
  if (in_parts) *in_parts = false;
  const unsigned parts = n < min_parallel_bytes ? 1u : std::min(std::max(threads, 1u), 16u);
  if (parts > 1) {
    struct Part {
      size_t start = SIZE_MAX, end = SIZE_MAX;
      std::vector<uint64_t> begin;
      std::vector<uint32_t> isize;
    };
    std::vector<Part> part(parts);
    part[0].start = 0;
    for (unsigned i = 1; i < parts; ++i) {
      const size_t cut = n / parts * i, limit = std::min(n, cut + (1u << 20));
      for (size_t pos = cut; pos + 28 <= limit; ++pos) {
        const void* hit = memchr(d + pos, 0x1f, limit - pos);
        if (!hit) break;
        pos = (size_t)(static_cast<const uint8_t*>(hit) - d);
        if (pos + 28 > limit) break;
        size_t at = pos;
        int chained = 0;
        for (; chained < 3 && at < n; ++chained) {
          const size_t b = bgzf_block_size(d, n, at);
          if (!b) break;
          at += b;
        }
        if (chained == 3 || (chained > 0 && at == n)) {
          part[i].start = pos;
          break;
        }
      }
      if (part[i].start == SIZE_MAX) break;  // no block start near a cut: not worth guessing further
    }
    bool guessed = true;
    for (unsigned i = 1; i < parts; ++i) guessed &= part[i].start != SIZE_MAX && part[i].start > part[i - 1].start;
    if (guessed) {
      std::vector<std::thread> pool;
      auto walk = [&](unsigned i) {
        part[i].end = bgzf_walk(d, n, part[i].start, i + 1 < parts ? part[i + 1].start : n, part[i].begin, part[i].isize);
      };
      for (unsigned i = 1; i < parts; ++i) pool.emplace_back(walk, i);
      walk(0);
      for (auto& t : pool) t.join();
      bool confirmed = true;
      for (unsigned i = 0; i < parts; ++i) confirmed &= part[i].end == (i + 1 < parts ? part[i + 1].start : n);
      if (confirmed) {
        size_t total = 0;
        for (auto& p : part) total += p.isize.size();
        begin.reserve(total + 1);
        isize.reserve(total);
        for (auto& p : part) {
          begin.insert(begin.end(), p.begin.begin(), p.begin.end());
          isize.insert(isize.end(), p.isize.begin(), p.isize.end());
        }
        begin.push_back(n);
        if (in_parts) *in_parts = true;
        return !isize.empty();
      }
      begin.clear();
      isize.clear();
    }
  }
  if (bgzf_walk(d, n, 0, n, begin, isize) != n) return false;
  begin.push_back(n);
  return !isize.empty();
}

size_t fastq_last_record_end(const char* text, size_t n) {
  std::vector<size_t> nl;  // newline positions
  for (const char* p = text; (p = static_cast<const char*>(memchr(p, '\n', (size_t)(text + n - p)))) != nullptr; ++p)
    nl.push_back((size_t)(p - text));
  // line j (j >= 1) is text(nl[j-1], nl[j]); the bytes before nl[0] are a line of unknown start
  for (size_t i = nl.size() >= 5 ? nl.size() - 4 : 0; i >= 1; --i) {
    const size_t h = nl[i - 1] + 1, s = nl[i] + 1, plus = nl[i + 1] + 1, q = nl[i + 2] + 1;
    if (text[h] == '@' && plus < n && text[plus] == '+' && nl[i + 1] - s == nl[i + 3] - q) return nl[i + 3] + 1;
  }
  return SIZE_MAX;
}

FastxReader::FastxReader(const std::string& path, unsigned inflate_threads) : src_(path, inflate_threads) {}

void FastxReader::sniff(const char* line, size_t len) {
  if (len == 0) throw FastxError("empty first line: not FASTA/FASTQ");
  if (line[0] == '>')
    lines_per_record_ = 2;
  else if (line[0] == '@')
    lines_per_record_ = 4;
  else
    throw FastxError("first byte is neither '>' nor '@'");
}

// Sequence line only, as a view into the read buffer (valid until the next call): the '+' and
// quality lines of a FASTQ record are dropped at the start of the NEXT call.
bool FastxReader::next_seq(const char*& seq, size_t& seq_len) {
  const char* line;
  size_t len;
  if (pending_skip_) {
    if (!src_.next(line, len) || !src_.next(line, len)) throw FastxError("truncated FASTQ record");
    pending_skip_ = false;
  }
  if (!src_.next(line, len)) return false;
  if (lines_per_record_ == 0) sniff(line, len);
  if (!src_.next(seq, seq_len)) throw FastxError("truncated record: header without a sequence line");
  pending_skip_ = lines_per_record_ == 4;
  return true;
}

bool FastxReader::next(const char*& id, size_t& id_len, const char*& seq, size_t& seq_len) {
  const char* line;
  size_t len;
  if (pending_skip_) {
    if (!src_.next(line, len) || !src_.next(line, len)) throw FastxError("truncated FASTQ record");
    pending_skip_ = false;
  }
  if (!src_.next(line, len)) return false;
  if (lines_per_record_ == 0) sniff(line, len);
  id_.assign(len ? line + 1 : line, len ? len - 1 : 0);
  if (!src_.next(line, len)) throw FastxError("truncated record: header without a sequence line");
  seq_.assign(line, len);
  if (lines_per_record_ == 4) {
    if (!src_.next(line, len) || !src_.next(line, len)) throw FastxError("truncated FASTQ record");
  }
  id = id_.data();
  id_len = id_.size();
  seq = seq_.data();
  seq_len = seq_.size();
  return true;
}

}  // namespace sgh
