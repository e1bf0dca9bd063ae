#include "fastx.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

namespace sgh {

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
  GzSource(const std::string& path, unsigned threads) {
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
    if (threads_ > 1 && size_ > (1u << 16)) find_candidates();
    if (cand_.size() <= 1) threads_ = 1;  // a single member: nothing to run in parallel
    if (threads_ > 1) {
      jobs_.resize(cand_.size());
      window_ = 2 * threads_;
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
    if (threads_ > 1 && !sequential_) {
      if (pos_ >= size_) return false;
      // the candidate that starts exactly at pos_
      auto it = std::lower_bound(cand_.begin(), cand_.end(), pos_);
      if (it != cand_.end() && *it == pos_) {
        const size_t j = (size_t)(it - cand_.begin());
        std::unique_lock<std::mutex> lk(mu_);
        consumer_at_ = j;
        cv_work_.notify_all();
        cv_done_.wait(lk, [&] { return jobs_[j].done; });
        Job& job = jobs_[j];
        if (job.ok) {
          std::vector<char> data = std::move(job.out);
          const size_t end = job.end;
          lk.unlock();
          if (out.empty()) {
            out.swap(data);
          } else if (!data.empty()) {
            const size_t at = out.size();
            out.resize(at + data.size());
            memcpy(out.data() + at, data.data(), data.size());
          }
          pos_ = end;
          if (data.empty() && out.empty()) return read_more(out);  // an empty member
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
    return read_sequential(out);
  }

 private:
  struct Job {
    std::vector<char> out;
    size_t end = 0;
    bool ok = false, done = false, taken = false;
  };

  void find_candidates() {
    const unsigned char* p = data_;
    const unsigned char* const last = data_ + size_ - 18;  // smallest member: 10 header + 8 trailer
    while (p < last) {
      p = static_cast<const unsigned char*>(memchr(p, 0x1f, (size_t)(last - p)));
      if (!p) break;
      if (p[1] == 0x8b && p[2] == 0x08 && (p[3] & 0xE0) == 0) cand_.push_back((size_t)(p - data_));
      ++p;
    }
    if (cand_.empty() || cand_[0] != 0) cand_.clear();  // not a gzip file: let zlib report it
  }

  // inflate ONE member starting at `from`; false if the bytes there are not a complete member
  bool inflate_member(size_t from, std::vector<char>& out, size_t& end) {
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
      std::vector<char> out;
      // size hint: ISIZE of a member that ends where the next candidate starts
      const size_t nxt = j + 1 < cand_.size() ? cand_[j + 1] : size_;
      if (nxt >= cand_[j] + 18) {
        uint32_t isize;
        memcpy(&isize, data_ + nxt - 4, 4);
        if (isize < (1u << 30)) out.reserve((size_t)isize + 64);
      }
      size_t end = 0;
      const bool ok = inflate_member(cand_[j], out, end);
      {
        std::lock_guard<std::mutex> lk(mu_);
        jobs_[j].out = std::move(out);
        jobs_[j].end = end;
        jobs_[j].ok = ok;
        jobs_[j].done = true;
      }
      cv_done_.notify_all();
    }
  }

  bool read_sequential(std::vector<char>& out) {
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
