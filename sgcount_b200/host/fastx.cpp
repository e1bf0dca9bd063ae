#include "fastx.h"

#include <cstring>

namespace sgh {

namespace {
constexpr size_t kInChunk = 1 << 20;
constexpr size_t kBufChunk = 4 << 20;
bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
}  // namespace

LineSource::LineSource(const std::string& path) {
  fp_ = fopen(path.c_str(), "rb");
  if (!fp_) throw FastxError("cannot open " + path);
  gz_ = ends_with(path, ".gz");
  buf_.resize(kBufChunk);
  if (gz_) {
    in_.resize(kInChunk);
    if (inflateInit2(&zs_, 15 + 16) != Z_OK) throw FastxError("inflateInit2 failed");
    z_init_ = true;
  }
}

LineSource::~LineSource() {
  if (z_init_) inflateEnd(&zs_);
  if (fp_) fclose(fp_);
}

// Moves the unread tail to the front and appends more bytes.  Returns false when no byte
// could be added (end of input).
bool LineSource::refill() {
  if (eof_) return false;
  if (pos_ > 0) {
    memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
    end_ -= pos_;
    pos_ = 0;
  }
  if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);  // one line longer than the buffer
  size_t got = 0;
  if (!gz_) {
    got = fread(buf_.data() + end_, 1, buf_.size() - end_, fp_);
    if (got == 0) eof_ = true;
  } else {
    while (got == 0 && !z_eof_) {
      if (zs_.avail_in == 0) {
        zs_.next_in = in_.data();
        zs_.avail_in = (uInt)fread(in_.data(), 1, in_.size(), fp_);
        if (zs_.avail_in == 0) {
          z_eof_ = true;
          break;
        }
      }
      zs_.next_out = reinterpret_cast<Bytef*>(buf_.data() + end_);
      zs_.avail_out = (uInt)std::min<size_t>(buf_.size() - end_, 1u << 30);
      const uInt before = zs_.avail_out;
      int rc = inflate(&zs_, Z_NO_FLUSH);
      got = before - zs_.avail_out;
      if (rc == Z_STREAM_END) {
        // next member of a multi-member file (flate2 MultiGzDecoder), if any bytes remain
        if (zs_.avail_in == 0) {
          zs_.next_in = in_.data();
          zs_.avail_in = (uInt)fread(in_.data(), 1, in_.size(), fp_);
        }
        if (zs_.avail_in == 0)
          z_eof_ = true;
        else if (inflateReset(&zs_) != Z_OK)
          throw FastxError("inflateReset failed");
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        throw FastxError(std::string("gzip stream is corrupt: ") + (zs_.msg ? zs_.msg : "inflate error"));
      }
    }
    if (got == 0) eof_ = true;
  }
  end_ += got;
  return got > 0;
}

bool LineSource::next(const char*& begin, size_t& len) {
  for (;;) {
    const char* p = buf_.data() + pos_;
    const char* nl = static_cast<const char*>(memchr(p, '\n', end_ - pos_));
    if (nl) {
      begin = p;
      len = (size_t)(nl - p);
      pos_ += len + 1;
      return true;
    }
    if (!refill()) {
      if (pos_ == end_) return false;
      begin = buf_.data() + pos_;  // last line without a newline
      len = end_ - pos_;
      pos_ = end_;
      return true;
    }
  }
}

FastxReader::FastxReader(const std::string& path) : src_(path) {}

bool FastxReader::next(const char*& id, size_t& id_len, const char*& seq, size_t& seq_len) {
  const char* line;
  size_t len;
  if (!src_.next(line, len)) return false;
  if (lines_per_record_ == 0) {
    if (len == 0) throw FastxError("empty first line: not FASTA/FASTQ");
    if (line[0] == '>')
      lines_per_record_ = 2;
    else if (line[0] == '@')
      lines_per_record_ = 4;
    else
      throw FastxError("first byte is neither '>' nor '@'");
  }
  id_.assign(len ? line + 1 : line, len ? len - 1 : 0);
  if (!src_.next(seq, seq_len)) throw FastxError("truncated record: header without a sequence line");
  if (lines_per_record_ == 4) {
    // the sequence view must survive two more reads: they only advance inside the buffer
    // unless a refill moves it, so copy when that can happen is avoided by reading ahead here
    static thread_local std::string hold;
    hold.assign(seq, seq_len);
    const char* skip;
    size_t skip_len;
    if (!src_.next(skip, skip_len) || !src_.next(skip, skip_len))
      throw FastxError("truncated FASTQ record");
    seq = hold.data();
    seq_len = hold.size();
  }
  id = id_.data();
  id_len = id_.size();
  return true;
}

}  // namespace sgh
