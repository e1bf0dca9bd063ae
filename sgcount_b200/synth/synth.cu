// synth.cu — see synth.h.  One generator function compiled for host and device.
#include "synth.h"

#include <cuda_runtime.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#define SGS_HD __host__ __device__ __forceinline__

namespace {

thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return 1;
}

SGS_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// counter-based draw: independent 64-bit value for (seed, stream, index, slot)
SGS_HD uint64_t draw(uint64_t seed, uint64_t stream, uint64_t index, uint64_t slot) {
  return splitmix64(splitmix64(seed ^ splitmix64(stream * 0xD1B54A32D192ED03ull + slot)) + index);
}

SGS_HD uint8_t comp(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return c;
  }
}

struct SampleView {
  uint64_t seed;
  uint32_t sample, n, k, read_len, offset, reverse;
  const uint8_t* library;   // n*k
  const uint32_t* cdf;      // n thresholds, cdf[g] = floor(2^32 * P(guide <= g)), last = 0xFFFFFFFF
  const uint8_t* prefix;    // offset + 1 bytes
  const uint8_t* scaffold;  // read_len + 2 bytes
};

// a base different from `c`, chosen by r in {0,1,2}
SGS_HD uint8_t other_base(uint8_t c, uint32_t r) {
  const char* acgt = "ACGT";
  uint32_t skip = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
  uint32_t j = r % 3;
  if (j >= skip) ++j;
  return (uint8_t)acgt[j];
}

// writes read_len bytes + '\n'
SGS_HD void make_read(const SampleView& s, uint64_t idx, uint8_t* out) {
  const uint32_t L = s.read_len, k = s.k;
  const uint64_t st = (uint64_t)s.sample + 1;
  const uint64_t rc = draw(s.seed, st, idx, 0);
  const uint32_t cls = (uint32_t)(rc % 100u);  // class
  uint8_t win[32];
  // guide by abundance: first g with cdf[g] > r
  const uint32_t r32 = (uint32_t)(draw(s.seed, st, idx, 1) >> 32);
  uint32_t lo = 0, hi = s.n - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (s.cdf[mid] > r32) hi = mid; else lo = mid + 1;
  }
  const uint8_t* g = s.library + (size_t)lo * k;
  for (uint32_t j = 0; j < k; ++j) win[j] = g[j];
  int shift = 0;
  int cut = -1;
  bool junk = false;
  const uint64_t rm = draw(s.seed, st, idx, 2);
  if (cls < 80) {
  } else if (cls < 88 || (cls < 90 && s.reverse)) {  // one ACGT substitution
    uint32_t p = (uint32_t)(rm % k);
    win[p] = other_base(win[p], (uint32_t)(rm >> 32));
  } else if (cls < 90) {  // one N
    win[rm % k] = 'N';
  } else if (cls < 92) {  // two substitutions
    uint32_t p = (uint32_t)(rm % k), q = (uint32_t)((rm >> 20) % (k - 1));
    if (q >= p) ++q;
    win[p] = other_base(win[p], (uint32_t)(rm >> 40));
    win[q] = other_base(win[q], (uint32_t)(rm >> 50));
  } else if (cls < 94) {
    shift = 1;
  } else if (cls < 96) {
    shift = -1;
  } else if (cls < 97) {  // truncated: N from the cut point on
    uint32_t span = s.offset + k >= 3 ? s.offset + k - 2 : 1;
    cut = (int)(rm % span);
  } else {
    junk = true;
  }
  // forward read: prefix[:offset+shift] + window + scaffold
  const int o = (int)s.offset + shift;
  for (uint32_t i = 0; i < L; ++i) {
    uint8_t c;
    int rel = (int)i - o;
    if (junk) {
      c = (uint8_t)"ACGT"[(draw(s.seed, st, idx, 8 + (i >> 5)) >> (2 * (i & 31))) & 3];
    } else if (rel < 0) {
      c = s.prefix[i];
    } else if (rel < (int)k) {
      c = win[rel];
    } else {
      c = s.scaffold[rel - (int)k];
    }
    if (cut >= 0 && (int)i >= cut) c = 'N';
    if (s.reverse)
      out[L - 1 - i] = comp(c);
    else
      out[i] = c;
  }
  out[L] = '\n';
}

__global__ void fill_kernel(SampleView s, uint64_t first, uint64_t n_reads, uint8_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  make_read(s, first + i, out + i * (s.read_len + 1));
}

}  // namespace

struct sgs_sample {
  uint64_t seed;
  uint32_t sample, n, k, read_len, offset, reverse;
  std::vector<uint8_t> library, prefix, scaffold;
  std::vector<uint32_t> cdf;
  // lazily created device copies
  int device = -1;
  uint8_t *d_library = nullptr, *d_prefix = nullptr, *d_scaffold = nullptr;
  uint32_t* d_cdf = nullptr;

  SampleView host_view() const {
    return SampleView{seed, sample, n, k, read_len, offset, reverse, library.data(), cdf.data(), prefix.data(), scaffold.data()};
  }
};

extern "C" {

const char* sgs_last_error(void) { return g_err.c_str(); }

int sgs_make_library(uint64_t seed, uint32_t n, uint32_t k, uint8_t* out) {
  if (!out || k == 0 || k > 32) return fail("bad arguments");
  std::unordered_set<std::string> seen;
  seen.reserve(n * 2);
  std::string s(k, 'A');
  for (uint32_t i = 0; i < n; ++i) {
    for (uint64_t attempt = 0;; ++attempt) {
      const uint64_t st = 0x4C4942ull;  // "LIB"
      const uint64_t kind = draw(seed, st, i, attempt * 64) % 1000;
      if (i > 0 && kind < 10) {
        // planted neighbour of an earlier guide: Hamming 1 (kind < 5) or Hamming 2
        uint64_t r = draw(seed, st, i, attempt * 64 + 1);
        uint32_t parent = (uint32_t)(r % i);
        s.assign((const char*)out + (size_t)parent * k, k);
        uint32_t p = (uint32_t)((r >> 24) % k);
        s[p] = (char)other_base((uint8_t)s[p], (uint32_t)(r >> 40));
        if (kind >= 5 && k > 1) {
          uint32_t q = (uint32_t)((r >> 48) % (k - 1));
          if (q >= p) ++q;
          s[q] = (char)other_base((uint8_t)s[q], (uint32_t)(r >> 56));
        }
      } else {
        for (uint32_t j = 0; j < k; ++j) {
          uint64_t r = draw(seed, st, i, attempt * 64 + 2 + j);
          double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
          double pg = 0.55 - 0.30 * (double)j / (double)(k > 1 ? k - 1 : 1);
          s[j] = u < pg ? 'G' : "ACT"[(r & 0x7FF) % 3];
        }
      }
      if (seen.insert(s).second) break;
    }
    memcpy(out + (size_t)i * k, s.data(), k);
  }
  return 0;
}

int sgs_sample_create(uint64_t seed, uint32_t sample_idx, const uint8_t* library, uint32_t n, uint32_t k,
                      uint32_t read_len, uint32_t offset, int reverse, sgs_sample** out) {
  if (!library || !out || n == 0 || k == 0 || k > 32) return fail("bad arguments");
  if (offset + k + 1 > read_len) return fail("offset + k + 1 must fit in the read");
  sgs_sample* s = new sgs_sample();
  s->seed = seed;
  s->sample = sample_idx;
  s->n = n;
  s->k = k;
  s->read_len = read_len;
  s->offset = offset;
  s->reverse = reverse != 0;
  s->library.assign(library, library + (size_t)n * k);
  const uint64_t st = 0x534D50ull + sample_idx;  // "SMP"
  s->prefix.resize(offset + 1);
  for (uint32_t i = 0; i <= offset; ++i) s->prefix[i] = (uint8_t)"ACGT"[draw(seed, st, i, 0) & 3];
  s->scaffold.resize(read_len + 2);
  // one scaffold for every sample of a seed (a vector backbone is constant across samples)
  for (uint32_t i = 0; i < read_len + 2; ++i) s->scaffold[i] = (uint8_t)"ACGT"[draw(seed, 0x534346ull, i, 0) & 3];
  // log-normal(0, 1) abundances -> 32-bit CDF thresholds
  std::vector<double> w(n);
  double total = 0.0;
  for (uint32_t g = 0; g < n; ++g) {
    uint64_t a = draw(seed, st, g, 1), b = draw(seed, st, g, 2);
    double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    w[g] = std::exp(z);
    total += w[g];
  }
  s->cdf.resize(n);
  double acc = 0.0;
  for (uint32_t g = 0; g < n; ++g) {
    acc += w[g];
    double t = acc / total * 4294967296.0;
    s->cdf[g] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  }
  s->cdf[n - 1] = 0xFFFFFFFFu;
  *out = s;
  return 0;
}

void sgs_sample_destroy(sgs_sample* s) {
  if (!s) return;
  if (s->device >= 0) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    cudaFree(s->d_library);
    cudaFree(s->d_prefix);
    cudaFree(s->d_scaffold);
    cudaFree(s->d_cdf);
    cudaSetDevice(prev);
  }
  delete s;
}

int sgs_sample_fill_host(const sgs_sample* s, uint64_t first_read, uint64_t n_reads, uint8_t* out, int n_threads) {
  if (!s || !out) return fail("NULL argument");
  if (n_threads < 1) n_threads = 1;
  const SampleView v = s->host_view();
  const uint64_t stride = s->read_len + 1;
  auto work = [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) make_read(v, first_read + i, out + i * stride);
  };
  if (n_threads == 1) {
    work(0, n_reads);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
      pool.emplace_back(work, n_reads * t / n_threads, n_reads * (t + 1) / n_threads);
    for (auto& th : pool) th.join();
  }
  return 0;
}

#define SGS_TRY(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(std::string("CUDA error in " #expr ": ") + cudaGetErrorString(_e)); \
  } while (0)

int sgs_sample_fill_device(sgs_sample* s, int device, uint64_t first_read, uint64_t n_reads, uint8_t* d_out,
                           void* stream) {
  if (!s || !d_out) return fail("NULL argument");
  if (n_reads == 0) return 0;
  int prev = 0;
  SGS_TRY(cudaGetDevice(&prev));
  SGS_TRY(cudaSetDevice(device));
  if (s->device != device) {
    if (s->device >= 0) return fail("sample already bound to another device");
    SGS_TRY(cudaMalloc(&s->d_library, s->library.size()));
    SGS_TRY(cudaMalloc(&s->d_prefix, s->prefix.size()));
    SGS_TRY(cudaMalloc(&s->d_scaffold, s->scaffold.size()));
    SGS_TRY(cudaMalloc(&s->d_cdf, s->cdf.size() * sizeof(uint32_t)));
    SGS_TRY(cudaMemcpy(s->d_library, s->library.data(), s->library.size(), cudaMemcpyHostToDevice));
    SGS_TRY(cudaMemcpy(s->d_prefix, s->prefix.data(), s->prefix.size(), cudaMemcpyHostToDevice));
    SGS_TRY(cudaMemcpy(s->d_scaffold, s->scaffold.data(), s->scaffold.size(), cudaMemcpyHostToDevice));
    SGS_TRY(cudaMemcpy(s->d_cdf, s->cdf.data(), s->cdf.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s->device = device;
  }
  SampleView v{s->seed, s->sample, s->n, s->k, s->read_len, s->offset, s->reverse,
               s->d_library, s->d_cdf, s->d_prefix, s->d_scaffold};
  const unsigned T = 256;
  const uint64_t blocks = (n_reads + T - 1) / T;
  if (blocks > 0x7FFFFFFFull) return fail("too many reads for one fill call");
  fill_kernel<<<(unsigned)blocks, T, 0, (cudaStream_t)stream>>>(v, first_read, n_reads, d_out);
  SGS_TRY(cudaGetLastError());
  SGS_TRY(cudaSetDevice(prev));
  return 0;
}

int sgs_sample_write_fastq(const sgs_sample* s, uint64_t first_read, uint64_t n_reads, const char* path,
                           uint64_t reads_per_member, int gz_level, int n_threads) {
  if (!s || !path) return fail("NULL argument");
  if (reads_per_member == 0) reads_per_member = 1 << 20;
  if (n_threads < 1) n_threads = 1;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("cannot create ") + path);
  const SampleView v = s->host_view();
  const uint32_t L = s->read_len;
  const uint64_t n_members = (n_reads + reads_per_member - 1) / reads_per_member;
  // members are produced n_threads at a time and written in order
  for (uint64_t m0 = 0; m0 < n_members; m0 += n_threads) {
    const uint64_t m1 = std::min<uint64_t>(n_members, m0 + n_threads);
    std::vector<std::string> blobs(m1 - m0);
    std::vector<int> status(m1 - m0, 0);
    std::vector<std::thread> pool;
    for (uint64_t m = m0; m < m1; ++m) {
      pool.emplace_back([&, m]() {
        const uint64_t lo = m * reads_per_member, hi = std::min(n_reads, lo + reads_per_member);
        std::string text;
        text.reserve((hi - lo) * (2 * L + 24));
        std::vector<uint8_t> line(L + 1);
        const std::string qual(L, 'I');
        char hdr[32];
        for (uint64_t i = lo; i < hi; ++i) {
          make_read(v, first_read + i, line.data());
          int hn = snprintf(hdr, sizeof hdr, "@r%llu\n", (unsigned long long)(first_read + i));
          text.append(hdr, hn);
          text.append((const char*)line.data(), L + 1);
          text.append("+\n");
          text.append(qual);
          text.push_back('\n');
        }
        if (gz_level <= 0) {
          blobs[m - m0].swap(text);
          return;
        }
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, gz_level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
          status[m - m0] = 1;
          return;
        }
        std::string& outb = blobs[m - m0];
        outb.resize(deflateBound(&zs, text.size()) + 64);
        zs.next_in = (Bytef*)text.data();
        zs.avail_in = (uInt)text.size();
        zs.next_out = (Bytef*)&outb[0];
        zs.avail_out = (uInt)outb.size();
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) status[m - m0] = 1;
        outb.resize(zs.total_out);
        deflateEnd(&zs);
      });
    }
    for (auto& th : pool) th.join();
    for (size_t i = 0; i < blobs.size(); ++i) {
      if (status[i] || fwrite(blobs[i].data(), 1, blobs[i].size(), f) != blobs[i].size()) {
        fclose(f);
        return fail("gzip/write failed");
      }
    }
  }
  fclose(f);
  return 0;
}

// BGZF: the blocked gzip that bgzip and sequencers write — members of at most 64 KB whose extra field
// ('B','C', block size - 1) lets a reader walk the blocks without inflating them.  Like bgzip, blocks
// are cut every `block_bytes` of text wherever that falls, not at record boundaries.
int sgs_sample_write_fastq_bgzf(const sgs_sample* s, uint64_t first_read, uint64_t n_reads, const char* path, int gz_level,
                                int n_threads, uint32_t block_bytes) {
  if (!s || !path) return fail("NULL argument");
  if (n_threads < 1) n_threads = 1;
  if (block_bytes == 0 || block_bytes > 65280) block_bytes = 65280;
  if (gz_level < 1) gz_level = 1;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("cannot create ") + path);
  const SampleView v = s->host_view();
  const uint32_t L = s->read_len;
  const uint64_t seg_reads = 32768;  // a thread's unit of work
  const uint64_t n_segs = (n_reads + seg_reads - 1) / seg_reads;
  std::string carry;  // text of a segment's tail that did not fill a block: prepended to the next segment
  for (uint64_t g0 = 0; g0 < n_segs; g0 += n_threads) {
    const uint64_t g1 = std::min<uint64_t>(n_segs, g0 + n_threads);
    std::vector<std::string> texts(g1 - g0), blobs(g1 - g0);
    std::vector<int> status(g1 - g0, 0);
    {
      std::vector<std::thread> pool;
      for (uint64_t g = g0; g < g1; ++g)
        pool.emplace_back([&, g]() {
          const uint64_t lo = g * seg_reads, hi = std::min(n_reads, lo + seg_reads);
          std::string& text = texts[g - g0];
          text.reserve((hi - lo) * (2 * L + 24));
          std::vector<uint8_t> line(L + 1);
          const std::string qual(L, 'I');
          char hdr[32];
          for (uint64_t i = lo; i < hi; ++i) {
            make_read(v, first_read + i, line.data());
            int hn = snprintf(hdr, sizeof hdr, "@r%llu\n", (unsigned long long)(first_read + i));
            text.append(hdr, hn);
            text.append((const char*)line.data(), L + 1);
            text.append("+\n");
            text.append(qual);
            text.push_back('\n');
          }
        });
      for (auto& th : pool) th.join();
    }
    // block boundaries run through the whole file: segment g starts with what g - 1 left over
    std::vector<size_t> head(g1 - g0, 0);
    for (uint64_t g = g0; g < g1; ++g) {
      std::string& text = texts[g - g0];
      text.insert(0, carry);
      const bool last = g + 1 == n_segs;
      const size_t whole = last ? text.size() : text.size() / block_bytes * block_bytes;
      carry.assign(text, whole, std::string::npos);
      text.resize(whole);
    }
    {
      std::vector<std::thread> pool;
      for (uint64_t g = g0; g < g1; ++g)
        pool.emplace_back([&, g]() {
          const std::string& text = texts[g - g0];
          std::string& outb = blobs[g - g0];
          std::vector<unsigned char> cbuf(block_bytes + 1024);
          for (size_t at = 0; at < text.size(); at += block_bytes) {
            const size_t n = std::min<size_t>(block_bytes, text.size() - at);
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, gz_level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
              status[g - g0] = 1;
              return;
            }
            zs.next_in = (Bytef*)text.data() + at;
            zs.avail_in = (uInt)n;
            zs.next_out = cbuf.data();
            zs.avail_out = (uInt)cbuf.size();
            const int rc = deflate(&zs, Z_FINISH);
            const size_t clen = zs.total_out;
            deflateEnd(&zs);
            if (rc != Z_STREAM_END || clen + 26 > 65536) {
              status[g - g0] = 1;
              return;
            }
            const uint32_t bsize = (uint32_t)(18 + clen + 8 - 1), crc = (uint32_t)crc32(0L, (const Bytef*)text.data() + at, (uInt)n);
            const unsigned char hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                                           (unsigned char)(bsize & 0xff), (unsigned char)(bsize >> 8)};
            outb.append((const char*)hdr, 18);
            outb.append((const char*)cbuf.data(), clen);
            const uint32_t tail[2] = {crc, (uint32_t)n};
            outb.append((const char*)tail, 8);
          }
        });
      for (auto& th : pool) th.join();
    }
    for (size_t i = 0; i < blobs.size(); ++i)
      if (status[i] || fwrite(blobs[i].data(), 1, blobs[i].size(), f) != blobs[i].size()) {
        fclose(f);
        return fail("bgzf/write failed");
      }
  }
  // the empty end-of-file block of the BGZF specification
  static const unsigned char eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (fwrite(eof_block, 1, 28, f) != 28) {
    fclose(f);
    return fail("bgzf/write failed");
  }
  fclose(f);
  return 0;
}

}  // extern "C"
