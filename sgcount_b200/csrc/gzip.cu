// gzip.cu — FASTQ straight from BGZF blocks: inflate, record framing and counting on the device.
//
// The reference reads its samples through fxread::initialize_reader (count.rs:24): gzip inflate and
// line splitting on one host thread per sample.  The C++ host spreads that over its cores
// (host/fastx.cpp), which is where it stops: ~0.2 core-seconds per million reads whatever the GPU
// does.  Blocked gzip — BGZF, what bgzip and sequencers write: independent members of at most
// 64 KB, each announcing its size in a 'BC' extra field — needs no host core at all:
//
//   inflate_blocks_kernel   ONE THREAD per block runs the sequential DEFLATE decoder of
//                           inflate_core.h (canonical Huffman, 7- and 5-bit look-ahead tables in
//                           bank-interleaved shared memory, the rest of the code tables in local
//                           memory), writing its text where the prefix sum of the blocks' ISIZE
//                           fields puts it.  Tens of thousands of blocks are in flight, so the
//                           serial bit-by-bit dependency of one stream does not matter.
//   newline_count_kernel    newlines per 4 KB chunk of the contiguous text           } record
//   (exclusive scan)        line number of every chunk's first newline               } framing:
//   span_extract_kernel     every line 4r+1 is the sequence of record r: its guide   } FASTQ is
//                           window span (sgc_span_geometry) goes to span record r    } 4 lines
//   count_stream_kernel     the ordinary count of fixed-stride span records (count.cu).
//
// BGZF cuts blocks anywhere, so a wave of blocks ends inside a record: the bytes after the last
// complete record are carried in front of the next wave's text.  Only fixed-length 4-line FASTQ
// takes this path (the head of the sample, which the host reads for the offset detector, says so);
// anything irregular is reported and the caller counts the sample through the host path instead.
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "inflate_core.h"
#include "internal.h"

namespace sgc {
namespace {

constexpr int kInflateThreads = 64;
constexpr uint32_t kHeadroom = 1u << 16;  // room in front of a wave's text for the carried bytes
constexpr uint32_t kChunk = 4096;         // bytes per newline-count chunk
constexpr uint32_t kChunkThreads = 256;   // 16 bytes per thread

// look-ahead tables and per-length code counts in shared memory, one column per thread: element e
// of thread t at [e * T + t]; the symbol lists stay in local memory (one access per long code).
// The handle is two pointers, passed by value (inflate_core.h).
struct DeviceSymbols {
  uint16_t lsym[inflate::kLitLenSyms], dsym[inflate::kDistSyms];
};
struct DeviceTables {
  uint16_t* col;
  DeviceSymbols* syms;
  static constexpr int kCounts = (1 << inflate::kFastBits) + (1 << inflate::kDistFastBits);
  __host__ __device__ __forceinline__ uint16_t get_lcount(int i) const { return col[(kCounts + i) * kInflateThreads]; }
  __host__ __device__ __forceinline__ void set_lcount(int i, uint16_t v) const { col[(kCounts + i) * kInflateThreads] = v; }
  __host__ __device__ __forceinline__ uint16_t get_lsym(int i) const { return syms->lsym[i]; }
  __host__ __device__ __forceinline__ void set_lsym(int i, uint16_t v) const { syms->lsym[i] = v; }
  __host__ __device__ __forceinline__ uint16_t get_dcount(int i) const { return col[(kCounts + 16 + i) * kInflateThreads]; }
  __host__ __device__ __forceinline__ void set_dcount(int i, uint16_t v) const { col[(kCounts + 16 + i) * kInflateThreads] = v; }
  __host__ __device__ __forceinline__ uint16_t get_dsym(int i) const { return syms->dsym[i]; }
  __host__ __device__ __forceinline__ void set_dsym(int i, uint16_t v) const { syms->dsym[i] = v; }
  __host__ __device__ __forceinline__ uint16_t get_lfast(int i) const { return col[i * kInflateThreads]; }
  __host__ __device__ __forceinline__ void set_lfast(int i, uint16_t v) const { col[i * kInflateThreads] = v; }
  __host__ __device__ __forceinline__ uint16_t get_dfast(int i) const { return col[((1 << inflate::kFastBits) + i) * kInflateThreads]; }
  __host__ __device__ __forceinline__ void set_dfast(int i, uint16_t v) const { col[((1 << inflate::kFastBits) + i) * kInflateThreads] = v; }
};
constexpr size_t kInflateSmem = (DeviceTables::kCounts + 32) * kInflateThreads * sizeof(uint16_t);

// what the kernels of a stream hand from wave to wave, and back to the host
struct StreamState {
  unsigned int bad_block;       // first block (index within its wave) that did not decode, or 0xFFFFFFFF
  int bad_status;               // its inflate::Status; 100: size / length disagreement; 101: CRC-32 mismatch
  unsigned int format_error;    // 1 irregular read length, 2 not FASTQ framing, 4 a record longer than the headroom
  unsigned int tail;            // bytes carried in front of the next wave
  unsigned int next_tail;
  unsigned int head_skip;       // bytes at the start of this wave's text that belong to the previous wave
  unsigned int self_contained;  // the caller cut this wave at record boundaries: nothing may be left over
  unsigned long long lines;     // newlines of the current wave's region
  unsigned long long records;   // records of the current wave
  unsigned long long records_total;
};

__global__ void __launch_bounds__(kInflateThreads) inflate_blocks_kernel(const uint8_t* __restrict__ gz,
                                                                        const uint64_t* __restrict__ begin,
                                                                        const uint64_t* __restrict__ out_off, uint32_t n,
                                                                        uint8_t* __restrict__ text, StreamState* st) {
  extern __shared__ uint16_t sm[];
  const uint32_t m = blockIdx.x * kInflateThreads + threadIdx.x;
  if (m >= n) return;
  DeviceSymbols symbols;
  const DeviceTables t{sm + threadIdx.x, &symbols};
  const uint64_t base = begin[0];
  const uint8_t* in = gz + (begin[m] - base);
  const size_t in_len = (size_t)(begin[m + 1] - begin[m]);
  uint8_t* out = text + out_off[m];
  const size_t cap = (size_t)(out_off[m + 1] - out_off[m]);
  size_t consumed = 0, produced = 0;
  uint32_t crc = 0, isize = 0;
  int rc = inflate::gunzip_member(in, in_len, out, cap, t, &consumed, &produced, &crc, &isize);
  if (rc == inflate::kOk && (consumed != in_len || produced != cap || isize != (uint32_t)cap)) rc = 100;
  if (rc != inflate::kOk) {
    const unsigned int prev = atomicMin(&st->bad_block, m);
    if (m < prev) st->bad_status = rc;  // (benign race between two bad blocks: either status will do)
  }
}

// CRC-32 of every block's text against the trailer of its member (RFC 1952; flate2 checks it for
// the reference): one warp per block, 32 pieces joined as inflate_core.h describes.
constexpr int kCrcWarps = 8;
__global__ void __launch_bounds__(kCrcWarps * 32) crc_blocks_kernel(const uint8_t* __restrict__ gz,
                                                                   const uint64_t* __restrict__ begin,
                                                                   const uint64_t* __restrict__ out_off, uint32_t n,
                                                                   const uint8_t* __restrict__ text, StreamState* st) {
  __shared__ uint32_t table[256];
  table[threadIdx.x] = inflate::crc_table_entry(threadIdx.x);
  __syncthreads();
  const uint32_t m = blockIdx.x * kCrcWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= n) return;
  const uint8_t* d = text + out_off[m];
  const size_t len = (size_t)(out_off[m + 1] - out_off[m]);
  const size_t c = inflate::crc_piece_len(len), first = len - 31 * c;
  const size_t at = lane == 0 ? 0 : first + (lane - 1) * c;
  uint32_t crc = inflate::crc_bytes(d + at, lane == 0 ? first : c, [&](uint32_t i) { return table[i]; });
  uint32_t p = inflate::crc_x8n(c);  // the same in every lane
#pragma unroll
  for (int s = 1; s < 32; s *= 2) {
    const uint32_t other = __shfl_down_sync(0xffffffffu, crc, s);
    if ((lane & (2 * s - 1)) == 0) crc = inflate::crc_multmodp(p, crc) ^ other;
    p = inflate::crc_multmodp(p, p);
  }
  if (lane == 0) {
    const uint8_t* t = gz + (begin[m + 1] - begin[0]) - 8;
    const uint32_t stored = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    if (crc != stored) {
      const unsigned int prev = atomicMin(&st->bad_block, m);
      if (m < prev) st->bad_status = 101;
    }
  }
}

// bytes [lo, hi) of the text are the current region: the carried tail, then this wave's text
__device__ __forceinline__ void region_of(const StreamState* st, uint64_t n_text, uint64_t* lo, uint64_t* hi) {
  *lo = kHeadroom - st->tail + st->head_skip;
  *hi = (uint64_t)kHeadroom + n_text;
}

__device__ __forceinline__ uint32_t newline_mask16(const uint8_t* __restrict__ text, uint64_t at, uint64_t lo, uint64_t hi) {
  // 16 bytes at `at` (16-byte aligned); bit j set iff byte at + j is a newline inside [lo, hi)
  const uint4 v = *reinterpret_cast<const uint4*>(text + at);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t mask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i] ^ 0x0A0A0A0Au;                                 // zero byte where a newline is
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;  // bit 7 of every zero byte
    mask |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * i);
  }
  if (at < lo) mask &= lo - at >= 16 ? 0u : ~0u << (uint32_t)(lo - at);
  if (at + 16 > hi) mask &= hi <= at ? 0u : (1u << (uint32_t)(hi - at)) - 1u;
  return mask;
}

__global__ void __launch_bounds__(kChunkThreads) newline_count_kernel(const uint8_t* __restrict__ text, uint64_t n_text,
                                                                     const StreamState* st, uint32_t* __restrict__ counts) {
  uint64_t lo, hi;
  region_of(st, n_text, &lo, &hi);
  const uint64_t at = (uint64_t)blockIdx.x * kChunk + threadIdx.x * 16;
  uint32_t c = at < hi && at + 16 > lo ? __popc(newline_mask16(text, at, lo, hi)) : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ uint32_t warp_sums[kChunkThreads / 32];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (uint32_t i = 0; i < kChunkThreads / 32; ++i) s += warp_sums[i];
    counts[blockIdx.x] = s;
  }
}

// counts[n_chunks] (set to 0 by the caller) scans to the number of lines of the region
__global__ void wave_totals_kernel(const uint32_t* __restrict__ first_line, uint32_t n_chunks, StreamState* st) {
  st->lines = first_line[n_chunks];
  st->records = st->lines / 4;
  st->next_tail = 0;  // set by span_extract_kernel when a last complete record exists
}

__global__ void __launch_bounds__(kChunkThreads) span_extract_kernel(const uint8_t* __restrict__ text, uint64_t n_text,
                                                                    StreamState* st, const uint32_t* __restrict__ first_line,
                                                                    uint32_t read_len, uint32_t span_start, uint32_t span_len,
                                                                    uint32_t span_stride, uint8_t* __restrict__ spans,
                                                                    uint32_t* __restrict__ seq_start, uint32_t* __restrict__ seq_end) {
  uint64_t lo, hi;
  region_of(st, n_text, &lo, &hi);
  const uint64_t lines = st->lines, records = st->records;
  const uint64_t at = (uint64_t)blockIdx.x * kChunk + threadIdx.x * 16;
  uint32_t mask = at < hi && at + 16 > lo ? newline_mask16(text, at, lo, hi) : 0u;
  // line number of this thread's first newline: chunk base + newlines of the threads before it
  const uint32_t mine = __popc(mask);
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
    if ((threadIdx.x & 31) >= (uint32_t)d) inc += v;
  }
  __shared__ uint32_t warp_sums[kChunkThreads / 32];
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = inc;
  __syncthreads();
  uint32_t before = first_line[blockIdx.x] + inc - mine;
  for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += warp_sums[w];
  if (blockIdx.x == 0 && threadIdx.x == 0 && hi > lo && text[lo] != '@') atomicOr(&st->format_error, 2u);
  uint64_t line = before;
  while (mask) {
    const uint32_t j = __ffs((int)mask) - 1;
    mask &= mask - 1;
    const uint64_t p = at + j;  // a newline: the end of line `line` of the region
    const uint64_t r = line >> 2;
    if (read_len == 0 && r < records) {
      // reads of any length: the sequence line stays where it is, [seq_start[r], seq_end[r]) of the text
      if ((line & 3) == 0) {
        seq_start[r] = (uint32_t)(p + 1);
      } else if ((line & 3) == 1) {
        seq_end[r] = (uint32_t)p;
        if (text[p + 1] != '+') atomicOr(&st->format_error, 2u);  // (p + 1 < hi: the record is complete)
      } else if ((line & 3) == 3 && p + 1 < hi && text[p + 1] != '@') {
        atomicOr(&st->format_error, 2u);
      }
    } else if ((line & 3) == 0 && r < records) {
      // the header of record r ends here; its sequence is the next read_len bytes, then "\n+"
      const uint64_t s = p + 1;
      if (s + read_len + 1 >= hi || text[s + read_len] != '\n' || text[s + read_len + 1] != '+') {
        atomicOr(&st->format_error, 1u);
      } else {
        uint8_t* dst = spans + r * span_stride;
        for (uint32_t b = 0; b < span_len; ++b) dst[b] = text[s + span_start + b];
      }
    }
    if ((line & 3) == 3 && r + 1 == records) {
      st->next_tail = (unsigned int)(hi - (p + 1));  // what follows the last complete record
    }
    ++line;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && records == 0) {
    // no complete record in the region: all of it is carried on (it must fit the headroom)
    if (hi - lo > kHeadroom) atomicOr(&st->format_error, 4u);
    st->next_tail = (unsigned int)min((uint64_t)kHeadroom, hi - lo);
  }
  (void)lines;
}

// the bytes after the last complete record, parked in `tail_buf`, then put in front of the next
// wave's text (two steps: source and destination may overlap when a wave is tiny)
__global__ void tail_save_kernel(const uint8_t* __restrict__ text, uint64_t n_text, StreamState* st, uint8_t* __restrict__ tail_buf) {
  const uint64_t hi = (uint64_t)kHeadroom + n_text;
  const uint32_t tail = st->next_tail;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x) tail_buf[i] = text[hi - tail + i];
}
__global__ void wave_begin_kernel(StreamState* st, unsigned int head_skip, unsigned int self_contained) {
  st->head_skip = head_skip;
  st->self_contained = self_contained;
}
__global__ void wave_commit_kernel(StreamState* st) {
  if (st->self_contained && st->next_tail) atomicOr(&st->format_error, 8u);  // the cut was not a record boundary
  st->tail = st->next_tail;
  st->head_skip = 0;
  st->records_total += st->records;
}
__global__ void tail_restore_kernel(uint8_t* __restrict__ text, const StreamState* st, const uint8_t* __restrict__ tail_buf) {
  const uint32_t tail = st->tail;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x)
    text[kHeadroom - tail + i] = tail_buf[i];
}

// Device scratch of the ingest streams comes from one memory pool per device whose unused memory stays
// cached, through the stream-ordered allocator.  cudaFree waits for ALL work on the device and cudaMalloc
// queues up behind it: with several samples in flight on one device (the CLI runs four) every sample that
// finished — a dozen buffers to release — stalled the others' next allocation until their kernels had
// drained, and the samples ran one after the other.  cudaFreeAsync is ordered on the stream only.
struct ScratchPools {
  std::mutex mu;
  std::map<int, cudaMemPool_t> pool;  // NULL: the device has no memory pools (cudaMalloc / cudaFree instead)
};
cudaMemPool_t scratch_pool(int device) {
  static ScratchPools pools;
  std::lock_guard<std::mutex> lk(pools.mu);
  auto it = pools.pool.find(device);
  if (it != pools.pool.end()) return it->second;
  cudaMemPool_t pool = nullptr;
  int supported = 0;
  if (cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && supported) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
      uint64_t keep = 8ull << 30;  // up to 8 GiB of freed memory stay with the pool for the next wave / sample
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      pool = nullptr;
      cudaGetLastError();
    }
  }
  pools.pool[device] = pool;
  return pool;
}
int scratch_alloc(void** p, size_t bytes, cudaMemPool_t pool, cudaStream_t stream) {
  if (pool)
    SGC_CUDA_TRY(cudaMallocFromPoolAsync(p, bytes, pool, stream));
  else
    SGC_CUDA_TRY(cudaMalloc(p, bytes));
  return SGC_OK;
}
void scratch_free(void* p, cudaMemPool_t pool, cudaStream_t stream) {
  if (!p) return;
  if (pool)
    cudaFreeAsync(p, stream);
  else
    cudaFree(p);
}

template <typename T>
int grow(T** p, size_t* cap, size_t need, cudaMemPool_t pool, cudaStream_t stream) {
  if (need <= *cap) return SGC_OK;
  scratch_free(*p, pool, stream);  // (stream order: the kernels that still read it come first)
  *p = nullptr;
  *cap = 0;
  const size_t want = need + need / 4;
  int rc = scratch_alloc(reinterpret_cast<void**>(p), want * sizeof(T), pool, stream);
  if (rc) return rc;
  *cap = want;
  return SGC_OK;
}

}  // namespace
}  // namespace sgc

using namespace sgc;

struct sgc_fastq_stream {
  sgc_counter* counter = nullptr;  // NULL once released (its counter was destroyed first)
  int device = 0;
  uint32_t read_len = 0, span_start = 0, span_len = 0, span_stride = 0;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;  // scratch_pool(device)
  StreamState* d_state = nullptr;
  uint8_t *d_gz = nullptr, *d_text = nullptr, *d_spans = nullptr, *d_tail = nullptr;
  uint32_t *d_seq_start = nullptr, *d_seq_end = nullptr;  // variable-length mode
  size_t seq_cap_a = 0, seq_cap_b = 0;
  uint64_t *d_begin = nullptr, *d_outoff = nullptr;
  uint32_t *d_counts = nullptr, *d_first = nullptr, *d_sums = nullptr;
  size_t gz_cap = 0, text_cap = 0, spans_cap = 0, begin_cap = 0, outoff_cap = 0, counts_cap = 0, first_cap = 0, sums_cap = 0;
  std::vector<uint64_t> h_begin, h_outoff;
  uint64_t blocks_total = 0, records_total = 0;
  bool failed = false;
};

namespace {

int stream_error(sgc_fastq_stream* s, const StreamState& st, uint64_t first_block) {
  s->failed = true;
  char buf[200];
  if (st.bad_block != 0xFFFFFFFFu) {
    if (st.bad_status == 101)
      snprintf(buf, sizeof buf, "gzip block %llu: CRC-32 of the inflated bytes differs from the member's trailer",
               (unsigned long long)(first_block + st.bad_block));
    else
      snprintf(buf, sizeof buf, "gzip block %llu did not inflate on the device (status %d)",
               (unsigned long long)(first_block + st.bad_block), st.bad_status);
    return set_error(SGC_ERR_GZIP, buf);
  }
  snprintf(buf, sizeof buf, "not fixed-length 4-line FASTQ (%s)",
           (st.format_error & 2) ? "a record does not start with '@'"
                                 : ((st.format_error & 4) ? "a record longer than 64 KB"
                                                          : ((st.format_error & 8) ? "a wave was not cut at a record boundary"
                                                                                   : "a read of another length, or a line that is not '+'")));
  return set_error(SGC_ERR_FASTQ_FORMAT, buf);
}

// frames and counts the region made of the carried tail and n_text fresh bytes at d_text + kHeadroom
int frame_and_count(sgc_fastq_stream* s, uint64_t n_text, uint64_t first_block) {
  const uint32_t n_chunks = (uint32_t)(((uint64_t)kHeadroom + n_text + kChunk - 1) / kChunk);
  int rc = grow(&s->d_counts, &s->counts_cap, (size_t)n_chunks + 1, s->pool, s->stream);
  if (rc == SGC_OK) rc = grow(&s->d_first, &s->first_cap, (size_t)n_chunks + 1, s->pool, s->stream);
  if (rc == SGC_OK) rc = grow(&s->d_sums, &s->sums_cap, (size_t)n_chunks / 2048 + 2, s->pool, s->stream);
  if (rc) return rc;
  // the carried bytes go in front of the fresh text (they are kept in d_tail: d_text may have moved)
  tail_restore_kernel<<<8, 256, 0, s->stream>>>(s->d_text, s->d_state, s->d_tail);
  SGC_CUDA_TRY(cudaMemsetAsync(s->d_counts + n_chunks, 0, sizeof(uint32_t), s->stream));
  newline_count_kernel<<<n_chunks, kChunkThreads, 0, s->stream>>>(s->d_text, n_text, s->d_state, s->d_counts);
  rc = exclusive_scan_u32(s->d_counts, n_chunks + 1, s->d_first, s->d_sums, s->stream);
  if (rc) return rc;
  wave_totals_kernel<<<1, 1, 0, s->stream>>>(s->d_first, n_chunks, s->d_state);
  // how many records the region holds decides the size of the span / offset buffers
  StreamState pre;
  SGC_CUDA_TRY(cudaMemcpyAsync(&pre, s->d_state, sizeof pre, cudaMemcpyDeviceToHost, s->stream));
  SGC_CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (pre.bad_block != 0xFFFFFFFFu) return stream_error(s, pre, first_block);
  const size_t max_records = (size_t)pre.records + 1;
  if (s->read_len) {
    rc = grow(&s->d_spans, &s->spans_cap, max_records * s->span_stride + 256, s->pool, s->stream);
  } else {
    rc = grow(&s->d_seq_start, &s->seq_cap_a, max_records, s->pool, s->stream);
    if (rc == SGC_OK) rc = grow(&s->d_seq_end, &s->seq_cap_b, max_records, s->pool, s->stream);
  }
  if (rc) return rc;
  span_extract_kernel<<<n_chunks, kChunkThreads, 0, s->stream>>>(s->d_text, n_text, s->d_state, s->d_first, s->read_len,
                                                                 s->span_start, s->span_len, s->span_stride, s->d_spans,
                                                                 s->d_seq_start, s->d_seq_end);
  tail_save_kernel<<<8, 256, 0, s->stream>>>(s->d_text, n_text, s->d_state, s->d_tail);
  wave_commit_kernel<<<1, 1, 0, s->stream>>>(s->d_state);
  SGC_CUDA_TRY(cudaGetLastError());
  StreamState st;
  SGC_CUDA_TRY(cudaMemcpyAsync(&st, s->d_state, sizeof st, cudaMemcpyDeviceToHost, s->stream));
  SGC_CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (st.bad_block != 0xFFFFFFFFu || st.format_error) return stream_error(s, st, first_block);
  if (st.records) {
    if (s->read_len)
      rc = sgc_counter_submit_device(s->counter, s->d_spans, st.records * s->span_stride + 64, nullptr, s->span_stride,
                                     s->span_len, st.records, nullptr);
    else  // the sequence lines in place, through the line kernel
      rc = count_gathered_lines(s->counter, s->d_text, (uint64_t)kHeadroom + n_text, s->d_seq_start, s->d_seq_end, st.records);
    if (rc) return rc;
  }
  s->records_total = st.records_total;
  return SGC_OK;
}

}  // namespace

// Frees the device side and detaches the stream from its counter; the handle itself stays valid.
void sgc::fastq_stream_release(sgc_fastq_stream* s) {
  if (!s || !s->counter) return;
  DeviceGuard guard(s->device);
  cudaStreamSynchronize(s->stream);
  auto& list = s->counter->fastq_streams;
  list.erase(std::remove(list.begin(), list.end(), s), list.end());
  s->counter = nullptr;
  s->failed = true;  // nothing more can be submitted
  for (void* p : {(void*)s->d_state, (void*)s->d_gz, (void*)s->d_text, (void*)s->d_spans, (void*)s->d_tail, (void*)s->d_seq_start,
                  (void*)s->d_seq_end, (void*)s->d_begin, (void*)s->d_outoff, (void*)s->d_counts, (void*)s->d_first, (void*)s->d_sums})
    scratch_free(p, s->pool, s->stream);
  s->d_state = nullptr;
  s->d_gz = s->d_text = s->d_spans = s->d_tail = nullptr;
  s->d_seq_start = s->d_seq_end = nullptr;
  s->d_begin = s->d_outoff = nullptr;
  s->d_counts = s->d_first = s->d_sums = nullptr;
}

extern "C" {

int sgc_fastq_stream_create(sgc_counter* counter, uint32_t read_len, uint32_t span_start, uint32_t span_len,
                            sgc_fastq_stream** out) {
  if (!counter || !out) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  if (read_len != 0 && (span_len == 0 || (uint64_t)span_start + span_len > read_len || read_len + 4u > kHeadroom))
    return set_error(SGC_ERR_INVALID_ARG, "the span does not fit the read");
  DeviceGuard guard(counter->lib->device);
  sgc_fastq_stream* s = new sgc_fastq_stream();
  s->counter = counter;
  s->device = counter->lib->device;
  s->read_len = read_len;
  s->span_start = span_start;
  s->span_len = span_len;
  s->span_stride = (span_len + 7u) & ~7u;
  s->device = counter->lib->device;
  s->stream = counter->stream;  // one stream: the count of a wave follows its framing
  s->pool = scratch_pool(s->device);
  struct Cleanup {
    sgc_fastq_stream* s;
    ~Cleanup() {
      if (s) sgc_fastq_stream_destroy(s);
    }
  } cleanup{s};
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&s->d_state), sizeof(StreamState), s->pool, s->stream)) return rc;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&s->d_tail), kHeadroom, s->pool, s->stream)) return rc;
  StreamState st0{};
  st0.bad_block = 0xFFFFFFFFu;
  SGC_CUDA_TRY(cudaMemcpyAsync(s->d_state, &st0, sizeof st0, cudaMemcpyHostToDevice, s->stream));
  SGC_CUDA_TRY(cudaStreamSynchronize(s->stream));
  SGC_CUDA_TRY(cudaFuncSetAttribute(inflate_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kInflateSmem));
  counter->fastq_streams.push_back(s);
  cleanup.s = nullptr;
  *out = s;
  return SGC_OK;
}

void sgc_fastq_stream_destroy(sgc_fastq_stream* s) {
  if (!s) return;
  sgc::fastq_stream_release(s);
  delete s;
}

int sgc_fastq_stream_submit(sgc_fastq_stream* s, const uint8_t* gz, const uint64_t* block_begin, const uint32_t* block_isize,
                            uint32_t n_blocks) {
  return sgc_fastq_stream_submit_range(s, gz, block_begin, block_isize, n_blocks, 0, 0, 0);
}

int sgc_fastq_stream_submit_range(sgc_fastq_stream* s, const uint8_t* gz, const uint64_t* block_begin,
                                  const uint32_t* block_isize, uint32_t n_blocks, uint32_t head_skip, uint32_t tail_skip,
                                  int self_contained) {
  if (!s || !gz || !block_begin || !block_isize) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  if (s->failed || !s->counter) return set_error(SGC_ERR_INVALID_ARG, "the stream has failed or its counter is gone");
  if (n_blocks == 0) return SGC_OK;
  DeviceGuard guard(s->device);
  // where every block's text goes: the prefix sum of the ISIZE fields
  s->h_outoff.resize((size_t)n_blocks + 1);
  uint64_t n_text = 0;
  for (uint32_t i = 0; i < n_blocks; ++i) {
    if (block_begin[i + 1] <= block_begin[i]) return set_error(SGC_ERR_INVALID_ARG, "block offsets must increase");
    s->h_outoff[i] = kHeadroom + n_text;
    n_text += block_isize[i];
  }
  s->h_outoff[n_blocks] = kHeadroom + n_text;
  if (n_text >= (64ull << 30)) return set_error(SGC_ERR_BATCH_TOO_LARGE, "a wave of blocks must inflate to less than 64 GiB");
  if (s->read_len == 0 && n_text >= (1ull << 32) - 2 * kHeadroom)
    return set_error(SGC_ERR_BATCH_TOO_LARGE, "in variable-length mode a wave of blocks must inflate to less than 4 GiB");
  const uint64_t gz_bytes = block_begin[n_blocks] - block_begin[0];
  int rc = grow(&s->d_gz, &s->gz_cap, (size_t)gz_bytes + 64, s->pool, s->stream);  // the decoder prefetches up to 47 bytes past a block
  if (rc == SGC_OK) rc = grow(&s->d_text, &s->text_cap, (size_t)kHeadroom + n_text + 64, s->pool, s->stream);
  if (rc == SGC_OK) rc = grow(&s->d_begin, &s->begin_cap, (size_t)n_blocks + 1, s->pool, s->stream);
  if (rc == SGC_OK) rc = grow(&s->d_outoff, &s->outoff_cap, (size_t)n_blocks + 1, s->pool, s->stream);
  if (rc) return rc;
  SGC_CUDA_TRY(cudaMemcpyAsync(s->d_gz, gz + block_begin[0], gz_bytes, cudaMemcpyHostToDevice, s->stream));
  SGC_CUDA_TRY(cudaMemcpyAsync(s->d_begin, block_begin, ((size_t)n_blocks + 1) * 8, cudaMemcpyHostToDevice, s->stream));
  SGC_CUDA_TRY(cudaMemcpyAsync(s->d_outoff, s->h_outoff.data(), ((size_t)n_blocks + 1) * 8, cudaMemcpyHostToDevice, s->stream));
  inflate_blocks_kernel<<<(n_blocks + kInflateThreads - 1) / kInflateThreads, kInflateThreads, kInflateSmem, s->stream>>>(
      s->d_gz, s->d_begin, s->d_outoff, n_blocks, s->d_text, s->d_state);
  crc_blocks_kernel<<<(n_blocks + kCrcWarps - 1) / kCrcWarps, kCrcWarps * 32, 0, s->stream>>>(s->d_gz, s->d_begin, s->d_outoff,
                                                                                           n_blocks, s->d_text, s->d_state);
  SGC_CUDA_TRY(cudaGetLastError());
  if ((uint64_t)head_skip + tail_skip > n_text) return set_error(SGC_ERR_INVALID_ARG, "the skips exceed the wave's text");
  wave_begin_kernel<<<1, 1, 0, s->stream>>>(s->d_state, head_skip, self_contained ? 1u : 0u);
  rc = frame_and_count(s, n_text - tail_skip, s->blocks_total);
  if (rc) return rc;
  s->blocks_total += n_blocks;
  return SGC_OK;
}

int sgc_fastq_stream_finish(sgc_fastq_stream* s, uint64_t* n_records) {
  if (!s) return set_error(SGC_ERR_INVALID_ARG, "stream is NULL");
  if (s->failed || !s->counter) return set_error(SGC_ERR_INVALID_ARG, "the stream has failed or its counter is gone");
  DeviceGuard guard(s->device);
  StreamState st;
  SGC_CUDA_TRY(cudaMemcpyAsync(&st, s->d_state, sizeof st, cudaMemcpyDeviceToHost, s->stream));
  SGC_CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (st.tail) {
    // the file does not end in a newline: the last line still counts (as the host reader has it)
    if (!s->d_text) return set_error(SGC_ERR_INVALID_ARG, "no text");
    SGC_CUDA_TRY(cudaMemsetAsync(s->d_text + kHeadroom, '\n', 1, s->stream));
    int rc = frame_and_count(s, 1, s->blocks_total);
    if (rc) return rc;
    SGC_CUDA_TRY(cudaMemcpyAsync(&st, s->d_state, sizeof st, cudaMemcpyDeviceToHost, s->stream));
    SGC_CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (st.tail) {
      s->failed = true;
      return set_error(SGC_ERR_FASTQ_FORMAT, "truncated FASTQ record at the end of the input");
    }
  }
  if (n_records) *n_records = s->records_total;
  return SGC_OK;
}

}  // extern "C"
