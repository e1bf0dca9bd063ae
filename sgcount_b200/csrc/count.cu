// count.cu — the per-read match-and-count loop (kernel K3) and the sgc_counter entry points.
//
// Replaces /root/reference/src/counter.rs:36-236 (Counter::new/count/assign/bounds/trim_*).
//
// Two kernels share one decision procedure (common.cuh assign_span):
//   count_staged_kernel  : fixed-stride sequence lines.  Persistent CTAs stream tiles of reads
//                          HBM -> shared memory with 1-D bulk async copies (TMA engine,
//                          cp.async.bulk + mbarrier complete_tx) through a multi-stage ring;
//                          each thread lifts the 4-byte words that cover its read's guide
//                          span out of shared memory, the stage is handed back to the copy
//                          engine at once, and the table probe + count atomics run from
//                          registers while the next tiles are in flight.
//   count_generic_kernel : any layout (variable-length lines via u32 offsets, unaligned
//                          buffers, tile remainders); one thread per read, byte loads.
// Per-guide counts are 64-bit atomics in the L2-resident state vector; matched reads are
// accumulated in registers and flushed once per warp.
#include <algorithm>
#include <vector>

#include "internal.h"

namespace sgc {
namespace {

struct CountParams {
  TableView table;
  const uint8_t* lines;
  const uint32_t* line_off;  // NULL => fixed stride
  uint64_t n_reads;
  uint64_t first_read;       // index of the first read this launch handles
  uint32_t stride, read_len;
  int offset;
  uint8_t with_perm, reverse, recursion, wild_byte;
  unsigned long long* state;  // counts[n_guides], total, matched
  uint32_t n_guides;
  int32_t* assign_out;
};

__device__ __forceinline__ void record_hit(const CountParams& p, int32_t hit, uint64_t read_idx, uint32_t& matched) {
  if (p.assign_out) p.assign_out[read_idx] = hit;
  if (hit >= 0) {
    ++matched;
    atomicAdd(p.state + hit, 1ull);  // counter.rs:232-235
  }
}

__device__ __forceinline__ void flush_matched(const CountParams& p, uint32_t matched) {
  matched = __reduce_add_sync(0xffffffffu, matched);
  if ((threadIdx.x & 31) == 0 && matched) atomicAdd(p.state + p.n_guides + 1, (unsigned long long)matched);
  // total_reads counts every record this launch walked (counter.rs:223-226)
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.state + p.n_guides, (unsigned long long)p.n_reads);
}

// Oriented span geometry of a read of length n: the bases a Centered/Plus/Minus window can
// touch are oriented positions [max(offset,1)-1, min(offset+k+1, n)).
struct SpanGeom {
  int base;  // oriented position of span base 0
  int m;     // number of bases
  int src;   // position in the read (as stored) of the first span byte
};
__device__ __forceinline__ SpanGeom span_geom(int n, int offset, int k, bool reverse) {
  SpanGeom g;
  g.base = offset > 0 ? offset - 1 : 0;
  int end = min(offset + k + 1, n);
  g.m = max(end - g.base, 0);
  g.src = reverse ? n - end : g.base;  // revcomp(r)[a:b] == comp(reverse(r[n-b : n-a]))
  return g;
}

__device__ __forceinline__ void orient(Span& sp, int m, bool reverse) {
  if (reverse && m > 0) {
    sp.codes = revcomp_codes(sp.codes, m);
    sp.bad = reverse_bits(sp.bad, m);
    sp.wild = reverse_bits(sp.wild, m);
  }
}

// ------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_generic_kernel(CountParams p) {
  uint32_t matched = 0;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_reads; i += nthreads) {
    const uint64_t r = p.first_read + i;
    uint64_t start;
    int n;
    if (p.line_off) {
      start = p.line_off[r];
      n = (int)(p.line_off[r + 1] - p.line_off[r]) - 1;
    } else {
      start = r * p.stride;
      n = (int)p.read_len;
    }
    const SpanGeom g = span_geom(n, p.offset, (int)p.table.k, p.reverse);
    Span sp{0, 0, 0};
    const uint8_t* s = p.lines + start + g.src;
    for (int j = 0; j < g.m; ++j) {
      uint8_t c = s[j];
      sp.codes |= (uint64_t)code_of(c) << (2 * j);
      if (!is_acgt(c)) {
        sp.bad |= 1u << j;
        if (c == p.wild_byte) sp.wild |= 1u << j;
      }
    }
    orient(sp, g.m, p.reverse);
    int32_t hit = assign_span(p.table, p.with_perm, sp, g.base, n, p.offset, p.recursion, nullptr);
    record_hit(p, hit, r, matched);
  }
  flush_matched(p, matched);
}

// ------------------------------------------------------------------------------------------
// staged kernel
// ------------------------------------------------------------------------------------------
constexpr int kTileReads = 256;  // one read per thread per tile
constexpr int kStages = 3;
constexpr int kSpanWords = 9;    // words covering a <= 32 byte span at any alignment

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// pack up to 32 span bytes held in aligned words aw[0..7]
__device__ __forceinline__ void pack_words(const uint32_t (&aw)[8], int m, uint8_t wild_byte, Span& sp) {
  uint64_t codes = 0;
  uint32_t bad = 0, wild = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (4 * i < m) {
      const uint32_t w = aw[i];
      const uint32_t c = (w >> 1) & 0x03030303u;            // code of each byte
      codes |= (uint64_t)((c * 0x01041040u) >> 24) << (8 * i);  // gather 4 x 2 bits
      // rebuild the ASCII each code stands for (A 41, C 43, T 54, G 47) and compare
      const uint32_t b0 = c & 0x01010101u, b1 = (c >> 1) & 0x01010101u;
      const uint32_t expect = 0x41414141u + 2u * b0 + 0x13u * b1 - 0x0Fu * (b0 & b1);
      const uint32_t x = w ^ expect;
      if (x) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((x >> (8 * j)) & 0xFFu) {
            bad |= 1u << (4 * i + j);
            if (((w >> (8 * j)) & 0xFFu) == wild_byte) wild |= 1u << (4 * i + j);
          }
        }
      }
    }
  }
  const uint32_t mmask = m >= 32 ? ~0u : ((1u << m) - 1);
  sp.codes = codes;
  sp.bad = bad & mmask;
  sp.wild = wild & mmask;
}

__global__ void __launch_bounds__(kTileReads) count_staged_kernel(CountParams p, uint64_t n_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages];

  const uint32_t tile_bytes = kTileReads * p.stride;            // multiple of 16
  const uint32_t stage_bytes = (tile_bytes + 16 + 127) & ~127u;  // +16: the last span may read past the tile
  const int tid = threadIdx.x;
  const int k = (int)p.table.k;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  uint64_t policy = 0;
  const uint64_t first_tile = blockIdx.x;
  const uint64_t tile_step = gridDim.x;
  if (tid == 0) {
    policy = l2_evict_first_policy();
    for (int s = 0; s < kStages; ++s) {
      uint64_t t = first_tile + (uint64_t)s * tile_step;
      if (t < n_tiles) {
        mbar_expect_tx(&full_bar[s], tile_bytes);
        bulk_load(smem + (size_t)s * stage_bytes, p.lines + (p.first_read + t * kTileReads) * p.stride, tile_bytes,
                  &full_bar[s], policy);
      }
    }
  }

  // every read has the same length, so the span geometry is uniform
  const int n = (int)p.read_len;
  const SpanGeom g = span_geom(n, p.offset, k, p.reverse);
  const uint32_t byte0 = (uint32_t)tid * p.stride + (uint32_t)g.src;  // tile-local address of the span
  const uint32_t word0 = byte0 >> 2;
  const uint32_t shift = (byte0 & 3u) * 8;
  const int n_words = (int)(((byte0 & 3u) + (uint32_t)g.m + 3u) >> 2);

  uint32_t matched = 0;
  uint32_t it = 0;
  for (uint64_t t = first_tile; t < n_tiles; t += tile_step, ++it) {
    const int s = it % kStages;
    const uint32_t parity = (it / kStages) & 1u;
    mbar_wait(&full_bar[s], parity);

    const uint32_t* tile = reinterpret_cast<const uint32_t*>(smem + (size_t)s * stage_bytes);
    uint32_t raw[kSpanWords];
#pragma unroll
    for (int i = 0; i < kSpanWords; ++i) raw[i] = (i < n_words) ? tile[word0 + i] : 0u;

    __syncthreads();  // all spans are in registers: the stage can be refilled
    if (tid == 0) {
      uint64_t nt = t + (uint64_t)kStages * tile_step;
      if (nt < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full_bar[s], tile_bytes);
        bulk_load(smem + (size_t)s * stage_bytes, p.lines + (p.first_read + nt * kTileReads) * p.stride, tile_bytes,
                  &full_bar[s], policy);
      }
    }

    uint32_t aw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[i] = __funnelshift_r(raw[i], raw[i + 1], shift);
    Span sp;
    pack_words(aw, g.m, p.wild_byte, sp);
    orient(sp, g.m, p.reverse);
    int32_t hit = assign_span(p.table, p.with_perm, sp, g.base, n, p.offset, p.recursion, nullptr);
    record_hit(p, hit, p.first_read + t * kTileReads + tid, matched);
  }
  flush_matched(p, matched);
}

}  // namespace
}  // namespace sgc

// ------------------------------------------------------------------------------------------
// sgc_counter
// ------------------------------------------------------------------------------------------
struct sgc_counter {
  const sgc_library* lib = nullptr;
  int is_reverse = 0;
  uint32_t offset = 0;
  int recursion = 1;
  int rc_mode = SGC_RC_BITTRICK;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  unsigned long long* d_state = nullptr;
  bool own_state = false;
  // host-batch staging (sgc_counter_submit)
  cudaStream_t copy_stream = nullptr;
  uint8_t* d_stage[2] = {nullptr, nullptr};
  uint32_t* d_stage_off[2] = {nullptr, nullptr};
  size_t stage_cap = 0, stage_off_cap = 0;
  cudaEvent_t copy_done[2] = {nullptr, nullptr}, kernel_done[2] = {nullptr, nullptr};
  uint64_t chunks_submitted = 0;
  sgc_launch_info last{};
  int staged_blocks_per_sm = 0;
};

using namespace sgc;

namespace {

constexpr size_t kChunkBytes = 64ull << 20;

uint8_t wild_byte_for(const sgc_counter* c) {
  // the byte that reads as 'N' to the lookup: under the fxread bit trick a reverse-complemented
  // 'J' becomes 'N' and 'N' becomes 'J' 
  return (c->is_reverse && c->rc_mode == SGC_RC_BITTRICK) ? (uint8_t)'J' : (uint8_t)'N';
}

CountParams make_params(const sgc_counter* c, const uint8_t* d_lines, const uint32_t* d_off, uint32_t stride,
                        uint32_t read_len, int32_t* d_assign) {
  CountParams p{};
  p.table = c->lib->view();
  p.lines = d_lines;
  p.line_off = d_off;
  p.stride = stride;
  p.read_len = read_len;
  p.offset = (int)c->offset;
  p.with_perm = c->lib->with_perm;
  p.reverse = c->is_reverse != 0;
  p.recursion = c->recursion != 0;
  p.wild_byte = wild_byte_for(c);
  p.state = c->d_state;
  p.n_guides = c->lib->n;
  p.assign_out = d_assign;
  return p;
}

// Enqueue the kernels for one device-resident batch.
int launch_count(sgc_counter* c, const uint8_t* d_lines, uint64_t n_bytes, const uint32_t* d_off, uint32_t stride,
                 uint32_t read_len, uint64_t n_reads, int32_t* d_assign, cudaStream_t stream) {
  if (n_reads == 0) return SGC_OK;
  CountParams p = make_params(c, d_lines, d_off, stride, read_len, d_assign);
  uint64_t done = 0;
  const uint64_t launches_before = c->last.launches_total;
  c->last = sgc_launch_info{};
  c->last.launches_total = launches_before;
  c->last.kernel = 1;
  const bool stageable = d_off == nullptr && ((uintptr_t)d_lines & 15u) == 0 && stride >= read_len &&
                         (uint64_t)kTileReads * stride + 16 <= 48 * 1024;
  if (stageable && n_reads >= (uint64_t)kTileReads && n_bytes >= (uint64_t)kTileReads * stride) {
    // whole tiles only, and never a bulk copy that would run past n_bytes
    const uint64_t n_tiles = std::min(n_reads / kTileReads, n_bytes / ((uint64_t)kTileReads * stride));
    const uint32_t tile_bytes = kTileReads * stride;
    const uint32_t stage_bytes = (tile_bytes + 16 + 127) & ~127u;
    const size_t smem = (size_t)kStages * stage_bytes;
    if (c->staged_blocks_per_sm == 0) {
      SGC_CUDA_TRY(cudaFuncSetAttribute(count_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    int per_sm = 0;
    SGC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_staged_kernel, kTileReads, smem));
    if (per_sm < 1) return set_error(SGC_ERR_CUDA, "staged kernel does not fit on an SM");
    c->staged_blocks_per_sm = per_sm;
    uint64_t grid = (uint64_t)c->lib->sm_count * per_sm;  // persistent: every CTA resident
    if (grid > n_tiles) grid = n_tiles;
    p.n_reads = n_tiles * kTileReads;
    p.first_read = 0;
    count_staged_kernel<<<(unsigned)grid, kTileReads, smem, stream>>>(p, n_tiles);
    SGC_CUDA_TRY(cudaGetLastError());
    done = n_tiles * kTileReads;
    c->last.grid = (uint32_t)grid;
    c->last.block = kTileReads;
    c->last.smem_bytes = (uint32_t)smem;
    c->last.kernel = 0;
    c->last.launches_total += 1;
  }
  if (done < n_reads) {
    p.first_read = done;
    p.n_reads = n_reads - done;
    uint64_t blocks = (p.n_reads + 255) / 256;
    uint64_t cap = (uint64_t)c->lib->sm_count * 8;
    if (blocks > cap) blocks = cap;
    count_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
    SGC_CUDA_TRY(cudaGetLastError());
    if (done == 0) {
      c->last.grid = (uint32_t)blocks;
      c->last.block = 256;
    }
    c->last.launches_total += 1;
  }
  return SGC_OK;
}

int check_batch(const sgc_counter* c, const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off,
                uint32_t stride, uint32_t read_len, uint64_t n_reads) {
  if (!c) return set_error(SGC_ERR_INVALID_ARG, "counter is NULL");
  if (n_reads == 0) return SGC_OK;
  if (!lines) return set_error(SGC_ERR_INVALID_ARG, "lines is NULL");
  if (line_off) {
    if (n_bytes >= (1ull << 32)) return set_error(SGC_ERR_BATCH_TOO_LARGE, "variable-length batch must stay below 4 GiB");
  } else {
    if (stride == 0 || read_len > stride) return set_error(SGC_ERR_INVALID_ARG, "need 0 < read_len <= stride");
    if ((n_reads - 1) * (uint64_t)stride + read_len > n_bytes)
      return set_error(SGC_ERR_INVALID_ARG, "fixed-stride batch does not fit n_bytes");
  }
  return SGC_OK;
}

}  // namespace

extern "C" {

int sgc_counter_create(const sgc_library* lib, int is_reverse, uint32_t offset, int position_recursion, int rc_mode,
                       void* stream, uint64_t* d_state, sgc_counter** out) {
  if (!lib || !out) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  if (rc_mode != SGC_RC_BITTRICK && rc_mode != SGC_RC_KEEP_N) return set_error(SGC_ERR_INVALID_ARG, "bad rc_mode");
  if (offset > 0x3FFFFFFFu) return set_error(SGC_ERR_INVALID_ARG, "offset too large");
  DeviceGuard guard(lib->device);
  if (!guard.ok()) return set_error(SGC_ERR_CUDA, "cudaSetDevice failed");
  sgc_counter* c = new sgc_counter();
  c->lib = lib;
  c->is_reverse = is_reverse != 0;
  c->offset = offset;
  c->recursion = position_recursion != 0;
  c->rc_mode = rc_mode;
  struct Cleanup {
    sgc_counter* c;
    ~Cleanup() {
      if (c) sgc_counter_destroy(c);
    }
  } cleanup{c};
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    SGC_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  const size_t words = (size_t)lib->n + 2;
  if (d_state) {
    c->d_state = reinterpret_cast<unsigned long long*>(d_state);
  } else {
    SGC_CUDA_TRY(cudaMalloc(&c->d_state, words * sizeof(uint64_t)));
    c->own_state = true;
  }
  SGC_CUDA_TRY(cudaMemsetAsync(c->d_state, 0, words * sizeof(uint64_t), c->stream));
  cleanup.c = nullptr;
  *out = c;
  return SGC_OK;
}

void sgc_counter_destroy(sgc_counter* c) {
  if (!c) return;
  DeviceGuard guard(c->lib->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->d_stage[i]);
    cudaFree(c->d_stage_off[i]);
    if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
    if (c->kernel_done[i]) cudaEventDestroy(c->kernel_done[i]);
  }
  if (c->own_state) cudaFree(c->d_state);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int sgc_counter_submit_device(sgc_counter* c, const uint8_t* d_lines, uint64_t n_bytes, const uint32_t* d_line_off,
                              uint32_t stride, uint32_t read_len, uint64_t n_reads, int32_t* d_assign_out) {
  int rc = check_batch(c, d_lines, n_bytes, d_line_off, stride, read_len, n_reads);
  if (rc) return rc;
  DeviceGuard guard(c->lib->device);
  return launch_count(c, d_lines, n_bytes, d_line_off, stride, read_len, n_reads, d_assign_out, c->stream);
}

int sgc_counter_submit(sgc_counter* c, const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off,
                       uint32_t stride, uint32_t read_len, uint64_t n_reads) {
  int rc = check_batch(c, lines, n_bytes, line_off, stride, read_len, n_reads);
  if (rc) return rc;
  if (n_reads == 0) return SGC_OK;
  if (line_off && line_off[n_reads] > n_bytes) return set_error(SGC_ERR_INVALID_ARG, "line offsets exceed n_bytes");
  DeviceGuard guard(c->lib->device);
  if (!c->copy_stream) {
    SGC_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SGC_CUDA_TRY(cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming));
      SGC_CUDA_TRY(cudaEventCreateWithFlags(&c->kernel_done[i], cudaEventDisableTiming));
    }
  }
  // chunk geometry: whole tiles of reads, about kChunkBytes each
  uint64_t r0 = 0;
  while (r0 < n_reads) {
    uint64_t r1;
    if (line_off) {
      // advance until the chunk holds about kChunkBytes
      uint64_t lo = r0, hi = n_reads;
      const uint64_t limit = (uint64_t)line_off[r0] + kChunkBytes;
      while (lo < hi) {  // last r with line_off[r] <= limit
        uint64_t mid = (lo + hi + 1) / 2;
        if (line_off[mid] <= limit) lo = mid; else hi = mid - 1;
      }
      r1 = lo > r0 ? lo : r0 + 1;
    } else {
      uint64_t per = std::max<uint64_t>(kTileReads, (kChunkBytes / stride) / kTileReads * kTileReads);
      r1 = std::min(n_reads, r0 + per);
    }
    const int b = (int)(c->chunks_submitted & 1);
    const uint64_t byte0 = line_off ? line_off[r0] : r0 * stride;
    const uint64_t byte1 = line_off ? line_off[r1] : std::min<uint64_t>(n_bytes, r1 * stride);
    const size_t bytes = byte1 - byte0;
    // (re)size this staging buffer; +64 so bulk copies of the last tile stay inside the allocation
    if (bytes + 64 > c->stage_cap) {
      SGC_CUDA_TRY(cudaStreamSynchronize(c->stream));
      SGC_CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
      size_t cap = std::max<size_t>(bytes + 64, std::min<size_t>(kChunkBytes + (1 << 20), n_bytes + 64));
      for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_stage[i]);
        c->d_stage[i] = nullptr;
        SGC_CUDA_TRY(cudaMalloc(&c->d_stage[i], cap));
      }
      c->stage_cap = cap;
    }
    const size_t n_off = line_off ? (r1 - r0 + 1) : 0;
    if (n_off > c->stage_off_cap) {
      SGC_CUDA_TRY(cudaStreamSynchronize(c->stream));
      SGC_CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
      size_t cap = n_off * 2;
      for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_stage_off[i]);
        c->d_stage_off[i] = nullptr;
        SGC_CUDA_TRY(cudaMalloc(&c->d_stage_off[i], cap * sizeof(uint32_t)));
      }
      c->stage_off_cap = cap;
    }
    // copy stream: wait until the kernel that last read this buffer is done, then copy
    if (c->chunks_submitted >= 2) SGC_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->kernel_done[b], 0));
    SGC_CUDA_TRY(cudaMemcpyAsync(c->d_stage[b], lines + byte0, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    if (line_off)
      SGC_CUDA_TRY(cudaMemcpyAsync(c->d_stage_off[b], line_off + r0, n_off * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                   c->copy_stream));
    SGC_CUDA_TRY(cudaEventRecord(c->copy_done[b], c->copy_stream));
    SGC_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->copy_done[b], 0));
    // offsets stay relative to the batch start: shift the line base instead of rewriting them
    const uint8_t* d_lines = line_off ? c->d_stage[b] - byte0 : c->d_stage[b];
    rc = launch_count(c, d_lines, bytes, line_off ? c->d_stage_off[b] : nullptr, stride, read_len, r1 - r0, nullptr,
                      c->stream);
    if (rc) return rc;
    SGC_CUDA_TRY(cudaEventRecord(c->kernel_done[b], c->stream));
    c->chunks_submitted += 1;
    r0 = r1;
  }
  return SGC_OK;
}

int sgc_counter_sync(sgc_counter* c) {
  if (!c) return set_error(SGC_ERR_INVALID_ARG, "counter is NULL");
  DeviceGuard guard(c->lib->device);
  SGC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return SGC_OK;
}

int sgc_counter_reset(sgc_counter* c) {
  if (!c) return set_error(SGC_ERR_INVALID_ARG, "counter is NULL");
  DeviceGuard guard(c->lib->device);
  SGC_CUDA_TRY(cudaMemsetAsync(c->d_state, 0, ((size_t)c->lib->n + 2) * sizeof(uint64_t), c->stream));
  return SGC_OK;
}

int sgc_counter_finish(sgc_counter* c, uint64_t* counts, uint64_t* total, uint64_t* matched) {
  if (!c) return set_error(SGC_ERR_INVALID_ARG, "counter is NULL");
  DeviceGuard guard(c->lib->device);
  const size_t n = c->lib->n;
  std::vector<uint64_t> host(n + 2);
  SGC_CUDA_TRY(cudaMemcpyAsync(host.data(), c->d_state, (n + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  SGC_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (counts) std::copy(host.begin(), host.begin() + n, counts);
  if (total) *total = host[n];
  if (matched) *matched = host[n + 1];
  return SGC_OK;
}

int sgc_counter_state(sgc_counter* c, uint64_t** d_state, uint64_t* n_words) {
  if (!c || !d_state || !n_words) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  *d_state = reinterpret_cast<uint64_t*>(c->d_state);
  *n_words = (uint64_t)c->lib->n + 2;
  return SGC_OK;
}

int sgc_counter_launch_info(const sgc_counter* c, sgc_launch_info* out) {
  if (!c || !out) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  *out = c->last;
  return SGC_OK;
}

}  // extern "C"
