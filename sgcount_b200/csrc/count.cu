// count.cu — the per-read match-and-count loop (kernel K3) and the sgc_counter entry points.
//
// Replaces /root/reference/src/counter.rs:36-236 (Counter::new/count/assign/bounds/trim_*).
//
// Two kernels share one decision procedure (common.cuh assign_span):
//   count_stream_kernel  : fixed-stride sequence lines.  Persistent warps stream tiles of reads
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
#include <cstdlib>
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
  // reads the streaming kernel could not settle, as structure-of-arrays records of park_cap
  // entries each: window words [NW], misc, read index (see WarpQueueT)
  uint32_t* park_rec;
  unsigned int* park_count;
  uint32_t park_cap;
  uint32_t debug;  // SGC_DEBUG bit mask (tuning only): 1 no count atomics, 2 no table probe, 4 no slow path
};

__device__ __forceinline__ void record_hit(const CountParams& p, int32_t hit, uint64_t read_idx, uint32_t& matched) {
  if (p.assign_out) p.assign_out[read_idx] = hit;
  if (hit >= 0) {
    ++matched;
    if (!(p.debug & 1u)) atomicAdd(p.state + hit, 1ull);  // counter.rs:232-235
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
__host__ __device__ __forceinline__ SpanGeom span_geom(int n, int offset, int k, bool reverse) {
  SpanGeom g;
  g.base = offset > 0 ? offset - 1 : 0;
  int end = offset + k + 1 < n ? offset + k + 1 : n;
  g.m = end > g.base ? end - g.base : 0;
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
// staged kernel: warp-private streaming rings
//
// Every warp owns a ring of kStages shared-memory buffers of ONE warp tile (32 reads =
// 32*stride bytes, always a multiple of 16) and one mbarrier per buffer; lane 0 keeps the
// ring full with 1-D bulk async copies, so warps never synchronise with each other and each
// SM keeps (warps x (kStages-1)) tiles in flight.  Per read the hot path is
//   LDS the span words -> funnel-align -> 2-bit pack + ASCII round-trip validity check ->
//   one 32-byte table probe for the Centered window -> RED.64 on the guide's counter.
// Reads the Centered probe does not settle (mismatch beyond the table, N, other bytes; about
// 10-20 %) are parked in a warp-private shared-memory queue and walked 32 at a time through
// the full Counter::assign procedure, so the rare path runs with full warps instead of
// dragging every warp through it.
// ------------------------------------------------------------------------------------------
constexpr int kWarpReads = 32;
constexpr int kMaxStages = 4;
constexpr int kQueueCap = 64;  // <= 31 parked + 32 new

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
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
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

// 2-bit pack NW window-aligned words (4 bases each).  `xs[i]` receives each word's XOR against
// the ASCII its codes stand for (A 41, C 43, T 54, G 47), masked to the bytes that belong to
// the window: all zero iff every window byte is A/C/G/T.  Returns the OR of xs.
template <int NW>
__device__ __forceinline__ uint32_t pack_window(const uint32_t (&aw)[NW], uint32_t last_mask, int n_words,
                                                uint64_t& codes, uint32_t (&xs)[NW]) {
  uint32_t lo = 0, hi = 0, any = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    const uint32_t w = aw[i];
    const uint32_t c = (w >> 1) & 0x03030303u;
    const uint32_t packed = (c * 0x01041040u) >> 24;  // gather 4 x 2 bits
    if (i < 4) lo |= packed << (8 * i); else hi |= packed << (8 * (i - 4));
    const uint32_t b0 = c & 0x01010101u, b1 = (c >> 1) & 0x01010101u;
    const uint32_t expect = 0x41414141u + 2u * b0 + 0x13u * b1 - 0x0Fu * (b0 & b1);
    // NW is the compile-time bound; words at or past n_words hold no window byte
    const uint32_t m = (NW == 5 || i < n_words - 1) ? ((i == NW - 1 && NW == 5) ? last_mask : ~0u)
                                                    : (i == n_words - 1 ? last_mask : 0u);
    xs[i] = (w ^ expect) & m;
    any |= xs[i];
  }
  codes = ((uint64_t)hi << 32) | lo;
  return any;
}

// 4-bit mask of the non-zero bytes of x
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {
  const uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
  return ((nz >> 7) * 0x10204080u) >> 28;
}

// Parked reads keep the raw window words; everything else is re-derived when a full warp of
// them is drained, so parking costs a handful of shared-memory stores.
// Layout per warp (words): w[NW][kQueueCap] | misc[kQueueCap] | read[kQueueCap]
//   misc = before | after << 8 | flags << 16   (flag 1: the Centered probe was a definite miss)
template <int NW>
struct WarpQueueT {
  uint32_t w[NW][kQueueCap];
  uint32_t misc[kQueueCap];
  uint32_t read[kQueueCap];  // read index relative to the launch's first read
};

// Everything a parked read needs, uniform across the warp.
struct WalkGeom {
  int k, n, o;
  bool reverse, recursion, with_perm, has_before, has_after;
  uint32_t last_mask, wild_byte;
  int n_words;
};

__device__ __forceinline__ WalkGeom make_geom(const CountParams& p) {
  WalkGeom g;
  g.k = (int)p.table.k;
  g.n = (int)p.read_len;
  g.o = p.offset;
  g.reverse = p.reverse;
  g.recursion = p.recursion;
  g.with_perm = p.with_perm;
  // stored (as-read) coordinates of the Centered window: forward [o, o+k), reverse [n-o-k, n-o)
  const int win_src = g.reverse ? g.n - g.o - g.k : g.o;
  g.has_before = win_src > 0;         // the bytes just outside the window are needed
  g.has_after = win_src + g.k < g.n;  // only by the Plus / Minus positions
  g.n_words = (g.k + 3) >> 2;
  g.last_mask = (g.k & 3) ? ((1u << (8 * (g.k & 3))) - 1) : ~0u;
  g.wild_byte = p.wild_byte;
  return g;
}

// Full Counter::assign (counter.rs:96-140) for one parked read.
template <int NW, bool WIDE>
__device__ __forceinline__ int32_t walk_parked(const TableView& t, const WalkGeom& g, const uint32_t (&aw)[NW],
                                               uint32_t misc) {
  uint32_t xs[NW];
  uint64_t codes;
  const uint32_t any = pack_window<NW>(aw, g.last_mask, g.n_words, codes, xs);
  const uint64_t kmask = (1ull << (2 * g.k)) - 1;
  codes &= kmask;
  if (g.reverse) codes = revcomp_codes(codes, g.k);
  const uint32_t b_before = misc & 0xFFu, b_after = (misc >> 8) & 0xFFu;
  const int first_pos = (misc >> 16) & 1u;
  // oriented neighbours: reverse swaps and complements them
  const uint32_t prev = g.reverse ? b_after : b_before, next = g.reverse ? b_before : b_after;
  const bool has_prev = g.reverse ? g.has_after : g.has_before, has_next = g.reverse ? g.has_before : g.has_after;
  const uint32_t cflip = g.reverse ? 2u : 0u;
  const bool prev_bad = !has_prev || !is_acgt((uint8_t)prev), next_bad = !has_next || !is_acgt((uint8_t)next);
  Span sp;
  sp.codes = (uint64_t)(code_of((uint8_t)prev) ^ cflip) | (codes << 2) |
             ((uint64_t)(code_of((uint8_t)next) ^ cflip) << (2 * g.k + 2));
  if (any == 0 && !(has_prev && prev_bad) && !(has_next && next_bad) && t.bloom != nullptr) {
    // Clean bases (the common parked read).  Everything the walk may consult that lives in
    // L2 -- the Bloom words of the three windows and the front-table buckets of Plus and
    // Minus -- is fetched in one go; the main table is touched only for keys the filter
    // cannot rule out.  Resolution keeps the reference's order: Centered, Plus, Minus, a
    // library member before a variant at each (counter.rs:111-135); a window whose trim
    // would fail is never consulted.
    const bool try_c = first_pos == 0;
    const bool front_c_open = (misc >> 17) & 1u;  // the streaming probe could not decide membership
    const bool try_p = g.recursion && g.o + 1 + g.k <= g.n;
    const bool try_m = try_p && g.o > 0;
    const uint64_t key_c = (sp.codes >> 2) & kmask, key_p = (sp.codes >> 4) & kmask, key_m = sp.codes & kmask;
    uint64_t fp[4], fm[4], bc = 0, bp = 0, bm = 0, mc, mp, mm, meta;
    uint32_t wc, wp, wm;
    bloom_locate(key_c, t.n_bloom_words, wc, mc);
    bloom_locate(key_p, t.n_bloom_words, wp, mp);
    bloom_locate(key_m, t.n_bloom_words, wm, mm);
    if (try_c) bc = __ldg(t.bloom + wc);
    if (try_p) {
      load_bucket(t.front_slots + (size_t)bucket_of(key_p, t.front_buckets) * 4, fp);
      bp = __ldg(t.bloom + wp);
    }
    if (try_m) {
      load_bucket(t.front_slots + (size_t)bucket_of(key_m, t.front_buckets) * 4, fm);
      bm = __ldg(t.bloom + wm);
    }
    if (try_c) {
      if (front_c_open) {
        const int32_t hit = meta_hit(table_find_t<WIDE>(t.front_slots, t.front_buckets, key_c));
        if (hit != kMiss) return hit;
      }
      if ((bc & mc) == mc) {
        const int32_t hit = meta_hit(table_find_t<WIDE>(t.slots, t.n_buckets, key_c));
        if (hit != kMiss) return hit;
      }
    }
    if (!try_p) return kMiss;  // no recursion, or the Plus trim fails: return (counter.rs:105-108)
    // one window: member in the front table, else (filter permitting) the main table
    auto settle = [&](const uint64_t (&f)[4], bool maybe, uint64_t key) -> int32_t {
      const int r = bucket_match<WIDE>(f, key, meta);
      if (r == kFound) return meta_hit(meta);
      if (!maybe) return kMiss;
      return meta_hit(table_find_t<WIDE>(t.slots, t.n_buckets, key));
    };
    const int32_t hit_p = settle(fp, (bp & mp) == mp, key_p);
    if (hit_p != kMiss) return hit_p;
    if (!try_m) return kMiss;  // checked_sub(1) -> None
    return settle(fm, (bm & mm) == mm, key_m);
  }
  uint32_t badw = 0, wildw = 0;
  if (any) {
    const uint32_t wild4 = 0x01010101u * g.wild_byte;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      badw |= nonzero_bytes(xs[i]) << (4 * i);
      wildw |= (nonzero_bytes(aw[i] ^ wild4) ^ 0xFu) << (4 * i);
    }
    wildw &= badw;
    if (g.reverse) {
      badw = reverse_bits(badw, g.k);
      wildw = reverse_bits(wildw, g.k);
    }
  }
  sp.bad = (prev_bad ? 1u : 0u) | (badw << 1) | ((next_bad ? 1u : 0u) << (g.k + 1));
  sp.wild = ((has_prev && prev == g.wild_byte) ? 1u : 0u) | (wildw << 1) |
            (((has_next && next == g.wild_byte) ? 1u : 0u) << (g.k + 1));
  return assign_span(t, g.with_perm, sp, g.o - 1, g.n, g.o, g.recursion, nullptr, first_pos);
}

// Pass 1.  NW = words that hold the k window bytes: 5 (k = 17..20, the common guide lengths) or 8.
template <int NW, bool WIDE>
__global__ void __launch_bounds__(384, 2) count_stream_kernel(const CountParams p, uint64_t n_wtiles, int n_stages,
                                                              uint32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  // shared layout: [warp][stage] tile buffers | [warp][stage] mbarriers | [warp] queues
  uint8_t* my_tiles = smem + (size_t)warp * n_stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps_per_cta * n_stages * stage_bytes);
  uint64_t* my_bar = bars + warp * kMaxStages;
  WarpQueueT<NW>* q = reinterpret_cast<WarpQueueT<NW>*>(bars + warps_per_cta * kMaxStages) + warp;

  const uint32_t tile_bytes = kWarpReads * p.stride;  // multiple of 16
  const uint64_t gwarp = (uint64_t)blockIdx.x * warps_per_cta + warp;
  const uint64_t gwarps = (uint64_t)gridDim.x * warps_per_cta;
  const uint64_t src_step = gwarps * tile_bytes;
  const uint8_t* next_src = p.lines + gwarp * tile_bytes;  // source of the next tile to request

  uint64_t policy = 0;
  uint64_t requested = gwarp;  // tile index of the next request
  if (lane == 0) {
    for (int s = 0; s < n_stages; ++s) mbar_init(&my_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    policy = l2_evict_first_policy();
    for (int s = 0; s < n_stages; ++s) {
      if (requested < n_wtiles) {
        mbar_expect_tx(&my_bar[s], tile_bytes);
        bulk_load(my_tiles + (size_t)s * stage_bytes, next_src, tile_bytes, &my_bar[s], policy);
        requested += gwarps;
        next_src += src_step;
      }
    }
  }
  __syncwarp();

  // every read has the same length n, so the geometry is uniform
  const WalkGeom g = make_geom(p);
  const int k = g.k;
  const bool centered_fits = g.o + k <= g.n;  // else every read fails its first trim (counter.rs:105-108)
  const int win_src = g.reverse ? g.n - g.o - k : g.o;
  const uint32_t wbyte = (uint32_t)lane * p.stride + (uint32_t)(centered_fits ? win_src : 0);
  const uint32_t word0 = wbyte >> 2;
  const uint32_t shift = (wbyte & 3u) * 8;
  const uint64_t kmask = (1ull << (2 * k)) - 1;
  const uint64_t* __restrict__ front = p.table.front_slots;
  const uint32_t front_buckets = p.table.front_buckets;
  const bool park_misses = g.recursion;
  const uint64_t table_policy = l2_evict_last_policy();

  uint32_t matched = 0;
  uint32_t qn = 0;  // parked reads (warp-uniform)

  // Parked reads leave the streaming loop: `count` queue entries starting at `from` are
  // appended to the global record buffer (coalesced, one reservation per flush) and walked
  // later by count_parked_kernel at full occupancy.
  auto flush = [&](uint32_t from, uint32_t count) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(p.park_count, count);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((uint32_t)lane < count) {
      const uint32_t e = from + lane;
      uint32_t* dst = p.park_rec + base + lane;
#pragma unroll
      for (int i = 0; i < NW; ++i) dst[(size_t)i * p.park_cap] = q->w[i][e];
      dst[(size_t)NW * p.park_cap] = q->misc[e];
      dst[(size_t)(NW + 1) * p.park_cap] = q->read[e];
    }
  };

  int s = 0;
  uint32_t parity = 0;
  uint32_t read_idx = (uint32_t)(gwarp * kWarpReads) + lane;
  const uint32_t read_step = (uint32_t)(gwarps * kWarpReads);
  for (uint64_t t = gwarp; t < n_wtiles; t += gwarps, read_idx += read_step) {
    // wait for the tile, lift this lane's window out of shared memory, hand the buffer back
    mbar_wait(&my_bar[s], parity);
    const uint8_t* tile8 = my_tiles + (size_t)s * stage_bytes;
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(tile8);
    uint32_t raw[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; ++i) raw[i] = tile[word0 + i];
    uint32_t misc = 0;
    if (g.has_before) misc = tile8[wbyte - 1];
    if (g.has_after) misc |= (uint32_t)tile8[wbyte + k] << 8;
    __syncwarp();  // the whole warp has its bytes in registers: refill this buffer
    if (lane == 0 && requested < n_wtiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&my_bar[s], tile_bytes);
      bulk_load(my_tiles + (size_t)s * stage_bytes, next_src, tile_bytes, &my_bar[s], policy);
      requested += gwarps;
      next_src += src_step;
    }
    if (++s == n_stages) {
      s = 0;
      parity ^= 1u;
    }
    if (!centered_fits) {  // every read fails its first trim: nothing is tried (counter.rs:105-108)
      if (p.assign_out) p.assign_out[read_idx] = kMiss;
      continue;
    }

    uint32_t aw[NW], xs[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) aw[i] = __funnelshift_r(raw[i], raw[i + 1], shift);
    uint64_t codes;
    const uint32_t any = pack_window<NW>(aw, g.last_mask, g.n_words, codes, xs);
    codes &= kmask;
    if (g.reverse) codes = revcomp_codes(codes, k);

    // Centered window against the FRONT table (library members only, L2 resident), one
    // bucket: settles every exact read whose guide sits in its home bucket.  Everything else
    // is parked for the full walk over the unified table.
    bool park = true;
    if (any == 0) {
      int32_t hit;
      int r;
      if (p.debug & 2u) {
        hit = (int32_t)(codes % p.n_guides);
        r = kFound;
      } else {
        uint64_t w[4], meta;
        load_bucket(front + (size_t)bucket_of(codes, front_buckets) * 4, w, table_policy);
        r = bucket_match<WIDE>(w, codes, meta);
        hit = meta_hit(meta);
      }
      if (r == kFound) {
        park = false;
        record_hit(p, hit, (uint64_t)read_idx, matched);
      } else if (r == kUndecided) {
        misc |= 1u << 17;  // the member may sit in a later bucket: the walk re-probes the front table
      } else if (!g.with_perm) {
        // no Permuter: the front table is the whole table, Centered is decided
        misc |= 1u << 16;
        park = park_misses;
        if (!park) record_hit(p, kMiss, (uint64_t)read_idx, matched);
      }
    }
    if (p.debug & 4u) park = false;
    const uint32_t pm = __ballot_sync(0xffffffffu, park);
    if (pm) {
      if (park) {
        const uint32_t e = qn + __popc(pm & ((1u << lane) - 1));
#pragma unroll
        for (int i = 0; i < NW; ++i) q->w[i][e] = aw[i];
        q->misc[e] = misc;
        q->read[e] = read_idx;
      }
      qn += __popc(pm);
      __syncwarp();
      if (qn >= 32) {
        qn -= 32;
        flush(qn, 32);
        __syncwarp();
      }
    }
  }
  if (qn) flush(0, qn);
  flush_matched(p, matched);
}

// Pass 2: one thread per parked read, the full Counter::assign walk (walk_parked).  Plain
// grid-stride kernel at full occupancy: the dependent table lookups of the rare reads are
// hidden by thread-level parallelism instead of stalling the streaming warps.
template <int NW, bool WIDE>
__global__ void __launch_bounds__(256) count_parked_kernel(CountParams p) {
  const WalkGeom g = make_geom(p);
  const unsigned int n = *p.park_count;
  uint32_t matched = 0;
  for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t* src = p.park_rec + e;
    uint32_t aw[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) aw[i] = src[(size_t)i * p.park_cap];
    const uint32_t misc = src[(size_t)NW * p.park_cap];
    const uint32_t read = src[(size_t)(NW + 1) * p.park_cap];
    const int32_t hit = walk_parked<NW, WIDE>(p.table, g, aw, misc);
    record_hit(p, hit, (uint64_t)read, matched);
  }
  matched = __reduce_add_sync(0xffffffffu, matched);
  if ((threadIdx.x & 31) == 0 && matched) atomicAdd(p.state + p.n_guides + 1, (unsigned long long)matched);
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
  unsigned long long* d_state = nullptr;
  bool own_state = false;
  // host-batch staging (sgc_counter_submit)
  cudaStream_t copy_stream = nullptr;
  uint8_t* d_stage[2] = {nullptr, nullptr};
  uint32_t* d_stage_off[2] = {nullptr, nullptr};
  size_t stage_cap = 0, stage_off_cap = 0;
  cudaEvent_t copy_done[2] = {nullptr, nullptr}, kernel_done[2] = {nullptr, nullptr};
  uint64_t chunks_submitted = 0;
  // parked-read records of the streaming kernel (pass 1 -> pass 2)
  uint32_t* d_park = nullptr;
  unsigned int* d_park_count = nullptr;
  size_t park_cap = 0;  // entries
  sgc_launch_info last{};
};

using namespace sgc;

namespace {

constexpr size_t kChunkBytes = 64ull << 20;
constexpr uint64_t kChunkAlignReads = 256;  // chunk starts stay 16-byte aligned for any stride

int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

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
  p.debug = (uint32_t)env_int("SGC_DEBUG", 0);
  return p;
}

struct StreamConfig {
  int warps = 0, stages = 0, ctas_per_sm = 0;
};
size_t stream_smem_bytes(const StreamConfig& c, uint32_t stage_bytes, size_t queue_bytes) {
  return (size_t)c.warps * ((size_t)c.stages * stage_bytes + kMaxStages * sizeof(uint64_t) + queue_bytes);
}
// Largest ring that fits: prefer 2 resident CTAs of 12 warps with >= 3 buffers per warp.
// SGC_WARPS / SGC_STAGES / SGC_CTAS override the choice (tuning only).
StreamConfig pick_stream_config(uint32_t stage_bytes, size_t queue_bytes) {
  const size_t sm_budget = 227 * 1024;
  StreamConfig best{};
  const int want_warps = env_int("SGC_WARPS", 12), want_ctas = env_int("SGC_CTAS", 2);
  const int want_stages = env_int("SGC_STAGES", 3);
  for (int ctas = want_ctas; ctas >= 1 && !best.stages; --ctas) {
    for (int stages = std::min(want_stages, kMaxStages); stages >= 2; --stages) {
      StreamConfig c{want_warps, stages, ctas};
      if (c.warps < 1 || c.warps > 12) c.warps = 12;
      if ((stream_smem_bytes(c, stage_bytes, queue_bytes) + 1024) * ctas <= sm_budget) {
        best = c;
        break;
      }
    }
  }
  return best;
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
  // staged kernel: fixed stride, 16-byte aligned base, whole warp tiles, < 2^32 reads per launch
  const uint32_t tile_bytes = kWarpReads * stride;
  const uint32_t stage_bytes = (tile_bytes + 48 + 15) & ~15u;  // +48: span words may run past the tile
  const bool stageable = d_off == nullptr && ((uintptr_t)d_lines & 15u) == 0 && stride >= read_len;
  const bool nw5 = !c->lib->wide && c->lib->k > 16;
  const size_t queue_bytes = nw5 ? sizeof(WarpQueueT<5>) : sizeof(WarpQueueT<8>);
  StreamConfig cfg = stageable ? pick_stream_config(stage_bytes, queue_bytes) : StreamConfig{};
  // whole tiles only, and never a bulk copy that would run past n_bytes
  uint64_t n_wtiles = stageable && cfg.stages ? std::min(n_reads / kWarpReads, n_bytes / tile_bytes) : 0;
  n_wtiles = std::min<uint64_t>(n_wtiles, 0xFFFFFFFFull / kWarpReads);
  if (n_wtiles > 0) {
    auto kernel = c->lib->wide ? count_stream_kernel<8, true>
                               : (nw5 ? count_stream_kernel<5, false> : count_stream_kernel<8, false>);
    auto parked = c->lib->wide ? count_parked_kernel<8, true>
                               : (nw5 ? count_parked_kernel<5, false> : count_parked_kernel<8, false>);
    // worst case every read is parked: (NW + 2) words per read
    const size_t need = n_wtiles * kWarpReads;
    const size_t rec_words = (nw5 ? 5 : 8) + 2;
    if (need > c->park_cap) {
      SGC_CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(c->d_park);
      c->d_park = nullptr;
      SGC_CUDA_TRY(cudaMalloc(&c->d_park, need * rec_words * sizeof(uint32_t)));
      c->park_cap = need;
    }
    if (!c->d_park_count) SGC_CUDA_TRY(cudaMalloc(&c->d_park_count, sizeof(unsigned int)));
    SGC_CUDA_TRY(cudaMemsetAsync(c->d_park_count, 0, sizeof(unsigned int), stream));
    p.park_rec = c->d_park;
    p.park_count = c->d_park_count;
    p.park_cap = (uint32_t)c->park_cap;
    const size_t smem = stream_smem_bytes(cfg, stage_bytes, queue_bytes);
    SGC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t grid = (uint64_t)c->lib->sm_count * cfg.ctas_per_sm;  // persistent: every CTA resident
    const uint64_t ctas_needed = (n_wtiles + cfg.warps - 1) / cfg.warps;
    if (grid > ctas_needed) grid = ctas_needed;
    p.n_reads = n_wtiles * kWarpReads;
    p.first_read = 0;
    kernel<<<(unsigned)grid, cfg.warps * 32, smem, stream>>>(p, n_wtiles, cfg.stages, stage_bytes);
    SGC_CUDA_TRY(cudaGetLastError());
    if (!(p.debug & 4u)) {
      parked<<<c->lib->sm_count * 8, 256, 0, stream>>>(p);
      SGC_CUDA_TRY(cudaGetLastError());
      c->last.launches_total += 1;
    }
    done = n_wtiles * kWarpReads;
    c->last.grid = (uint32_t)grid;
    c->last.block = cfg.warps * 32;
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
  c->stream = (cudaStream_t)stream;  // NULL is the CUDA default stream
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
  cudaStreamSynchronize(c->stream);
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
  cudaFree(c->d_park);
  cudaFree(c->d_park_count);
  if (c->own_state) cudaFree(c->d_state);
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
      uint64_t per = std::max<uint64_t>(kChunkAlignReads, (kChunkBytes / stride) / kChunkAlignReads * kChunkAlignReads);
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
