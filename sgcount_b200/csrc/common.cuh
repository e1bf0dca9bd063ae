// common.cuh — 2-bit encoding, the device-side library structures and the per-window decision
// procedure shared by every kernel of libsgcount_cuda.
//
// Encoding: code(c) = (c >> 1) & 3 for c in {A,C,G,T}: A=0 C=1 T=2 G=3; complement = code ^ 2.
// The NATURAL key of a token of k bases is  sum_j code(tok[j]) << 2j  (base 0 in the low bits).
//
// Two structures stand for the reference's Library + Permuter maps (library.rs:9-62,
// permutes.rs:34-158).  Both hold the n library members only, so they stay a few MB and live
// in L2 whatever the read stream does; the 80 n variant strings the reference materialises
// are never stored.
//
// 1. SEED INDEX (exact semantics of the whole lookup; every kernel's slow path)
//    The k bases are cut into kSeeds = 3 contiguous parts.  A token within Hamming distance 1
//    of a member differs from it inside at most one part, so it agrees with the member on the
//    COMPLEMENT of that part.  Seed i is that complement (13-14 bases at k = 20): directory i
//    hashes the token with part i masked out to a bucket (start | count) of postings (member
//    key + guide index) sorted by bucket.  A member at distance 1 whose difference lies in part
//    i sits in list i and in no other list's matching set, the member itself (distance 0) sits
//    in all of them.  One window costs three directory loads plus about one posting load:
//      - a posting equal to the token                          -> library member (library.rs:34-46)
//      - else exactly ONE member at Hamming distance 1         -> that member (the Permuter's
//        map entry, permutes.rs:127-144); two or more          -> the Permuter's `_null` set
//        (permutes.rs:149-152), no match.  SURVEY.md A.2 shows this closed form equals the
//        reference's insertion-order-dependent build in every observable lookup.
//      - a token with exactly one N: its parents are the members equal to it everywhere else
//        (permutes.rs:3 puts N in the lexicon); they all sit in the list of the part the N is in.
//    Buckets are hashed, so a list can hold members of other seeds; every posting is checked
//    against the seed before it counts.
// 2. FRONT TABLE (accelerator of the streaming kernel's common case, members only)
//    One 32-byte bucket per hash value, no probing chain: a member that does not fit its home
//    bucket is left out and the bucket is flagged, and a flagged miss is re-resolved through
//    the seed index.  Keys are in the INTERLEAVED layout the streaming kernel packs for free:
//    window word i (bases 4i..4i+3, one per byte) contributes (word >> 1) & 0x03030303 shifted
//    left by 2 (i mod 4) into `lo` (i < 4) or `hi` (i >= 4).
//      narrow slot (k <= 20), 64 bit: [63:42] guide index [41] occupied [40] bucket flag
//                                      [39:32] hi gathered to 8 bits [31:0] lo
//      wide slot (k = 21..30), 2 x 64 bit: word0 = hi << 32 | lo, word1 = same meta in [63:40]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sgc {

constexpr uint32_t kMaxK = 30;
constexpr uint32_t kNarrowMaxK = 20;
constexpr int kSeeds = 3;
constexpr int32_t kMiss = -1;

// directory entry: [21:0] first posting of the bucket, [23:22] number of postings, [31:24] tag
// (the 8 hash bits below the bucket bits) shared by every posting of the bucket.  Count 3 marks
// a GENERAL bucket: more than two postings, or postings of different tags; its exact size is
// in LibView::dir_count and its tag is not used.
constexpr uint32_t kDirStartMask = 0x3FFFFFu;
constexpr int kDirCountShift = 22;
constexpr uint32_t kDirGeneral = 3u;
constexpr int kDirTagShift = 24;

// narrow posting: [39:0] natural key, [61:40] guide index.  Wide posting: {natural key, guide index}.
constexpr int kPostIdxShift = 40;
constexpr uint64_t kPostKeyMask = (1ull << kPostIdxShift) - 1;

// front-table meta bits (bit positions inside the 64-bit slot / meta word)
constexpr uint64_t kFrontFlag = 1ull << 40;      // set in slot 0: some member of this bucket was left out
constexpr uint64_t kFrontOccupied = 1ull << 41;
constexpr int kFrontIdxShift = 42;

struct LibView {
  uint32_t k, n;
  uint32_t wide;                  // k > 20: 16-byte postings and front slots
  uint32_t part_end[kSeeds];      // part i = bases [part_end[i-1], part_end[i])
  uint64_t keep[kSeeds];          // key bits seed i keeps (everything but part i)
  const uint32_t* __restrict__ dir[kSeeds];        // 1 << dir_bits entries each
  const uint32_t* __restrict__ dir_count[kSeeds];  // exact bucket sizes (read for saturated entries only)
  const uint64_t* __restrict__ post;               // kSeeds x n postings, list i at i * n, each sorted by bucket
                                                   // (2 words per posting when wide)
  uint32_t dir_shift;                              // bucket = seed_hash >> dir_shift (>= 8)
  const uint64_t* __restrict__ front;      // forward orientation (keys of the guides as written)
  const uint64_t* __restrict__ front_rev;  // keys of the guides' reverse complements
  uint32_t front_shift;                    // bucket = front_hash >> front_shift
};

// bucket hash of a masked natural key (top bits are used)
__host__ __device__ __forceinline__ uint32_t seed_hash(uint64_t mkey) {
  uint32_t h = (uint32_t)mkey * 0x9E3779B1u + (uint32_t)(mkey >> 32) * 0x85EBCA77u;
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}

__host__ __device__ __forceinline__ bool is_acgt(uint8_t c) {
  return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
__host__ __device__ __forceinline__ uint32_t code_of(uint8_t c) { return (c >> 1) & 3u; }

// bucket of an interleaved key: multiplicative hash, top bits
__host__ __device__ __forceinline__ uint32_t front_hash(uint32_t lo, uint32_t hi) {
  return lo * 0x9E3779B1u + hi * 0x85EBCA77u;
}

// Interleaved key of a natural key (build side).  Narrow: hi is the gathered 8-bit form.
__host__ __device__ __forceinline__ void interleave_key(uint64_t key, uint32_t k, bool wide, uint32_t& lo, uint32_t& hi) {
  lo = 0;
  hi = 0;
  for (uint32_t j = 0; j < k; ++j) {
    const uint32_t c = (uint32_t)(key >> (2 * j)) & 3u;
    const uint32_t word = j >> 2, byte = j & 3u;
    if (word < 4)
      lo |= c << (8 * byte + 2 * word);
    else if (wide)
      hi |= c << (8 * byte + 2 * (word - 4));
    else
      hi |= c << (2 * byte);  // word 4 only (k <= 20)
  }
}

#ifdef __CUDACC__

__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p, uint64_t policy) {
  uint64_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t* p, uint64_t policy) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
// one 32-byte bucket = one sector, fetched with a single 256-bit read-only load (LDG.256)
__device__ __forceinline__ void load_bucket(const uint64_t* p, uint64_t (&w)[4], uint64_t policy) {
  asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
               : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
               : "l"(p), "l"(policy));
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

// x = XOR of two natural keys: true iff they differ in exactly one base
__device__ __forceinline__ bool one_base_differs(uint64_t x) {
  const uint64_t y = (x | (x >> 1)) & 0x5555555555555555ull;
  return y != 0 && (y & (y - 1)) == 0;
}

// One directory probe: where the postings of the token's seed i would be.
struct SeedRun {
  uint32_t first;  // index of the first posting in LibView::post (list offset included)
  uint32_t count;
};
__device__ __forceinline__ uint32_t seed_bucket(const LibView& v, uint32_t h) { return h >> v.dir_shift; }
__device__ __forceinline__ SeedRun seed_run(const LibView& v, int seed, uint32_t h, uint32_t entry) {
  SeedRun r;
  r.first = (uint32_t)seed * v.n + (entry & kDirStartMask);
  r.count = (entry >> kDirCountShift) & 3u;
  if (r.count == kDirGeneral)
    r.count = v.dir_count[seed][h >> v.dir_shift];
  else if ((entry >> kDirTagShift) != ((h >> (v.dir_shift - 8)) & 0xFFu))
    r.count = 0;  // the bucket belongs to another seed
  return r;
}
template <bool WIDE>
__device__ __forceinline__ void load_posting(const LibView& v, uint32_t at, uint64_t policy, uint64_t& key, uint32_t& idx) {
  if (WIDE) {
    key = ldg_u64(v.post + 2 * (size_t)at, policy);
    idx = (uint32_t)ldg_u64(v.post + 2 * (size_t)at + 1, policy);
  } else {
    const uint64_t w = ldg_u64(v.post + at, policy);
    key = w & kPostKeyMask;
    idx = (uint32_t)(w >> kPostIdxShift);
  }
}

// Every posting of list `seed` that may share the seed of `key` (build-side checks; the count
// kernels use window_lookup_t).  `visit(key, idx)` returning true stops the walk.
template <bool WIDE, typename F>
__device__ __forceinline__ void for_each_posting(const LibView& v, int seed, uint64_t key, uint64_t policy, F&& visit) {
  const uint32_t h = seed_hash(key & v.keep[seed]);
  const SeedRun r = seed_run(v, seed, h, ldg_u32(v.dir[seed] + seed_bucket(v, h), policy));
#pragma unroll 1
  for (uint32_t c = 0; c < r.count; ++c) {
    uint64_t mk;
    uint32_t idx;
    load_posting<WIDE>(v, r.first + c, policy, mk, idx);
    if (visit(mk, idx)) break;
  }
}

// Decision for ONE window, given its natural key, the number of bytes in
// it that are not A/C/G/T (`nbad`), the position of the single bad byte (`bad_pos`) and
// whether that byte is the wildcard ('N' as seen by the lookup).
// Returns the guide index or kMiss; *kind = 1 library member, 2 one-mismatch variant.
template <bool WIDE>
__device__ __forceinline__ int32_t window_lookup_t(const LibView& v, bool with_perm, uint64_t key, int nbad,
                                                   int bad_pos, bool bad_is_wild, int* kind, uint64_t policy) {
  if (nbad == 0) {
    // The three directory entries are fetched together (without a Permuter only the first is
    // needed: a member sits in every list) and their runs are walked as ONE sequence, so a
    // warp iterates max-over-lanes of the candidates per window, not per list.
    uint32_t h[kSeeds], entry[kSeeds];
#pragma unroll
    for (int i = 0; i < kSeeds; ++i) {
      h[i] = seed_hash(key & v.keep[i]);
      entry[i] = (i == 0 || with_perm) ? ldg_u32(v.dir[i] + seed_bucket(v, h[i]), policy) : 0u;
    }
    SeedRun r[kSeeds];
#pragma unroll
    for (int i = 0; i < kSeeds; ++i) {
      r[i] = seed_run(v, i, h[i], entry[i]);
      if (i > 0 && !with_perm) r[i].count = 0;
    }
    static_assert(kSeeds == 3, "the merged walk below is written for three lists");
    const uint32_t end0 = r[0].count, end1 = end0 + r[1].count, total = end1 + r[2].count;
    const uint32_t base0 = r[0].first, base1 = r[1].first - end0, base2 = r[2].first - end1;
    int32_t found = kMiss;
    int parents = 0;
    bool member = false;
#pragma unroll 1
    for (uint32_t j = 0; j < total; ++j) {
      const bool in0 = j < end0, in1 = j < end1;
      const uint32_t at = (in0 ? base0 : (in1 ? base1 : base2)) + j;
      const uint64_t keep = in0 ? v.keep[0] : (in1 ? v.keep[1] : v.keep[2]);
      uint64_t mk;
      uint32_t idx;
      load_posting<WIDE>(v, at, policy, mk, idx);
      const uint64_t x = mk ^ key;
      if ((x & keep) != 0) continue;  // a posting of another seed that hashed to this bucket
      if (x == 0) {                   // Library::contains (counter.rs:111-112)
        member = true;
        found = (int32_t)idx;
        break;
      }
      if (one_base_differs(x)) {  // the difference lies inside the part this list leaves out
        ++parents;
        found = (int32_t)idx;
      }
    }
    if (member) {
      if (kind) *kind = 1;
      return found;
    }
    if (with_perm && parents == 1) {  // Permuter::contains -> Library::alias (counter.rs:113-116)
      if (kind) *kind = 2;
      return found;
    }
    return kMiss;  // no parent, or the Permuter's null set (permutes.rs:149-152)
  }
  if (nbad == 1 && bad_is_wild && with_perm) {
    // parents = members equal to the token everywhere but at the wildcard: the list of the
    // part the wildcard is in
    int i = 0;
    while (i < kSeeds - 1 && (uint32_t)bad_pos >= v.part_end[i]) ++i;
    const uint64_t hole = ~(3ull << (2 * bad_pos));
    int32_t found = kMiss;
    int parents = 0;
    for_each_posting<WIDE>(v, i, key, policy, [&](uint64_t mk, uint32_t idx) {
      if (((mk ^ key) & hole) == 0) {
        ++parents;
        found = (int32_t)idx;
      }
      return false;
    });
    if (parents == 1) {
      if (kind) *kind = 2;
      return found;
    }
  }
  return kMiss;
}
__device__ __forceinline__ int32_t window_lookup(const LibView& v, bool with_perm, uint64_t key, int nbad, int bad_pos,
                                                 bool bad_is_wild, int* kind, uint64_t policy) {
  return v.wide ? window_lookup_t<true>(v, with_perm, key, nbad, bad_pos, bad_is_wild, kind, policy)
                : window_lookup_t<false>(v, with_perm, key, nbad, bad_pos, bad_is_wild, kind, policy);
}

// A span of up to 32 consecutive bases around the guide window, packed from the read in
// ORIENTED coordinates (forward: the read itself; reverse: its reverse complement).
struct Span {
  uint64_t codes;  // base j of the span at bits 2j
  uint32_t bad;    // bit j: byte j is not A/C/G/T (or lies outside the read)
  uint32_t wild;   // bit j: byte j is the wildcard byte
};

// reverse the order of the m 2-bit groups of x and complement every base
__device__ __forceinline__ uint64_t revcomp_codes(uint64_t x, int m) {
  uint64_t r = __brevll(x);
  r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
  r >>= (64 - 2 * m);
  return r ^ (0xAAAAAAAAAAAAAAAAull >> (64 - 2 * m));
}
__device__ __forceinline__ uint32_t reverse_bits(uint32_t x, int m) { return __brev(x) >> (32 - m); }

// The per-read walk of Counter::assign (counter.rs:96-140) over a span.
//   n        : read length
//   offset   : Offset index;  recursion: Centered -> Plus -> Minus, else Null only
// The span must cover oriented positions [offset-1, offset+k+1) (clipped to the read);
// `span_base` is the oriented position of span base 0.
template <bool WIDE>
__device__ __forceinline__ int32_t assign_span_t(const LibView& v, bool with_perm, const Span& sp, int span_base, int n,
                                                 int offset, bool recursion, int* kind, uint64_t policy) {
  const int k = (int)v.k;
  const uint64_t kmask = (1ull << (2 * k)) - 1;
  const uint32_t wmask = (1u << k) - 1;
  const int npos = recursion ? 3 : 1;
#pragma unroll 1
  for (int p = 0; p < npos; ++p) {
    // Centered/Null: offset; Plus: offset+1; Minus: offset-1 (counter.rs:164-174)
    int lo;
    if (p == 0) {
      lo = offset;
    } else if (p == 1) {
      lo = offset + 1;
    } else {
      if (offset == 0) return kMiss;  // checked_sub(1) -> None
      lo = offset - 1;
    }
    if (lo + k > n) return kMiss;  // failed trim RETURNS (counter.rs:105-108,175-176)
    const int sh = lo - span_base;
    const uint64_t key = (sp.codes >> (2 * sh)) & kmask;
    const uint32_t badw = (sp.bad >> sh) & wmask;
    const int nbad = __popc(badw);
    int bad_pos = 0;
    bool wild = false;
    if (nbad == 1) {
      bad_pos = __ffs(badw) - 1;
      wild = ((sp.wild >> sh) >> bad_pos) & 1u;
    }
    if (nbad > 1 || (nbad == 1 && !(wild && with_perm))) continue;  // a lookup miss, not a trim failure
    const int32_t hit = window_lookup_t<WIDE>(v, with_perm, key, nbad, bad_pos, wild, kind, policy);
    if (hit != kMiss) return hit;
  }
  return kMiss;
}
__device__ __forceinline__ int32_t assign_span(const LibView& v, bool with_perm, const Span& sp, int span_base, int n,
                                               int offset, bool recursion, int* kind, uint64_t policy) {
  return v.wide ? assign_span_t<true>(v, with_perm, sp, span_base, n, offset, recursion, kind, policy)
                : assign_span_t<false>(v, with_perm, sp, span_base, n, offset, recursion, kind, policy);
}

#endif  // __CUDACC__

}  // namespace sgc
