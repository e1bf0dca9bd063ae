// common.cuh — 2-bit encoding, the device-side library structures and the per-window decision
// procedure shared by every kernel of libsgcount_cuda.
//
// Encoding: code(c) = (c >> 1) & 3 for c in {A,C,G,T}: A=0 C=1 T=2 G=3; complement = code ^ 2.
//
// KEY LAYOUT.  A token of k bases is held as two 32-bit words in the INTERLEAVED layout the
// streaming kernel packs in two instructions per word of sequence bytes: base j = 4 i + b
// (word i of the window, byte b of that word) sits at
//     lo bit 8 b + 2 i            for i < 4   (bases 0..15)
//     hi bit 8 b + 2 (i - 4)      for i >= 4  (bases 16..31; wide keys, k = 21..30)
//     hi bit 2 b                  for i == 4  (bases 16..19; narrow keys, k <= 20: `hi` gathered
//                                              to 8 bits so that key + guide index fit one word)
// i.e. word i contributes (word >> 1) & 0x03030303, shifted left by 2 (i mod 4).
//
// Two structures stand for the reference's Library + Permuter maps (library.rs:9-62,
// permutes.rs:34-158).  Both hold the n library members only, so they stay a few MB and live
// in L2 whatever the read stream does; the 80 n variant strings the reference materialises
// are never stored.  Each exists once per read orientation: the forward index holds the guides
// as written, the reverse index their reverse complements, so that the streaming kernel looks
// up the bytes of a reverse read AS STORED (counter.rs:196-204 reverse-complements the read
// instead; the two are the same comparison).
//
// 1. SEED INDEX (exact semantics of the whole lookup)
//    The bases are cut into kSeeds = 3 parts of (nearly) equal size, [0, e0), [e0, e1), [e1, k)
//    (SeedParts, worked out per library: with fixed parts a short guide would leave one part
//    empty and the other seeds too short to tell the members apart).  A token within Hamming
//    distance 1 of a member differs
//    from it inside at most one part, so it agrees with the member on the COMPLEMENT of that
//    part.  Seed i is that complement: directory i hashes the token with part i masked out to a
//    bucket: its only posting inline, or (first | count) of a run of postings (member key + guide
//    index) sorted by bucket.  A
//    member at distance 1 whose difference lies in part i sits in list i; the member itself
//    (distance 0) sits in all of them.  One window costs three directory loads plus about one
//    posting load:
//      - a posting equal to the token                          -> library member (library.rs:34-46)
//      - else exactly ONE member at Hamming distance 1         -> that member (the Permuter's
//        map entry, permutes.rs:127-144); two or more          -> the Permuter's `_null` set
//        (permutes.rs:149-152), no match.  SURVEY.md A.2 shows this closed form equals the
//        reference's insertion-order-dependent build in every observable lookup.
//      - a token with exactly one N: its parents are the members equal to it everywhere else
//        (permutes.rs:3 puts N in the lexicon); they all sit in the list of the part the N is in.
//    Buckets are hashed, so a list can hold members of other seeds; every posting is checked
//    against the seed before it counts.
// 2. FRONT TABLE (accelerator of the streaming kernel's common case)
//    One 32-byte bucket per hash value, no probing chain: a member that does not fit its home
//    bucket is left out and the bucket is flagged, and a flagged miss is re-resolved through
//    the seed index.
//      narrow slot, 64 bit: [63:42] guide index [41] occupied [40] bucket flag [39:32] hi [31:0] lo
//      wide slot, 2 x 64 bit: word0 = hi << 32 | lo, word1 = same meta in [63:40]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sgc {

constexpr uint32_t kMaxK = 30;
constexpr uint32_t kNarrowMaxK = 20;
constexpr int kSeeds = 3;
constexpr int32_t kMiss = -1;

// NARROW directory entry (k <= 20), 64 bit, kind in [63:62]:
//   0  empty bucket
//   1  the bucket's ONLY posting, inline: [31:0] lo, [39:32] hi, [61:40] guide index — the common
//      case is decided with ONE load, and the inline key is its own (exact) tag
//   2  a run of postings: [21:0] first posting, [53:22] count
constexpr uint64_t kDirKindInline = 1ull << 62, kDirKindRun = 2ull << 62;
constexpr int kDirKindShift = 62;
// WIDE directory entry (k = 21..30), 2 x 64 bit, the same three kinds in [63:62] of word 0:
//   1  the bucket's only posting inline: word 0 [61:0] = hi << 32 | lo (a 30-base key leaves the
//      two top bits of `hi` free), word 1 = guide index
//   2  a run of postings: word 1 = first posting | count << 22
constexpr uint32_t kDirStartMask = 0x3FFFFFu;

// narrow posting: [31:0] lo, [39:32] hi, [61:40] guide index.  Wide posting: {hi << 32 | lo, guide index}.
constexpr int kPostIdxShift = 40;

// front-table meta bits (bit positions inside the 64-bit slot / meta word)
constexpr uint64_t kFrontFlag = 1ull << 40;      // set in slot 0: some member of this bucket was left out
constexpr uint64_t kFrontOccupied = 1ull << 41;
constexpr int kFrontIdxShift = 42;

struct Key {
  uint32_t lo, hi;
};

// the 2-bit field of base j inside a key
__host__ __device__ __forceinline__ Key base_field(uint32_t j, bool wide) {
  const uint32_t word = j >> 2, byte = j & 3u;
  if (word < 4) return Key{3u << (8 * byte + 2 * word), 0u};
  return wide ? Key{0u, 3u << (8 * byte + 2 * (word - 4))} : Key{0u, 3u << (2 * byte)};
}
__host__ __device__ __forceinline__ void key_set_base(Key& k, uint32_t j, bool wide, uint32_t code) {
  const Key f = base_field(j, wide);
  // lowest set bit of the field times the code (fields are 2 aligned bits)
  k.lo = (k.lo & ~f.lo) | ((f.lo & (~f.lo + 1)) * code);
  k.hi = (k.hi & ~f.hi) | ((f.hi & (~f.hi + 1)) * code);
}
__host__ __device__ __forceinline__ uint32_t key_get_base(Key k, uint32_t j, bool wide) {
  const Key f = base_field(j, wide);
  return f.lo ? (k.lo & f.lo) / (f.lo & (~f.lo + 1)) : (k.hi & f.hi) / (f.hi & (~f.hi + 1));
}

// The three parts of the k bases and, per list, the key bits its seed KEEPS (everything outside
// its own part).
constexpr uint32_t kFixedPartsMinK = 17;
struct SeedParts {
  uint32_t end[2];  // part 0 = bases [0, end[0]), part 1 = [end[0], end[1]), part 2 = the rest
  uint32_t keep_lo[kSeeds], keep_hi[kSeeds];
};
inline SeedParts make_seed_parts(uint32_t k, bool wide) {
  SeedParts sp{};
  // k >= kFixedPartsMinK: two parts of 8 bases (whole window words, i.e. alternating nibbles of
  // `lo`) and the rest — the streaming kernel has these masks as immediates (seed_of_t<true>);
  // shorter guides: three parts of nearly equal size, masks read from here
  const bool fixed = k >= kFixedPartsMinK;
  const uint32_t p0 = fixed ? 8 : (k + 2) / 3, p1 = fixed ? 8 : (k - p0 + 1) / 2;
  sp.end[0] = p0;
  sp.end[1] = p0 + p1;
  for (uint32_t j = 0; j < k; ++j) {
    const int part = j < sp.end[0] ? 0 : (j < sp.end[1] ? 1 : 2);
    const Key f = base_field(j, wide);
    for (int i = 0; i < kSeeds; ++i)
      if (i != part) {
        sp.keep_lo[i] |= f.lo;
        sp.keep_hi[i] |= f.hi;
      }
  }
  return sp;
}
// what seed i keeps of a key
__host__ __device__ __forceinline__ Key seed_of(const SeedParts& sp, Key k, int i) {
  return Key{k.lo & sp.keep_lo[i], k.hi & sp.keep_hi[i]};
}
// part (= list) a base position belongs to
__host__ __device__ __forceinline__ int part_of_base(const SeedParts& sp, uint32_t j) {
  return j < sp.end[0] ? 0 : (j < sp.end[1] ? 1 : 2);
}

// The same with the parts of a k >= kFixedPartsMinK library as compile-time constants (FIXED),
// for the streaming kernel's hot instantiations; FIXED = false reads them from the library.
template <bool FIXED>
__host__ __device__ __forceinline__ Key seed_of_t(const SeedParts& sp, Key k, int i) {
  if (FIXED) return i == 0 ? Key{k.lo & 0xF0F0F0F0u, k.hi} : (i == 1 ? Key{k.lo & 0x0F0F0F0Fu, k.hi} : Key{k.lo, 0u});
  return seed_of(sp, k, i);
}
template <bool FIXED>
__host__ __device__ __forceinline__ int part_of_base_t(const SeedParts& sp, uint32_t j) {
  if (FIXED) return j < 8 ? 0 : (j < 16 ? 1 : 2);
  return part_of_base(sp, j);
}
// the bits of a key difference that lie in what list `list` keeps
template <bool FIXED>
__host__ __device__ __forceinline__ uint32_t kept_difference(const SeedParts& sp, Key x, int list) {
  if (FIXED) {
    const uint32_t keep_lo = list == 0 ? 0xF0F0F0F0u : (list == 1 ? 0x0F0F0F0Fu : 0xFFFFFFFFu);
    return (x.lo & keep_lo) | (list < 2 ? x.hi : 0u);
  }
  return (x.lo & sp.keep_lo[list]) | (x.hi & sp.keep_hi[list]);
}

// bucket hash of a (masked) key; the top bits are used
__host__ __device__ __forceinline__ uint32_t seed_hash(Key s) {
  uint32_t h = s.lo * 0x9E3779B1u + s.hi * 0x85EBCA77u;
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}
// bucket of the front table
__host__ __device__ __forceinline__ uint32_t front_hash(uint32_t lo, uint32_t hi) {
  return lo * 0x9E3779B1u + hi * 0x85EBCA77u;
}

__host__ __device__ __forceinline__ bool is_acgt(uint8_t c) {
  return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
__host__ __device__ __forceinline__ uint32_t code_of(uint8_t c) { return (c >> 1) & 3u; }

// One orientation's structures.
struct IndexView {
  const uint64_t* __restrict__ dir64[kSeeds];      // narrow: 1 << dir_bits entries of one word each
  const ulonglong2* __restrict__ dir128[kSeeds];   // wide: 1 << dir_bits entries of two words each
  const uint64_t* __restrict__ post;               // kSeeds x n postings, list i at i * n, each sorted by bucket
                                                   // (2 words per posting when wide)
  const uint64_t* __restrict__ front;              // front table
};

struct LibView {
  uint32_t k, n;
  uint32_t wide;         // k > 20: 16-byte postings and front slots
  uint32_t dir_shift;    // bucket = seed_hash >> dir_shift (>= 8)
  uint32_t front_shift;  // bucket = front_hash >> front_shift
  SeedParts parts;
  IndexView fwd;         // guides as written
  IndexView rev;         // reverse complements of the guides
};

#ifdef __CUDACC__

// L1 policy of the table loads (tuning knob): the tables have no L1 locality worth keeping
#ifndef SGC_L1_HINT
#define SGC_L1_HINT ""
#endif
__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p, uint64_t policy) {
  uint64_t v;
  asm volatile("ld.global.nc" SGC_L1_HINT ".L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t* p, uint64_t policy) {
  uint32_t v;
  asm volatile("ld.global.nc" SGC_L1_HINT ".L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ ulonglong2 ldg_u128(const ulonglong2* p, uint64_t policy) {
  ulonglong2 v;
  asm volatile("ld.global.nc" SGC_L1_HINT ".L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
               : "=l"(v.x), "=l"(v.y)
               : "l"(p), "l"(policy));
  return v;
}
// one 32-byte bucket = one sector, fetched with a single 256-bit read-only load (LDG.256)
__device__ __forceinline__ void load_bucket(const uint64_t* p, uint64_t (&w)[4], uint64_t policy) {
  asm volatile("ld.global.nc" SGC_L1_HINT ".L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
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

// number of bases in which two keys differ, given their XOR
__device__ __forceinline__ int bases_differing(Key x) {
  return __popc((x.lo | (x.lo >> 1)) & 0x55555555u) + __popc((x.hi | (x.hi >> 1)) & 0x55555555u);
}

// One directory probe: where the postings of a seed would be.
struct SeedRun {
  uint32_t first;  // index of the first posting in IndexView::post (list offset included)
  uint32_t count;
};
// wide entries of kind 2 (any other kind: an empty run)
__device__ __forceinline__ SeedRun seed_run128(const LibView& v, int seed, ulonglong2 entry) {
  SeedRun r;
  r.first = (uint32_t)seed * v.n + ((uint32_t)entry.y & kDirStartMask);
  r.count = (entry.x >> kDirKindShift) == 2 ? (uint32_t)(entry.y >> 22) : 0u;
  return r;
}
// narrow entries of kind 2 (any other kind: an empty run)
__device__ __forceinline__ SeedRun seed_run64(const LibView& v, int seed, uint64_t entry) {
  SeedRun r;
  r.first = (uint32_t)seed * v.n + ((uint32_t)entry & kDirStartMask);
  r.count = (entry >> kDirKindShift) == 2 ? (uint32_t)(entry >> 22) : 0u;
  return r;
}
template <bool WIDE>
__device__ __forceinline__ void load_posting(const IndexView& ix, uint32_t at, uint64_t policy, Key& key, uint32_t& idx) {
  if (WIDE) {
    const uint64_t w = ldg_u64(ix.post + 2 * (size_t)at, policy);
    key = Key{(uint32_t)w, (uint32_t)(w >> 32)};
    idx = (uint32_t)ldg_u64(ix.post + 2 * (size_t)at + 1, policy);
  } else {
    const uint64_t w = ldg_u64(ix.post + at, policy);
    key = Key{(uint32_t)w, (uint32_t)(w >> 32) & 0xFFu};
    idx = (uint32_t)(w >> kPostIdxShift) & 0x3FFFFFu;
  }
}

// Every posting of list `seed` that may share the seed of `key`.  `visit(key, idx)` returning
// true stops the walk.
template <bool WIDE, typename F>
__device__ __forceinline__ void for_each_posting(const LibView& v, const IndexView& ix, int seed, Key key,
                                                 uint64_t policy, F&& visit) {
  const uint32_t h = seed_hash(seed_of(v.parts, key, seed));
  SeedRun r;
  if (WIDE) {
    const ulonglong2 e = ldg_u128(ix.dir128[seed] + (h >> v.dir_shift), policy);
    if ((e.x >> kDirKindShift) == 1) {
      visit(Key{(uint32_t)e.x, (uint32_t)(e.x >> 32) & 0x3FFFFFFFu}, (uint32_t)e.y);
      return;
    }
    r = seed_run128(v, seed, e);
  } else {
    const uint64_t e = ldg_u64(ix.dir64[seed] + (h >> v.dir_shift), policy);
    if ((e >> kDirKindShift) == 1) {
      visit(Key{(uint32_t)e, (uint32_t)(e >> 32) & 0xFFu}, (uint32_t)(e >> kPostIdxShift) & 0x3FFFFFu);
      return;
    }
    r = seed_run64(v, seed, e);
  }
#pragma unroll 1
  for (uint32_t c = 0; c < r.count; ++c) {
    Key mk;
    uint32_t idx;
    load_posting<WIDE>(ix, r.first + c, policy, mk, idx);
    if (visit(mk, idx)) break;
  }
}

// ONE token against the seed index, both kinds of token through the same straight-line code so
// that the lanes of a warp stay together (counter.rs:111-117):
//   hole_list < 0   every byte is A/C/G/T: Library::contains, then Permuter::contains.  A posting
//                   equal to the token is the member; otherwise exactly one member at Hamming
//                   distance 1 is the Permuter's parent, two or more its null set.
//   hole_list >= 0  exactly one byte is the wildcard; `hole` is its 2-bit field, already cleared
//                   in `key`, and hole_list the list of the part it lies in.  Its parents are the
//                   members equal to it everywhere else; they all sit in that one list (the other
//                   lists keep the hole base in their seed and are not consulted).
//   !active         nothing is loaded, the answer is kMiss.
// Returns the guide index or kMiss; *kind = 1 member, 2 one-mismatch variant.
template <bool WIDE, bool FIXED = false>
__device__ __forceinline__ int32_t lookup_token(const LibView& v, const IndexView& ix, bool with_perm, Key key, Key hole,
                                                int hole_list, bool active, int* kind, uint64_t policy) {
  static_assert(kSeeds == 3, "written for three lists");
  const bool wild = hole_list >= 0;
  int32_t found = kMiss;
  int parents = 0;
  bool member = false;
  // one candidate of list `list`: a posting of another seed that hashed to this bucket differs
  // on what the list keeps; a difference of one base lies inside the part this list leaves out
  auto consider = [&](int list, Key mk, uint32_t idx) {
    const Key x{(mk.lo ^ key.lo) & ~hole.lo, (mk.hi ^ key.hi) & ~hole.hi};
    if (kept_difference<FIXED>(v.parts, x, list) != 0) return;
    if ((x.lo | x.hi) == 0) {
      if (wild) {
        ++parents;
      } else {
        member = true;  // Library::contains (counter.rs:111-112)
      }
      found = (int32_t)idx;
    } else if (!wild && !member && bases_differing(x) == 1) {
      ++parents;
      found = (int32_t)idx;
    }
  };
  // The directory entries are fetched together (without a Permuter only the first is needed: a
  // member sits in every list).  Narrow keys: a bucket with one posting carries it inline, so
  // the usual window is decided by this one round trip.  What is left are runs of postings,
  // walked as ONE sequence so that a warp iterates max-over-lanes of the candidates per window,
  // not per list.
  bool use[kSeeds];
  uint32_t h[kSeeds];
  SeedRun r[kSeeds];
#pragma unroll
  for (int i = 0; i < kSeeds; ++i) {
    use[i] = active && (wild ? i == hole_list : (i == 0 || with_perm));
    h[i] = seed_hash(seed_of_t<FIXED>(v.parts, key, i));
  }
  if (WIDE) {
    ulonglong2 entry[kSeeds];
#pragma unroll
    for (int i = 0; i < kSeeds; ++i)
      entry[i] = use[i] ? ldg_u128(ix.dir128[i] + (h[i] >> v.dir_shift), policy) : ulonglong2{0ull, 0ull};
#pragma unroll
    for (int i = 0; i < kSeeds; ++i) {
      if ((entry[i].x >> kDirKindShift) == 1)
        consider(i, Key{(uint32_t)entry[i].x, (uint32_t)(entry[i].x >> 32) & 0x3FFFFFFFu}, (uint32_t)entry[i].y);
      r[i] = seed_run128(v, i, entry[i]);
    }
  } else {
    uint64_t entry[kSeeds];
#pragma unroll
    for (int i = 0; i < kSeeds; ++i) entry[i] = use[i] ? ldg_u64(ix.dir64[i] + (h[i] >> v.dir_shift), policy) : 0ull;
#pragma unroll
    for (int i = 0; i < kSeeds; ++i) {
      if ((entry[i] >> kDirKindShift) == 1)
        consider(i, Key{(uint32_t)entry[i], (uint32_t)(entry[i] >> 32) & 0xFFu},
                 (uint32_t)(entry[i] >> kPostIdxShift) & 0x3FFFFFu);
      r[i] = seed_run64(v, i, entry[i]);
    }
  }
  const uint32_t end0 = r[0].count, end1 = end0 + r[1].count, total = end1 + r[2].count;
  const uint32_t base0 = r[0].first, base1 = r[1].first - end0, base2 = r[2].first - end1;
#pragma unroll 1
  for (uint32_t j = 0; j < total && !member; ++j) {
    const bool in0 = j < end0, in1 = j < end1;
    const uint32_t at = (in0 ? base0 : (in1 ? base1 : base2)) + j;
    Key mk;
    uint32_t idx;
    load_posting<WIDE>(ix, at, policy, mk, idx);
    consider(in0 ? 0 : (in1 ? 1 : 2), mk, idx);
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

// A token all of whose bytes are A/C/G/T.
template <bool WIDE>
__device__ __forceinline__ int32_t lookup_clean(const LibView& v, const IndexView& ix, bool with_perm, Key key,
                                                int* kind, uint64_t policy) {
  return lookup_token<WIDE>(v, ix, with_perm, key, Key{0u, 0u}, -1, true, kind, policy);
}

// A token with exactly one wildcard byte at base `pos` (its 2-bit field in `key` is ignored):
// the parents are the members equal to it everywhere else.  Only asked with a Permuter.
template <bool WIDE>
__device__ __forceinline__ int32_t lookup_wild(const LibView& v, const IndexView& ix, Key key, uint32_t pos, int* kind,
                                               uint64_t policy) {
  const Key hole = base_field(pos, WIDE);
  key.lo &= ~hole.lo;
  key.hi &= ~hole.hi;
  return lookup_token<WIDE>(v, ix, true, key, hole, part_of_base(v.parts, pos), true, kind, policy);
}

// Decision for ONE window given its key, the number of bytes in it that
// are not A/C/G/T (`nbad`), the base position of the single bad byte and whether that byte is
// the wildcard ('N' as seen by the lookup).
template <bool WIDE>
__device__ __forceinline__ int32_t window_lookup_t(const LibView& v, const IndexView& ix, bool with_perm, Key key,
                                                   int nbad, uint32_t bad_pos, bool bad_is_wild, int* kind,
                                                   uint64_t policy) {
  if (nbad == 0) return lookup_clean<WIDE>(v, ix, with_perm, key, kind, policy);
  if (nbad == 1 && bad_is_wild && with_perm) return lookup_wild<WIDE>(v, ix, key, bad_pos, kind, policy);
  return kMiss;
}

#endif  // __CUDACC__

}  // namespace sgc
