// opaque.cu — libraries the 2-bit tables cannot hold: a sequence byte outside A,C,G,T, or guides
// longer than 30 bases.
//
// The reference keeps library sequences as opaque byte strings of any length (library.rs:65-99)
// and builds the one-mismatch variants by substituting the lexicon A,C,G,T,N at every position
// (permutes.rs:3,78-107), so such libraries are legal input.  They are rare (a masked base, a
// long construct), so this path is written for exactness, not for the roofline: the members'
// bytes stay in device memory as written and are indexed by their two HALVES.  A token within
// Hamming distance 1 of a member equals it on at least one half; list h (a directory of buckets
// over a hash of half h, postings = guide indices) therefore yields every member whose half h
// equals the token's, and each candidate is compared byte for byte:
//   - no difference                                   -> library member (library.rs:34-46);
//   - one difference, lying in the OTHER half (so that every distance-1 member is seen exactly
//     once), the token's byte there in the lexicon    -> a parent (permutes.rs:127-144);
//   exactly one parent -> the Permuter's answer, two or more -> its null set (permutes.rs:149-152).
// Same decision procedure as common.cuh's lookup_token, on bytes.
//
// Reverse orientation compares the reverse complement of the stored read, byte by byte as fxread
// produces it (counter.rs:196-204; rc_mode as everywhere else).
#include "internal.h"

namespace sgc {
namespace {

struct OpaqueView {
  const uint8_t* __restrict__ seqs;        // n * k bytes, library order
  const uint32_t* __restrict__ start[2];   // buckets + 1 posting offsets per half
  const uint32_t* __restrict__ post[2];    // n guide indices per half, grouped by bucket
  uint32_t k, n, shift;
};

__host__ __device__ __forceinline__ bool in_lexicon(uint8_t c) {  // permutes.rs:3
  return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N';
}

__device__ __forceinline__ uint32_t bucket_of(uint64_t h, uint32_t shift) {
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h >> shift;
}

// half h of a k-byte token is [h ? k/2 : 0, h ? k : k/2)
template <typename Tok>
__device__ __forceinline__ uint32_t half_bucket(Tok tok, uint32_t k, int h, uint32_t shift) {
  const uint32_t a = h ? k / 2 : 0, b = h ? k : k / 2;
  uint64_t x = 1469598103934665603ull;
  for (uint32_t j = a; j < b; ++j) x = (x ^ tok(j)) * 1099511628211ull;
  return bucket_of(x, shift);
}

// Library::contains, then Permuter::contains -> Library::alias (counter.rs:111-117) for one token
// given as an accessor of its bytes.  *kind = 1 member, 2 one-mismatch variant.  `parents` /
// `smallest`, when asked for, receive the number of distance-1 parents and the smallest of them.
template <typename Tok>
__device__ int32_t opaque_lookup(const OpaqueView& v, Tok tok, bool with_perm, int* kind, uint32_t* parents_out = nullptr,
                                 uint32_t* smallest_out = nullptr, bool* member_out = nullptr) {
  const uint32_t k = v.k, mid = k / 2;
  int32_t parent = kMiss;
  uint32_t parents = 0, smallest = 0xFFFFFFFFu;
  if (member_out) *member_out = false;
  for (int h = 0; h < 2; ++h) {
    const uint32_t a = h ? mid : 0, b = h ? k : mid;
    const uint32_t bucket = half_bucket(tok, k, h, v.shift);
    for (uint32_t p = v.start[h][bucket], end = v.start[h][bucket + 1]; p < end; ++p) {
      const uint32_t idx = v.post[h][p];
      const uint8_t* m = v.seqs + (size_t)idx * k;
      uint32_t mism = 0, pos = 0;
      for (uint32_t j = 0; j < k && mism < 2; ++j)
        if (m[j] != tok(j)) {
          ++mism;
          pos = j;
        }
      if (mism == 0) {
        if (kind) *kind = 1;
        if (member_out) *member_out = true;
        return (int32_t)idx;
      }
      if (mism == 1 && with_perm && !(pos >= a && pos < b) && in_lexicon(tok(pos))) {
        ++parents;
        parent = (int32_t)idx;
        smallest = min(smallest, idx);
      }
    }
  }
  if (parents_out) *parents_out = parents;
  if (smallest_out) *smallest_out = smallest;
  if (with_perm && parents == 1) {
    if (kind) *kind = 2;
    return parent;
  }
  return kMiss;
}

// Record::seq_rev_comp of the fxread crate on one byte
__device__ __forceinline__ uint8_t complement_byte(uint8_t c, int rc_mode) {
  if (rc_mode == SGC_RC_BITTRICK) return (c & 2) ? (c ^ 4) : (c ^ 21);
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return c;
  }
}

// ---- build ----------------------------------------------------------------------------------
struct OpaqueStatus {
  unsigned int dup_guide;
  unsigned long long n_variants, n_ambiguous;
};

__global__ void opaque_count_kernel(const uint8_t* __restrict__ seqs, uint32_t n, uint32_t k, uint32_t shift,
                                    uint32_t* cnt0, uint32_t* cnt1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* s = seqs + (size_t)i * k;
  auto tok = [&](uint32_t j) { return s[j]; };
  atomicAdd(cnt0 + half_bucket(tok, k, 0, shift), 1u);
  atomicAdd(cnt1 + half_bucket(tok, k, 1, shift), 1u);
}

__global__ void opaque_fill_kernel(const uint8_t* __restrict__ seqs, uint32_t n, uint32_t k, uint32_t shift,
                                   uint32_t* cur0, uint32_t* cur1, uint32_t* post0, uint32_t* post1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* s = seqs + (size_t)i * k;
  auto tok = [&](uint32_t j) { return s[j]; };
  post0[atomicAdd(cur0 + half_bucket(tok, k, 0, shift), 1u)] = i;
  post1[atomicAdd(cur1 + half_bucket(tok, k, 1, shift), 1u)] = i;
}

// duplicate sequences (library.rs:91-95): the later of two equal records reports itself
__global__ void opaque_duplicates_kernel(OpaqueView v, OpaqueStatus* st) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  const uint8_t* s = v.seqs + (size_t)i * v.k;
  auto tok = [&](uint32_t j) { return s[j]; };
  const uint32_t bucket = half_bucket(tok, v.k, 0, v.shift);
  for (uint32_t p = v.start[0][bucket], end = v.start[0][bucket + 1]; p < end; ++p) {
    const uint32_t idx = v.post[0][p];
    if (idx >= i) continue;
    const uint8_t* m = v.seqs + (size_t)idx * v.k;
    bool same = true;
    for (uint32_t j = 0; j < v.k && same; ++j) same = m[j] == s[j];
    if (same) {
      atomicMin(&st->dup_guide, i);
      return;
    }
  }
}

// Permuter statistics in the terms of sgc_library_info: A,C,G,T substitutions of every member
// (one thread per member and position) that resolve to exactly one parent / to several.
__global__ void opaque_variant_stats_kernel(OpaqueView v, OpaqueStatus* st) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned var = 0, amb = 0;
  if (t < (uint64_t)v.n * v.k) {
    const uint32_t i = (uint32_t)(t / v.k), pos = (uint32_t)(t % v.k);
    const uint8_t* s = v.seqs + (size_t)i * v.k;
    for (int c = 0; c < 4; ++c) {
      const uint8_t sub = "ACGT"[c];
      if (sub == s[pos]) continue;
      auto tok = [&](uint32_t j) { return j == pos ? sub : s[j]; };
      uint32_t parents = 0, smallest = 0;
      bool member = false;
      opaque_lookup(v, tok, true, nullptr, &parents, &smallest, &member);
      if (member) continue;
      if (parents == 1) ++var;
      if (parents > 1 && smallest == i) ++amb;
    }
  }
  var = __reduce_add_sync(0xffffffffu, var);
  amb = __reduce_add_sync(0xffffffffu, amb);
  if ((threadIdx.x & 31) == 0) {
    if (var) atomicAdd(&st->n_variants, (unsigned long long)var);
    if (amb) atomicAdd(&st->n_ambiguous, (unsigned long long)amb);
  }
}

__global__ void opaque_lookup_tokens_kernel(OpaqueView v, bool with_perm, const uint8_t* __restrict__ tokens,
                                            uint64_t n_tokens, int32_t* __restrict__ idx_out,
                                            uint8_t* __restrict__ kind_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tokens) return;
  const uint8_t* s = tokens + i * v.k;
  auto tok = [&](uint32_t j) { return s[j]; };
  int kind = 0;
  const int32_t hit = opaque_lookup(v, tok, with_perm, &kind);
  idx_out[i] = hit;
  if (kind_out) kind_out[i] = hit == kMiss ? 0 : (uint8_t)kind;
}

// ---- count: Counter::assign (counter.rs:96-140) spelled out on the oriented bytes -----------------
struct OpaqueCountParams {
  OpaqueView v;
  const uint8_t* lines;
  const uint32_t* line_off;
  uint64_t n_reads, first_read;
  uint32_t stride, read_len, off_base;
  int offset;
  uint8_t with_perm, reverse, recursion, rc_mode;
  unsigned long long* state;
  int32_t* assign_out;
};

__global__ void __launch_bounds__(256) count_opaque_kernel(OpaqueCountParams p) {
  uint32_t matched = 0;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const int k = (int)p.v.k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_reads; i += nthreads) {
    const uint64_t r = p.first_read + i;
    uint64_t start;
    int n;
    if (p.line_off) {
      start = p.line_off[r] - p.off_base;
      n = (int)(p.line_off[r + 1] - p.line_off[r]) - 1;
    } else {
      start = r * p.stride;
      n = (int)p.read_len;
    }
    const uint8_t* s = p.lines + start;
    int32_t hit = kMiss;
    const int npos = p.recursion ? 3 : 1;
    for (int pos = 0; pos < npos && hit == kMiss; ++pos) {
      // Centered/Null: offset; Plus: offset+1; Minus: offset-1 (counter.rs:164-174)
      int lo;
      if (pos == 0) {
        lo = p.offset;
      } else if (pos == 1) {
        lo = p.offset + 1;
      } else {
        if (p.offset == 0) break;  // checked_sub(1) -> None
        lo = p.offset - 1;
      }
      if (lo + k > n) break;  // a failed trim RETURNS (counter.rs:105-108,175-176)
      // forward: the read itself; reverse: byte lo+j of the reverse complement (counter.rs:196-204)
      auto tok = [&](uint32_t j) -> uint8_t {
        return p.reverse ? complement_byte(s[n - 1 - (lo + (int)j)], p.rc_mode) : s[lo + (int)j];
      };
      hit = opaque_lookup(p.v, tok, p.with_perm != 0, nullptr);
    }
    if (p.assign_out) p.assign_out[r] = hit;
    if (hit >= 0) {
      ++matched;
      atomicAdd(p.state + hit, 1ull);
    }
  }
  matched = __reduce_add_sync(0xffffffffu, matched);
  if ((threadIdx.x & 31) == 0 && matched) atomicAdd(p.state + p.v.n + 1, (unsigned long long)matched);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.state + p.v.n, (unsigned long long)p.n_reads);
}

OpaqueView view_of(const sgc_library* lib) {
  OpaqueView v{};
  v.seqs = lib->d_oseqs;
  for (int h = 0; h < 2; ++h) {
    v.start[h] = lib->d_ostart[h];
    v.post[h] = lib->d_opost[h];
  }
  v.k = lib->k;
  v.n = lib->n;
  v.shift = lib->oshift;
  return v;
}

inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

int opaque_build(sgc_library* lib, const uint8_t* d_seqs, uint64_t* n_variants, uint64_t* n_ambiguous,
                 uint32_t* dup_guide) {
  const uint32_t n = lib->n, k = lib->k;
  uint32_t bits = 8;
  while (((uint64_t)1 << bits) < 4ull * n) ++bits;
  lib->oshift = 32 - bits;
  const uint32_t entries = 1u << bits;
  SGC_CUDA_TRY(cudaMalloc(&lib->d_oseqs, (size_t)n * k));
  SGC_CUDA_TRY(cudaMemcpyAsync(lib->d_oseqs, d_seqs, (size_t)n * k, cudaMemcpyDeviceToDevice, 0));
  uint32_t *cnt[2] = {nullptr, nullptr}, *cur[2] = {nullptr, nullptr}, *sums = nullptr;
  OpaqueStatus* d_st = nullptr;
  struct Free {
    uint32_t **a, **b, **s;
    OpaqueStatus** st;
    ~Free() {
      for (int h = 0; h < 2; ++h) {
        cudaFree(a[h]);
        cudaFree(b[h]);
      }
      cudaFree(*s);
      cudaFree(*st);
    }
  } free_all{cnt, cur, &sums, &d_st};
  for (int h = 0; h < 2; ++h) {
    SGC_CUDA_TRY(cudaMalloc(&cnt[h], ((size_t)entries + 1) * 4));
    SGC_CUDA_TRY(cudaMalloc(&cur[h], ((size_t)entries + 1) * 4));
    SGC_CUDA_TRY(cudaMalloc(&lib->d_ostart[h], ((size_t)entries + 1) * 4));
    SGC_CUDA_TRY(cudaMalloc(&lib->d_opost[h], (size_t)n * 4));
    SGC_CUDA_TRY(cudaMemsetAsync(cnt[h], 0, ((size_t)entries + 1) * 4, 0));
  }
  SGC_CUDA_TRY(cudaMalloc(&sums, ((size_t)entries / 2048 + 2) * 4));
  SGC_CUDA_TRY(cudaMalloc(&d_st, sizeof(OpaqueStatus)));
  const OpaqueStatus st0{0xFFFFFFFFu, 0, 0};
  SGC_CUDA_TRY(cudaMemcpyAsync(d_st, &st0, sizeof st0, cudaMemcpyHostToDevice, 0));
  opaque_count_kernel<<<blocks_for(n, 256), 256>>>(lib->d_oseqs, n, k, lib->oshift, cnt[0], cnt[1]);
  for (int h = 0; h < 2; ++h) {
    int rc = exclusive_scan_u32(cnt[h], entries + 1, lib->d_ostart[h], sums);  // entry `entries` = n
    if (rc) return rc;
    SGC_CUDA_TRY(cudaMemcpyAsync(cur[h], lib->d_ostart[h], ((size_t)entries + 1) * 4, cudaMemcpyDeviceToDevice, 0));
  }
  opaque_fill_kernel<<<blocks_for(n, 256), 256>>>(lib->d_oseqs, n, k, lib->oshift, cur[0], cur[1], lib->d_opost[0],
                                                  lib->d_opost[1]);
  const OpaqueView v = view_of(lib);
  opaque_duplicates_kernel<<<blocks_for(n, 256), 256>>>(v, d_st);
  if (lib->with_perm) opaque_variant_stats_kernel<<<blocks_for((uint64_t)n * k, 256), 256>>>(v, d_st);
  SGC_CUDA_TRY(cudaGetLastError());
  OpaqueStatus st;
  SGC_CUDA_TRY(cudaMemcpy(&st, d_st, sizeof st, cudaMemcpyDeviceToHost));
  *n_variants = st.n_variants;
  *n_ambiguous = st.n_ambiguous;
  *dup_guide = st.dup_guide;
  lib->info.table_bytes = (size_t)n * k + 2 * (((size_t)entries + 1) * 4 + (size_t)n * 4);
  return SGC_OK;
}

void opaque_destroy(sgc_library* lib) {
  cudaFree(lib->d_oseqs);
  for (int h = 0; h < 2; ++h) {
    cudaFree(lib->d_ostart[h]);
    cudaFree(lib->d_opost[h]);
  }
}

int opaque_lookup_tokens(const sgc_library* lib, const uint8_t* d_tokens, uint64_t n_tokens, int32_t* d_idx,
                         uint8_t* d_kind) {
  opaque_lookup_tokens_kernel<<<blocks_for(n_tokens, 256), 256>>>(view_of(lib), lib->with_perm, d_tokens, n_tokens, d_idx,
                                                                   d_kind);
  SGC_CUDA_TRY(cudaGetLastError());
  return SGC_OK;
}

int opaque_count(const sgc_counter* c, const uint8_t* d_lines, const uint32_t* d_off, uint32_t off_base, uint32_t stride,
                 uint32_t read_len, uint64_t first, uint64_t n_reads, int32_t* d_assign, cudaStream_t stream,
                 uint32_t* grid_out) {
  OpaqueCountParams p{};
  p.v = view_of(c->lib);
  p.lines = d_lines;
  p.line_off = d_off;
  p.n_reads = n_reads - first;
  p.first_read = first;
  p.stride = stride;
  p.read_len = read_len;
  p.off_base = off_base;
  p.offset = (int)c->offset;
  p.with_perm = c->lib->with_perm;
  p.reverse = c->is_reverse != 0;
  p.recursion = c->recursion != 0;
  p.rc_mode = (uint8_t)c->rc_mode;
  p.state = c->d_state;
  p.assign_out = d_assign;
  uint64_t blocks = (p.n_reads + 255) / 256;
  const uint64_t cap = (uint64_t)c->lib->sm_count * 8;
  if (blocks > cap) blocks = cap;
  count_opaque_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
  SGC_CUDA_TRY(cudaGetLastError());
  *grid_out = (uint32_t)blocks;
  return SGC_OK;
}

}  // namespace sgc
