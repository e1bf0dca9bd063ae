// library.cu — Library::from_reader + Permuter::new as device table build (kernels K2a/K2b),
// and the composed token lookup.
//
// Replaces /root/reference/src/library.rs:17-99 and permutes.rs:47-158.  The reference's
// stateful insert algorithm is order independent in its observable effect:
// a token that is not a library member resolves iff exactly ONE library sequence lies at
// Hamming distance 1.  The build below realises that directly and in parallel:
//   1. pack every guide 2-bit, reject bytes outside A,C,G,T                    (pack_library)
//   2. insert the guides as `library member` slots, detect duplicates          (insert_exact)
//   3. insert the 3k ACGT variants of every guide; a variant whose key already
//      belongs to a member is dropped, one that meets a different parent is
//      marked AMBIG in place                                                    (insert_variants)
// Variants with an 'N' are not stored: a read window with one N is answered by four member
// probes (common.cuh window_lookup), which is exactly the set of parents of that token.
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "internal.h"

namespace sgc {

thread_local std::string g_last_error;

int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  cudaGetLastError();  // clear the sticky flag of non-fatal errors
  return set_error(SGC_ERR_CUDA, buf);
}

namespace {

constexpr uint64_t kWideEmptyKey = ~0ull;

struct BuildStatus {
  unsigned int bad_guide;   // smallest guide index holding a non-ACGT byte, or 0xFFFFFFFF
  unsigned int dup_guide;   // smallest guide index that duplicates another sequence
  unsigned long long n_variants, n_ambiguous;
};

// K2a: one thread per guide.
__global__ void pack_library_kernel(const uint8_t* __restrict__ seqs, uint32_t n, uint32_t k,
                                    uint64_t* __restrict__ keys, BuildStatus* st) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* s = seqs + (size_t)i * k;
  uint64_t key = 0;
  bool bad = false;
  for (uint32_t j = 0; j < k; ++j) {
    uint8_t c = s[j];
    bad |= !is_acgt(c);
    key |= (uint64_t)code_of(c) << (2 * j);
  }
  keys[i] = key;
  if (bad) atomicMin(&st->bad_guide, i);
}

// ---- narrow table (k <= 20): key and meta share one word, one CAS claims a slot ---------
__device__ __forceinline__ void narrow_insert(uint64_t* slots, uint32_t n_buckets, uint64_t key, uint32_t idx,
                                              bool variant, BuildStatus* st) {
  const uint64_t val = make_meta(idx, variant) | key;
  uint32_t b = bucket_of(key, n_buckets);
  for (;;) {
    unsigned long long* base = reinterpret_cast<unsigned long long*>(slots + (size_t)b * 4);
    for (int s = 0; s < 4; ++s) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(base + s);
      if (cur == 0) {
        cur = atomicCAS(base + s, 0ull, (unsigned long long)val);
        if (cur == 0) return;  // claimed
      }
      if ((cur & kKeyMaskNarrow) == key) {
        if (!variant) {
          atomicMin(&st->dup_guide, idx);  // library.rs:92
        } else if (meta_variant(cur) && meta_idx(cur) != idx) {
          atomicOr(base + s, (unsigned long long)kAmbig << kMetaShift);  // permutes.rs:149-152
        }
        // a variant that equals a library member is never stored (unreachable in the reference)
        return;
      }
    }
    b = (b + 1 == n_buckets) ? 0 : b + 1;
  }
}

// ---- wide table (k = 21..30): two words per slot, built in two passes -------------------
// pass 1 claims key words, pass 2 (a later launch, so every key is visible) fills metas.
__device__ __forceinline__ unsigned long long* wide_claim(uint64_t* slots, uint32_t n_buckets, uint64_t key,
                                                          bool insert) {
  uint32_t b = bucket_of(key, n_buckets);
  for (;;) {
    unsigned long long* base = reinterpret_cast<unsigned long long*>(slots + (size_t)b * 4);
    for (int s = 0; s < 2; ++s) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(base + 2 * s);
      if (cur == kWideEmptyKey) {
        if (!insert) return nullptr;
        cur = atomicCAS(base + 2 * s, (unsigned long long)kWideEmptyKey, (unsigned long long)key);
        if (cur == kWideEmptyKey) return base + 2 * s + 1;
      }
      if (cur == key) return base + 2 * s + 1;
    }
    b = (b + 1 == n_buckets) ? 0 : b + 1;
  }
}

__device__ __forceinline__ void wide_set_meta(unsigned long long* meta, uint32_t idx, bool variant, BuildStatus* st) {
  unsigned long long old = atomicCAS(meta, 0ull, (unsigned long long)make_meta(idx, variant));
  if (old == 0) return;
  if (!variant) {
    atomicMin(&st->dup_guide, idx);
  } else if (meta_variant(old) && meta_idx(old) != idx) {
    atomicOr(meta, (unsigned long long)kAmbig << kMetaShift);
  }
}

__global__ void wide_init_kernel(uint64_t* slots, size_t n_slots) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots) {
    slots[2 * i] = kWideEmptyKey;
    slots[2 * i + 1] = 0;
  }
}

// mode 0: narrow insert; 1: wide pass 1 (claim keys); 2: wide pass 2 (metas)
template <int MODE>
__global__ void insert_exact_kernel(uint64_t* slots, uint32_t n_buckets, const uint64_t* __restrict__ keys,
                                    uint32_t n, BuildStatus* st) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (MODE == 0) narrow_insert(slots, n_buckets, keys[i], i, false, st);
  if (MODE == 1) wide_claim(slots, n_buckets, keys[i], true);
  if (MODE == 2) wide_set_meta(wide_claim(slots, n_buckets, keys[i], false), i, false, st);
}

// K2b: one thread per (guide, position): the three ACGT substitutions at that position
// (permutes.rs:78-107 restricted to A,C,G,T; the N variants need no storage).
template <int MODE>
__global__ void insert_variants_kernel(uint64_t* slots, uint32_t n_buckets, const uint64_t* __restrict__ keys,
                                       uint32_t n, uint32_t k, BuildStatus* st) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)n * k) return;
  uint32_t i = (uint32_t)(t / k), pos = (uint32_t)(t % k);
  uint64_t key = keys[i];
  for (uint64_t d = 1; d < 4; ++d) {
    uint64_t v = key ^ (d << (2 * pos));
    if (MODE == 0) narrow_insert(slots, n_buckets, v, i, true, st);
    if (MODE == 1) wide_claim(slots, n_buckets, v, true);
    if (MODE == 2) wide_set_meta(wide_claim(slots, n_buckets, v, false), i, true, st);
  }
}

__global__ void table_stats_kernel(const uint64_t* __restrict__ slots, size_t n_slots, bool wide, BuildStatus* st) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long var = 0, amb = 0;
  if (i < n_slots) {
    uint64_t m = wide ? slots[2 * i + 1] : slots[i];
    if (m != 0 && meta_variant(m)) {
      if (meta_idx(m) == kAmbig)
        amb = 1;
      else
        var = 1;
    }
  }
  var = __reduce_add_sync(0xffffffffu, (unsigned)var);
  amb = __reduce_add_sync(0xffffffffu, (unsigned)amb);
  if ((threadIdx.x & 31) == 0) {
    if (var) atomicAdd(&st->n_variants, var);
    if (amb) atomicAdd(&st->n_ambiguous, amb);
  }
}

// one thread per slot: every occupied key sets its four bits
__global__ void bloom_build_kernel(const uint64_t* __restrict__ slots, size_t n_slots, bool wide,
                                   unsigned long long* bloom, uint32_t n_words) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  uint64_t key;
  if (wide) {
    if (slots[2 * i + 1] == 0) return;
    key = slots[2 * i];
  } else {
    if (slots[i] == 0) return;
    key = slots[i] & kKeyMaskNarrow;
  }
  uint32_t word;
  uint64_t mask;
  bloom_locate(key, n_words, word, mask);
  atomicOr(bloom + word, (unsigned long long)mask);
}

// composed lookup of raw k-byte tokens (sgc_library_lookup)
__global__ void lookup_tokens_kernel(TableView t, bool with_perm, const uint8_t* __restrict__ tokens,
                                     uint64_t n_tokens, int32_t* __restrict__ idx_out, uint8_t* __restrict__ kind_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tokens) return;
  const uint8_t* s = tokens + i * t.k;
  uint64_t key = 0;
  int nbad = 0, bad_pos = 0;
  bool wild = false;
  for (uint32_t j = 0; j < t.k; ++j) {
    uint8_t c = s[j];
    key |= (uint64_t)code_of(c) << (2 * j);
    if (!is_acgt(c)) {
      ++nbad;
      bad_pos = (int)j;
      wild = (c == 'N');
    }
  }
  int kind = 0;
  int32_t hit = window_lookup(t, with_perm, key, nbad, bad_pos, wild, &kind);
  idx_out[i] = hit;
  if (kind_out) kind_out[i] = hit == kMiss ? 0 : (uint8_t)kind;
}

template <typename T>
struct DeviceBuffer {
  T* p = nullptr;
  ~DeviceBuffer() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)); }
};

inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace
}  // namespace sgc

using namespace sgc;

extern "C" {

const char* sgc_last_error(void) { return g_last_error.c_str(); }
int sgc_abi_version(void) { return SGC_ABI_VERSION; }

int sgc_device_count(int* n) {
  if (!n) return set_error(SGC_ERR_INVALID_ARG, "n is NULL");
  SGC_CUDA_TRY(cudaGetDeviceCount(n));
  return SGC_OK;
}

int sgc_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return set_error(SGC_ERR_INVALID_ARG, "ptr is NULL");
  SGC_CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
  return SGC_OK;
}
int sgc_host_free(void* ptr) {
  SGC_CUDA_TRY(cudaFreeHost(ptr));
  return SGC_OK;
}

void sgc_library_destroy(sgc_library* lib) {
  if (!lib) return;
  DeviceGuard g(lib->device);
  if (lib->d_front != lib->d_slots) cudaFree(lib->d_front);
  cudaFree(lib->d_bloom);
  cudaFree(lib->d_slots);
  cudaFree(lib->d_keys);
  cudaFree(lib->d_lib_hist);
  delete lib;
}

int sgc_library_create(int device, const uint8_t* seqs, uint32_t n, uint32_t k, int with_permutations,
                       sgc_library** out) {
  if (!seqs || !out) return set_error(SGC_ERR_INVALID_ARG, "seqs/out is NULL");
  if (n == 0) return set_error(SGC_ERR_EMPTY_READER, "empty library (library.rs:74 unwraps on an empty table)");
  if (k == 0 || k > kMaxK) return set_error(SGC_ERR_K_UNSUPPORTED, "guide length must be 1..30");
  if (n > SGC_MAX_GUIDES) return set_error(SGC_ERR_TOO_MANY_GUIDES, "too many guides");
  int ndev = 0;
  SGC_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_error(SGC_ERR_INVALID_ARG, "no such device");
  DeviceGuard guard(device);
  if (!guard.ok()) return set_error(SGC_ERR_CUDA, "cudaSetDevice failed");

  sgc_library* lib = new sgc_library();
  lib->device = device;
  lib->n = n;
  lib->k = k;
  lib->with_perm = with_permutations != 0;
  lib->wide = k > kNarrowMaxK;
  struct Cleanup {
    sgc_library* l;
    ~Cleanup() {
      if (l) sgc_library_destroy(l);
    }
  } cleanup{lib};
  SGC_CUDA_TRY(cudaDeviceGetAttribute(&lib->sm_count, cudaDevAttrMultiProcessorCount, device));

  // table geometry: <= 50 % load, whole 32-byte buckets
  const uint64_t entries = lib->with_perm ? (uint64_t)n * (1 + 3ull * k) : n;
  const uint32_t per_bucket = lib->wide ? 2 : 4;
  uint64_t n_buckets = (entries * 2 + per_bucket - 1) / per_bucket;
  if (n_buckets < 64) n_buckets = 64;
  if (n_buckets > 0xFFFFFFF0ull) return set_error(SGC_ERR_TOO_MANY_GUIDES, "table too large");
  lib->n_buckets = (uint32_t)n_buckets;
  const size_t n_slots = (size_t)n_buckets * per_bucket;
  const size_t table_bytes = (size_t)n_buckets * 32;

  DeviceBuffer<uint8_t> d_seqs;
  DeviceBuffer<BuildStatus> d_st;
  SGC_CUDA_TRY(d_seqs.alloc((size_t)n * k));
  SGC_CUDA_TRY(d_st.alloc(1));
  SGC_CUDA_TRY(cudaMalloc(&lib->d_keys, (size_t)n * sizeof(uint64_t)));
  SGC_CUDA_TRY(cudaMalloc(&lib->d_slots, table_bytes));
  SGC_CUDA_TRY(cudaMalloc(&lib->d_lib_hist, (size_t)k * 4 * sizeof(uint32_t)));
  SGC_CUDA_TRY(cudaMemcpy(d_seqs.p, seqs, (size_t)n * k, cudaMemcpyHostToDevice));
  BuildStatus st0{0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0};
  SGC_CUDA_TRY(cudaMemcpy(d_st.p, &st0, sizeof st0, cudaMemcpyHostToDevice));

  cudaEvent_t e0, e1;
  SGC_CUDA_TRY(cudaEventCreate(&e0));
  SGC_CUDA_TRY(cudaEventCreate(&e1));
  SGC_CUDA_TRY(cudaEventRecord(e0, 0));
  const unsigned T = 256;
  if (lib->wide)
    wide_init_kernel<<<blocks_for(n_slots, T), T>>>(lib->d_slots, n_slots);
  else
    SGC_CUDA_TRY(cudaMemsetAsync(lib->d_slots, 0, table_bytes, 0));
  pack_library_kernel<<<blocks_for(n, T), T>>>(d_seqs.p, n, k, lib->d_keys, d_st.p);
  const uint64_t nv = (uint64_t)n * k;
  if (!lib->wide) {
    insert_exact_kernel<0><<<blocks_for(n, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, d_st.p);
    if (lib->with_perm)
      insert_variants_kernel<0><<<blocks_for(nv, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, k, d_st.p);
  } else {
    insert_exact_kernel<1><<<blocks_for(n, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, d_st.p);
    if (lib->with_perm)
      insert_variants_kernel<1><<<blocks_for(nv, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, k, d_st.p);
    insert_exact_kernel<2><<<blocks_for(n, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, d_st.p);
    if (lib->with_perm)
      insert_variants_kernel<2><<<blocks_for(nv, T), T>>>(lib->d_slots, lib->n_buckets, lib->d_keys, n, k, d_st.p);
  }
  table_stats_kernel<<<blocks_for(n_slots, T), T>>>(lib->d_slots, n_slots, lib->wide, d_st.p);
  // front table: members only, <= 25 % load (skipped when the main table already is that)
  if (lib->with_perm) {
    uint64_t fb = ((uint64_t)n * 4 + per_bucket - 1) / per_bucket;
    if (fb < 64) fb = 64;
    lib->front_buckets = (uint32_t)fb;
    const size_t f_slots = (size_t)fb * per_bucket;
    SGC_CUDA_TRY(cudaMalloc(&lib->d_front, (size_t)fb * 32));
    if (lib->wide) {
      wide_init_kernel<<<blocks_for(f_slots, T), T>>>(lib->d_front, f_slots);
      insert_exact_kernel<1><<<blocks_for(n, T), T>>>(lib->d_front, lib->front_buckets, lib->d_keys, n, d_st.p);
      insert_exact_kernel<2><<<blocks_for(n, T), T>>>(lib->d_front, lib->front_buckets, lib->d_keys, n, d_st.p);
    } else {
      SGC_CUDA_TRY(cudaMemsetAsync(lib->d_front, 0, (size_t)fb * 32, 0));
      insert_exact_kernel<0><<<blocks_for(n, T), T>>>(lib->d_front, lib->front_buckets, lib->d_keys, n, d_st.p);
    }
    // Bloom filter: one 64-bit word per 4 expected keys (2 bytes per key, false positives < 1 %)
    lib->n_bloom_words = (uint32_t)std::max<uint64_t>(entries / 4, 64);
    SGC_CUDA_TRY(cudaMalloc(&lib->d_bloom, (size_t)lib->n_bloom_words * 8));
    SGC_CUDA_TRY(cudaMemsetAsync(lib->d_bloom, 0, (size_t)lib->n_bloom_words * 8, 0));
    bloom_build_kernel<<<blocks_for(n_slots, T), T>>>(lib->d_slots, n_slots, lib->wide,
                                                      reinterpret_cast<unsigned long long*>(lib->d_bloom),
                                                      lib->n_bloom_words);
  } else {
    lib->d_front = lib->d_slots;
    lib->front_buckets = lib->n_buckets;
  }
  // library positional histogram for the offset detector: records 1..n-1 (offsetter.rs:57,190-191)
  SGC_CUDA_TRY(cudaMemsetAsync(lib->d_lib_hist, 0, (size_t)k * 4 * sizeof(uint32_t), 0));
  if (n > 1) {
    int rc = position_counts_device(d_seqs.p + k, nullptr, k, k, n - 1, k, lib->d_lib_hist, 0);
    if (rc) return rc;
  }
  SGC_CUDA_TRY(cudaEventRecord(e1, 0));
  SGC_CUDA_TRY(cudaGetLastError());
  SGC_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0;
  SGC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);

  BuildStatus st;
  SGC_CUDA_TRY(cudaMemcpy(&st, d_st.p, sizeof st, cudaMemcpyDeviceToHost));
  if (st.bad_guide != 0xFFFFFFFFu) {
    char buf[160];
    snprintf(buf, sizeof buf, "library sequence %u holds a byte outside A,C,G,T", st.bad_guide);
    return set_error(SGC_ERR_NON_ACGT_LIBRARY, buf);
  }
  if (st.dup_guide != 0xFFFFFFFFu) {
    std::string s((const char*)seqs + (size_t)st.dup_guide * k, k);
    return set_error(SGC_ERR_DUPLICATE_SEQUENCE, "Unexpected duplicate sequence in library found: " + s);
  }
  lib->info.n_guides = n;
  lib->info.k = k;
  lib->info.with_permutations = lib->with_perm;
  lib->info.device = device;
  lib->info.n_variants = st.n_variants;
  lib->info.n_ambiguous = st.n_ambiguous;
  lib->info.n_slots = n_slots;
  lib->info.table_bytes =
      table_bytes + (lib->with_perm ? (size_t)lib->front_buckets * 32 + (size_t)lib->n_bloom_words * 8 : 0);
  lib->info.build_ms = ms;
  cleanup.l = nullptr;
  *out = lib;
  return SGC_OK;
}

int sgc_library_get_info(const sgc_library* lib, sgc_library_info* out) {
  if (!lib || !out) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  *out = lib->info;
  return SGC_OK;
}

int sgc_library_lookup(const sgc_library* lib, const uint8_t* tokens, uint64_t n_tokens, int32_t* idx_out,
                       uint8_t* kind_out) {
  if (!lib || !tokens || !idx_out) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  if (n_tokens == 0) return SGC_OK;
  DeviceGuard guard(lib->device);
  DeviceBuffer<uint8_t> d_tok, d_kind;
  DeviceBuffer<int32_t> d_idx;
  SGC_CUDA_TRY(d_tok.alloc(n_tokens * lib->k));
  SGC_CUDA_TRY(d_idx.alloc(n_tokens));
  if (kind_out) SGC_CUDA_TRY(d_kind.alloc(n_tokens));
  SGC_CUDA_TRY(cudaMemcpy(d_tok.p, tokens, n_tokens * lib->k, cudaMemcpyHostToDevice));
  lookup_tokens_kernel<<<blocks_for(n_tokens, 256), 256>>>(lib->view(), lib->with_perm, d_tok.p, n_tokens, d_idx.p,
                                                            d_kind.p);
  SGC_CUDA_TRY(cudaGetLastError());
  SGC_CUDA_TRY(cudaMemcpy(idx_out, d_idx.p, n_tokens * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (kind_out) SGC_CUDA_TRY(cudaMemcpy(kind_out, d_kind.p, n_tokens, cudaMemcpyDeviceToHost));
  return SGC_OK;
}

}  // extern "C"
