// offset.cu — the entropy offset/orientation detector (kernels K1/K1b).
//
// Replaces /root/reference/src/offsetter.rs:37-163,185-210.
//   K1  position_counts_kernel : per-position base histogram of the subsampled reads
//                                (integer exact; offsetter.rs:55-79)
//   K1b entropy_argmin_kernel  : normalise -> Shannon entropy (ln) -> windowed MSE against the
//                                library entropy, forward and reversed -> first argmin of each
//                                -> Forward iff min_f < min_r (offsetter.rs:82-150)
// K1b runs in f64 with explicit round-to-nearest adds/multiplies (no FMA contraction) and
// the same summation order as the reference, one warp, lanes over candidate windows.
#include <vector>

#include "internal.h"

namespace sgc {
namespace {

constexpr int kHistThreads = 128;

// One thread per read position (column); a block walks a slab of reads, so consecutive
// threads read consecutive bytes of one read and no atomics are needed until the flush.
__global__ void position_counts_kernel(const uint8_t* __restrict__ lines, const uint32_t* __restrict__ line_off,
                                       uint32_t stride, uint32_t read_len, uint64_t n_reads, uint32_t size,
                                       uint32_t reads_per_block, uint32_t* __restrict__ hist) {
  const uint64_t r0 = (uint64_t)blockIdx.x * reads_per_block;
  uint64_t r1 = r0 + reads_per_block;
  if (r1 > n_reads) r1 = n_reads;
  for (uint32_t col = threadIdx.x; col < size; col += blockDim.x) {
    uint32_t a = 0, c = 0, g = 0, t = 0;
    for (uint64_t r = r0; r < r1; ++r) {
      uint64_t start;
      uint32_t len;
      if (line_off) {
        start = line_off[r];
        len = line_off[r + 1] - line_off[r] - 1;
      } else {
        start = r * stride;
        len = read_len;
      }
      if (col >= len) continue;  // .take(size) over a shorter read (offsetter.rs:63)
      uint8_t b = lines[start + col];
      // base_map, offsetter.rs:42-50; anything else counts for all four (offsetter.rs:70-74)
      bool other = !is_acgt(b);
      a += (b == 'A') | other;
      c += (b == 'C') | other;
      g += (b == 'G') | other;
      t += (b == 'T') | other;
    }
    if (a) atomicAdd(&hist[col * 4 + 0], a);
    if (c) atomicAdd(&hist[col * 4 + 1], c);
    if (g) atomicAdd(&hist[col * 4 + 2], g);
    if (t) atomicAdd(&hist[col * 4 + 3], t);
  }
}

struct OffsetResult {
  int is_reverse;
  unsigned int index;
  int nan;
  double min_forward, min_reverse;
};

// -(sum_j p_j ln p_j), columns left to right, 0 ln 0 := 0 (offsetter.rs:82-94)
__device__ double row_entropy(const uint32_t* h) {
  double m0 = (double)h[0], m1 = (double)h[1], m2 = (double)h[2], m3 = (double)h[3];
  double sum = __dadd_rn(__dadd_rn(__dadd_rn(m0, m1), m2), m3);
  double m[4] = {m0, m1, m2, m3};
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double p = __ddiv_rn(m[j], sum);  // 0/0 = NaN, as in the reference
    double term = (p == 0.0) ? 0.0 : __dmul_rn(p, log(p));
    acc = __dadd_rn(acc, term);
  }
  return -acc;
}

// One warp.  Shared: href[k], hcmp[size].
__global__ void entropy_argmin_kernel(const uint32_t* __restrict__ lib_hist, uint32_t k,
                                      const uint32_t* __restrict__ cmp_hist, uint32_t size, OffsetResult* out) {
  extern __shared__ double sh[];
  double* href = sh;
  double* hcmp = sh + k;
  const int lane = threadIdx.x;
  for (uint32_t i = lane; i < k; i += 32) href[i] = row_entropy(lib_hist + 4 * i);
  for (uint32_t i = lane; i < size; i += 32) hcmp[i] = row_entropy(cmp_hist + 4 * i);
  __syncwarp();

  const uint32_t n_win = size - k + 1;
  double best[2] = {0.0, 0.0};
  uint32_t arg[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
  int nan = 0;
  for (int dir = 0; dir < 2; ++dir) {
    for (uint32_t x = lane; x < n_win; x += 32) {
      // mean_sq_err of href against window x of hcmp (dir 0) or of reversed hcmp (dir 1)
      double acc = 0.0;
      for (uint32_t i = 0; i < k; ++i) {
        double b = dir == 0 ? hcmp[x + i] : hcmp[size - 1 - (x + i)];
        double d = __dadd_rn(href[i], -b);
        acc = __dadd_rn(acc, __dmul_rn(d, d));
      }
      double mse = __ddiv_rn(acc, (double)k);
      if (mse != mse) nan = 1;
      if (arg[dir] == 0xFFFFFFFFu || mse < best[dir]) {  // increasing x: keeps the first minimum
        best[dir] = mse;
        arg[dir] = x;
      }
    }
    // warp argmin, ties to the smaller index (QuantileExt::argmin keeps the first minimum)
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best[dir], o);
      uint32_t oa = __shfl_xor_sync(0xffffffffu, arg[dir], o);
      bool take = oa != 0xFFFFFFFFu && (arg[dir] == 0xFFFFFFFFu || ob < best[dir] || (ob == best[dir] && oa < arg[dir]));
      if (take) {
        best[dir] = ob;
        arg[dir] = oa;
      }
    }
  }
  nan = __any_sync(0xffffffffu, nan);
  if (lane == 0) {
    out->nan = nan;
    out->min_forward = best[0];
    out->min_reverse = best[1];
    if (best[0] < best[1]) {  // strict: ties go to Reverse (offsetter.rs:143-149)
      out->is_reverse = 0;
      out->index = arg[0];
    } else {
      out->is_reverse = 1;
      out->index = arg[1];
    }
  }
}

}  // namespace

int position_counts_device(const uint8_t* d_lines, const uint32_t* d_line_off, uint32_t stride, uint32_t read_len,
                           uint64_t n_reads, uint32_t size, uint32_t* d_hist, cudaStream_t stream) {
  if (n_reads == 0 || size == 0) return SGC_OK;
  // ~64 reads per block keeps every SM busy on a 5000-read subsample
  uint32_t reads_per_block = 64;
  uint64_t blocks = (n_reads + reads_per_block - 1) / reads_per_block;
  if (blocks > 65535) {
    reads_per_block = (uint32_t)((n_reads + 65534) / 65535);
    blocks = (n_reads + reads_per_block - 1) / reads_per_block;
  }
  position_counts_kernel<<<(unsigned)blocks, kHistThreads, 0, stream>>>(d_lines, d_line_off, stride, read_len,
                                                                         n_reads, size, reads_per_block, d_hist);
  SGC_CUDA_TRY(cudaGetLastError());
  return SGC_OK;
}

namespace {

struct Staged {
  uint8_t* d_lines = nullptr;
  uint32_t* d_off = nullptr;
  uint32_t* d_hist = nullptr;
  OffsetResult* d_res = nullptr;
  ~Staged() {
    cudaFree(d_lines);
    cudaFree(d_off);
    cudaFree(d_hist);
    cudaFree(d_res);
  }
};

// Validates a host batch description and returns the first read's length.
int first_read_len(const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off, uint32_t stride,
                   uint32_t read_len, uint64_t n_reads, uint32_t* first) {
  if (!lines) return set_error(SGC_ERR_INVALID_ARG, "lines is NULL");
  if (n_reads == 0) return set_error(SGC_ERR_EMPTY_READER, "empty reader (offsetter.rs:38)");
  if (line_off) {
    if (n_bytes >= (1ull << 32)) return set_error(SGC_ERR_BATCH_TOO_LARGE, "batch must stay below 4 GiB");
    if (line_off[n_reads] > n_bytes || line_off[1] <= line_off[0])
      return set_error(SGC_ERR_INVALID_ARG, "line offsets are inconsistent with n_bytes");
    *first = line_off[1] - line_off[0] - 1;
  } else {
    if (read_len > stride || (n_reads - 1) * (uint64_t)stride + read_len > n_bytes)
      return set_error(SGC_ERR_INVALID_ARG, "fixed-stride batch does not fit n_bytes");
    *first = read_len;
  }
  return SGC_OK;
}

// Copies the batch to the device and histograms reads 1..n-1 against size = len(read 0).
int stage_and_count(int device, const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off, uint32_t stride,
                    uint32_t read_len, uint64_t n_reads, uint32_t size, Staged& s) {
  (void)device;
  SGC_CUDA_TRY(cudaMalloc(&s.d_lines, n_bytes ? n_bytes : 1));
  SGC_CUDA_TRY(cudaMemcpy(s.d_lines, lines, n_bytes, cudaMemcpyHostToDevice));
  if (line_off) {
    SGC_CUDA_TRY(cudaMalloc(&s.d_off, (n_reads + 1) * sizeof(uint32_t)));
    SGC_CUDA_TRY(cudaMemcpy(s.d_off, line_off, (n_reads + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  SGC_CUDA_TRY(cudaMalloc(&s.d_hist, (size_t)(size ? size : 1) * 4 * sizeof(uint32_t)));
  SGC_CUDA_TRY(cudaMemset(s.d_hist, 0, (size_t)(size ? size : 1) * 4 * sizeof(uint32_t)));
  if (n_reads > 1) {
    // the first record is consumed for its length and not counted (offsetter.rs:57)
    const uint8_t* d_first = line_off ? s.d_lines : s.d_lines + stride;
    const uint32_t* d_off = line_off ? s.d_off + 1 : nullptr;
    int rc = position_counts_device(d_first, d_off, stride, read_len, n_reads - 1, size, s.d_hist, 0);
    if (rc) return rc;
  }
  return SGC_OK;
}

}  // namespace
}  // namespace sgc

using namespace sgc;

extern "C" {

int sgc_position_counts(int device, const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off, uint32_t stride,
                        uint32_t read_len, uint64_t n_reads, uint32_t* out, uint32_t out_cap, uint32_t* size) {
  if (!size) return set_error(SGC_ERR_INVALID_ARG, "size is NULL");
  uint32_t first = 0;
  int rc = first_read_len(lines, n_bytes, line_off, stride, read_len, n_reads, &first);
  if (rc) return rc;
  *size = first;
  if (!out) return SGC_OK;  // size query
  if (out_cap < first) return set_error(SGC_ERR_INVALID_ARG, "out_cap is smaller than the first read");
  DeviceGuard guard(device);
  if (!guard.ok()) return set_error(SGC_ERR_CUDA, "cudaSetDevice failed");
  Staged s;
  rc = stage_and_count(device, lines, n_bytes, line_off, stride, read_len, n_reads, first, s);
  if (rc) return rc;
  SGC_CUDA_TRY(cudaMemcpy(out, s.d_hist, (size_t)first * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return SGC_OK;
}

int sgc_offset_detect(const sgc_library* lib, const uint8_t* lines, uint64_t n_bytes, const uint32_t* line_off,
                      uint32_t stride, uint32_t read_len, uint64_t n_reads, int* is_reverse, uint32_t* index) {
  if (!lib || !is_reverse || !index) return set_error(SGC_ERR_INVALID_ARG, "NULL argument");
  uint32_t size = 0;
  int rc = first_read_len(lines, n_bytes, line_off, stride, read_len, n_reads, &size);
  if (rc) return rc;
  if (size < lib->k)  // minimize_mse bails before any arithmetic (offsetter.rs:154-156)
    return set_error(SGC_ERR_READ_TOO_SHORT,
                     "Sequences in reference library are larger than the sequences in input.");
  DeviceGuard guard(lib->device);
  if (!guard.ok()) return set_error(SGC_ERR_CUDA, "cudaSetDevice failed");
  Staged s;
  rc = stage_and_count(lib->device, lines, n_bytes, line_off, stride, read_len, n_reads, size, s);
  if (rc) return rc;
  SGC_CUDA_TRY(cudaMalloc(&s.d_res, sizeof(OffsetResult)));
  size_t smem = (size_t)(lib->k + size) * sizeof(double);
  if (smem > 48 * 1024)
    SGC_CUDA_TRY(cudaFuncSetAttribute(entropy_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  entropy_argmin_kernel<<<1, 32, smem>>>(lib->d_lib_hist, lib->k, s.d_hist, size, s.d_res);
  SGC_CUDA_TRY(cudaGetLastError());
  OffsetResult res;
  SGC_CUDA_TRY(cudaMemcpy(&res, s.d_res, sizeof res, cudaMemcpyDeviceToHost));
  if (res.nan) return set_error(SGC_ERR_NAN_ENTROPY, "Unexpected minmax error in entropy (NaN; offsetter.rs:123-141)");
  *is_reverse = res.is_reverse;
  *index = res.index;
  return SGC_OK;
}

}  // extern "C"
