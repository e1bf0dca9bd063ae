// Microbenchmark (tuning aid): does a global -> L1 prefetch stick on sm_100a?  Measures the
// latency of a dependent 8-byte load from an L2-resident table after (0) nothing, (1)
// prefetch.global.L1, (2) cp.async.ca of the same word into shared memory, (3) a plain earlier
// load of the same word (.nc), each followed by a ~3000-cycle pause.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t ld_nc(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_ca(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int MODE, bool CA>
__global__ void probe(const uint64_t* table, uint32_t mask, uint32_t iters, unsigned long long* out) {
  __shared__ uint64_t dump[256];
  uint32_t x = 0x9E3779B1u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long total = 0, sink = 0;
  for (uint32_t i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    const uint64_t* p = table + ((x >> 4) & mask);
    if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p) : "memory");
    if (MODE == 2) {
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(&dump[threadIdx.x]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(p) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (MODE == 3) sink += ld_nc(p);
    if (MODE == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
    if (MODE != 0) {
      const long long t = clock64();
      while (clock64() - t < 3000) {}
      if (MODE == 2) asm volatile("cp.async.wait_all;" ::: "memory");
    }
    const long long t0 = clock64();
    const uint64_t v = CA ? ld_ca(p) : ld_nc(p);
    sink += v;
    // a real instruction whose address needs v: the clock read behind it issues after v arrived
    ((volatile uint64_t*)dump)[((uint32_t)(v >> 60) + threadIdx.x) & 255] = i;
    const long long t1 = clock64();
    total += (unsigned long long)(t1 - t0);
  }
  if (threadIdx.x == 0) {
    out[blockIdx.x * 2] = total;
    out[blockIdx.x * 2 + 1] = sink + dump[0];
  }
}

int main() {
  const uint32_t words = 1u << 22;  // 32 MB
  uint64_t* t;
  unsigned long long* out;
  cudaMalloc(&t, (size_t)words * 8);
  cudaMemset(t, 1, (size_t)words * 8);
  cudaMalloc(&out, 4096);
  const uint32_t iters = 2000;
  auto run = [&](const char* name, auto kern) {
    for (int rep = 0; rep < 2; ++rep) kern<<<8, 32>>>(t, words - 1, iters, out);
    unsigned long long h[16];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s %7.1f cycles per load  (%s)\n", name, (double)h[0] / iters, cudaGetErrorString(e));
  };
  run("none, ld.nc", probe<0, false>);
  run("none, ld.ca", probe<0, true>);
  run("prefetch.global.L1, ld.nc", probe<1, false>);
  run("prefetch.global.L1, ld.ca", probe<1, true>);
  run("cp.async.ca 8B, ld.nc", probe<2, false>);
  run("cp.async.ca 8B, ld.ca", probe<2, true>);
  run("earlier ld.nc, ld.nc", probe<3, false>);
  run("earlier ld.nc, ld.ca", probe<3, true>);
  run("prefetch.global.L2, ld.nc", probe<4, false>);
  return 0;
}
