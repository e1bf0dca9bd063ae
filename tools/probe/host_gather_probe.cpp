// Tuning aid: how fast can the host cut the guide-window span (24 of 76 bytes per read) out of
// sequence lines into a compact buffer, with T threads?  Decides whether a span-only host-to-device
// path could beat the 55 GB/s PCIe copy of the whole lines.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const size_t n = argc > 1 ? atoll(argv[1]) : 50000000;
  const size_t stride = 76, span = 24, at = 4;
  std::vector<char> src(n * stride + 64), dst(n * span + 64);
  memset(src.data(), 'A', src.size());
  memset(dst.data(), 0, dst.size());
  for (int T : {1, 4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> pool;
      for (int t = 0; t < T; ++t)
        pool.emplace_back([&, t] {
          const size_t a = n * t / T, b = n * (t + 1) / T;
          const char* s = src.data() + a * stride + at;
          char* d = dst.data() + a * span;
          for (size_t i = a; i < b; ++i, s += stride, d += span) {
            memcpy(d, s, 8);
            memcpy(d + 8, s + 8, 8);
            memcpy(d + 16, s + 16, 8);
          }
        });
      for (auto& th : pool) th.join();
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep) printf("threads %2d: %.1f ms  (%.1f GB/s of lines)\n", T, dt * 1e3, n * stride / dt / 1e9);
    }
  }
  return 0;
}
