// fastx_dump <path> <inflate threads> [seq] — record count and FNV-1a of ids+sequences as the
// host reader yields them (test helper for tests/test_host_reader.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fastx.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    sgh::FastxReader r(argv[1], (unsigned)atoi(argv[2]));
    const bool seq_only = argc > 3 && !strcmp(argv[3], "seq");
    const char *id = nullptr, *seq;
    size_t il = 0, sl;
    unsigned long long n = 0, h = 1469598103934665603ull;
    while (seq_only ? r.next_seq(seq, sl) : r.next(id, il, seq, sl)) {
      ++n;
      for (size_t i = 0; i < sl; ++i) h = (h ^ (unsigned char)seq[i]) * 1099511628211ull;
      h = (h ^ 0xFFu) * 1099511628211ull;
      if (!seq_only)
        for (size_t i = 0; i < il; ++i) h = (h ^ (unsigned char)id[i]) * 1099511628211ull;
    }
    printf("%llu %llx\n", n, h);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
}
