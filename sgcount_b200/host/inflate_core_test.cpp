// inflate_core_test <file.gz> — decodes every member of a gzip file with the device decoder's core
// (csrc/inflate_core.h compiled for the host) and prints "<members> <bytes> <fnv1a>", or the
// failing status.  Test helper for tests/test_device_inflate_core.py; zlib is the comparison.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../csrc/inflate_core.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<unsigned char> data;
  unsigned char buf[1 << 16];
  for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + got);
  fclose(f);
  const size_t n_data = data.size();
  data.resize(n_data + 8);  // the decoder loads aligned words: up to 3 bytes past the end are touched
  const size_t cap = argc > 2 ? (size_t)atoll(argv[2]) : (size_t)1 << 30;
  std::vector<unsigned char> out(cap < ((size_t)64 << 20) ? cap : ((size_t)64 << 20));
  sgc::inflate::PlainTables t;
  size_t pos = 0, members = 0;
  unsigned long long total = 0, h = 1469598103934665603ull;
  while (pos < n_data) {
    size_t used = 0, made = 0;
    uint32_t crc = 0, isize = 0;
    int rc;
    for (;;) {
      rc = sgc::inflate::gunzip_member(data.data() + pos, n_data - pos, out.data(), out.size(), t, &used, &made, &crc, &isize);
      if (rc != sgc::inflate::kOutputFull || out.size() >= cap) break;
      out.resize(out.size() * 2 < cap ? out.size() * 2 : cap);
    }
    if (rc != sgc::inflate::kOk) {
      printf("status %d at member %zu\n", rc, members);
      return 4;
    }
    if ((uint32_t)made != isize) {
      printf("isize mismatch at member %zu\n", members);
      return 5;
    }
    for (size_t i = 0; i < made; ++i) h = (h ^ out[i]) * 1099511628211ull;
    total += made;
    pos += used;
    ++members;
  }
  printf("%zu %llu %llx\n", members, total, h);
  return 0;
}
