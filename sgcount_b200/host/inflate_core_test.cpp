// inflate_core_test <file.gz> [capacity [alignment]] — decodes every member of a gzip file with the device decoder's core
// (csrc/inflate_core.h compiled for the host) and prints "<members> <bytes> <fnv1a>", or the
// failing status.  Test helper for tests/test_device_inflate_core.py; zlib is the comparison.
#include <cstdio>
#include <algorithm>
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
  data.resize(n_data + 64);  // the decoder prefetches aligned 16-byte blocks: up to 47 bytes past the end are touched
  const size_t cap = argc > 2 ? (size_t)atoll(argv[2]) : (size_t)1 << 30;
  // The member is decoded at every alignment in turn, between guard bytes: on the device its
  // neighbours in the text are written by other threads, so not one byte outside
  // [out, out + produced) may be touched (the writer stores aligned words).
  const size_t kGuard = 24;
  const size_t align0 = argc > 3 ? (size_t)atoll(argv[3]) : 0;  // alignment of the first member's output
  size_t room = cap < ((size_t)64 << 20) ? cap : ((size_t)64 << 20);
  std::vector<unsigned char> store(room + 2 * kGuard + 16);
  sgc::inflate::PlainTableStorage table_storage;
  sgc::inflate::PlainTables t{&table_storage};
  size_t pos = 0, members = 0;
  unsigned long long total = 0, h = 1469598103934665603ull;
  while (pos < n_data) {
    size_t used = 0, made = 0;
    uint32_t crc = 0, isize = 0;
    int rc;
    unsigned char* out = nullptr;
    for (;;) {
      std::fill(store.begin(), store.end(), (unsigned char)0xAB);
      out = store.data() + kGuard + ((members + align0) % 8);
      rc = sgc::inflate::gunzip_member(data.data() + pos, n_data - pos, out, room, t, &used, &made, &crc, &isize);
      if (rc != sgc::inflate::kOutputFull || room >= cap) break;
      room = room * 2 < cap ? room * 2 : cap;
      store.resize(room + 2 * kGuard + 16);
    }
    if (rc != sgc::inflate::kOk) {
      printf("status %d at member %zu\n", rc, members);
      return 4;
    }
    if ((uint32_t)made != isize) {
      printf("isize mismatch at member %zu\n", members);
      return 5;
    }
    for (unsigned char* p = store.data(); p < store.data() + store.size(); ++p)
      if ((p < out || p >= out + made) && *p != 0xAB) {
        printf("a byte outside the member's output was written (member %zu, offset %td)\n", members, p - out);
        return 6;
      }
    for (size_t i = 0; i < made; ++i) h = (h ^ out[i]) * 1099511628211ull;
    total += made;
    pos += used;
    ++members;
  }
  printf("%zu %llu %llx\n", members, total, h);
  return 0;
}
