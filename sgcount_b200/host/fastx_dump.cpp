// fastx_dump <path> <inflate threads> [seq|blocks] — record count and FNV-1a of ids+sequences as the
// host reader yields them (test helper for tests/test_host_reader.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <vector>

#include "fastx.h"
#include "inflate.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    if (argc > 3 && !strcmp(argv[3], "inflate")) {  // the whole-member decoder alone, member after member
      FILE* f = fopen(argv[1], "rb");
      if (!f) return 2;
      std::vector<unsigned char> data;
      unsigned char buf[1 << 16];
      for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + got);
      fclose(f);
      size_t pos = 0;
      unsigned long long total = 0, h = 1469598103934665603ull;
      while (pos < data.size()) {
        sgh::Bytes out;
        size_t used = 0;
        if (!sgh::gunzip_member(data.data() + pos, data.size() - pos, out, used)) {
          printf("declined at %zu\n", pos);
          return 4;
        }
        for (char c : out) h = (h ^ (unsigned char)c) * 1099511628211ull;
        total += out.size();
        pos += used;
      }
      printf("%llu %llx\n", total, h);
      return 0;
    }
    if (argc > 3 && (!strcmp(argv[3], "blockcount") || !strcmp(argv[3], "seqcount"))) {  // timing: framing only
      unsigned long long n = 0, bytes = 0;
      if (!strcmp(argv[3], "blockcount")) {
        sgh::SeqBlockReader br(argv[1], (unsigned)atoi(argv[2]));
        sgh::SeqBlock b;
        while (br.next(b)) n += b.n, bytes += b.lines.size();
      } else {
        sgh::FastxReader r(argv[1], (unsigned)atoi(argv[2]));
        const char* seq;
        size_t sl;
        while (r.next_seq(seq, sl)) ++n, bytes += sl + 1;
      }
      printf("%llu %llu\n", n, bytes);
      return 0;
    }
    if (argc > 4 && !strcmp(argv[3], "start")) {  // fastx_first_record_start of the raw bytes of a file: fastx_dump <path> 1 start <2|4>
      FILE* f = fopen(argv[1], "rb");
      if (!f) return 2;
      std::vector<char> data;
      char buf[1 << 16];
      for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + got);
      fclose(f);
      const size_t at = sgh::fastx_first_record_start(data.data(), data.size(), atoi(argv[4]));
      if (at == SIZE_MAX)
        printf("none\n");
      else
        printf("%zu\n", at);
      return 0;
    }
    if (argc > 4 && !strcmp(argv[3], "bgzfindex")) {  // bgzf_index: fastx_dump <path> <threads> bgzfindex <min bytes for the walk in parts>
      FILE* f = fopen(argv[1], "rb");
      if (!f) return 2;
      std::vector<unsigned char> data;
      unsigned char buf[1 << 16];
      for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + got);
      fclose(f);
      std::vector<uint64_t> begin;
      std::vector<uint32_t> isize;
      bool in_parts = false;
      if (!sgh::bgzf_index(data.data(), data.size(), begin, isize, (unsigned)atoi(argv[2]), (size_t)atoll(argv[4]), &in_parts)) {
        printf("not bgzf\n");
        return 0;
      }
      unsigned long long h = 1469598103934665603ull;
      for (uint64_t b : begin) h = (h ^ b) * 1099511628211ull;
      for (uint32_t s : isize) h = (h ^ s) * 1099511628211ull;
      printf("%zu %llx %d\n", isize.size(), h, in_parts ? 1 : 0);
      return 0;
    }
    if (argc > 3 && !strcmp(argv[3], "cut")) {  // fastq_last_record_end of the raw bytes of a file
      FILE* f = fopen(argv[1], "rb");
      if (!f) return 2;
      std::vector<char> data;
      char buf[1 << 16];
      for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + got);
      fclose(f);
      const size_t cut = sgh::fastq_last_record_end(data.data(), data.size());
      if (cut == SIZE_MAX)
        printf("none\n");
      else
        printf("%zu\n", cut);
      return 0;
    }
    if (argc > 4 && !strcmp(argv[3], "spans")) {  // span framing: fastx_dump <path> <threads> spans read_len,start,len,stride
      sgh::SpanSpec spec;
      if (sscanf(argv[4], "%u,%u,%u,%u", &spec.read_len, &spec.start, &spec.len, &spec.stride) != 4) return 2;
      sgh::SeqBlockReader br(argv[1], (unsigned)atoi(argv[2]), &spec);
      sgh::SeqBlock b;
      // hash of what a counter would look at: the span of every read of read_len bytes (whether it
      // came as a span record or inside a whole line), the whole sequence of any other read
      unsigned long long n = 0, n_span = 0, h = 1469598103934665603ull;
      auto mix = [&](const char* p, size_t l, unsigned mark) {
        for (size_t j = 0; j < l; ++j) h = (h ^ (unsigned char)p[j]) * 1099511628211ull;
        h = (h ^ mark) * 1099511628211ull;
      };
      while (br.next(b)) {
        if (b.spans) {
          if (b.first_len != spec.len || b.stride != spec.stride || b.lines.size() != b.n * spec.stride || !b.uniform) return 3;
          for (unsigned long long i = 0; i < b.n; ++i) mix(b.lines.data() + i * spec.stride, spec.len, 0xFEu);
          n_span += b.n;
        } else {
          size_t at = 0;
          for (unsigned long long i = 0; i < b.n; ++i) {
            const size_t l = b.len[i];
            if (l == spec.read_len)
              mix(b.lines.data() + at + spec.start, spec.len, 0xFEu);
            else
              mix(b.lines.data() + at, l, 0xFFu);
            at += l + 1;
          }
        }
        n += b.n;
      }
      printf("%llu %llx %llu\n", n, h, n_span);
      fprintf(stderr, "adopted %llu reframed %llu\n", (unsigned long long)br.members_adopted(), (unsigned long long)br.members_reframed());
      return 0;
    }
    if (argc > 3 && !strcmp(argv[3], "blocks")) {  // the packed-sequence-line path of count_sample
      sgh::SeqBlockReader br(argv[1], (unsigned)atoi(argv[2]));
      sgh::SeqBlock b;
      unsigned long long n = 0, h = 1469598103934665603ull;
      while (br.next(b)) {
        size_t at = 0;
        bool uniform = true;
        for (unsigned long long i = 0; i < b.n; ++i) {
          const size_t l = b.len[i];
          uniform &= l == b.first_len;
          for (size_t j = 0; j < l; ++j) h = (h ^ (unsigned char)b.lines[at + j]) * 1099511628211ull;
          if (b.lines[at + l] != '\n') return 3;
          h = (h ^ 0xFFu) * 1099511628211ull;
          at += l + 1;
        }
        if (at != b.lines.size() || b.len.size() != b.n || uniform != b.uniform) return 3;
        n += b.n;
      }
      printf("%llu %llx\n", n, h);
      fprintf(stderr, "adopted %llu reframed %llu\n", (unsigned long long)br.members_adopted(), (unsigned long long)br.members_reframed());
      return 0;
    }
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
