// oracle.cpp — CPU parity oracle for the sgcount read->guide matching/counting path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT (see oracle.h).  Every function cites the lines of
// /root/reference/src it restates.  The data-structure shape follows the reference:
// byte-string keyed hash maps for library / permuter / per-sample results, one heap
// token per probe, one thread per sample.  Nothing here is shared with the CUDA path.
//
// Parity status: pinned against the reference's unit-test vectors and the
// header labels of the example fixtures; third-party crate behaviour (fxread, ndarray-stats)
// is restated from their published behaviour and marked [3P] where it matters.

#include "oracle.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// ---------------------------------------------------------------------------------
// ByteMap: stand-in for hashbrown::HashMap<Vec<u8>, V>.  Open addressing, one heap
// string per key (the reference stores Vec<u8> keys), cached 64-bit hash, backward-shift
// deletion (needed by Permuter's table.remove, permutes.rs:150).
// ---------------------------------------------------------------------------------
inline uint64_t hash_bytes(const uint8_t* p, size_t n) {
  // 64-bit multiply-fold hash (same family as the foldhash hashbrown 0.15 defaults to)
  const uint64_t k0 = 0x9E3779B97F4A7C15ull, k1 = 0xD6E8FEB86659FD93ull;
  uint64_t h = k0 ^ (uint64_t)n;
  auto fold = [](uint64_t a, uint64_t b) {
    __uint128_t m = (__uint128_t)a * b;
    return (uint64_t)m ^ (uint64_t)(m >> 64);
  };
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    h = fold(h ^ w, k1);
    p += 8;
    n -= 8;
  }
  if (n) {
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = fold(h ^ w, k1);
  }
  return fold(h, k0);
}

template <typename V>
class ByteMap {
  struct Slot {
    uint64_t hash = 0;
    std::unique_ptr<std::string> key;  // heap key, like Vec<u8>
    V value{};
  };
  std::vector<Slot> slots_;
  size_t len_ = 0, mask_ = 0;

  void grow() {
    size_t cap = slots_.empty() ? 16 : slots_.size() * 2;
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.resize(cap);
    mask_ = cap - 1;
    for (auto& s : old)
      if (s.key) {
        size_t i = s.hash & mask_;
        while (slots_[i].key) i = (i + 1) & mask_;
        slots_[i] = std::move(s);
      }
  }
  // index of the slot holding key, or of the empty slot where it would go
  size_t probe(const uint8_t* p, size_t n, uint64_t h) const {
    size_t i = h & mask_;
    while (slots_[i].key) {
      const std::string& k = *slots_[i].key;
      if (slots_[i].hash == h && k.size() == n && memcmp(k.data(), p, n) == 0) return i;
      i = (i + 1) & mask_;
    }
    return i;
  }

 public:
  size_t len() const { return len_; }
  const V* get(const uint8_t* p, size_t n) const {
    if (slots_.empty()) return nullptr;
    size_t i = probe(p, n, hash_bytes(p, n));
    return slots_[i].key ? &slots_[i].value : nullptr;
  }
  V* get_mut(const uint8_t* p, size_t n) { return const_cast<V*>(get(p, n)); }
  bool contains_key(const uint8_t* p, size_t n) const { return get(p, n) != nullptr; }
  // returns true if the key was new (HashMap::insert(..) == None)
  bool insert(const uint8_t* p, size_t n, V v) {
    if ((len_ + 1) * 8 > slots_.size() * 7) grow();
    uint64_t h = hash_bytes(p, n);
    size_t i = probe(p, n, h);
    if (slots_[i].key) {
      slots_[i].value = std::move(v);
      return false;
    }
    slots_[i].hash = h;
    slots_[i].key.reset(new std::string((const char*)p, n));
    slots_[i].value = std::move(v);
    ++len_;
    return true;
  }
  bool remove(const uint8_t* p, size_t n) {
    if (slots_.empty()) return false;
    size_t i = probe(p, n, hash_bytes(p, n));
    if (!slots_[i].key) return false;
    slots_[i] = Slot{};
    --len_;
    size_t j = i;
    for (;;) {  // backward-shift the rest of the cluster
      j = (j + 1) & mask_;
      if (!slots_[j].key) break;
      size_t home = slots_[j].hash & mask_;
      bool between = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
      if (!between) {
        slots_[i] = std::move(slots_[j]);
        slots_[j] = Slot{};
        i = j;
      }
    }
    return true;
  }
  template <typename F>
  void for_each(F f) const {
    for (auto& s : slots_)
      if (s.key) f(*s.key, s.value);
  }
};

inline const uint8_t* u8(const std::string& s) { return (const uint8_t*)s.data(); }

}  // namespace

// ---------------------------------------------------------------------------------
// records — restates what the reference consumes from fxread ^0.2.5 [3P]:
// initialize_reader (count.rs:24,64,87; offsetter.rs:172-173,190,195): gzip iff path ends
// ".gz" (multi-member), format from the first byte ('>' FASTA 2 lines/record, '@' FASTQ
// 4 lines/record); Record::id() = header line without marker/newline, Record::seq() = raw
// bytes, case preserved (pinned by library.rs:126-130 and counter.rs:283-288).
// ---------------------------------------------------------------------------------
struct orc_records {
  std::vector<std::string> ids, seqs;
};

namespace {

int inflate_all(const uint8_t* buf, size_t len, std::string& out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 15 + 32) != Z_OK) return fail(ORC_ERR_IO, "inflateInit2 failed");
  zs.next_in = const_cast<Bytef*>(buf);
  zs.avail_in = (uInt)len;  // fixtures and synthetic members are < 4 GiB
  std::vector<uint8_t> chunk(1 << 20);
  for (;;) {
    zs.next_out = chunk.data();
    zs.avail_out = (uInt)chunk.size();
    int rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) {
      inflateEnd(&zs);
      return fail(ORC_ERR_IO, "inflate failed");
    }
    out.append((const char*)chunk.data(), chunk.size() - zs.avail_out);
    if (rc == Z_STREAM_END) {
      if (zs.avail_in == 0) break;
      if (inflateReset(&zs) != Z_OK) {  // next gzip member (MultiGzDecoder)
        inflateEnd(&zs);
        return fail(ORC_ERR_IO, "inflateReset failed");
      }
    }
  }
  inflateEnd(&zs);
  return ORC_OK;
}

// one line without its '\n'; returns false at end of input
bool next_line(const char*& p, const char* end, const char*& line, size_t& n) {
  if (p >= end) return false;
  const char* nl = (const char*)memchr(p, '\n', end - p);
  line = p;
  if (nl) {
    n = nl - p;
    p = nl + 1;
  } else {
    n = end - p;
    p = end;
  }
  return true;
}

int parse_fastx(const std::string& text, orc_records* r) {
  const char* p = text.data();
  const char* end = p + text.size();
  if (p == end) return ORC_OK;
  const char marker = *p;
  if (marker != '>' && marker != '@') return fail(ORC_ERR_IO, "unrecognised fastx format");
  const int extra = marker == '@' ? 2 : 0;  // '+' and quality lines
  const char* line;
  size_t n;
  while (next_line(p, end, line, n)) {
    if (n == 0 && p >= end) break;  // trailing blank line
    if (n == 0 || line[0] != marker) return fail(ORC_PANIC_MALFORMED, "malformed record header");
    r->ids.emplace_back(line + 1, n - 1);
    if (!next_line(p, end, line, n)) return fail(ORC_PANIC_MALFORMED, "truncated record");
    r->seqs.emplace_back(line, n);
    for (int i = 0; i < extra; ++i)
      if (!next_line(p, end, line, n)) return fail(ORC_PANIC_MALFORMED, "truncated fastq record");
  }
  return ORC_OK;
}

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_records_from_memory(const uint8_t* buf, size_t len, int gz, orc_records** out) {
  std::string text;
  if (gz) {
    int rc = inflate_all(buf, len, text);
    if (rc) return rc;
  } else {
    text.assign((const char*)buf, len);
  }
  auto r = std::make_unique<orc_records>();
  int rc = parse_fastx(text, r.get());
  if (rc) return rc;
  *out = r.release();
  return ORC_OK;
}

int orc_records_from_path(const char* path, orc_records** out) {
  FILE* f = fopen(path, "rb");
  if (!f) return fail(ORC_ERR_IO, std::string("cannot open ") + path);
  std::string raw;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, n);
  fclose(f);
  size_t pl = strlen(path);
  int gz = pl >= 3 && strcmp(path + pl - 3, ".gz") == 0;
  return orc_records_from_memory((const uint8_t*)raw.data(), raw.size(), gz, out);
}

int orc_records_from_seqs(const uint8_t* seqs, const uint64_t* off, uint64_t n, orc_records** out) {
  auto r = std::make_unique<orc_records>();
  r->ids.reserve(n);
  r->seqs.reserve(n);
  for (uint64_t i = 0; i < n; ++i) {
    r->ids.push_back("r" + std::to_string(i));
    r->seqs.emplace_back((const char*)seqs + off[i], off[i + 1] - off[i]);
  }
  *out = r.release();
  return ORC_OK;
}

uint64_t orc_records_len(const orc_records* r) { return r->seqs.size(); }
const uint8_t* orc_records_seq(const orc_records* r, uint64_t i, uint64_t* len) {
  *len = r->seqs[i].size();
  return u8(r->seqs[i]);
}
const uint8_t* orc_records_id(const orc_records* r, uint64_t i, uint64_t* len) {
  *len = r->ids[i].size();
  return u8(r->ids[i]);
}
uint64_t orc_records_seq_bytes(const orc_records* r) {
  uint64_t t = 0;
  for (auto& s : r->seqs) t += s.size();
  return t;
}
void orc_records_export_lines(const orc_records* r, uint8_t* lines, uint64_t* off) {
  uint64_t pos = 0;
  for (size_t i = 0; i < r->seqs.size(); ++i) {
    off[i] = pos;
    memcpy(lines + pos, r->seqs[i].data(), r->seqs[i].size());
    pos += r->seqs[i].size();
    lines[pos++] = '\n';
  }
  off[r->seqs.size()] = pos;
}
void orc_records_free(orc_records* r) { delete r; }

// fxread::Record::seq_rev_comp, used at counter.rs:203 [3P, appendix D.1]
void orc_seq_rev_comp(const uint8_t* seq, size_t len, int rc_mode, uint8_t* out) {
  for (size_t i = 0; i < len; ++i) {
    uint8_t c = seq[len - 1 - i];
    if (rc_mode == ORC_RC_BITTRICK) {
      out[i] = (c & 2) ? (uint8_t)(c ^ 4) : (uint8_t)(c ^ 21);
    } else {
      switch (c) {
        case 'A': out[i] = 'T'; break;
        case 'C': out[i] = 'G'; break;
        case 'G': out[i] = 'C'; break;
        case 'T': out[i] = 'A'; break;
        default: out[i] = c;
      }
    }
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// Library — library.rs:9-99
// ---------------------------------------------------------------------------------
struct orc_library {
  ByteMap<std::string> table;      // seq -> alias            (library.rs:10)
  ByteMap<uint64_t> index;         // seq -> insertion index  (oracle bookkeeping only)
  std::vector<std::string> seqs;   // insertion order
  std::vector<std::string> aliases;
  uint64_t size = 0;

  // library.rs:34-40: contains_key, then a second lookup through alias()
  const std::string* contains(const uint8_t* t, size_t n) const {
    if (table.contains_key(t, n)) return alias(t, n);
    return nullptr;
  }
  // library.rs:44-46
  const std::string* alias(const uint8_t* t, size_t n) const { return table.get(t, n); }
};

extern "C" {

int orc_library_from_records(const orc_records* r, orc_library** out) {
  auto lib = std::make_unique<orc_library>();
  // table_from_reader, library.rs:89-99: insert seq -> id, panic on a duplicate sequence
  for (size_t i = 0; i < r->seqs.size(); ++i) {
    const std::string& s = r->seqs[i];
    if (!lib->table.insert(u8(s), s.size(), r->ids[i]))
      return fail(ORC_PANIC_DUPLICATE_SEQ, "Unexpected duplicate sequence in library found: " + s);
    lib->index.insert(u8(s), s.size(), (uint64_t)i);
    lib->seqs.push_back(s);
    lib->aliases.push_back(r->ids[i]);
  }
  // calculate_base_size, library.rs:65-85 (get_key_size unwraps on an empty table)
  if (lib->seqs.empty()) return fail(ORC_PANIC_EMPTY_READER, "empty library");
  for (size_t i = 1; i < lib->seqs.size(); ++i)
    if (lib->seqs[i].size() != lib->seqs[i - 1].size())
      return fail(ORC_ERR_INCONSISTENT_SIZE, "Library sequence sizes are inconsistent");
  lib->size = lib->seqs[0].size();
  *out = lib.release();
  return ORC_OK;
}

uint64_t orc_library_len(const orc_library* l) { return l->seqs.size(); }
uint64_t orc_library_size(const orc_library* l) { return l->size; }
int64_t orc_library_contains(const orc_library* l, const uint8_t* t, size_t n) {
  if (!l->contains(t, n)) return -1;
  return (int64_t)*l->index.get(t, n);
}
const uint8_t* orc_library_seq(const orc_library* l, uint64_t i, uint64_t* len) {
  *len = l->seqs[i].size();
  return u8(l->seqs[i]);
}
const uint8_t* orc_library_alias(const orc_library* l, uint64_t i, uint64_t* len) {
  *len = l->aliases[i].size();
  return u8(l->aliases[i]);
}
void orc_library_free(orc_library* l) { delete l; }

}  // extern "C"

// ---------------------------------------------------------------------------------
// Permuter — permutes.rs:3-158, the literal stateful insert algorithm
// ---------------------------------------------------------------------------------
struct orc_permuter {
  ByteMap<std::string> map;  // variant -> parent sequence   (permutes.rs:35)
  ByteMap<char> null;        // parents + ambiguous variants (permutes.rs:36)
  const orc_library* lib = nullptr;
};

namespace {

const uint8_t LEXICON[5] = {'A', 'C', 'G', 'T', 'N'};  // permutes.rs:3

// permutes.rs:127-144
void insert_sequence(const std::string& sequence, const std::string& permutation, orc_permuter* p) {
  if (!p->null.contains_key(u8(sequence), sequence.size()))
    p->null.insert(u8(sequence), sequence.size(), 1);
  if (!p->null.contains_key(u8(permutation), permutation.size())) {
    if (p->map.contains_key(u8(permutation), permutation.size())) {
      // insert_to_null, permutes.rs:149-152
      p->map.remove(u8(permutation), permutation.size());
      p->null.insert(u8(permutation), permutation.size(), 1);
    } else {
      // insert_to_table, permutes.rs:156-158
      p->map.insert(u8(permutation), permutation.size(), sequence);
    }
  }
}

// permute_sequence / build_permutations / build_permutation, permutes.rs:78-117
std::vector<std::string> permute_sequence(const std::string& sequence) {
  std::vector<std::string> all;
  for (size_t idx = 0; idx < sequence.size(); ++idx)
    for (uint8_t y : LEXICON) {
      if (y == (uint8_t)sequence[idx]) continue;
      std::string v;
      v.append(sequence, 0, idx);
      v.push_back((char)y);
      v.append(sequence, idx + 1, std::string::npos);
      all.push_back(std::move(v));
    }
  return all;
}

}  // namespace

extern "C" {

int orc_permuter_new(const orc_library* lib, const uint64_t* order, orc_permuter** out) {
  auto p = std::make_unique<orc_permuter>();
  p->lib = lib;
  // build, permutes.rs:63-75: fold over the key iterator
  for (size_t i = 0; i < lib->seqs.size(); ++i) {
    const std::string& seq = lib->seqs[order ? order[i] : i];
    for (const std::string& v : permute_sequence(seq)) insert_sequence(seq, v, p.get());
  }
  *out = p.release();
  return ORC_OK;
}

int64_t orc_permuter_contains(const orc_permuter* p, const uint8_t* t, size_t n) {
  const std::string* parent = p->map.get(t, n);  // permutes.rs:55-57
  if (!parent) return -1;
  return (int64_t)*p->lib->index.get(u8(*parent), parent->size());
}
uint64_t orc_permuter_map_len(const orc_permuter* p) { return p->map.len(); }
uint64_t orc_permuter_null_len(const orc_permuter* p) { return p->null.len(); }
int orc_permuter_null_contains(const orc_permuter* p, const uint8_t* t, size_t n) {
  return p->null.contains_key(t, n) ? 1 : 0;
}
void orc_permuter_free(orc_permuter* p) { delete p; }

}  // extern "C"

// ---------------------------------------------------------------------------------
// Counter — counter.rs:7-252
// ---------------------------------------------------------------------------------
namespace {

enum Position { PLUS = 0, MINUS = 1, CENTERED = 2, PNULL = 3 };  // counter.rs:7-12

// counter.rs:158-180
bool bounds(size_t seq_len, size_t offset, size_t size, Position pos, size_t& mn, size_t& mx) {
  switch (pos) {
    case PLUS:
      mn = offset + 1;
      mx = offset + 1 + size;
      break;
    case MINUS:
      if (offset == 0) return false;  // checked_sub(1)
      mn = offset - 1;
      mx = mn + size;
      break;
    default:
      mn = offset;
      mx = offset + size;
  }
  return mx <= seq_len;
}

// apply_trim / trim_forward_sequence / trim_reverse_sequence, counter.rs:144-204.
// One fresh token per probe; Reverse reverse-complements the WHOLE read first (counter.rs:203).
bool apply_trim(const std::string& seq, bool is_reverse, size_t offset, size_t size, Position pos,
                int rc_mode, std::string& token) {
  size_t mn, mx;
  if (!bounds(seq.size(), offset, size, pos, mn, mx)) return false;
  if (!is_reverse) {
    token.assign(seq, mn, mx - mn);
  } else {
    std::string rc(seq.size(), '\0');
    orc_seq_rev_comp(u8(seq), seq.size(), rc_mode, (uint8_t*)&rc[0]);
    token.assign(rc, mn, mx - mn);
  }
  return true;
}

// counter.rs:96-140
const std::string* assign(const std::string& seq, const orc_library* lib, const orc_permuter* perm,
                          bool is_reverse, size_t offset, size_t size, Position pos, int rc_mode) {
  std::string token;
  if (!apply_trim(seq, is_reverse, offset, size, pos, rc_mode, token)) return nullptr;  // :105-108
  const std::string* alias = lib->contains(u8(token), token.size());                    // :111
  if (!alias && perm) {                                                                  // :113-116
    const std::string* parent = perm->map.get(u8(token), token.size());
    if (parent) alias = lib->alias(u8(*parent), parent->size());
  }
  if (!alias) {  // :120-135
    if (pos == CENTERED) return assign(seq, lib, perm, is_reverse, offset, size, PLUS, rc_mode);
    if (pos == PLUS) return assign(seq, lib, perm, is_reverse, offset, size, MINUS, rc_mode);
    return nullptr;
  }
  return alias;
}

// oracle-only helper: library index of the sequence that produced `alias` for this read.
// Re-derives the hit the same way assign() found it (indices are not part of the reference).
int64_t assign_index(const std::string& seq, const orc_library* lib, const orc_permuter* perm,
                     bool is_reverse, size_t offset, size_t size, bool recursion, int rc_mode) {
  const Position order_rec[3] = {CENTERED, PLUS, MINUS};
  const Position order_null[1] = {PNULL};
  const Position* order = recursion ? order_rec : order_null;
  int n = recursion ? 3 : 1;
  std::string token;
  for (int i = 0; i < n; ++i) {
    if (!apply_trim(seq, is_reverse, offset, size, order[i], rc_mode, token)) return -1;
    if (lib->contains(u8(token), token.size())) return (int64_t)*lib->index.get(u8(token), token.size());
    if (perm) {
      const std::string* parent = perm->map.get(u8(token), token.size());
      if (parent && lib->alias(u8(*parent), parent->size()))
        return (int64_t)*lib->index.get(u8(*parent), parent->size());
    }
  }
  return -1;
}

}  // namespace

struct orc_counter {
  ByteMap<uint64_t> results;  // alias -> count (counter.rs:18)
  uint64_t total_reads = 0, matched_reads = 0;
};

namespace {

// Counter::count, counter.rs:211-236, over records [lo, hi)
void count_range(const orc_records* recs, size_t lo, size_t hi, const orc_library* lib,
                 const orc_permuter* perm, bool is_reverse, size_t offset, bool recursion,
                 int rc_mode, int32_t* assign_out, orc_counter* c) {
  const Position start = recursion ? CENTERED : PNULL;  // counter.rs:44-48
  const size_t size = lib->size;                        // count.rs:31
  for (size_t i = lo; i < hi; ++i) {
    const std::string& seq = recs->seqs[i];
    c->total_reads += 1;  // :223-226
    const std::string* alias = assign(seq, lib, perm, is_reverse, offset, size, start, rc_mode);
    if (assign_out)
      assign_out[i] = (int32_t)assign_index(seq, lib, perm, is_reverse, offset, size, recursion, rc_mode);
    if (!alias) continue;
    c->matched_reads += 1;  // :228-231
    std::string key = *alias;  // x.clone(), :233
    uint64_t* v = c->results.get_mut(u8(key), key.size());
    if (v)
      *v += 1;
    else
      c->results.insert(u8(key), key.size(), 1);
  }
}

}  // namespace

extern "C" {

int orc_bounds(uint64_t seq_len, uint64_t offset, uint64_t size, int position, uint64_t* mn, uint64_t* mx) {
  size_t a = 0, b = 0;
  bool ok = bounds(seq_len, offset, size, (Position)position, a, b);
  if (ok) {
    *mn = a;
    *mx = b;
  }
  return ok ? 1 : 0;
}

int64_t orc_assign(const orc_library* lib, const orc_permuter* perm, const uint8_t* read, size_t len,
                   int is_reverse, uint64_t offset, int position_recursion, int rc_mode) {
  std::string seq((const char*)read, len);
  const std::string* alias = assign(seq, lib, perm, is_reverse != 0, offset, lib->size,
                                    position_recursion ? CENTERED : PNULL, rc_mode);
  int64_t idx = assign_index(seq, lib, perm, is_reverse != 0, offset, lib->size, position_recursion != 0, rc_mode);
  if ((alias == nullptr) != (idx < 0)) abort();  // the two walks must agree
  return idx;
}

int orc_counter_new(const orc_records* recs, const orc_library* lib, const orc_permuter* perm,
                    int is_reverse, uint64_t offset, int position_recursion, int rc_mode, int n_threads,
                    int32_t* assign_out, orc_counter** out) {
  auto c = std::make_unique<orc_counter>();
  const size_t n = recs->seqs.size();
  if (n_threads <= 1) {
    count_range(recs, 0, n, lib, perm, is_reverse != 0, offset, position_recursion != 0, rc_mode,
                assign_out, c.get());
  } else {
    std::vector<orc_counter> parts(n_threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
      size_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
      pool.emplace_back(count_range, recs, lo, hi, lib, perm, is_reverse != 0, (size_t)offset,
                        position_recursion != 0, rc_mode, assign_out, &parts[t]);
    }
    for (auto& th : pool) th.join();
    for (auto& p : parts) {
      c->total_reads += p.total_reads;
      c->matched_reads += p.matched_reads;
      p.results.for_each([&](const std::string& k, const uint64_t& v) {
        uint64_t* cur = c->results.get_mut(u8(k), k.size());
        if (cur)
          *cur += v;
        else
          c->results.insert(u8(k), k.size(), v);
      });
    }
  }
  *out = c.release();
  return ORC_OK;
}

// counter.rs:71-76
uint64_t orc_counter_get_value(const orc_counter* c, const uint8_t* alias, size_t len) {
  const uint64_t* v = c->results.get(alias, len);
  return v ? *v : 0;
}
uint64_t orc_counter_total_reads(const orc_counter* c) { return c->total_reads; }
uint64_t orc_counter_matched_reads(const orc_counter* c) { return c->matched_reads; }
void orc_counter_counts_by_index(const orc_counter* c, const orc_library* lib, uint64_t* out) {
  for (size_t i = 0; i < lib->aliases.size(); ++i)
    out[i] = orc_counter_get_value(c, u8(lib->aliases[i]), lib->aliases[i].size());
}
void orc_counter_free(orc_counter* c) { delete c; }

}  // extern "C"

// ---------------------------------------------------------------------------------
// Offsetter — offsetter.rs:37-210
// ---------------------------------------------------------------------------------
namespace {

// offsetter.rs:42-50
int base_map(uint8_t c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

// offsetter.rs:37-39 + 55-79 over the first `take` records (Iterator::take, :173,:197)
int position_counts(const orc_records* r, uint64_t take, std::vector<double>& m, uint64_t& size) {
  uint64_t n = std::min<uint64_t>(take, r->seqs.size());
  if (n == 0) return fail(ORC_PANIC_EMPTY_READER, "empty reader");
  size = r->seqs[0].size();  // first record consumed, not counted
  m.assign(size * 4, 0.0);
  for (uint64_t i = 1; i < n; ++i) {
    const std::string& s = r->seqs[i];
    uint64_t lim = std::min<uint64_t>(size, s.size());  // .take(size)
    for (uint64_t idx = 0; idx < lim; ++idx) {
      int j = base_map((uint8_t)s[idx]);
      if (j >= 0) {
        m[idx * 4 + j] += 1.0;
      } else {
        for (int q = 0; q < 4; ++q) m[idx * 4 + q] += 1.0;  // :70-74
      }
    }
  }
  return ORC_OK;
}

// normalize_counts (:82-87) then ndarray-stats entropy per row (:92-94) [3P]:
// -(sum over the four columns, left to right, of x == 0 ? 0 : x * ln x)
void entropy_rows(const std::vector<double>& m, uint64_t size, double* out) {
  for (uint64_t i = 0; i < size; ++i) {
    double sum = ((m[i * 4] + m[i * 4 + 1]) + m[i * 4 + 2]) + m[i * 4 + 3];
    double acc = 0.0;
    for (int j = 0; j < 4; ++j) {
      double p = m[i * 4 + j] / sum;
      acc += (p == 0.0) ? 0.0 : p * std::log(p);
    }
    out[i] = -acc;
  }
}

// windowed_mse, offsetter.rs:109-120; mean_sq_err = (sum_i (a_i - b_i)^2, sequential) / n [3P]
std::vector<double> windowed_mse(const double* a1, uint64_t n1, const double* a2, uint64_t n2) {
  uint64_t size = n2 - n1 + 1;
  std::vector<double> out(size, 0.0);
  for (uint64_t x = 0; x < size; ++x) {
    double acc = 0.0;
    for (uint64_t i = 0; i < n1; ++i) {
      double d = a1[i] - a2[x + i];
      acc += d * d;
    }
    out[x] += acc / (double)n1;
  }
  return out;
}

// QuantileExt::argmin / min [3P]: first minimum; any undefined comparison (NaN) is an error
bool argmin_first(const std::vector<double>& v, uint64_t& arg, double& mn) {
  arg = 0;
  mn = v[0];
  if (std::isnan(mn)) return false;
  for (uint64_t i = 1; i < v.size(); ++i) {
    if (std::isnan(v[i])) return false;
    if (v[i] < mn) {
      mn = v[i];
      arg = i;
    }
  }
  return true;
}

}  // namespace

extern "C" {

int orc_position_counts(const orc_records* r, uint64_t take, double* out, uint64_t* size) {
  std::vector<double> m;
  int rc = position_counts(r, take, m, *size);
  if (rc) return rc;
  if (out) memcpy(out, m.data(), m.size() * sizeof(double));
  return ORC_OK;
}

int orc_entropy_from_counts(const double* counts, uint64_t size, double* out) {
  std::vector<double> m(counts, counts + size * 4);
  entropy_rows(m, size, out);
  return ORC_OK;
}

int orc_positional_entropy(const orc_records* r, uint64_t take, double* out, uint64_t* size) {
  std::vector<double> m;
  int rc = position_counts(r, take, m, *size);
  if (rc) return rc;
  if (out) entropy_rows(m, *size, out);
  return ORC_OK;
}

// minimize_mse + assign_offset, offsetter.rs:122-163
int orc_minimize_mse(const double* reference, uint64_t n_ref, const double* comparison, uint64_t n_cmp,
                     int* is_reverse, uint64_t* index) {
  if (n_cmp < n_ref)
    return fail(ORC_ERR_READ_TOO_SHORT,
                "Sequences in reference library are larger than the sequences in input.");
  std::vector<double> rev(comparison, comparison + n_cmp);
  std::reverse(rev.begin(), rev.end());
  std::vector<double> f = windowed_mse(reference, n_ref, comparison, n_cmp);
  std::vector<double> r = windowed_mse(reference, n_ref, rev.data(), n_cmp);
  uint64_t af, ar;
  double mf, mr;
  if (!argmin_first(f, af, mf) || !argmin_first(r, ar, mr))
    return fail(ORC_PANIC_NAN, "Unexpected minmax error in entropy");
  if (mf < mr) {  // strict: ties go to Reverse (:143-149)
    *is_reverse = 0;
    *index = af;
  } else {
    *is_reverse = 1;
    *index = ar;
  }
  return ORC_OK;
}

// entropy_offset / entropy_offset_group for one sample, offsetter.rs:167-210.
// The library reader is NOT subsampled (:190-191); the sample is .take(subsample) (:197).
int orc_entropy_offset(const orc_records* library, const orc_records* sample, uint64_t subsample,
                       int* is_reverse, uint64_t* index) {
  uint64_t ls = 0, ss = 0;
  std::vector<double> lm, sm;
  int rc = position_counts(library, UINT64_MAX, lm, ls);
  if (rc) return rc;
  rc = position_counts(sample, subsample, sm, ss);
  if (rc) return rc;
  std::vector<double> lh(ls), sh(ss);
  entropy_rows(lm, ls, lh.data());
  entropy_rows(sm, ss, sh.data());
  return orc_minimize_mse(lh.data(), ls, sh.data(), ss, is_reverse, index);
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// GeneMap + results + sample names — genemap.rs:53-86, results.rs:32-99, utils.rs:18-49
// ---------------------------------------------------------------------------------
namespace {

// genemap.rs:53-68 (bstr for_byte_line strips "\n" / "\r\n")
int build_genemap(const uint8_t* buf, size_t len, ByteMap<std::string>& map) {
  const char* p = (const char*)buf;
  const char* end = p + len;
  const char* line;
  size_t n;
  while (next_line(p, end, line, n)) {
    if (n && line[n - 1] == '\r') --n;
    const char* tab = (const char*)memchr(line, '\t', n);
    if (!tab) return fail(ORC_PANIC_GENEMAP, "Missing '\\t' in gene map");
    std::string gene(line, tab - line), sgrna(tab + 1, line + n - (tab + 1));
    if (!map.insert(u8(sgrna), sgrna.size(), gene))
      return fail(ORC_PANIC_GENEMAP, "Duplicate sgRNA key found in gene map: " + sgrna);
  }
  return ORC_OK;
}

std::string trim_end_matches(std::string s, const std::string& suf) {
  while (s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0)
    s.erase(s.size() - suf.size());
  return s;
}

char* dup_cstr(const std::string& s) {
  char* out = (char*)malloc(s.size() + 1);
  memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

}  // namespace

extern "C" {

int orc_render_results(const orc_counter* const* counters, const char* const* names, uint64_t n_samples,
                       const orc_library* lib, const uint8_t* genemap_buf, size_t genemap_len,
                       int include_zero, char** out) {
  ByteMap<std::string> gm;
  const bool has_gm = genemap_buf != nullptr;
  if (has_gm) {
    int rc = build_genemap(genemap_buf, genemap_len, gm);
    if (rc) return rc;
    // missing_aliases, genemap.rs:81-86 -> bail at count.rs:90-95
    for (auto& a : lib->aliases)
      if (!gm.get(u8(a), a.size())) return fail(ORC_ERR_IO, "Missing sgRNA aliases in gene map: \"" + a + "\"");
  }
  // generate_columns, results.rs:32-43
  std::string text = "Guide";
  for (uint64_t i = 0; i < n_samples; ++i) {
    if (i == 0 && has_gm) text += "\tGene";
    text += "\t";
    text += names[i];
  }
  text += "\n";
  // write_results, results.rs:79-95 (library.values() order is hash order; we use insertion order)
  for (size_t g = 0; g < lib->aliases.size(); ++g) {
    const std::string& alias = lib->aliases[g];
    uint64_t total = 0;
    std::string row = alias;
    for (uint64_t i = 0; i < n_samples; ++i) {
      if (i == 0 && has_gm) row += "\t" + *gm.get(u8(alias), alias.size());  // append_gene :46-62
      uint64_t v = orc_counter_get_value(counters[i], u8(alias), alias.size());
      row += "\t" + std::to_string(v);  // append_count :65-67
      total += v;
    }
    if (include_zero || total > 0) text += row + "\n";  // :90-94
  }
  *out = dup_cstr(text);
  return ORC_OK;
}

// utils.rs:18-49
int orc_generate_sample_names(const char* const* paths, uint64_t n, char** out) {
  std::vector<std::string> base;
  for (uint64_t i = 0; i < n; ++i) {
    std::string p = paths[i];
    size_t slash = p.rfind('/');
    std::string b = slash == std::string::npos ? p : p.substr(slash + 1);
    for (const char* suf : {".gz", ".fasta", ".fastq", ".fa", ".fq"}) b = trim_end_matches(b, suf);
    base.push_back(b);
  }
  std::vector<std::string> uniq = base;
  std::sort(uniq.begin(), uniq.end());
  bool dup = std::adjacent_find(uniq.begin(), uniq.end()) != uniq.end();
  std::string text;
  for (uint64_t i = 0; i < n; ++i) {
    if (i) text += "\n";
    text += dup ? "Sample." + std::to_string(i) : base[i];
  }
  *out = dup_cstr(text);
  return ORC_OK;
}

void orc_free(void* p) { free(p); }

}  // extern "C"
