// inflate.cpp — whole-member gzip decoder (see inflate.h).  DEFLATE per RFC 1951, gzip framing
// per RFC 1952; written for this ingest path, not taken from zlib or any other decoder.
//
// Shape of the decoder:
//   * the whole member is decoded in one call into one contiguous buffer: a match is a copy from
//     `out - distance`, there is no sliding window and no suspend/resume state;
//   * a 64-bit bit buffer refilled with one unaligned 8-byte load (byte-wise only for the last
//     bytes of the file), so that up to three literals or one whole length/distance pair are
//     decoded per refill;
//   * table-driven Huffman decoding: an 11-bit root table for literals/lengths and an 8-bit one
//     for distances, longer codes through second-level tables; every entry already holds what
//     the symbol means (literal byte, or base value + number of extra bits);
//   * matches are copied 8 bytes at a time (a run, distance 1, is a fill).
// The member's CRC-32 and ISIZE are verified here; the caller treats `false` as "let zlib look
// at these bytes", so nothing this decoder gets wrong can reach the counts.
#include "inflate.h"

#include <zlib.h>  // crc32_z only

#include <cstring>

namespace sgh {
namespace {

constexpr int kLitRoot = 11, kDistRoot = 8, kPreRoot = 7, kMaxCodeLen = 15;
constexpr uint32_t kLitCap = (1u << kLitRoot) + 2560, kDistCap = (1u << kDistRoot) + 640;

// Table entry:
//   [7:0]   bits this step consumes (a second-level pointer consumes the root bits)
//   [11:8]  extra bits of a length / distance symbol; index bits of a second-level table
//   [15:12] kind flags
//   [31:16] literal byte | length base | distance base | start of the second-level table
constexpr uint32_t kLiteral = 1u << 15, kSpecial = 1u << 14, kEndOfBlock = 1u << 13, kSecondLevel = 1u << 12;
constexpr uint32_t kInvalid = kSpecial;

inline uint64_t load64(const unsigned char* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;  // little-endian hosts only (x86-64, aarch64)
}
inline void store64(unsigned char* p, uint64_t v) { memcpy(p, &v, 8); }

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// Canonical Huffman decode table (RFC 1951 3.2.2).  entry_of[s] is symbol s's entry without its
// bit count.  Like zlib, a code that over-subscribes the code space is an error and an
// incomplete one is accepted only if it is a single 1-bit code; no code at all gives a table of
// invalid entries.
bool build_table(const uint8_t* lens, int n, const uint32_t* entry_of, int root, uint32_t* table, uint32_t cap) {
  int count[kMaxCodeLen + 1] = {0};
  for (int s = 0; s < n; ++s) ++count[lens[s]];
  count[0] = 0;
  int max_len = 0, left = 1;
  for (int l = 1; l <= kMaxCodeLen; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
    if (count[l]) max_len = l;
  }
  if (left > 0 && max_len > 1) return false;
  for (uint32_t i = 0; i < (1u << root); ++i) table[i] = kInvalid;
  if (max_len == 0) return true;

  uint16_t start[kMaxCodeLen + 2], sorted[320];
  start[1] = 0;
  for (int l = 1; l <= kMaxCodeLen; ++l) start[l + 1] = (uint16_t)(start[l] + count[l]);
  {
    uint16_t at[kMaxCodeLen + 2];
    memcpy(at, start, sizeof at);
    for (int s = 0; s < n; ++s)
      if (lens[s]) sorted[at[lens[s]]++] = (uint16_t)s;
  }
  uint32_t next_free = 1u << root, code = 0;
  uint32_t cur_prefix = ~0u, sub_start = 0;
  int sub_bits = 0;
  for (int len = 1; len <= max_len; ++len) {
    for (int c = 0; c < count[len]; ++c) {
      const uint32_t e = entry_of[sorted[start[len] + c]];
      const uint32_t rev = reverse_bits(code, len);
      if (len <= root) {
        for (uint32_t i = rev; i < (1u << root); i += 1u << len) table[i] = e | (uint32_t)len;
      } else {
        const uint32_t prefix = rev & ((1u << root) - 1);
        if (prefix != cur_prefix) {
          // a new second-level table: wide enough for every code that shares this prefix — the
          // codes not yet placed fill it in order of length
          cur_prefix = prefix;
          sub_bits = len - root;
          int room = 1 << sub_bits, l2 = len, pending = count[len] - c;
          for (;;) {
            room -= pending;
            if (room <= 0 || l2 == max_len) break;
            ++l2;
            ++sub_bits;
            room <<= 1;
            pending = count[l2];
          }
          if (next_free + (1u << sub_bits) > cap) return false;
          sub_start = next_free;
          next_free += 1u << sub_bits;
          for (uint32_t i = 0; i < (1u << sub_bits); ++i) table[sub_start + i] = kInvalid;
          table[prefix] = (sub_start << 16) | kSpecial | kSecondLevel | ((uint32_t)sub_bits << 8) | (uint32_t)root;
        }
        const int sub_len = len - root;
        for (uint32_t i = rev >> root; i < (1u << sub_bits); i += 1u << sub_len)
          table[sub_start + i] = e | (uint32_t)sub_len;
      }
      ++code;
    }
    code <<= 1;
  }
  return true;
}

// what the symbols mean (RFC 1951 3.2.5)
struct SymbolEntries {
  uint32_t lit[288], dist[32], pre[19];
  SymbolEntries() {
    static const uint16_t len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                          31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                           193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (uint32_t s = 0; s < 256; ++s) lit[s] = (s << 16) | kLiteral;
    lit[256] = kSpecial | kEndOfBlock;
    for (int s = 257; s < 286; ++s) lit[s] = ((uint32_t)len_base[s - 257] << 16) | ((uint32_t)len_extra[s - 257] << 8);
    lit[286] = lit[287] = kInvalid;
    for (int s = 0; s < 30; ++s) dist[s] = ((uint32_t)dist_base[s] << 16) | ((uint32_t)dist_extra[s] << 8);
    dist[30] = dist[31] = kInvalid;
    for (uint32_t s = 0; s < 19; ++s) pre[s] = s << 16;
  }
};
const SymbolEntries kSymbols;

struct FixedTables {
  uint32_t lit[kLitCap], dist[kDistCap];
  bool ok;
  FixedTables() {
    uint8_t lens[288];
    for (int s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
    ok = build_table(lens, 288, kSymbols.lit, kLitRoot, lit, kLitCap);
    for (int s = 0; s < 32; ++s) lens[s] = 5;
    ok = ok && build_table(lens, 32, kSymbols.dist, kDistRoot, dist, kDistCap);
  }
};
const FixedTables kFixed;

struct Decoder {
  const unsigned char* ip;
  const unsigned char* const in_end;
  uint64_t bits = 0;
  int n_bits = 0;  // may go negative in the last bytes of the input: the stream is truncated then
  Bytes& out;
  unsigned char *op, *out_begin, *out_end;
  uint32_t lit[kLitCap], dist[kDistCap];

  Decoder(const unsigned char* in, const unsigned char* end, Bytes& o) : ip(in), in_end(end), out(o) {
    out.resize(out.capacity() > (1u << 16) ? out.capacity() : (1u << 16));
    out_begin = op = reinterpret_cast<unsigned char*>(out.data());
    out_end = out_begin + out.size();
  }
  void grow(size_t need) {
    const size_t used = (size_t)(op - out_begin);
    size_t cap = out.size();
    while (cap - used < need) cap += cap / 2;
    out.resize(cap);
    out_begin = reinterpret_cast<unsigned char*>(out.data());
    op = out_begin + used;
    out_end = out_begin + cap;
  }
  // at least 56 valid bits while eight input bytes are left; whatever there is after that
  inline void refill() {
    if (in_end - ip >= 8) {
      bits |= load64(ip) << n_bits;
      ip += (63 - n_bits) >> 3;
      n_bits |= 56;
    } else {
      while (n_bits < 56 && ip < in_end) {
        bits |= (uint64_t)*ip++ << n_bits;
        n_bits += 8;
      }
    }
  }
  inline uint32_t take(int n) {
    const uint32_t v = (uint32_t)(bits & ((1ull << n) - 1));
    bits >>= n;
    n_bits -= n;
    return v;
  }

  bool stored_block() {
    take(n_bits & 7);      // to the byte boundary
    ip -= n_bits >> 3;     // hand the whole bytes back
    bits = 0;
    n_bits = 0;
    if (in_end - ip < 4) return false;
    const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
    if ((len ^ nlen) != 0xFFFFu) return false;
    ip += 4;
    if ((size_t)(in_end - ip) < len) return false;
    if ((size_t)(out_end - op) < len) grow(len);
    memcpy(op, ip, len);
    op += len;
    ip += len;
    return true;
  }

  bool dynamic_header() {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    refill();
    const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
    if (n_bits < 0 || hlit > 286 || hdist > 30) return false;
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < hclen; ++i) {
      if (n_bits < 3) refill();
      pre_lens[order[i]] = (uint8_t)take(3);
    }
    if (n_bits < 0) return false;
    uint32_t pre[1u << kPreRoot];
    if (!build_table(pre_lens, 19, kSymbols.pre, kPreRoot, pre, 1u << kPreRoot)) return false;
    uint8_t lens[286 + 30 + 138];
    int i = 0;
    const int total = hlit + hdist;
    while (i < total) {
      refill();
      const uint32_t e = pre[bits & ((1u << kPreRoot) - 1)];
      if (e & kSpecial) return false;
      take((int)(e & 0xFF));
      const uint32_t sym = e >> 16;
      if (sym < 16) {
        lens[i++] = (uint8_t)sym;
      } else {
        int rep;
        uint8_t v = 0;
        if (sym == 16) {
          if (i == 0) return false;
          v = lens[i - 1];
          rep = 3 + (int)take(2);
        } else if (sym == 17) {
          rep = 3 + (int)take(3);
        } else {
          rep = 11 + (int)take(7);
        }
        if (i + rep > total) return false;
        memset(lens + i, v, (size_t)rep);
        i += rep;
      }
      if (n_bits < 0) return false;
    }
    if (lens[256] == 0) return false;  // no end-of-block code
    return build_table(lens, hlit, kSymbols.lit, kLitRoot, lit, kLitCap) &&
           build_table(lens + hlit, hdist, kSymbols.dist, kDistRoot, dist, kDistCap);
  }

  // The symbols of one block, up to and including its end-of-block code.
  //
  // Fast loop, while at least 16 input bytes and a worst-case symbol of output room are left:
  // the state lives in locals (byte stores through `op` may alias the decoder's own fields, so
  // members would be reloaded after every literal), every iteration starts with a full bit
  // buffer and the NEXT symbol's table entry already loaded — the load is issued before the
  // match copy of the current symbol, off the critical path — and a match first copies 32 bytes
  // unconditionally, which covers most matches without a data-dependent loop.  Everything else
  // (the last bytes of the input, a nearly full buffer) goes through the careful loop below.
  bool block_body(const uint32_t* lt, const uint32_t* dt) {
    constexpr uint32_t lit_mask = (1u << kLitRoot) - 1, dist_mask = (1u << kDistRoot) - 1;
    constexpr size_t kSlack = 3 + 258 + 40;  // three literals, or the longest match + copy overshoot
    for (;;) {
      if ((size_t)(out_end - op) < 2 * kSlack) grow(2 * kSlack);
      if (in_end - ip >= 16) {
        uint64_t b = bits;
        int nb = n_bits;
        const unsigned char* in = ip;
        unsigned char* o = op;
        const unsigned char* const in_fast_end = in_end - 16;
        unsigned char* const out_fast_end = out_end - kSlack;
        int status = 0;  // 1 end of block, -1 invalid
#define SGH_REFILL()                 \
  do {                               \
    b |= load64(in) << nb;           \
    in += (63 - nb) >> 3;            \
    nb |= 56;                        \
  } while (0)
        SGH_REFILL();
        uint32_t e = lt[b & lit_mask];
        while (in <= in_fast_end && o <= out_fast_end) {
          if (e & kSecondLevel) {
            b >>= kLitRoot;
            nb -= kLitRoot;
            e = lt[(e >> 16) + (b & ((1u << ((e >> 8) & 15)) - 1))];
          }
          if (e & kLiteral) {
            b >>= (e & 0xFF);
            nb -= (int)(e & 0xFF);
            *o++ = (unsigned char)(e >> 16);
            // a second literal straight away when the bits allow it (>= 41 are left)
            e = lt[b & lit_mask];
            if ((e & (kLiteral | kSecondLevel)) == kLiteral) {
              b >>= (e & 0xFF);
              nb -= (int)(e & 0xFF);
              *o++ = (unsigned char)(e >> 16);
            }
            SGH_REFILL();
            e = lt[b & lit_mask];
            continue;
          }
          if (e & kSpecial) {
            if (e & kEndOfBlock) {
              b >>= (e & 0xFF);
              nb -= (int)(e & 0xFF);
              status = 1;
            } else {
              status = -1;
            }
            break;
          }
          // a match: <= 20 bits of length code + extra, <= 28 of distance code + extra: the 56
          // bits every iteration starts with cover both (a second-level entry re-shifted above)
          b >>= (e & 0xFF);
          nb -= (int)(e & 0xFF);
          const uint32_t len_extra = (e >> 8) & 15;
          const uint32_t length = (e >> 16) + (uint32_t)(b & ((1u << len_extra) - 1));
          b >>= len_extra;
          nb -= (int)len_extra;
          if (nb < 32) SGH_REFILL();  // only after a second-level length code
          uint32_t d = dt[b & dist_mask];
          if (d & kSecondLevel) {
            b >>= kDistRoot;
            nb -= kDistRoot;
            d = dt[(d >> 16) + (b & ((1u << ((d >> 8) & 15)) - 1))];
          }
          if (d & kSpecial) {
            status = -1;
            break;
          }
          b >>= (d & 0xFF);
          nb -= (int)(d & 0xFF);
          const uint32_t dist_extra = (d >> 8) & 15;
          const size_t distance = (d >> 16) + (size_t)(b & ((1u << dist_extra) - 1));
          b >>= dist_extra;
          nb -= (int)dist_extra;
          if (distance > (size_t)(o - out_begin)) {
            status = -1;
            break;
          }
          // next symbol's entry: in flight during the copy
          SGH_REFILL();
          e = lt[b & lit_mask];
          const unsigned char* src = o - distance;
          unsigned char* dst = o;
          o += length;
          if (distance >= 8) {
            store64(dst, load64(src));
            store64(dst + 8, load64(src + 8));
            store64(dst + 16, load64(src + 16));
            store64(dst + 24, load64(src + 24));
            if (length > 32) {
              dst += 32;
              src += 32;
              do {
                store64(dst, load64(src));
                store64(dst + 8, load64(src + 8));
                dst += 16;
                src += 16;
              } while (dst < o);
            }
          } else if (distance == 1) {
            const uint64_t v = 0x0101010101010101ull * src[0];
            do {
              store64(dst, v);
              store64(dst + 8, v);
              dst += 16;
            } while (dst < o);
          } else {
            do {
              *dst++ = *src++;
            } while (dst < o);
          }
        }
#undef SGH_REFILL
        // the entry in `e` was only looked up, not consumed: the bit buffer is exactly where the
        // next symbol starts
        bits = b;
        n_bits = nb;
        ip = in;
        op = o;
        if (status < 0) return false;
        if (status > 0) return true;
        continue;  // out of input or output room for the fast loop: re-check, then go careful
      }
      // ---- careful: the last bytes of the input (n_bits may run negative: truncated stream) ----
      auto lookup = [&](const uint32_t* table, uint32_t mask, int root) {
        uint32_t e = table[bits & mask];
        if (e & kSecondLevel) {
          bits >>= root;
          n_bits -= root;
          e = table[(e >> 16) + (bits & ((1u << ((e >> 8) & 15)) - 1))];
        }
        return e;
      };
      refill();
      uint32_t e = lookup(lt, lit_mask, kLitRoot);
      if (e & kLiteral) {
        take((int)(e & 0xFF));
        *op++ = (unsigned char)(e >> 16);
        if (n_bits < 0) return false;
        continue;
      }
      if (e & kSpecial) {
        if (!(e & kEndOfBlock)) return false;
        take((int)(e & 0xFF));
        return n_bits >= 0;
      }
      take((int)(e & 0xFF));
      const uint32_t length = (e >> 16) + take((int)((e >> 8) & 15));
      refill();
      const uint32_t d = lookup(dt, dist_mask, kDistRoot);
      if (d & kSpecial) return false;
      take((int)(d & 0xFF));
      const size_t distance = (d >> 16) + take((int)((d >> 8) & 15));
      if (n_bits < 0 || distance > (size_t)(op - out_begin)) return false;
      const unsigned char* src = op - distance;
      unsigned char* dst = op;
      op += length;
      do {
        *dst++ = *src++;
      } while (dst < op);
    }
  }

  bool run() {
    if (!kFixed.ok) return false;
    for (;;) {
      refill();
      const uint32_t final_block = take(1), type = take(2);
      if (n_bits < 0) return false;
      bool ok;
      if (type == 0)
        ok = stored_block();
      else if (type == 1)
        ok = block_body(kFixed.lit, kFixed.dist);
      else if (type == 2)
        ok = dynamic_header() && block_body(lit, dist);
      else
        ok = false;
      if (!ok) return false;
      if (final_block) break;
    }
    take(n_bits & 7);   // the trailer starts on a byte boundary
    ip -= n_bits >> 3;  // bytes fetched but not used
    bits = 0;
    n_bits = 0;
    return true;
  }
};

// ---- CRC-32 (the gzip trailer) -----------------------------------------------------------------
// zlib's crc32 runs at ~3 GB/s, a fifth of a member's decode time.  On x86-64 with PCLMULQDQ the
// bulk is folded 64 bytes per step with carry-less multiplies (Gopal et al., "Fast CRC
// Computation for Generic Polynomials Using PCLMULQDQ", reflected polynomial 0xEDB88320); head
// and tail go through zlib.  The routine is self-checked against zlib once per process and not
// used if the two ever disagree.
#if defined(__x86_64__)
}  // namespace
}  // namespace sgh
#include <immintrin.h>
namespace sgh {
namespace {
// acc * x^128 mod P, plus the next 16 bytes
__attribute__((target("pclmul,sse4.1"))) inline __m128i fold128(__m128i acc, __m128i k3k4, __m128i next) {
  const __m128i lo = _mm_clmulepi64_si128(acc, k3k4, 0x00);
  return _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(acc, k3k4, 0x11), lo), next);
}
// raw CRC state in, raw state out; len a multiple of 16 and >= 64
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_fold(uint32_t state, const unsigned char* buf, size_t len) {
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);  // x^(4*128+32), x^(4*128-32) mod P
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);  // x^(128+32), x^(128-32) mod P
  const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);               // x^64 mod P
  const __m128i poly_mu = _mm_set_epi64x(0x01f7011641, 0x01db710641);
  const __m128i* p = reinterpret_cast<const __m128i*>(buf);
  __m128i x1 = _mm_loadu_si128(p + 0), x2 = _mm_loadu_si128(p + 1), x3 = _mm_loadu_si128(p + 2),
          x4 = _mm_loadu_si128(p + 3);
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
  p += 4;
  len -= 64;
  while (len >= 64) {
    const __m128i y1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), y2 = _mm_clmulepi64_si128(x2, k1k2, 0x00),
                  y3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), y4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
    x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
    x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
    x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, y1), _mm_loadu_si128(p + 0));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, y2), _mm_loadu_si128(p + 1));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, y3), _mm_loadu_si128(p + 2));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, y4), _mm_loadu_si128(p + 3));
    p += 4;
    len -= 64;
  }
  x1 = fold128(x1, k3k4, x2);
  x1 = fold128(x1, k3k4, x3);
  x1 = fold128(x1, k3k4, x4);
  while (len >= 16) {
    x1 = fold128(x1, k3k4, _mm_loadu_si128(p));
    ++p;
    len -= 16;
  }
  // 128 -> 64 -> 32 bits, then Barrett reduction
  const __m128i mask32 = _mm_setr_epi32(-1, 0, -1, 0);
  __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
  t = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, mask32);
  x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5, 0x00), t);
  t = _mm_and_si128(x1, mask32);
  t = _mm_clmulepi64_si128(t, poly_mu, 0x10);
  t = _mm_and_si128(t, mask32);
  t = _mm_clmulepi64_si128(t, poly_mu, 0x00);
  x1 = _mm_xor_si128(x1, t);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
uint32_t crc32_with_fold(const unsigned char* buf, size_t n) {
  uint32_t crc = 0;
  const size_t bulk = n & ~(size_t)15;
  if (bulk >= 64) {
    crc = ~crc32_fold(~crc, buf, bulk);
    buf += bulk;
    n -= bulk;
  }
  return (uint32_t)crc32_z(crc, buf, n);
}
bool fold_is_trustworthy() {
  if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
  unsigned char probe[1000];
  uint32_t x = 0x2545F491u;
  for (auto& b : probe) {
    x = x * 1664525u + 1013904223u;
    b = (unsigned char)(x >> 24);
  }
  for (size_t n : {64u, 80u, 127u, 128u, 333u, 1000u})
    for (size_t skew : {0u, 1u, 7u})
      if (n + skew <= sizeof probe &&
          crc32_with_fold(probe + skew, n) != (uint32_t)crc32_z(0L, probe + skew, n))
        return false;
  return true;
}
const bool kFoldOk = fold_is_trustworthy();
uint32_t member_crc32(const unsigned char* buf, size_t n) {
  return kFoldOk ? crc32_with_fold(buf, n) : (uint32_t)crc32_z(0L, buf, n);
}
#else
uint32_t member_crc32(const unsigned char* buf, size_t n) { return (uint32_t)crc32_z(0L, buf, n); }
#endif

}  // namespace

bool crc32_fold_in_use() {
#if defined(__x86_64__)
  return kFoldOk;
#else
  return false;
#endif
}

bool gunzip_member(const unsigned char* in, size_t in_len, Bytes& out, size_t& consumed) {
  // RFC 1952 header
  if (in_len < 18 || in[0] != 0x1f || in[1] != 0x8b || in[2] != 8 || (in[3] & 0xE0)) return false;
  const unsigned flags = in[3];
  const unsigned char* p = in + 10;
  const unsigned char* const end = in + in_len;
  if (flags & 4) {  // FEXTRA
    if (end - p < 2) return false;
    const size_t xlen = p[0] | (p[1] << 8);
    p += 2;
    if ((size_t)(end - p) < xlen) return false;
    p += xlen;
  }
  for (unsigned f : {8u, 16u}) {  // FNAME, FCOMMENT: zero-terminated
    if (!(flags & f)) continue;
    const void* z = memchr(p, 0, (size_t)(end - p));
    if (!z) return false;
    p = static_cast<const unsigned char*>(z) + 1;
  }
  if (flags & 2) {  // FHCRC
    if (end - p < 2) return false;
    p += 2;
  }
  // the decoder is large (its tables): keep it off small thread stacks
  struct Holder {
    Decoder* d;
    ~Holder() { delete d; }
  } h{new Decoder(p, end, out)};
  Decoder& dec = *h.d;
  if (!dec.run()) return false;
  if (dec.in_end - dec.ip < 8) return false;
  const unsigned char* t = dec.ip;
  const uint32_t crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
  const uint32_t isize = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
  const size_t n = (size_t)(dec.op - dec.out_begin);
  out.resize(n);
  if ((uint32_t)n != isize) return false;
  if (member_crc32(reinterpret_cast<const unsigned char*>(out.data()), n) != crc) return false;
  consumed = (size_t)(t + 8 - in);
  return true;
}

}  // namespace sgh
