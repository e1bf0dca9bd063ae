// inflate.h — whole-member gzip decoder of the host ingest path.
//
// What the reference gets from the `flate2` crate (through `fxread`, count.rs:24): the bytes of
// every gzip member, CRC checked.  zlib's streaming inflate spends most of its time on generic
// machinery this path does not need (a sliding window, byte-wise input, resumability); here a
// member is decoded in ONE call from a memory-mapped file into ONE contiguous buffer, so
// matches copy straight from earlier output, the bit buffer is refilled 8 bytes at a time and
// literals are decoded several per refill.  The result is accepted only if the member's CRC-32
// and ISIZE trailer agree; on ANY doubt the caller falls back to zlib, so the bytes delivered
// are always zlib-identical.
#pragma once

#include <cstddef>
#include <cstdint>

#include "fastx.h"

namespace sgh {

// Decodes the gzip member that starts at `in`.  On success returns true, `out` holds the
// member's bytes and `consumed` the compressed size of the member (header + deflate stream +
// trailer).  Returns false if the bytes are not a complete, valid member (the caller decides
// what that means); `out` is unspecified then.  Never reads outside [in, in + in_len).
bool gunzip_member(const unsigned char* in, size_t in_len, Bytes& out, size_t& consumed);

// whether the carry-less-multiply CRC passed its self-check against zlib (diagnostic)
bool crc32_fold_in_use();

}  // namespace sgh
