// inflate.h — a zlib-stream decoder for the PNG reader (RFC 1950 / RFC 1951), written for throughput on the literal-heavy
// streams 16-bit camera images produce: a 64-bit bit buffer refilled with unaligned 8-byte loads, one table lookup per
// literal / length symbol (11 index bits, second-level tables for longer codes), 8-byte-chunk match copies.  With 80 PNGs
// per 16-camera frame, inflate on the host threads is the largest share of DerpCLI's files-in -> files-out time once the
// GPU work is 0.2 s (profiles/README.md).  The reader falls back to zlib's uncompress() whenever this decoder declines a
// stream (it never produces different bytes for a stream it accepts: tests/test_apps.py::test_png_inflate_equals_zlib checks
// it against zlib on stored / fixed / dynamic blocks, compression levels and strategies, long matches at the maximum
// distance and flush points; an AddressSanitizer run over 30 000 valid and mutated streams found no overrun).  Little-endian
// hosts only (8-byte loads into the bit buffer).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace io {
namespace inflate {

struct Table {  // entry: bits 0-3 code length consumed at this level (0 = invalid), bit 4 = link to a sub-table,
                // bits 5-8 sub-table index width, bit 9 = literal (symbol < 256, first-level entries only),
                // bits 16-31 symbol or sub-table offset
  std::vector<uint32_t> e;
  int rootBits = 0;
};

// canonical Huffman decoding table from code lengths (RFC 1951 3.2.2); false if the lengths over-subscribe the code space
inline bool buildTable(const uint8_t* lens, int n, int rootBits, Table* t) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) ++count[lens[i]];
  count[0] = 0;
  int maxLen = 0;
  for (int l = 1; l < 16; ++l)
    if (count[l]) maxLen = l;
  t->rootBits = rootBits;
  t->e.assign((size_t)1 << rootBits, 0);
  if (maxLen == 0) return true;  // no codes: every lookup is invalid
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
  }
  uint16_t next[16];
  {
    int code = 0;
    for (int l = 1; l < 16; ++l) {
      code = (code + count[l - 1]) << 1;
      next[l] = (uint16_t)code;
    }
  }
  // codes longer than rootBits share a root slot (their first rootBits bits, reversed) that links to a sub-table sized for
  // the longest code under that prefix
  std::vector<int> subLen((size_t)1 << rootBits, 0);
  std::vector<uint16_t> codeOf(n);
  for (int i = 0; i < n; ++i) {
    const int l = lens[i];
    if (!l) continue;
    const unsigned code = next[l]++;
    unsigned rev = 0;
    for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
    codeOf[i] = (uint16_t)rev;  // bit-reversed: deflate packs Huffman codes starting from the most significant bit
    if (l > rootBits) {
      const unsigned root = rev & ((1u << rootBits) - 1);
      if (l - rootBits > subLen[root]) subLen[root] = l - rootBits;
    }
  }
  for (size_t r = 0; r < subLen.size(); ++r)
    if (subLen[r]) {
      const size_t offset = t->e.size();
      if (offset > 0xFFFF) return false;
      t->e.resize(offset + ((size_t)1 << subLen[r]), 0);
      t->e[r] = (uint32_t)rootBits | 16u | ((uint32_t)subLen[r] << 5) | ((uint32_t)offset << 16);
    }
  for (int i = 0; i < n; ++i) {
    const int l = lens[i];
    if (!l) continue;
    const unsigned rev = codeOf[i];
    if (l <= rootBits) {
      const uint32_t entry = (uint32_t)l | (i < 256 ? 512u : 0u) | ((uint32_t)i << 16);
      for (unsigned k = rev; k < (1u << rootBits); k += 1u << l) t->e[k] = entry;
    } else {
      const unsigned root = rev & ((1u << rootBits) - 1);
      const uint32_t link = t->e[root];
      const int width = (int)((link >> 5) & 15u);
      const size_t offset = link >> 16;
      const uint32_t entry = (uint32_t)(l - rootBits) | ((uint32_t)i << 16);
      for (unsigned k = rev >> rootBits; k < (1u << width); k += 1u << (l - rootBits)) t->e[offset + k] = entry;
    }
  }
  return true;
}

struct Bits {
  const uint8_t *p, *end;
  uint64_t buf = 0;
  int n = 0;  // valid bits in buf
  void refill() {
    if (end - p >= 8) {  // branch-free bulk refill: top up to at least 56 bits
      uint64_t v;
      std::memcpy(&v, p, 8);
      buf |= v << n;
      const int take = (63 - n) >> 3;
      p += take;
      n += take * 8;
    } else {
      while (n <= 56 && p < end) {
        buf |= (uint64_t)*p++ << n;
        n += 8;
      }
    }
  }
  void refillBulk() {  // caller guarantees end - p >= 8
    uint64_t v;
    std::memcpy(&v, p, 8);
    buf |= v << n;
    const int take = (63 - n) >> 3;
    p += take;
    n += take * 8;
  }
  uint32_t peek(int bits) const { return (uint32_t)(buf & (((uint64_t)1 << bits) - 1)); }
  void drop(int bits) {
    buf >>= bits;
    n -= bits;
  }
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// one symbol: returns the symbol, or -1 for an invalid code / truncated input (needs <= 15 bits in the buffer)
inline int decode(const Table& t, Bits& in) {
  uint32_t e = t.e[in.peek(t.rootBits)];
  if (e & 16u) {
    in.drop((int)(e & 15u));
    e = t.e[(e >> 16) + in.peek((int)((e >> 5) & 15u))];
  }
  const int len = (int)(e & 15u);
  if (len == 0 || len > in.n) return -1;
  in.drop(len);
  return (int)(e >> 16);
}

inline uint32_t adler32(const uint8_t* d, size_t n) {
  uint32_t a = 1, b = 0;
  while (n) {
    size_t k = n < 5552 ? n : 5552;  // the largest run for which the 32-bit sums cannot overflow before the reduction
    n -= k;
    // b grows by a after every byte: over 8 bytes that is 8 a + 8 d0 + 7 d1 + ... + 1 d7
    for (; k >= 8; k -= 8, d += 8) {
      const uint32_t s = (uint32_t)d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7];
      b += 8 * a + 8u * d[0] + 7u * d[1] + 6u * d[2] + 5u * d[3] + 4u * d[4] + 3u * d[5] + 2u * d[6] + d[7];
      a += s;
    }
    for (; k; --k, ++d) {
      a += *d;
      b += a;
    }
    a %= 65521;
    b %= 65521;
  }
  return (b << 16) | a;
}

// zlib stream -> exactly outLen bytes.  false = declined (malformed, truncated, preset dictionary, wrong size or checksum).
inline bool zlibDecode(const uint8_t* src, size_t n, uint8_t* out, size_t outLen) {
  if (n < 6) return false;
  if ((src[0] & 15) != 8 || (src[0] >> 4) > 7 || ((src[0] << 8) | src[1]) % 31 != 0 || (src[1] & 32)) return false;
  Bits in{src + 2, src + n - 4};
  size_t at = 0;
  Table lit, dist;
  bool last = false;
  while (!last) {
    in.refill();
    if (in.n < 3) return false;
    last = in.peek(1);
    in.drop(1);
    const uint32_t type = in.peek(2);
    in.drop(2);
    if (type == 0) {  // stored: skip to the byte boundary, LEN / NLEN, raw bytes
      in.drop(in.n & 7);
      in.refill();
      if (in.n < 32) return false;
      const uint32_t len = in.peek(16);
      in.drop(16);
      const uint32_t nlen = in.peek(16);
      in.drop(16);
      if ((len ^ nlen) != 0xFFFFu) return false;
      // give the buffered whole bytes back to the byte stream
      in.p -= in.n >> 3;
      in.buf = 0;
      in.n = 0;
      if ((size_t)(in.end - in.p) < len || outLen - at < len) return false;
      std::memcpy(out + at, in.p, len);
      in.p += len;
      at += len;
      continue;
    }
    if (type == 3) return false;
    uint8_t lens[320];
    int nlit, ndist;
    if (type == 1) {
      nlit = 288;
      ndist = 30;
      for (int i = 0; i < 144; ++i) lens[i] = 8;
      for (int i = 144; i < 256; ++i) lens[i] = 9;
      for (int i = 256; i < 280; ++i) lens[i] = 7;
      for (int i = 280; i < 288; ++i) lens[i] = 8;
      for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
    } else {
      in.refill();
      if (in.n < 14) return false;
      nlit = (int)in.peek(5) + 257;
      in.drop(5);
      ndist = (int)in.peek(5) + 1;
      in.drop(5);
      const int ncode = (int)in.peek(4) + 4;
      in.drop(4);
      if (nlit > 286 || ndist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (int i = 0; i < ncode; ++i) {
        in.refill();
        if (in.n < 3) return false;
        cl[order[i]] = (uint8_t)in.peek(3);
        in.drop(3);
      }
      Table ct;
      if (!buildTable(cl, 19, 7, &ct)) return false;
      int i = 0;
      while (i < nlit + ndist) {
        in.refill();
        const int s = decode(ct, in);
        if (s < 0) return false;
        if (s < 16) {
          lens[i++] = (uint8_t)s;
          continue;
        }
        int rep, extra, base;
        uint8_t v = 0;
        if (s == 16) {
          if (i == 0) return false;
          v = lens[i - 1];
          extra = 2;
          base = 3;
        } else if (s == 17) {
          extra = 3;
          base = 3;
        } else {
          extra = 7;
          base = 11;
        }
        if (in.n < extra) return false;
        rep = base + (int)in.peek(extra);
        in.drop(extra);
        if (i + rep > nlit + ndist) return false;
        while (rep--) lens[i++] = v;
      }
      if (lens[256] == 0) return false;
      std::memmove(lens + 288, lens + nlit, (size_t)ndist);  // distance lengths to a fixed place
      std::memset(lens + nlit, 0, (size_t)(288 - nlit));
    }
    if (!buildTable(lens, 288, 11, &lit) || !buildTable(lens + 288, ndist, 8, &dist)) return false;

    for (;;) {
      in.refill();
      // fast path: enough input for a whole length/distance pair (<= 48 bits) after every refill and room for the longest
      // match plus the copy overshoot
      while (in.end - in.p >= 16 && outLen - at >= 258 + 8 + 4) {
        // up to four literals per refill: first-level literal codes take at most 11 bits each, the buffer holds >= 56
        const uint32_t* const T = lit.e.data();
        uint32_t e = T[in.buf & 2047];
        if (e & 512u) {
          in.drop((int)(e & 15u));
          out[at++] = (uint8_t)(e >> 16);
          e = T[in.buf & 2047];
          if (e & 512u) {
            in.drop((int)(e & 15u));
            out[at++] = (uint8_t)(e >> 16);
            e = T[in.buf & 2047];
            if (e & 512u) {
              in.drop((int)(e & 15u));
              out[at++] = (uint8_t)(e >> 16);
              e = T[in.buf & 2047];
              if (e & 512u) {
                in.drop((int)(e & 15u));
                out[at++] = (uint8_t)(e >> 16);
                in.refillBulk();
                continue;
              }
            }
          }
        }
        in.refillBulk();  // a whole length / distance pair needs at most 15 + 5 + 15 + 13 = 48 bits
        const int s = decode(lit, in);
        if (s < 0) return false;
        if (s < 256) {
          out[at++] = (uint8_t)s;
          in.refillBulk();
          continue;
        }
        if (s == 256) goto blockDone;
        if (s > 285) return false;
        const int len = kLenBase[s - 257] + (int)in.peek(kLenExtra[s - 257]);
        in.drop(kLenExtra[s - 257]);
        const int ds = decode(dist, in);
        if (ds < 0 || ds > 29) return false;
        const size_t d = kDistBase[ds] + (size_t)in.peek(kDistExtra[ds]);
        in.drop(kDistExtra[ds]);
        if (d > at) return false;
        uint8_t* to = out + at;
        const uint8_t* from = to - d;
        if (d >= 8) {
          for (int k = 0; k < len; k += 8) std::memcpy(to + k, from + k, 8);  // may write up to 7 bytes past the match
        } else {
          for (int k = 0; k < len; ++k) to[k] = from[k];
        }
        at += (size_t)len;
        in.refillBulk();
      }
      // careful path near the ends
      in.refill();
      const int s = decode(lit, in);
      if (s < 0) return false;
      if (s < 256) {
        if (at >= outLen) return false;
        out[at++] = (uint8_t)s;
        continue;
      }
      if (s == 256) break;
      if (s > 285) return false;
      in.refill();
      if (in.n < kLenExtra[s - 257]) return false;
      const size_t len = kLenBase[s - 257] + (size_t)in.peek(kLenExtra[s - 257]);
      in.drop(kLenExtra[s - 257]);
      in.refill();
      const int ds = decode(dist, in);
      if (ds < 0 || ds > 29) return false;
      if (in.n < kDistExtra[ds]) return false;
      const size_t d = kDistBase[ds] + (size_t)in.peek(kDistExtra[ds]);
      in.drop(kDistExtra[ds]);
      if (d > at || outLen - at < len) return false;
      for (size_t k = 0; k < len; ++k) out[at + k] = out[at + k - d];
      at += len;
    }
  blockDone:;
  }
  if (at != outLen) return false;
  const uint8_t* tail = src + n - 4;
  const uint32_t want = ((uint32_t)tail[0] << 24) | ((uint32_t)tail[1] << 16) | ((uint32_t)tail[2] << 8) | tail[3];
  return adler32(out, outLen) == want;
}

}  // namespace inflate
}  // namespace io
