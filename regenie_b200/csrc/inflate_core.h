// DEFLATE (RFC 1951) / zlib (RFC 1950) decoder for BGEN genotype payloads, written once for the device and for the host.
//
// BGEN v1.2 stores every variant's probability block as one zlib stream (compression flag 1); the reference inflates
// them on the host inside its OpenMP loop (`uncompress`, src/Geno.cpp:1608, :2207).  Here one WARP owns one stream:
// lane 0 walks the bit stream (Huffman decode is sequential), stores literals itself and hands every match
// (length, distance) to the whole warp, which copies it with coalesced byte accesses.  The same source compiles for
// the host with a "warp" of one lane, which is how tests/test_host_cpu.py checks it against zlib without a GPU
// (`rgb200_hostprobe inflate-bgen`).
//
// Tables per stream (the caller provides the storage: shared memory on the device, the stack on the host):
//   fast tables   lit/len: 2^10 entries, distance: 2^8 entries; entry = (symbol << 4) | code length, 0 = "longer code"
//   canonical     count[16] + symbols sorted by code (the classic bit-serial decode) for codes longer than the fast index
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define RGI_HD __device__ __forceinline__
#define RGI_LANE ((int)(threadIdx.x & 31))
#define RGI_LANES 32
#define RGI_BCAST(x) __shfl_sync(0xffffffffu, (x), 0)
#define RGI_SYNC() __syncwarp()
#else
#define RGI_HD inline
#define RGI_LANE 0
#define RGI_LANES 1
#define RGI_BCAST(x) (x)
#define RGI_SYNC() ((void)0)
#endif

// copy loops of the ring variant: lane l takes elements first + l, first + l + LANES, ...  On the host the elements of a
// 32-wide batch can be visited in descending order instead (RGI_HOST_REVERSED_BATCHES, used by tools/inflate_fuzz.cpp) to
// show that the result does not depend on the order in which the lanes of a warp get to their loads and stores.
#if defined(__CUDA_ARCH__) || !defined(RGI_HOST_REVERSED_BATCHES)
#define RGI_FOR_LANES(i, first, n) for (uint32_t i = (first) + (uint32_t)RGI_LANE; i < (n); i += RGI_LANES)
#else
#define RGI_FOR_LANES(i, first, n)                                                               \
  for (uint32_t i##_b = (first); i##_b < (n); i##_b += 32)                                       \
    for (uint32_t i##_k = ((n) - i##_b < 32 ? (n) - i##_b : 32), i = i##_b + i##_k - 1; i##_k > 0; --i##_k, --i)
#endif

namespace rgi {

constexpr int kLitBits = 10, kDistBits = 8;
constexpr int kMaxBits = 15, kMaxLitCodes = 288, kMaxDistCodes = 30;

enum Status : int {
  kOk = 0,
  kErrHeader = 1,        // not a zlib stream (CM != 8, window > 32K, FDICT set, or header checksum)
  kErrBlockType = 2,     // reserved block type 3
  kErrStored = 3,        // LEN / NLEN mismatch
  kErrCodeLengths = 4,   // over-subscribed or incomplete code-length / literal / distance code
  kErrSymbol = 5,        // invalid literal/length or distance symbol, or no end-of-block code
  kErrDistance = 6,      // distance reaches before the start of the output
  kErrOutput = 7,        // stream inflates to more bytes than the declared length
  kErrInput = 8,         // ran past the end of the compressed bytes
  kErrLength = 9,        // stream ended with fewer bytes than the declared length
  kErrAdler = 10,        // Adler-32 of the output does not match the trailer
};

struct Tables {
  uint16_t lit_fast[1 << kLitBits];
  uint16_t dist_fast[1 << kDistBits];
  uint16_t lit_count[kMaxBits + 1], dist_count[kMaxBits + 1];
  uint16_t lit_sym[kMaxLitCodes], dist_sym[kMaxDistCodes];
};

// LSB-first bit reader over [in, in + n); reading past the end yields zero bits and sets `over`
struct Bits {
  const uint8_t* in;
  uint32_t n, pos;
  uint64_t buf;
  int cnt;
  bool over;
};

RGI_HD void bits_init(Bits& b, const uint8_t* in, uint32_t n) {
  b.in = in; b.n = n; b.pos = 0; b.buf = 0; b.cnt = 0; b.over = false;
}
RGI_HD void bits_fill(Bits& b) {                 // at least 56 valid bits afterwards (zeros past the end)
  while (b.cnt <= 56) {
    uint64_t v = 0;
    if (b.pos < b.n) v = b.in[b.pos];
    else if (b.pos >= b.n + 8) b.over = true;    // a well-formed stream never needs more than its trailer
    ++b.pos;
    b.buf |= v << b.cnt;
    b.cnt += 8;
  }
}
RGI_HD uint32_t bits_peek(const Bits& b, int k) { return (uint32_t)(b.buf & ((1ull << k) - 1)); }
RGI_HD void bits_drop(Bits& b, int k) { b.buf >>= k; b.cnt -= k; }
RGI_HD uint32_t bits_get(Bits& b, int k) {       // k <= 16
  if (b.cnt < k) bits_fill(b);
  const uint32_t v = bits_peek(b, k);
  bits_drop(b, k);
  return v;
}

RGI_HD uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1); code >>= 1; }
  return r;
}

// canonical Huffman code from `n` code lengths: count / sorted symbols (bit-serial decode) and the fast table
// indexed by the next `fast_bits` stream bits.  Returns 0 for a complete code, <0 over-subscribed, >0 incomplete.
RGI_HD int build_code(const uint8_t* length, int n, uint16_t* count, uint16_t* symbol, uint16_t* fast, int fast_bits) {
  for (int l = 0; l <= kMaxBits; ++l) count[l] = 0;
  for (int s = 0; s < n; ++s) ++count[length[s]];
  for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
  if (count[0] == n) return 0;                   // no codes at all: complete, decoding any symbol fails later
  int left = 1;
  for (int l = 1; l <= kMaxBits; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return left;
  }
  uint16_t offs[kMaxBits + 1];
  offs[1] = 0;
  for (int l = 1; l < kMaxBits; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
  for (int s = 0; s < n; ++s)
    if (length[s] != 0) symbol[offs[length[s]]++] = (uint16_t)s;
  // fast table: walk the canonical codes in order
  uint32_t code = 0;
  int idx = 0;
  for (int l = 1; l <= fast_bits; ++l) {
    for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
      const uint32_t r = reverse_bits(code, l);
      const uint16_t e = (uint16_t)((symbol[idx] << 4) | l);
      for (uint32_t i = r; i < (1u << fast_bits); i += (1u << l)) fast[i] = e;
    }
    code <<= 1;
  }
  return left;
}

// next symbol of a code: fast path on the table, bit-serial canonical decode for longer codes; -1 = invalid
RGI_HD int decode_symbol(Bits& b, const uint16_t* count, const uint16_t* symbol, const uint16_t* fast, int fast_bits) {
  if (b.cnt < kMaxBits) bits_fill(b);
  const uint16_t e = fast[bits_peek(b, fast_bits)];
  if (e != 0) {
    bits_drop(b, e & 15);
    return e >> 4;
  }
  int code = 0, first = 0, index = 0;
  uint64_t bitbuf = b.buf;
  for (int len = 1; len <= kMaxBits; ++len) {
    code |= (int)(bitbuf & 1);
    bitbuf >>= 1;
    const int cnt = count[len];
    if (code - cnt < first) {
      bits_drop(b, len);
      return symbol[index + (code - first)];
    }
    index += cnt;
    first += cnt;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

RGI_HD void build_fixed(Tables& t) {
  uint8_t len[kMaxLitCodes];
  int s = 0;
  for (; s < 144; ++s) len[s] = 8;
  for (; s < 256; ++s) len[s] = 9;
  for (; s < 280; ++s) len[s] = 7;
  for (; s < kMaxLitCodes; ++s) len[s] = 8;
  build_code(len, kMaxLitCodes, t.lit_count, t.lit_sym, t.lit_fast, kLitBits);
  for (s = 0; s < kMaxDistCodes; ++s) len[s] = 5;
  build_code(len, kMaxDistCodes, t.dist_count, t.dist_sym, t.dist_fast, kDistBits);
}

RGI_HD int build_dynamic(Bits& b, Tables& t) {
  const int nlen = (int)bits_get(b, 5) + 257, ndist = (int)bits_get(b, 5) + 1, ncode = (int)bits_get(b, 4) + 4;
  if (nlen > 286 || ndist > kMaxDistCodes) return kErrCodeLengths;
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint8_t len[kMaxLitCodes + kMaxDistCodes];
  for (int i = 0; i < 19; ++i) len[i] = 0;
  for (int i = 0; i < ncode; ++i) len[order[i]] = (uint8_t)bits_get(b, 3);
  // the code-length code is decoded bit-serially through the literal slots of the tables (rebuilt right after)
  if (build_code(len, 19, t.lit_count, t.lit_sym, t.dist_fast, 7) != 0) return kErrCodeLengths;
  int idx = 0;
  while (idx < nlen + ndist) {
    const int sym = decode_symbol(b, t.lit_count, t.lit_sym, t.dist_fast, 7);
    if (sym < 0) return kErrCodeLengths;
    if (sym < 16) {
      len[idx++] = (uint8_t)sym;
    } else {
      int prev = 0, rep;
      if (sym == 16) {
        if (idx == 0) return kErrCodeLengths;
        prev = len[idx - 1];
        rep = 3 + (int)bits_get(b, 2);
      } else if (sym == 17) {
        rep = 3 + (int)bits_get(b, 3);
      } else {
        rep = 11 + (int)bits_get(b, 7);
      }
      if (idx + rep > nlen + ndist) return kErrCodeLengths;
      while (rep--) len[idx++] = (uint8_t)prev;
    }
  }
  if (len[256] == 0) return kErrSymbol;          // no end-of-block code
  int err = build_code(len, nlen, t.lit_count, t.lit_sym, t.lit_fast, kLitBits);
  if (err < 0 || (err > 0 && nlen - t.lit_count[0] != 1)) return kErrCodeLengths;     // incomplete only if a single code
  err = build_code(len + nlen, ndist, t.dist_count, t.dist_sym, t.dist_fast, kDistBits);
  if (err < 0 || (err > 0 && ndist - t.dist_count[0] != 1)) return kErrCodeLengths;
  return kOk;
}

// Adler-32 over out[0, n) by the whole warp: lane l takes bytes l, l + LANES, ...; (a, b) are combined from the per-lane
// sums  a = 1 + sum d_i,  b = n + sum (n - i) d_i  (mod 65521)
RGI_HD uint32_t adler32_warp(const uint8_t* out, uint32_t n) {
  const uint32_t kMod = 65521;
  uint64_t sa = 0, sb = 0;
  uint32_t chunk = 0;
  for (uint32_t i = (uint32_t)RGI_LANE; i < n; i += RGI_LANES) {
    const uint64_t d = out[i];
    sa += d;
    sb += d * (uint64_t)((n - i) % kMod);
    if (++chunk == (1u << 20)) { sa %= kMod; sb %= kMod; chunk = 0; }
  }
  sa %= kMod; sb %= kMod;
#if defined(__CUDA_ARCH__)
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  sa %= kMod; sb %= kMod;
#endif
  const uint32_t a = (uint32_t)((1 + sa) % kMod), b2 = (uint32_t)((n % kMod + sb) % kMod);
  return (b2 << 16) | a;
}

// Inflate one zlib stream into out[0, out_len).  Called by all lanes of a warp with identical arguments (one lane on the
// host); returns the same status in every lane.  `t` is per-warp scratch.
RGI_HD int inflate_zlib(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& t, bool check_adler) {
  const int lane = RGI_LANE;
  Bits b;
  bits_init(b, in, in_len);
  uint32_t pos = 0;                               // bytes produced (kept by lane 0, broadcast at every hand-over)
  int status = kOk;
  if (lane == 0) {
    if (in_len < 6) status = kErrHeader;
    else {
      const uint32_t cmf = in[0], flg = in[1];
      if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32) || ((cmf << 8) | flg) % 31 != 0) status = kErrHeader;
      b.pos = 2;
    }
  }
  status = RGI_BCAST(status);
  if (status != kOk) return status;

  const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

  // lane-0 state machine: 0 = read a block header, 1 = inside a Huffman block, 2 = after the final block
  int mode = 0, last = 0;
  for (;;) {
    // ---- lane 0 runs until it has something for the warp: a match / stored run to copy (kind 1 / 2), the end (3)
    int kind = 0;
    uint32_t a0 = 0, a1 = 0;                       // match: (length, distance); stored: (length, input offset)
    if (lane == 0) {
      while (kind == 0 && status == kOk) {
        if (mode == 0) {
          if (last) { mode = 2; kind = 3; break; }
          last = (int)bits_get(b, 1);
          const int type = (int)bits_get(b, 2);
          if (type == 0) {
            bits_drop(b, b.cnt & 7);               // to the byte boundary; whole bytes still buffered are given back
            b.pos -= (uint32_t)(b.cnt >> 3);
            b.buf = 0; b.cnt = 0;
            if (b.pos + 4 > in_len) { status = kErrInput; break; }
            const uint32_t len = in[b.pos] | ((uint32_t)in[b.pos + 1] << 8);
            const uint32_t nlen = in[b.pos + 2] | ((uint32_t)in[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xffffu) != nlen) { status = kErrStored; break; }
            if (b.pos + len > in_len) { status = kErrInput; break; }
            if (pos + len > out_len) { status = kErrOutput; break; }
            if (len > 0) { kind = 2; a0 = len; a1 = b.pos; }
            b.pos += len;
          } else if (type == 1) {
            build_fixed(t);
            mode = 1;
          } else if (type == 2) {
            status = build_dynamic(b, t);
            mode = 1;
          } else {
            status = kErrBlockType;
          }
        } else {                                   // mode 1: symbols of the current block
          const int sym = decode_symbol(b, t.lit_count, t.lit_sym, t.lit_fast, kLitBits);
          if (sym < 0) { status = kErrSymbol; break; }
          if (sym < 256) {
            if (pos >= out_len) { status = kErrOutput; break; }
            out[pos++] = (uint8_t)sym;
          } else if (sym == 256) {
            mode = 0;
          } else {
            const int li = sym - 257;
            if (li >= 29) { status = kErrSymbol; break; }
            const uint32_t len = len_base[li] + bits_get(b, len_extra[li]);
            const int ds = decode_symbol(b, t.dist_count, t.dist_sym, t.dist_fast, kDistBits);
            if (ds < 0 || ds >= 30) { status = kErrSymbol; break; }
            uint32_t dist = dist_base[ds];
            const int de = dist_extra[ds];
            if (de > 0) {
              if (b.cnt < de) bits_fill(b);
              dist += bits_peek(b, de);
              bits_drop(b, de);
            }
            if (dist > pos) { status = kErrDistance; break; }
            if (pos + len > out_len) { status = kErrOutput; break; }
            kind = 1; a0 = len; a1 = dist;
          }
        }
        if (b.over) status = kErrInput;
      }
      if (status != kOk) kind = 3;
    }
    kind = RGI_BCAST(kind);
    if (kind == 3) break;
    a0 = RGI_BCAST(a0);
    a1 = RGI_BCAST(a1);
    const uint32_t p0 = RGI_BCAST(pos);
    RGI_SYNC();                                    // lane 0's literal stores are visible to the copying lanes
    if (kind == 1) {
      const uint8_t* src = out + p0 - a1;          // every source byte lies before p0: for distance < length the pattern
      if (a1 >= a0) {                              // of `distance` bytes repeats, so index modulo the distance
        for (uint32_t i = (uint32_t)lane; i < a0; i += RGI_LANES) out[p0 + i] = src[i];
      } else {
        for (uint32_t i = (uint32_t)lane; i < a0; i += RGI_LANES) out[p0 + i] = src[i % a1];
      }
    } else {
      const uint8_t* src = in + a1;
      for (uint32_t i = (uint32_t)lane; i < a0; i += RGI_LANES) out[p0 + i] = src[i];
    }
    RGI_SYNC();
    if (lane == 0) pos = p0 + a0;
  }
  status = RGI_BCAST(status);
  pos = RGI_BCAST(pos);
  if (status != kOk) return status;
  if (pos != out_len) return kErrLength;
  if (check_adler) {
    uint32_t want = 0;
    int st = kOk;
    if (lane == 0) {
      // the trailer follows the last block at the next byte boundary; whole bytes still in the bit buffer are given back
      const uint32_t tp = b.pos - (uint32_t)(b.cnt >> 3);
      if (tp + 4 > in_len) st = kErrInput;
      else want = ((uint32_t)in[tp] << 24) | ((uint32_t)in[tp + 1] << 16) | ((uint32_t)in[tp + 2] << 8) | in[tp + 3];
    }
    st = RGI_BCAST(st);
    if (st != kOk) return st;
    want = RGI_BCAST(want);
    RGI_SYNC();
    if (adler32_warp(out, out_len) != want) return kErrAdler;
  }
  return kOk;
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant with a per-stream output window in shared memory.  inflate_zlib above writes every byte straight to global
// memory and reads match sources back from it, so each match costs round trips to L2 and literals are single-byte
// global stores.  Here the last kWinBytes of output live in a ring the caller provides (shared memory on the device):
// lane 0 stores literals into the ring, matches up to kWinMaxDist back are copied ring -> ring by the warp, and the ring
// is written out to global memory in coalesced runs (`flush`).  Matches that reach further back (about 8 % of them at a
// 16 KB ring on BGEN payloads) flush first and read their source from global memory.  Same contract and status codes as
// inflate_zlib; the host build (one lane, `win` on the heap) goes through exactly the same ring arithmetic.
constexpr uint32_t kWinBytes = 16384;                  // power of two
constexpr uint32_t kWinMask = kWinBytes - 1;
constexpr uint32_t kWinMaxDist = kWinBytes - 258;      // a match of up to 258 bytes never overwrites its own source slots
constexpr uint32_t kWinPending = kWinBytes / 2;        // lane 0 asks for a flush once this many bytes are unwritten
constexpr uint32_t kWinChunk = 4096;                   // stored blocks are moved in chunks of this size

RGI_HD int inflate_zlib_window(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& t, uint8_t* win,
                               bool check_adler) {
  const int lane = RGI_LANE;
  Bits b;
  bits_init(b, in, in_len);
  uint32_t pos = 0, flushed = 0;                  // bytes produced / bytes already in global memory (lane 0's copy is the truth)
  int status = kOk;
  if (lane == 0) {
    if (in_len < 6) status = kErrHeader;
    else {
      const uint32_t cmf = in[0], flg = in[1];
      if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32) || ((cmf << 8) | flg) % 31 != 0) status = kErrHeader;
      b.pos = 2;
    }
  }
  status = RGI_BCAST(status);
  if (status != kOk) return status;

  const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

  // events lane 0 hands to the warp: 1 = match inside the ring, 2 = stored bytes, 3 = end, 4 = flush, 5 = match whose
  // source is read from global memory.  Events 2 and 5 are preceded by a flush (lane 0 emits event 4 first).
  int mode = 0, last = 0;
  uint32_t stored_left = 0;                       // bytes of the current stored block not yet handed over
  bool want_flush = false;                        // the next event must be a flush, then the held event is replayed
  int held_kind = 0;
  uint32_t held_a0 = 0, held_a1 = 0;
  for (;;) {
    int kind = 0;
    uint32_t a0 = 0, a1 = 0;
    if (lane == 0) {
      while (kind == 0 && status == kOk) {
        if (want_flush || pos - flushed > kWinPending) {
          want_flush = false;
          if (pos > flushed) { kind = 4; break; }
        }
        if (held_kind != 0) {                     // the event that had to wait for its flush
          kind = held_kind; a0 = held_a0; a1 = held_a1; held_kind = 0;
          break;
        }
        if (stored_left > 0) {                    // rest of a stored block, chunk by chunk
          const uint32_t n = stored_left < kWinChunk ? stored_left : kWinChunk;
          held_kind = 2; held_a0 = n; held_a1 = b.pos;
          b.pos += n; stored_left -= n;
          if (pos > flushed) { want_flush = true; continue; }
          kind = held_kind; a0 = held_a0; a1 = held_a1; held_kind = 0;
          break;
        }
        if (mode == 0) {
          if (last) { mode = 2; kind = 3; break; }
          last = (int)bits_get(b, 1);
          const int type = (int)bits_get(b, 2);
          if (type == 0) {
            bits_drop(b, b.cnt & 7);
            b.pos -= (uint32_t)(b.cnt >> 3);
            b.buf = 0; b.cnt = 0;
            if (b.pos + 4 > in_len) { status = kErrInput; break; }
            const uint32_t len = in[b.pos] | ((uint32_t)in[b.pos + 1] << 8);
            const uint32_t nlen = in[b.pos + 2] | ((uint32_t)in[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xffffu) != nlen) { status = kErrStored; break; }
            if (b.pos + len > in_len) { status = kErrInput; break; }
            if (pos + len > out_len) { status = kErrOutput; break; }
            stored_left = len;
          } else if (type == 1) {
            build_fixed(t);
            mode = 1;
          } else if (type == 2) {
            status = build_dynamic(b, t);
            mode = 1;
          } else {
            status = kErrBlockType;
          }
        } else {
          const int sym = decode_symbol(b, t.lit_count, t.lit_sym, t.lit_fast, kLitBits);
          if (sym < 0) { status = kErrSymbol; break; }
          if (sym < 256) {
            if (pos >= out_len) { status = kErrOutput; break; }
            win[pos & kWinMask] = (uint8_t)sym;
            ++pos;
          } else if (sym == 256) {
            mode = 0;
          } else {
            const int li = sym - 257;
            if (li >= 29) { status = kErrSymbol; break; }
            const uint32_t len = len_base[li] + bits_get(b, len_extra[li]);
            const int ds = decode_symbol(b, t.dist_count, t.dist_sym, t.dist_fast, kDistBits);
            if (ds < 0 || ds >= 30) { status = kErrSymbol; break; }
            uint32_t dist = dist_base[ds];
            const int de = dist_extra[ds];
            if (de > 0) {
              if (b.cnt < de) bits_fill(b);
              dist += bits_peek(b, de);
              bits_drop(b, de);
            }
            if (dist > pos) { status = kErrDistance; break; }
            if (pos + len > out_len) { status = kErrOutput; break; }
            if (dist <= kWinMaxDist) { kind = 1; a0 = len; a1 = dist; }
            else {                                // source in global memory: everything produced so far must be there
              held_kind = 5; held_a0 = len; held_a1 = dist;
              if (pos > flushed) { want_flush = true; }
              else { kind = 5; a0 = len; a1 = dist; held_kind = 0; }
            }
          }
        }
        if (b.over) status = kErrInput;
      }
      if (status != kOk) kind = 3;
    }
    kind = RGI_BCAST(kind);
    if (kind == 3) break;
    a0 = RGI_BCAST(a0);
    a1 = RGI_BCAST(a1);
    const uint32_t p0 = RGI_BCAST(pos);
    const uint32_t f0 = RGI_BCAST(flushed);
    RGI_SYNC();                                    // lane 0's ring stores are visible to the warp
    if (kind == 1) {                               // ring -> ring; sources lie before p0 (index modulo the distance when
      const uint32_t s0 = p0 - a1;                 // the match overlaps its own output)
      if (a1 >= a0) {
        RGI_FOR_LANES(i, 0u, a0) win[(p0 + i) & kWinMask] = win[(s0 + i) & kWinMask];
      } else {
        RGI_FOR_LANES(i, 0u, a0) win[(p0 + i) & kWinMask] = win[(s0 + i % a1) & kWinMask];
      }
    } else if (kind == 4) {                        // ring -> global, [f0, p0)
      RGI_FOR_LANES(i, f0, p0) out[i] = win[i & kWinMask];
    } else if (kind == 2) {                        // stored bytes: input -> global and ring (f0 == p0 here)
      const uint8_t* src = in + a1;
      RGI_FOR_LANES(i, 0u, a0) {
        const uint8_t v = src[i];
        out[p0 + i] = v;
        win[(p0 + i) & kWinMask] = v;
      }
    } else {                                       // kind 5: far match, source already in global memory (f0 == p0, a1 > a0)
      const uint8_t* src = out + (p0 - a1);
      RGI_FOR_LANES(i, 0u, a0) {
        const uint8_t v = src[i];
        out[p0 + i] = v;
        win[(p0 + i) & kWinMask] = v;
      }
    }
    RGI_SYNC();
    if (lane == 0) {
      if (kind == 4) flushed = p0;
      else {
        pos = p0 + a0;
        if (kind != 1) flushed = pos;
      }
    }
  }
  status = RGI_BCAST(status);
  pos = RGI_BCAST(pos);
  flushed = RGI_BCAST(flushed);
  if (status != kOk) return status;
  RGI_SYNC();
  RGI_FOR_LANES(i, flushed, pos) out[i] = win[i & kWinMask];   // the tail of the ring
  RGI_SYNC();
  if (pos != out_len) return kErrLength;
  if (check_adler) {
    uint32_t want = 0;
    int st = kOk;
    if (lane == 0) {
      const uint32_t tp = b.pos - (uint32_t)(b.cnt >> 3);
      if (tp + 4 > in_len) st = kErrInput;
      else want = ((uint32_t)in[tp] << 24) | ((uint32_t)in[tp + 1] << 16) | ((uint32_t)in[tp + 2] << 8) | in[tp + 3];
    }
    st = RGI_BCAST(st);
    if (st != kOk) return st;
    want = RGI_BCAST(want);
    RGI_SYNC();
    if (adler32_warp(out, out_len) != want) return kErrAdler;
  }
  return kOk;
}

}  // namespace rgi
