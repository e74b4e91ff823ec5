// Decoder core of PLINK 2 .pgen hard-call records -> PLINK 1 2-bit rows, written once for the device
// (csrc/pgen_decode.cu) and, with the lanes of a warp run one after the other, for the host (rgb200_hostprobe and the
// mock ABI of the tests), so that the arithmetic the GPU runs is checked bit for bit on the CPU against oracle/pgen.py.
//
// The reference reads every variant through the vendored pgenlib (`PgenReader::Read`, src/Geno.cpp:1773-1821 in Step 1,
// :2538-2594 / :2596-2712 in Step 2): a serial, per-variant, per-thread decode.  The record layouts handled here are the
// ones pgenlib documents (external_libs/pgenlib/include/pgenlib_read.cc: difflists :2177-2268, group byte counts :9617):
//   type 0      n 2-bit values
//   type 1      1 code byte (low value << 2 | step), n bits, then a difflist of the exceptions
//   type 4/6/7  every sample 0 / 2 / missing except the entries of a difflist;  type 5: every sample 0
//   type 2/3    a difflist against the most recent record that is not of type 2/3; type 3 swaps 0 <-> 2 afterwards
// A difflist: vint length L; per group of 64 entries the first sample id (1-4 bytes); per group but the last one byte =
// byte length of the group's delta stream - 63; L 2-bit values; the deltas as vints.  The per-group byte counts make the
// groups independent: one lane per group.
//
// Output coding: ALT count 0 / 1 / 2 / missing -> PLINK 1 codes 11 / 10 / 00 / 01 (ref-last), i.e. what host/pgen.cpp emits.
// A row is decoded in two steps: every 32-bit word (16 samples) is FILLED from the record that carries the dense part
// (the record itself, or the base of an LD record), then the difflists are PATCHED in (base first, then the record's own).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RGP_HD __host__ __device__ __forceinline__
#else
#define RGP_HD inline
#endif

namespace rgp {

constexpr uint32_t kGroup = 64;
enum : int { kOk = 0, kErrTruncated = 1, kErrSampleIdx = 2, kErrListLen = 3, kErrType = 4 };

struct Rec {                 // one record as staged for the device
  const uint8_t* p;          // first byte (4-byte aligned in the staging buffer)
  uint32_t len;              // bytes
  uint32_t type;             // low 3 bits of the variant record type
};

// 16 values 0..3 -> 16 PLINK 1 codes; inv swaps 0 <-> 2 first (type 3)
RGP_HD uint32_t bed_word(uint32_t w, bool inv) {
  const uint32_t m = 0x55555555u;
  uint32_t hi = (w >> 1) & m;
  const uint32_t lo = w & m;
  if (inv) hi ^= (~lo & m);
  return ((~hi & m) << 1) | (~(hi ^ lo) & m);
}

RGP_HD uint32_t bed_code(uint32_t v, bool inv) {
  if (inv && !(v & 1)) v ^= 2;
  return (0x4Bu >> (2 * v)) & 3;           // {3, 2, 0, 1}
}

RGP_HD uint32_t spread16(uint32_t x) {      // bit i -> bit 2i
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

// word w (samples 16 w .. 16 w + 15) of the dense part of record r; bits of samples >= n are zero
RGP_HD uint32_t fill_word(const Rec& r, uint32_t n, uint32_t w, bool inv, int* err) {
  const uint32_t s0 = w * 16;
  if (s0 >= n) return 0;
  uint32_t vals = 0;
  switch (r.type) {
    case 0: {
      const uint32_t nb = (n + 3) / 4;
      if (r.len < nb) { *err = kErrTruncated; return 0; }
      if (4 * w + 4 <= nb) {
        vals = *reinterpret_cast<const uint32_t*>(r.p + 4 * w);
      } else {
        for (uint32_t k = 4 * w; k < nb; ++k) vals |= (uint32_t)r.p[k] << (8 * (k - 4 * w));
      }
      break;
    }
    case 1: {
      const uint32_t nb = (n + 7) / 8;
      if (r.len < 1 + nb) { *err = kErrTruncated; return 0; }
      const uint32_t code = r.p[0], lo = code >> 2, step = code & 3;
      if (lo + step > 3) { *err = kErrType; return 0; }
      uint32_t bits = r.p[1 + 2 * w];
      if (2 * w + 1 < nb) bits |= (uint32_t)r.p[2 + 2 * w] << 8;
      vals = lo * 0x55555555u + spread16(bits) * step;
      break;
    }
    case 4: case 6: case 7: vals = (r.type & 3) * 0x55555555u; break;
    case 5: vals = 0; break;
    default: *err = kErrType; return 0;
  }
  uint32_t out = bed_word(vals, inv);
  const uint32_t rem = n - s0;
  if (rem < 16) out &= (1u << (2 * rem)) - 1u;
  return out;
}

// offset of the difflist inside a record of this type (0xFFFFFFFF: the type has none)
RGP_HD uint32_t difflist_pos(uint32_t type, uint32_t n) {
  if (type == 1) return 1 + (n + 7) / 8;
  if (type == 0 || type == 5) return 0xFFFFFFFFu;
  return 0;
}

RGP_HD void put(uint32_t* row, uint32_t id, uint32_t code) {
  const uint32_t sh = 2 * (id & 15);
#if defined(__CUDA_ARCH__)
  atomicAnd(row + (id >> 4), ~(3u << sh));           // other lanes own other samples of the same word
  atomicOr(row + (id >> 4), code << sh);
#else
  row[id >> 4] = (row[id >> 4] & ~(3u << sh)) | (code << sh);
#endif
}

// vint at r.p[*q], bounded by the record; returns 0 and sets err past the end
RGP_HD uint32_t vint(const Rec& r, uint32_t* q, int* err) {
  uint32_t v = 0;
  for (uint32_t shift = 0; shift < 35; shift += 7) {
    if (*q >= r.len) { *err = kErrTruncated; return 0; }
    const uint32_t b = r.p[(*q)++];
    v |= (b & 0x7Fu) << shift;
    if (!(b & 0x80u)) return v;
  }
  *err = kErrTruncated;
  return 0;
}

// This lane's share (groups lane, lane + nlanes, ...) of the difflist at r.p[pos]: writes the entries into the row.
// All lanes of a row must have finished an earlier list before any lane starts the next one (the caller's barrier).
RGP_HD void patch_difflist(const Rec& r, uint32_t pos, uint32_t n, bool inv, uint32_t* row, uint32_t lane, uint32_t nlanes,
                           int* err) {
  uint32_t q = pos;
  const uint32_t L = vint(r, &q, err);
  if (*err || L == 0) return;
  if (L > n) { *err = kErrListLen; return; }
  const uint32_t ng = (L + kGroup - 1) / kGroup;
  const uint32_t sb = n <= 0xFFu ? 1 : n <= 0xFFFFu ? 2 : n <= 0xFFFFFFu ? 3 : 4;
  const uint32_t first = q, sizes = first + ng * sb, vals = sizes + (ng - 1), deltas = vals + (L + 3) / 4;
  if (deltas > r.len || deltas < q) { *err = kErrTruncated; return; }
  uint32_t run = 0;                         // byte offset of group gb's delta stream (the same on every lane)
  for (uint32_t gb = 0; gb < ng; gb += nlanes) {
    const uint32_t g = gb + lane;
    uint32_t off = run;
    for (uint32_t h = gb; h < gb + nlanes && h + 1 < ng; ++h) {
      const uint32_t sz = (uint32_t)r.p[sizes + h] + (kGroup - 1);
      if (h < g) off += sz;
      run += sz;
    }
    if (g >= ng) continue;
    uint32_t id = 0;
    for (uint32_t k = 0; k < sb; ++k) id |= (uint32_t)r.p[first + g * sb + k] << (8 * k);
    const uint32_t cnt = (L - g * kGroup < kGroup) ? L - g * kGroup : kGroup;
    uint32_t p = deltas + off;
    for (uint32_t k = 0; k < cnt; ++k) {
      if (k) id += vint(r, &p, err);
      if (*err) return;
      if (id >= n) { *err = kErrSampleIdx; return; }
      const uint32_t e = g * kGroup + k;
      put(row, id, bed_code((r.p[vals + (e >> 2)] >> (2 * (e & 3))) & 3u, inv));
    }
  }
}

// One row, the way a warp of the device kernel produces it (fill, then base list, then own list), with the lanes run one
// after the other: used by the device kernel's host twin (rgb200_hostprobe `pgen-rows`, the tests' mock ABI).
// own / base: the variant's record and the record its dense part comes from (NULL unless own is of type 2 / 3).
#if !defined(__CUDA_ARCH__)
inline int decode_row_serial(const Rec& own, const Rec* base, uint32_t n, uint32_t* row, uint32_t words, uint32_t nlanes) {
  int err = kOk;
  const bool ld = (own.type & 6) == 2, inv = own.type == 3;
  if (ld && !base) return kErrType;
  const Rec& dense = ld ? *base : own;
  if ((dense.type & 6) == 2) return kErrType;
  for (uint32_t w = 0; w < words; ++w) row[w] = fill_word(dense, n, w, inv, &err);
  if (err) return err;
  if (ld && difflist_pos(dense.type, n) != 0xFFFFFFFFu)
    for (uint32_t l = 0; l < nlanes; ++l) {
      int e = kOk;
      patch_difflist(dense, difflist_pos(dense.type, n), n, inv, row, l, nlanes, &e);
      if (e) err = e;
    }
  if (err) return err;
  if (difflist_pos(own.type, n) != 0xFFFFFFFFu)
    for (uint32_t l = 0; l < nlanes; ++l) {
      int e = kOk;
      patch_difflist(own, difflist_pos(own.type, n), n, inv, row, l, nlanes, &e);
      if (e) err = e;
    }
  return err;
}
#endif

}  // namespace rgp
