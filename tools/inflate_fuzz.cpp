// Differential fuzz of csrc/inflate_core.h (the decoder the GPU runs) against zlib, host build:
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 tools/inflate_fuzz.cpp -o /tmp/inflate_fuzz -lz && /tmp/inflate_fuzz
// 20000 streams (random / low-entropy / periodic / probability-like payloads of up to 120 KB, zlib levels 0-9), each
// damaged by 0-3 bit flips, a random truncation and a random declared output length, inputs and outputs in exact-size
// heap blocks so that the sanitizers see any out-of-bounds access (an overrun on the device would be a fault).  Both
// variants of the decoder (direct, and with the 16 KB output ring) must accept exactly the streams zlib accepts at that
// output length, with identical bytes.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../regenie_b200/csrc/inflate_core.h"

int main() {
  std::mt19937_64 rng(7);
  static rgi::Tables t;
  size_t accepted = 0, rejected = 0, disagree = 0;
  for (int iter = 0; iter < 20000; ++iter) {
    const size_t n = 1 + rng() % (iter % 8 == 0 ? 120000 : 5000);
    std::vector<uint8_t> raw(n);
    const int kind = rng() % 4;
    for (size_t i = 0; i < n; ++i)
      raw[i] = kind == 0 ? (uint8_t)rng() : kind == 1 ? (uint8_t)(rng() % 3) : kind == 2 ? (uint8_t)((i / 7) & 255)
                                                                                       : (uint8_t)((rng() % 16 == 0) ? rng() : 255 * (i & 1));
    uLongf cl = compressBound(n);
    std::vector<uint8_t> c(cl);
    compress2(c.data(), &cl, raw.data(), n, (int)(rng() % 10));
    const int nflip = rng() % 4;
    size_t len = cl;
    if (rng() % 5 == 0) len = rng() % (cl + 1);
    uint8_t* in = (uint8_t*)malloc(len ? len : 1);
    memcpy(in, c.data(), len);
    for (int k = 0; k < nflip && len; ++k) in[rng() % len] ^= (uint8_t)(1u << (rng() % 8));
    size_t out_len = n;
    if (rng() % 7 == 0) out_len = rng() % (2 * n + 1);
    uint8_t* out = (uint8_t*)malloc(out_len ? out_len : 1);
    static std::vector<uint8_t> win(rgi::kWinBytes);
    const bool window = iter & 1;
    const int st = window ? rgi::inflate_zlib_window(in, (uint32_t)len, out, (uint32_t)out_len, t, win.data(), true)
                          : rgi::inflate_zlib(in, (uint32_t)len, out, (uint32_t)out_len, t, true);
    std::vector<uint8_t> z(out_len ? out_len : 1);
    uLongf dl = out_len;
    const int zr = uncompress(z.data(), &dl, in, len);
    const bool zlib_ok = zr == Z_OK && dl == out_len;
    if ((st == 0) != zlib_ok || (st == 0 && memcmp(out, z.data(), out_len) != 0)) {
      ++disagree;
      printf("iter %d: status %d, zlib rc %d (%lu of %zu bytes)\n", iter, st, zr, (unsigned long)dl, out_len);
    }
    (st == 0 ? accepted : rejected)++;
    free(in);
    free(out);
  }
  printf("accepted %zu rejected %zu disagreements with zlib %zu\n", accepted, rejected, disagree);
  return disagree ? 1 : 0;
}
