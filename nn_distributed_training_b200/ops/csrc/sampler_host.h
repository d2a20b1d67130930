// Host twin of sampler.cuh (same keyed Feistel permutation) for the native batch loader.
#pragma once
#include <stdint.h>

namespace nndt {
namespace host {

inline uint32_t mix_key(uint32_t seed, uint32_t node, uint32_t epoch) {
  uint32_t x = seed * 0x9E3779B1u + node * 0x85EBCA77u + epoch * 0xC2B2AE3Du + 0x27D4EB2Fu;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
inline uint32_t feistel_round(uint32_t x, uint32_t k, uint32_t mask) {
  x = (x ^ k) * 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13;
  return x & mask;
}
inline uint32_t feistel_permute(uint32_t pos, uint32_t m, uint32_t key) {
  if (m <= 1) return 0;
  int bits = 32 - __builtin_clz(m - 1);
  if (bits < 2) bits = 2;
  const int h = (bits + 1) >> 1;
  const uint32_t mask = (1u << h) - 1u;
  static const uint32_t rk[4] = {0xA511E9B3u, 0x63D83595u, 0x1B873593u, 0xCC9E2D51u};
  uint32_t x = pos;
  do {
    uint32_t l = x >> h, r = x & mask;
    for (int i = 0; i < 4; ++i) {
      const uint32_t t = l ^ feistel_round(r, key + rk[i], mask);
      l = r; r = t;
    }
    x = (l << h) | r;
  } while (x >= m);
  return x;
}
struct BatchLoc { uint32_t epoch, start, size; };
inline BatchLoc locate_batch(uint32_t call, uint32_t m, uint32_t B) {
  const uint32_t bpe = (m + B - 1) / B;
  BatchLoc o;
  o.epoch = call / bpe; o.start = (call % bpe) * B;
  o.size = (B < m - o.start) ? B : m - o.start;
  return o;
}

}  // namespace host
}  // namespace nndt
