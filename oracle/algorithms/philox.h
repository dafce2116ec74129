// TEST INFRASTRUCTURE ONLY.  Philox4x32-10 (Salmon et al., SC'11) + Lemire rejection sampling: the injected
// random stream shared by the oracle's MCTS / rollout restatements and the device kernels
// (open_spiel_b200/csrc/common.cuh philox4 / philox_uniform define the same function).
#ifndef B2S_ORACLE_PHILOX_H_
#define B2S_ORACLE_PHILOX_H_
#include <cstdint>

namespace oracle {

inline void Philox4(uint64_t key, uint64_t lane, uint32_t ply, uint32_t stream, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)lane, (uint32_t)(lane >> 32), ply, stream};
  uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// Unbiased uniform integer in [0, n).
inline uint32_t PhiloxUniform(uint64_t key, uint64_t lane, uint32_t ply, uint32_t n) {
  uint32_t thresh = (uint32_t)(0u - n) % n;
  for (uint32_t stream = 0;; ++stream) {
    uint32_t r[4];
    Philox4(key, lane, ply, stream, r);
    for (int j = 0; j < 4; ++j) {
      uint64_t m = (uint64_t)r[j] * n;
      if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    }
  }
}

// MCTS random decisions: (key, a, b, c) -> [0, n);  a = simulation / expansion index, b = position, c = domain.
inline uint32_t RngUniform(uint64_t key, uint32_t a, uint32_t b, uint32_t c, uint32_t n) {
  return PhiloxUniform(key, (uint64_t)a | ((uint64_t)b << 32), c, n);
}

// Playout draws of one simulation share Philox blocks: draw number b uses word (b & 3) of the block keyed by
// (a, b >> 2, c) — four consecutive plies, one block.  A rejected word (probability < n / 2^32) falls back to streams
// 4 s + (b & 3), s = 1, 2, ..., of the same key, all four words in order.
inline uint32_t RngUniformShared(uint64_t key, uint32_t a, uint32_t b, uint32_t c, uint32_t n) {
  const uint64_t lane = (uint64_t)a | ((uint64_t)(b >> 2) << 32);
  const uint32_t thresh = (uint32_t)(0u - n) % n;
  uint32_t r[4];
  Philox4(key, lane, c, 0, r);
  uint64_t m = (uint64_t)r[b & 3] * n;
  if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
  for (uint32_t s = 1;; ++s) {
    Philox4(key, lane, c, 4 * s + (b & 3), r);
    for (int j = 0; j < 4; ++j) {
      m = (uint64_t)r[j] * n;
      if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    }
  }
}

}  // namespace oracle
#endif
