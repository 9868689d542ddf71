// ishara_b200 — counter-based dropout masks shared by the training kernels. A mask bit is a pure
// function of (seed, site, element index), so the forward and backward kernels agree without storing masks and a host
// can reproduce them (IsharaModel.dropout_masks). 16 random bits per element: keep iff bits >= p * 65536.
#pragma once
#include <cstdint>

namespace ishara {

__host__ __device__ inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t dropout_key(uint64_t seed, uint32_t site) {
  return mix64(seed ^ mix64((static_cast<uint64_t>(site) << 32) | 0x5bd1e995ull));
}
inline uint32_t dropout_thr16(float p) { return static_cast<uint32_t>(p * 65536.f); }

// attention probabilities [B, H, T, T]: element (row = (b*H + h)*T + i, j) lives in pair ((row * Tpair + j) >> 1),
// Tpair = T rounded up to even. A 32-bit finaliser (murmur3 fmix32) gives the pair's two 16-bit lanes: the T^2 masks per
// head are recomputed in three kernels, and the 64-bit mix costs ~3x the integer work.
__device__ __forceinline__ uint32_t attn_pair_hash(uint64_t key, uint64_t rowbase, int j) {
  const uint64_t idx = (rowbase + static_cast<uint64_t>(j)) >> 1;
  uint32_t x = static_cast<uint32_t>(idx) * 0x9E3779B1u + static_cast<uint32_t>(idx >> 32) * 0x85EBCA77u + static_cast<uint32_t>(key);
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x ^ static_cast<uint32_t>(key >> 32);
}
__device__ __forceinline__ float attn_keep(uint64_t key, uint64_t rowbase, int j, uint32_t thr16, float inv_keep) {
  const uint32_t u = (attn_pair_hash(key, rowbase, j) >> (16 * (j & 1))) & 0xFFFFu;
  return u >= thr16 ? inv_keep : 0.f;
}
// both elements of the pair (j even)
__device__ __forceinline__ void attn_keep2(uint64_t key, uint64_t rowbase, int j, uint32_t thr16, float inv_keep, float& k0, float& k1) {
  const uint32_t r = attn_pair_hash(key, rowbase, j);
  k0 = (r & 0xFFFFu) >= thr16 ? inv_keep : 0.f;
  k1 = (r >> 16) >= thr16 ? inv_keep : 0.f;
}

}  // namespace ishara
