// rng.cuh — device side of the engine's RNG contract (see include/mambacuda.h "RNG contract").
//
// The reference draws from Julia's global RNG (`rand()`, `randn()`, e.g. src/samplers/amwg.jl:102,107);
// the engine replaces it by a counter-based stream so that thousands of chains advance in
// lockstep with no RNG state in memory and with results independent of the chain→GPU mapping:
//
//   Philox4x32-10, key = (seed_lo, seed_hi), counter = (k >> 1, iter, chain, block | kind << 16 | stream << 24)
//   Two independent streams per block update: stream 0 feeds rand(), stream 1 feeds randn(); k counts the
//   draws of a stream in the order the reference consumes them.  One Philox block
//   gives two draws:
//     uniform k : u53(w0, w1) for even k, u53(w2, w3) for odd k;  u53(hi, lo) = (hi * 2^21 + (lo >> 11)) * 2^-53 in [0,1)
//     normal  k : rad = sqrt(-2 log(1 - u53(w0,w1))), ang = 2 pi u53(w2,w3);  rad cos(ang) for even k, rad sin(ang) for odd k
//   (both Box-Muller branches are used, so 26 normals cost 13 Philox blocks, logs and square roots).
// EXTERNAL mode reads uniforms sequentially from a caller-supplied stream (the north_star "shim"
// stream); a normal consumes two entries and uses the cosine branch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fastfn.cuh"

namespace mcu {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const unsigned long long bits = ((unsigned long long)hi << 21) | (unsigned long long)(lo >> 11);
  return (double)bits * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double box_muller(double ua, double ub) {
  return sqrt(-2.0 * log(1.0 - ua)) * cos(6.283185307179586476925286766559 * ub);
}

__device__ __forceinline__ double box_muller_sin(double ua, double ub) {
  return sqrt(-2.0 * log(1.0 - ua)) * sin(6.283185307179586476925286766559 * ub);
}

struct Draws {
  uint32_t k0, k1, chain, iter, blockkind, ku, kn;
  uint32_t z_blk, u_blk;      // Philox block whose second draw is held in z_next / u_next (0xffffffff: none)
  double z_next, u_next;
  const double* ext;          // EXTERNAL mode: this chain's stream, or nullptr
  unsigned long long ext_n;
  unsigned long long* ext_pos;  // this chain's cursor (persists across block updates and launches)

  __device__ __forceinline__ void seek(uint32_t it, uint32_t block, uint32_t kind) {
    iter = it; blockkind = block | (kind << 16); ku = 0; kn = 0; z_blk = 0xffffffffu; u_blk = 0xffffffffu;
  }
  __device__ __forceinline__ double next_ext() {
    unsigned long long p = *ext_pos;
    if (p >= ext_n) { *ext_pos = ~0ull; return 0.5; }   // exhausted: the cursor becomes a sentinel that mcu_run turns into MCU_ERR_STATE
    *ext_pos = p + 1;
    return ext[p];
  }
  __device__ __noinline__ double uniform() {
    if (ext) return next_ext();
    if ((ku & 1u) && u_blk == (ku >> 1)) { ++ku; return u_next; }   // second half of the block drawn a call ago
    uint32_t w[4];
    philox4x32_10(ku >> 1, iter, chain, blockkind, k0, k1, w);
    u_next = u53(w[2], w[3]); u_blk = ku >> 1;
    const double u = (ku & 1u) ? u_next : u53(w[0], w[1]);
    ++ku;
    return u;
  }
  __device__ __noinline__ double normal() {
    if (ext) { const double a = next_ext(); const double b = next_ext(); return box_muller(a, b); }
    if ((kn & 1u) && z_blk == (kn >> 1)) { ++kn; return z_next; }   // second draw of the block evaluated a call ago
    uint32_t w[4];
    philox4x32_10(kn >> 1, iter, chain, blockkind | (1u << 24), k0, k1, w);
    // both Box-Muller branches from one log / sqrt / sin-cos evaluation (the same arithmetic as the fused kernels' draw_normal_pair)
    const double rad = sqrt(-2.0 * fast_log(1.0 - u53(w[0], w[1])));
    const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
    z_next = rad * sc.a; z_blk = kn >> 1;
    const double z = (kn & 1u) ? z_next : rad * sc.b;
    ++kn;
    return z;
  }
};

}  // namespace mcu
