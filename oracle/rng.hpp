// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
// executed from the product (libmambacuda.so / mambacuda python package).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// rng.hpp — the RNG contract of the engine, restated for the CPU oracle.
//
// The reference draws from Julia's global MersenneTwister (`rand()`, `randn()`; e.g.
// src/samplers/amwg.jl:102,107).  That stream cannot be reproduced; the contract instead fixes a
// counter-based stream that a Julia shim can also produce:
//
//   Philox4x32-10 (Salmon et al., SC'11 — Random123), key = (seed_lo, seed_hi),
//   counter = (k >> 1, iter, chain, block | kind << 16 | stream << 24)
//     Every block update of every chain owns two independent streams: stream 0 feeds rand(),
//     stream 1 feeds randn().  k is the index of the draw inside its stream, in the order the
//     reference consumes draws; a vector draw randn(n) takes n consecutive k.
//     One Philox block (4 words) yields TWO draws:
//       uniform k : u53(w0, w1) if k is even, u53(w2, w3) if k is odd,  u53(hi, lo) = (hi * 2^21 + (lo >> 11)) * 2^-53 in [0,1)
//       normal  k : rad = sqrt(-2 log(1 - u53(w0,w1))), ang = 2 pi u53(w2,w3): rad cos(ang) if k is even, rad sin(ang) if odd
//     iter  : model.iter (1-based, src/model/simulation.jl:94)
//     chain : GLOBAL chain id
//     block : sampler index (0-based); kind 0 = sampler draws, 1 = init jitter
//
// EXTERNAL mode: draws are read sequentially from a caller-supplied uniform stream; a normal
// consumes two entries (ua, ub) with the same Box-Muller map.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace orc {

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline double u53(uint32_t hi, uint32_t lo) {
  uint64_t bits = ((uint64_t)hi << 21) | (uint64_t)(lo >> 11);
  return (double)bits * (1.0 / 9007199254740992.0);
}

inline double box_muller(double ua, double ub) {
  const double TWO_PI = 6.283185307179586476925286766559;
  return std::sqrt(-2.0 * std::log(1.0 - ua)) * std::cos(TWO_PI * ub);
}
inline double box_muller_sin(double ua, double ub) {
  const double TWO_PI = 6.283185307179586476925286766559;
  return std::sqrt(-2.0 * std::log(1.0 - ua)) * std::sin(TWO_PI * ub);
}

struct Rng {
  virtual ~Rng() {}
  virtual double uniform() = 0;
  virtual double normal() = 0;
  virtual void seek(uint32_t /*iter*/, uint32_t /*block*/, uint32_t /*kind*/) {}
};

struct PhiloxRng : Rng {
  uint32_t key[2];
  uint32_t chain, iter = 0, blockkind = 0, ku = 0, kn = 0;
  PhiloxRng(uint64_t seed, uint32_t chain_) : chain(chain_) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
  }
  void seek(uint32_t it, uint32_t block, uint32_t kind) override {
    iter = it; blockkind = block | (kind << 16); ku = 0; kn = 0;
  }
  void words(uint32_t k, uint32_t stream, uint32_t w[4]) const {
    uint32_t ctr[4] = {k >> 1, iter, chain, blockkind | (stream << 24)};
    philox4x32_10(ctr, key, w);
  }
  double uniform() override {
    uint32_t w[4]; words(ku, 0, w);
    const double u = (ku & 1u) ? u53(w[2], w[3]) : u53(w[0], w[1]);
    ++ku;
    return u;
  }
  double normal() override {
    uint32_t w[4]; words(kn, 1, w);
    const double ua = u53(w[0], w[1]), ub = u53(w[2], w[3]);
    const double z = (kn & 1u) ? box_muller_sin(ua, ub) : box_muller(ua, ub);
    ++kn;
    return z;
  }
};

struct ExternalRng : Rng {
  const double* u; size_t n, pos = 0;
  bool exhausted = false;
  ExternalRng(const double* u_, size_t n_) : u(u_), n(n_) {}
  double next() { if (pos >= n) { exhausted = true; return 0.5; } return u[pos++]; }
  double uniform() override { return next(); }
  double normal() override { double a = next(); double b = next(); return box_muller(a, b); }
};

}  // namespace orc
