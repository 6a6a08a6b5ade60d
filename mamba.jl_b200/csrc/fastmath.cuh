// fastmath.cuh — the paired Philox draws of the RNG contract and the MH-test helpers shared by the fused one-chain-per-thread kernels
// (seeds_fast.cu, rats_fast.cu, pumps_fast.cu); the FP64 exp / log / sincos they use live in fastfn.cuh (also used by rng.cuh).
#pragma once
#include "launch.hpp"

#ifndef MCU_FASTMATH_ESTRIN
#define MCU_FASTMATH_ESTRIN 0
#endif

namespace mcu {
namespace {

// rng.cuh contract: stream 0 = uniforms, stream 1 = normals, TWO draws per Philox block: draw k of a stream comes from
// block k >> 1 (first/second half for uniforms, cosine/sine branch of Box-Muller for normals).
MCU_D Pair draw_uniform_pair(const RunArgs& a, uint32_t chain, uint32_t iter, uint32_t block, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, block, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
  return {u53(w[0], w[1]), u53(w[2], w[3])};
}
MCU_D Pair draw_normal_pair(const RunArgs& a, uint32_t chain, uint32_t iter, uint32_t block, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, block | (1u << 24), (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
  const double rad = sqrt(-2.0 * fast_log(1.0 - u53(w[0], w[1])));
  const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
  return {rad * sc.b, rad * sc.a};   // (rad cos, rad sin)
}

// rand() < exp(delta)  <=>  log(rand()) < delta: the log of the uniform does not depend on the proposal, so it is evaluated next to the
// draw, off the critical path of the update (same decision up to rounding of the comparison)
MCU_D double log_uniform(double u) { return u > 0.0 ? fast_log(u) : -CUDART_INF; }

// Two-stage form of the same test.  log u is needed to full precision only when it is within rounding distance of delta; everywhere
// else a bracket decides.  Stage 1 (next to the draw, off the critical path): la = ln2 * MUFU.LG2(float(u)), |la - log u| <=
// 3e-7 (1 + |la|) (float rounding of u 6e-8, MUFU.LG2 2^-22 absolute / relative, float multiply).  Stage 2 (at the test): outside
// la +- 4e-6 (1 + |la|) — a 12x margin — the float comparison IS the double comparison; inside (probability ~1e-5 per test) the
// FP64 log decides.  The decisions are those of log_uniform(u) < delta, bit for bit; ~10 instructions instead of ~45 per test.
static __device__ __noinline__ bool logu_less_exact(double u, double delta) { return log_uniform(u) < delta; }   // cold path
#ifndef MCU_LOGU_DOUBLE
#define MCU_LOGU_DOUBLE 0
#endif
#if MCU_LOGU_DOUBLE
struct LogU { double u, lo, hi; };
MCU_D LogU logu_bracket(double u) {
  const float la = __log2f((float)u) * 0.693147181f;
  const float band = fmaf(fabsf(la), 4e-6f, 4e-6f);
  return {u, (double)(la - band), (double)(la + band)};
}
MCU_D bool logu_less(const LogU& l, double delta) {
  if (delta > l.hi) return true;
  if (!(delta > l.lo)) return false;
  return logu_less_exact(l.u, delta);
}
#else
struct LogU { double u; float la; };
MCU_D LogU logu_bracket(double u) { return {u, __log2f((float)u) * 0.693147181f}; }
MCU_D bool logu_less(const LogU& l, double delta) {
  const float band = fmaf(fabsf(l.la), 4e-6f, 4e-6f);
  const float df = (float)delta;
  if (df > l.la + band) return true;
  if (!(df > l.la - band)) return false;
  return logu_less_exact(l.u, delta);
}
#endif

// AMWG tune update every `batchsize` adaptive iterations: amwg.jl:74-80
MCU_D double amwg_delta(double m, int batchsize) { return fmin(0.01, pow(m / (double)batchsize, -0.5)); }

}  // namespace
}  // namespace mcu
