// samplers.cuh — the reference's samplers as device functions, one chain per thread.
//
// Each function is the stand-alone `sample!(v::XVariate, logf)` of the reference with the block
// vector `v` in thread-local memory, the tune state behind a strided reference into the
// chain-fastest tune array, and every random draw taken from `Draws` in exactly the order the
// reference consumes them:
//   AMWG   src/samplers/amwg.jl:68-115      Slice  src/samplers/slice.jl:66-117
//   RWM    src/samplers/rwm.jl:65-71        NUTS   src/samplers/nuts.jl:63-205
//   HMC    src/samplers/hmc.jl:72-111       AMM    src/samplers/amm.jl:66-108
// `T` is a block target: T::logf(x) is logpdf!(block, x) (src/samplers/sampler.jl:102-104) and
// T::logfgrad(x, g) is logpdfgrad!(block, x, dtype) (src/samplers/sampler.jl:106-111).
//
// NUTS: the reference's recursive buildtree (nuts.jl:139-180) is unrolled into an iterative
// leaf-by-leaf walk with one pending "first half" per tree level (binary-counter carries); the
// uniforms of the merge steps are drawn in the recursion's post-order, so a chain follows the
// same trajectory as the recursive form on the same stream.  The reference has no depth cap
// (nuts.jl:106-124); the engine stops doubling after `max_depth` doublings (documented deviation).
#pragma once
#include "models.cuh"
#include "rng.cuh"

namespace mcu {

constexpr int kMaxDepth = 10;
constexpr int kAmmMaxK = 8;

struct TuneRef {
  double* base; size_t stride;
  MCU_D double& operator[](int i) const { return base[(size_t)i * stride]; }
};

struct DevBlock {
  int kind, k, transform, adapt, batchsize, proposal, L, grad, max_depth, n_own;
  uint32_t mask;                       // parameter nodes in the block
  int own[8];                          // node ids, in the order given to the sampler
  int tune_off;                        // offset of this block's tune slots inside a chain's tune record
  double target, epsilon, beta, amm_scale;
  const int* elem;                     // [k] state index of each block element
  const int* elink;                    // [k] link code of each block element
  const double* ebound;                // [2k] (lo, hi) of each block element with a LINK_BOUNDED link, or nullptr when the block has none
  const double* scale;                 // [k] sigma / width / scale, expanded
  const double* SigmaL;                // [k*k] column-major lower Cholesky factor (HMC with Sigma, AMM) or nullptr
};

// -------------------------------------------------------------------------------- AMWG
// tune slots: 0 m, 1 adapt, 2.. sigma[k], 2+k.. accept[k]
template <int K, class T>
MCU_NOINL void amwg_sample(double* v, const DevBlock& b, TuneRef tn, T& tgt, Draws& rng, bool fresh, bool adapt) {
  const int k = b.k;
  if (fresh) {   // AMWGTune(x, sigma): amwg.jl:14-21 via sampler.jl:40-45
    tn[0] = 0.0; tn[1] = 0.0;
    for (int i = 0; i < k; ++i) { tn[2 + i] = b.scale[i]; tn[2 + k + i] = 0.0; }
  }
  bool was = tn[1] != 0.0;
  if (adapt && !was) { for (int i = 0; i < k; ++i) tn[2 + k + i] = 0.0; tn[0] = 0.0; }   // setadapt!: amwg.jl:88-96
  tn[1] = adapt ? 1.0 : 0.0;
  double m = tn[0];
  if (adapt) { m += 1.0; tn[0] = m; }
  // amwg_sub!: amwg.jl:99-115
  double logf0 = tgt.logf(v);
  double z[K];
  for (int i = 0; i < k; ++i) z[i] = tn[2 + i] * rng.normal();
  for (int i = 0; i < k; ++i) {
    const double x = v[i];
    v[i] += z[i];
    const double logfprime = tgt.logf_comp(v, i, logf0);   // one component moved: local terms only where the template provides them
    if (rng.uniform() < exp(logfprime - logf0)) {
      logf0 = logfprime;
      if (adapt) tn[2 + k + i] += 1.0;
    } else {
      v[i] = x;
      tgt.put(i, x);
    }
  }
  if (adapt) {
    const long long mi = (long long)m;
    if (mi % b.batchsize == 0) {   // amwg.jl:74-80
      const double delta = fmin(0.01, pow(m / (double)b.batchsize, -0.5));
      for (int i = 0; i < k; ++i) {
        const double epsilon = tn[2 + k + i] / m < b.target ? -delta : delta;
        tn[2 + i] *= exp(epsilon);
      }
    }
  }
}

// -------------------------------------------------------------------------------- Slice
template <int K, class T>
MCU_NOINL void slice_uni_sample(double* v, const DevBlock& b, T& tgt, Draws& rng) {   // slice.jl:66-92
  const int k = b.k;
  double logf0 = tgt.logf(v);
  double lower[K], upper[K];
  for (int i = 0; i < k; ++i) lower[i] = v[i] - b.scale[i] * rng.uniform();
  for (int i = 0; i < k; ++i) upper[i] = lower[i] + b.scale[i];
  for (int i = 0; i < k; ++i) {
    const double p0 = logf0 + log(rng.uniform());
    const double x = v[i];
    // the coordinate's candidates are all evaluated against the vector the loop starts from (value logf0, component x)
    const double base_lp = logf0;
    const bool local = tgt.comp_is_local(i, base_lp);
    const double t_base = local ? tgt.terms_at(i) : 0.0;
    v[i] = lower[i] + (upper[i] - lower[i]) * rng.uniform();
    while (true) {
      logf0 = local ? tgt.logf_comp_base(v, i, base_lp, t_base) : tgt.logf(v);
      if (!(logf0 < p0)) break;
      const double value = v[i];
      if (value < x) lower[i] = value; else upper[i] = value;
      v[i] = lower[i] + (upper[i] - lower[i]) * rng.uniform();
    }
  }
}
template <int K, class T>
MCU_NOINL void slice_multi_sample(double* v, const DevBlock& b, T& tgt, Draws& rng) {   // slice.jl:95-117
  const int k = b.k;
  const double p0 = tgt.logf(v) + log(rng.uniform());
  double lower[K], upper[K], x[K];
  for (int i = 0; i < k; ++i) lower[i] = v[i] - b.scale[i] * rng.uniform();
  for (int i = 0; i < k; ++i) upper[i] = lower[i] + b.scale[i];
  for (int i = 0; i < k; ++i) x[i] = b.scale[i] * rng.uniform() + lower[i];
  while (tgt.logf(x) < p0) {
    for (int i = 0; i < k; ++i) {
      const double value = x[i];
      if (value < v[i]) lower[i] = value; else upper[i] = value;
      x[i] = lower[i] + (upper[i] - lower[i]) * rng.uniform();
    }
  }
  for (int i = 0; i < k; ++i) v[i] = x[i];
}

// -------------------------------------------------------------------------------- Gibbs
// Gamma(shape a, scale 1), Marsaglia & Tsang (2000); draw order: include/mambacuda.h, MCU_GIBBS
static MCU_NOINL double rgamma_mt(double a, Draws& rng) {
  double boost = 1.0;
  if (a < 1.0) { boost = pow(rng.uniform(), 1.0 / a); a += 1.0; }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x, v;
    do { x = rng.normal(); v = 1.0 + c * x; } while (v <= 0.0);
    v = v * v * v;
    const double u = rng.uniform();
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (flog(u) < 0.5 * x2 + d * (1.0 - v + flog(v))) return boost * d * v;
  }
}

// -------------------------------------------------------------------------------- RWM
// rand(proposal(0.0, 1.0)) for the symmetric kernels of src/distributions/extensions.jl:51-53.  The reference calls Distributions.rand;
// the engine's draw order per component (shared with the oracle): Normal one normal; SymUniform, Cosine one uniform; SymTriangularDist two
// uniforms; Epanechnikov three uniforms (the middle-of-three rule, Devroye 1986 p. 236); Biweight / Triweight two Gamma(3) / Gamma(4) draws
// (z = 2 B - 1 with B ~ Beta(3, 3) / Beta(4, 4), since (1 - z^2)^k = (4 B (1 - B))^k).
static MCU_NOINL double rwm_draw(int proposal, Draws& rng) {
  switch (proposal) {
    case 1: return -1.0 + 2.0 * rng.uniform();                                   // SymUniform(0,1) = Uniform(-1,1)
    case 2: { const double a = rng.uniform(); return a - rng.uniform(); }        // SymTriangularDist(0,1)
    case 3: {                                                                     // Cosine(0,1): F(z) = (1 + z + sin(pi z) / pi) / 2 on [-1, 1], inverted by bisection
      const double u = rng.uniform();
      double lo = -1.0, hi = 1.0;
      for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (0.5 * (1.0 + mid + sinpi(mid) * 0.31830988618379067154) < u) lo = mid; else hi = mid;
      }
      return 0.5 * (lo + hi);
    }
    case 4: {                                                                     // Epanechnikov(0,1)
      const double u1 = -1.0 + 2.0 * rng.uniform(), u2 = -1.0 + 2.0 * rng.uniform(), u3 = -1.0 + 2.0 * rng.uniform();
      return (fabs(u3) >= fabs(u2) && fabs(u3) >= fabs(u1)) ? u2 : u3;
    }
    case 5: case 6: {                                                             // Biweight(0,1) / Triweight(0,1)
      const double a = proposal == 5 ? 3.0 : 4.0;
      const double g1 = rgamma_mt(a, rng), g2 = rgamma_mt(a, rng);
      return 2.0 * (g1 / (g1 + g2)) - 1.0;
    }
    default: return rng.normal();
  }
}
template <int K, class T>
MCU_NOINL void rwm_sample(double* v, const DevBlock& b, T& tgt, Draws& rng) {   // rwm.jl:65-71
  const int k = b.k;
  double x[K];
  for (int i = 0; i < k; ++i) x[i] = v[i] + b.scale[i] * rwm_draw(b.proposal, rng);
  const double u = rng.uniform();
  const double lx = tgt.logf(x);
  const double lv = tgt.logf(v);
  if (u < exp(lx - lv)) for (int i = 0; i < k; ++i) v[i] = x[i];
}

// -------------------------------------------------------------------------------- NUTS
template <class T>
MCU_NOINL double leapfrog(double* x, double* r, double* g, int k, double eps, T& tgt) {   // nuts.jl:129-136 (in place)
  for (int i = 0; i < k; ++i) r[i] += (0.5 * eps) * g[i];
  for (int i = 0; i < k; ++i) x[i] += eps * r[i];
  const double logf = tgt.logfgrad(x, g);
  for (int i = 0; i < k; ++i) r[i] += (0.5 * eps) * g[i];
  return logf;
}
MCU_D double dot_self(const double* a, int k) { double s = 0; for (int i = 0; i < k; ++i) s += a[i] * a[i]; return s; }
MCU_D bool nouturn(const double* xminus, const double* xplus, const double* rminus, const double* rplus, int k) {   // nuts.jl:183-187
  double a = 0, c = 0;
  for (int i = 0; i < k; ++i) { const double d = xplus[i] - xminus[i]; a += d * rminus[i]; c += d * rplus[i]; }
  return a >= 0 && c >= 0;
}
MCU_D void copyv(double* dst, const double* src, int k) { for (int i = 0; i < k; ++i) dst[i] = src[i]; }

// tune slots: 0 adapt, 1 alpha, 2 epsilon, 3 epsilonbar, 4 Hbar, 5 m, 6 mu, 7 nalpha
template <int K, class T>
MCU_NOINL void nuts_sub(double* v, int k, TuneRef tn, double eps, T& tgt, Draws& rng, int max_depth) {   // nuts.jl:95-126
  double xm[K], rm[K], gm[K], xp[K], rp[K], gp[K];
  double cx[K], cr[K], cg[K];
  double Txf[K], Trf[K], Txp[K];
  double Sxf[kMaxDepth][K], Srf[kMaxDepth][K], Sxp[kMaxDepth][K];
  double Sn[kMaxDepth];

  for (int i = 0; i < k; ++i) { cr[i] = rng.normal(); cx[i] = v[i]; cg[i] = 0.0; }
  const double logf_init = leapfrog(cx, cr, cg, k, 0.0, tgt);
  const double logp0 = logf_init - 0.5 * dot_self(cr, k);
  const double logu0 = logp0 + log(rng.uniform());
  copyv(xm, cx, k); copyv(xp, cx, k); copyv(rm, cr, k); copyv(rp, cr, k); copyv(gm, cg, k); copyv(gp, cg, k);
  int j = 0; double n = 1.0; bool s = true;
  double alpha = 0.0, nalpha = 0.0;
  while (s) {
    const int pm = rng.uniform() > 0.5 ? 1 : -1;
    if (pm == -1) { copyv(cx, xm, k); copyv(cr, rm, k); copyv(cg, gm, k); }
    else { copyv(cx, xp, k); copyv(cr, rp, k); copyv(cg, gp, k); }
    // ---- buildtree(.., pm, j, ..) unrolled: nuts.jl:139-180
    const unsigned nleaf = 1u << j;
    double Tn = 0.0; bool Ts = true;
    alpha = 0.0; nalpha = 0.0;
    for (unsigned t = 0; t < nleaf; ++t) {
      const double logf = leapfrog(cx, cr, cg, k, pm * eps, tgt);
      const double logpp = logf - 0.5 * dot_self(cr, k);
      Tn = logu0 < logpp ? 1.0 : 0.0;
      Ts = logu0 < logpp + 1000.0;
      alpha += fmin(1.0, exp(logpp - logp0));
      nalpha += 1.0;
      copyv(Txf, cx, k); copyv(Trf, cr, k); copyv(Txp, cx, k);
      int l = 0;
      while (l < j) {
        if ((t >> l) & 1u) {   // this subtree is a second half: merge with the pending first half
          const double u = rng.uniform();
          const double nA = Sn[l];
          if (!(u < Tn / (nA + Tn))) copyv(Txp, Sxp[l], k);
          Tn = nA + Tn;
          const bool ok = pm == 1 ? nouturn(Sxf[l], cx, Srf[l], cr, k) : nouturn(cx, Sxf[l], cr, Srf[l], k);
          Ts = Ts && ok;
          copyv(Txf, Sxf[l], k); copyv(Trf, Srf[l], k);
          ++l;
        } else if (Ts) {       // a good first half: park it and build its sibling
          copyv(Sxf[l], Txf, k); copyv(Srf[l], Trf, k); copyv(Sxp[l], Txp, k); Sn[l] = Tn;
          break;
        } else {
          ++l;                 // a failed first half: the parent returns it unchanged
        }
      }
      if (l == j) break;
    }
    if (pm == -1) { copyv(xm, cx, k); copyv(rm, cr, k); copyv(gm, cg, k); }
    else { copyv(xp, cx, k); copyv(rp, cr, k); copyv(gp, cg, k); }
    if (Ts) { if (rng.uniform() < Tn / n) copyv(v, Txp, k); }
    j += 1;
    n += Tn;
    s = Ts && nouturn(xm, xp, rm, rp, k);
    if (j >= max_depth) s = false;
  }
  tn[1] = alpha; tn[7] = nalpha;
}

template <int K, class T>
MCU_NOINL double nutsepsilon(const double* x0, int k, T& tgt, Draws& rng) {   // nuts.jl:192-205
  double x[K], r0[K], g0[K], r[K], g[K];
  for (int i = 0; i < k; ++i) { r0[i] = rng.normal(); x[i] = x0[i]; g0[i] = 0.0; }
  const double logf0 = leapfrog(x, r0, g0, k, 0.0, tgt);
  const double d0 = dot_self(r0, k);
  double eps = 1.0;
  auto trial = [&](double e) {
    for (int i = 0; i < k; ++i) { x[i] = x0[i]; r[i] = r0[i]; g[i] = g0[i]; }
    const double lf = leapfrog(x, r, g, k, e, tgt);
    return exp(lf - logf0 - 0.5 * (dot_self(r, k) - d0));
  };
  double prob = trial(eps);
  const int pm = prob > 0.5 ? 1 : -1;
  int guard = 0;
  while (pow(prob, (double)pm) > pow(0.5, (double)pm)) {
    eps *= pm == 1 ? 2.0 : 0.5;
    prob = trial(eps);
    if (++guard > 2000) break;
  }
  return eps;
}

template <int K, class T>
MCU_NOINL void nuts_sample(double* v, const DevBlock& b, TuneRef tn, T& tgt, Draws& rng, bool fresh, bool adapt) {   // nuts.jl:63-92
  const int k = b.k;
  const int max_depth = b.max_depth > 0 ? (b.max_depth < kMaxDepth ? b.max_depth : kMaxDepth) : kMaxDepth;
  if (fresh) {   // NUTSTune(x, nutsepsilon(x, f)): nuts.jl:17-30
    tn[0] = 0.0; tn[1] = 0.0; tn[3] = 1.0; tn[4] = 0.0; tn[5] = 0.0; tn[6] = CUDART_NAN; tn[7] = 0.0;
    tn[2] = b.epsilon > 0.0 ? b.epsilon : nutsepsilon<K>(v, k, tgt, rng);
  }
  const bool was = tn[0] != 0.0;
  if (adapt && !was) { tn[5] = 0.0; tn[6] = log(10.0 * tn[2]); }   // setadapt!: nuts.jl:84-92
  tn[0] = adapt ? 1.0 : 0.0;
  if (adapt) {
    const double m = tn[5] + 1.0; tn[5] = m;
    nuts_sub<K>(v, k, tn, tn[2], tgt, rng, max_depth);
    double p = 1.0 / (m + 10.0);                                   // t0 = 10
    const double Hbar = (1.0 - p) * tn[4] + p * (b.target - tn[1] / tn[7]);
    tn[4] = Hbar;
    const double eps = exp(tn[6] - sqrt(m) * Hbar / 0.05);         // gamma = 0.05
    tn[2] = eps;
    p = pow(m, -0.75);                                             // kappa = 0.75
    tn[3] = exp(p * log(eps) + (1.0 - p) * log(tn[3]));
  } else {
    if (tn[5] > 0.0) tn[2] = tn[3];
    nuts_sub<K>(v, k, tn, tn[2], tgt, rng, max_depth);
  }
}

// -------------------------------------------------------------------------------- MALA
template <int K, class T>
MCU_NOINL void mala_sample(double* v, const DevBlock& b, T& tgt, Draws& rng) {   // mala.jl:67-86
  const int k = b.k;
  const double se = sqrt(b.epsilon);
  const double* SL = b.SigmaL;               // nullptr = identity; else column-major lower Cholesky factor of Sigma
  double g0[K], g1[K], y[K], m0[K], m1[K], t[K], w[K];
  auto Lmul = [&](const double* z, double* r) {    // r = L z, L = sqrt(epsilon) SigmaL
    if (!SL) { for (int i = 0; i < k; ++i) r[i] = se * z[i]; return; }
    for (int i = 0; i < k; ++i) { double s = 0; for (int c = 0; c <= i; ++c) s += SL[i + c * k] * z[c]; r[i] = se * s; }
  };
  auto Ltmul = [&](const double* z, double* r) {   // r = L' z
    if (!SL) { for (int i = 0; i < k; ++i) r[i] = se * z[i]; return; }
    for (int i = 0; i < k; ++i) { double s = 0; for (int c = i; c < k; ++c) s += SL[c + i * k] * z[c]; r[i] = se * s; }
  };
  auto half_sq_Linv = [&](const double* x) {       // |inv(L) x|^2 / 2
    double acc = 0;
    if (!SL) { for (int i = 0; i < k; ++i) { const double r = x[i] / se; acc += r * r; } return 0.5 * acc; }
    for (int i = 0; i < k; ++i) { double s = x[i]; for (int c = 0; c < i; ++c) s -= se * SL[i + c * k] * t[c]; t[i] = s / (se * SL[i + i * k]); acc += t[i] * t[i]; }
    return 0.5 * acc;
  };
  const double logf0 = tgt.logfgrad(v, g0);
  for (int i = 0; i < k; ++i) w[i] = rng.normal();
  Ltmul(g0, t); Lmul(t, m0);                       // M2 grad0 = 0.5 L L' grad0
  for (int i = 0; i < k; ++i) m0[i] *= 0.5;
  Lmul(w, t);
  for (int i = 0; i < k; ++i) y[i] = v[i] + m0[i] + t[i];
  const double logf1 = tgt.logfgrad(y, g1);
  Ltmul(g1, t); Lmul(t, m1);
  for (int i = 0; i < k; ++i) m1[i] *= 0.5;
  for (int i = 0; i < k; ++i) w[i] = v[i] - y[i] - m1[i];
  const double q0 = -half_sq_Linv(w);
  for (int i = 0; i < k; ++i) w[i] = y[i] - v[i] - m0[i];
  const double q1 = -half_sq_Linv(w);
  if (rng.uniform() < exp((logf1 - q1) - (logf0 - q0))) copyv(v, y, k);
}

// -------------------------------------------------------------------------------- HMC
template <int K, class T>
MCU_NOINL void hmc_sample(double* v, const DevBlock& b, T& tgt, Draws& rng) {   // hmc.jl:72-111
  const int k = b.k;
  double x1[K], g0[K], g1[K], z[K], p0[K], p1[K];
  copyv(x1, v, k);
  const double logf0 = tgt.logfgrad(x1, g0);
  double logf1 = logf0; copyv(g1, g0, k);
  for (int i = 0; i < k; ++i) z[i] = rng.normal();
  const double* SL = b.SigmaL;
  if (!SL) copyv(p0, z, k);
  else for (int i = 0; i < k; ++i) { double a = 0; for (int c = 0; c <= i; ++c) a += SL[i + c * k] * z[c]; p0[i] = a; }
  for (int i = 0; i < k; ++i) p1[i] = p0[i] + 0.5 * b.epsilon * g0[i];
  for (int l = 0; l < b.L; ++l) {
    for (int i = 0; i < k; ++i) x1[i] += b.epsilon * p1[i];
    logf1 = tgt.logfgrad(x1, g1);
    for (int i = 0; i < k; ++i) p1[i] += b.epsilon * g1[i];
  }
  for (int i = 0; i < k; ++i) p1[i] -= 0.5 * b.epsilon * g1[i];
  for (int i = 0; i < k; ++i) p1[i] *= -1.0;
  auto kinetic = [&](const double* p) {
    if (!SL) return 0.5 * dot_self(p, k);
    double w[K];
    for (int i = 0; i < k; ++i) { double a = p[i]; for (int c = 0; c < i; ++c) a -= SL[i + c * k] * w[c]; w[i] = a / SL[i + i * k]; }
    return 0.5 * dot_self(w, k);
  };
  const double Kp0 = kinetic(p0), Kp1 = kinetic(p1);
  if (rng.uniform() < exp((logf1 - Kp1) - (logf0 - Kp0))) copyv(v, x1, k);
}

// -------------------------------------------------------------------------------- AMM
// cholfact(Hermitian(Sigma), Val{true}) restated (LAPACK dpstrf semantics: complete pivoting on the
// largest remaining diagonal, tolerance n*eps*max(diag)); returns the rank and writes P*L.
static MCU_NOINL int pivoted_chol_PL(const double* A_in, int n, double* PL) {
  double A[kAmmMaxK * kAmmMaxK], L[kAmmMaxK * kAmmMaxK], dots[kAmmMaxK]; int piv[kAmmMaxK];
  double amax = 0.0;
  for (int i = 0; i < n * n; ++i) { A[i] = A_in[i]; L[i] = 0.0; }
  for (int i = 0; i < n; ++i) { piv[i] = i; dots[i] = 0.0; amax = fmax(amax, A[i + i * n]); }
  for (int i = 0; i < n * n; ++i) PL[i] = 0.0;
  if (!(amax > 0.0)) return 0;
  const double tol = (double)n * 2.220446049250313e-16 * amax;
  int rank = n;
  for (int j = 0; j < n; ++j) {
    int pvt = j; double best = -1.0;
    for (int i = j; i < n; ++i) { const double d = A[i + i * n] - dots[i]; if (d > best) { best = d; pvt = i; } }
    if (best <= tol || isnan(best)) { rank = j; break; }
    if (pvt != j) {
      for (int c = 0; c < n; ++c) { const double t = A[j + c * n]; A[j + c * n] = A[pvt + c * n]; A[pvt + c * n] = t; }
      for (int c = 0; c < n; ++c) { const double t = A[c + j * n]; A[c + j * n] = A[c + pvt * n]; A[c + pvt * n] = t; }
      for (int c = 0; c < j; ++c) { const double t = L[j + c * n]; L[j + c * n] = L[pvt + c * n]; L[pvt + c * n] = t; }
      { const double t = dots[j]; dots[j] = dots[pvt]; dots[pvt] = t; }
      { const int t = piv[j]; piv[j] = piv[pvt]; piv[pvt] = t; }
    }
    const double ajj = sqrt(best); L[j + j * n] = ajj;
    for (int i = j + 1; i < n; ++i) {
      double a = A[i + j * n];
      for (int c = 0; c < j; ++c) a -= L[i + c * n] * L[j + c * n];
      L[i + j * n] = a / ajj;
      dots[i] += L[i + j * n] * L[i + j * n];
    }
  }
  for (int i = 0; i < n; ++i) for (int c = 0; c < n; ++c) PL[piv[i] + c * n] = L[i + c * n];
  return rank;
}
// tune slots: 0 adapt, 1 m, 2.. Mv[k], then Mvv[k*k], then SigmaLm[k*k]
template <int K, class T>
MCU_NOINL void amm_sample(double* v, const DevBlock& b, TuneRef tn, T& tgt, Draws& rng, bool fresh, bool adapt) {   // amm.jl:66-108
  const int k = b.k;
  const int oMv = 2, oMvv = 2 + k, oSLm = 2 + k + k * k;
  if (fresh) { tn[0] = 0.0; tn[1] = 0.0; for (int i = 0; i < k + 2 * k * k; ++i) tn[2 + i] = 0.0; }
  const bool was = tn[0] != 0.0;
  if (adapt && !was) {   // setadapt!: amm.jl:97-108
    tn[1] = 0.0;
    for (int i = 0; i < k; ++i) tn[oMv + i] = v[i];
    for (int i = 0; i < k; ++i) for (int c = 0; c < k; ++c) tn[oMvv + i + c * k] = v[i] * v[c];
    for (int i = 0; i < k * k; ++i) tn[oSLm + i] = 0.0;
  }
  tn[0] = adapt ? 1.0 : 0.0;
  double m = tn[1];
  double z[kAmmMaxK], x[kAmmMaxK];
  for (int i = 0; i < k; ++i) z[i] = rng.normal();
  for (int i = 0; i < k; ++i) { double a = 0; for (int c = 0; c <= i; ++c) a += b.SigmaL[i + c * k] * z[c]; x[i] = a; }
  if (m > 2.0 * k) {
    double z2[kAmmMaxK];
    for (int i = 0; i < k; ++i) z2[i] = rng.normal();
    for (int i = 0; i < k; ++i) {
      double a = 0; for (int c = 0; c < k; ++c) a += tn[oSLm + i + c * k] * z2[c];
      x[i] = b.beta * x[i] + (1.0 - b.beta) * a;
    }
  }
  for (int i = 0; i < k; ++i) x[i] += v[i];
  const double u = rng.uniform();
  const double lx = tgt.logf(x);
  const double lv = tgt.logf(v);
  if (u < exp(lx - lv)) for (int i = 0; i < k; ++i) v[i] = x[i];
  if (adapt) {
    m += 1.0; tn[1] = m;
    const double p = m / (m + 1.0);
    double Sigma[kAmmMaxK * kAmmMaxK], PL[kAmmMaxK * kAmmMaxK], Mv[kAmmMaxK];
    for (int i = 0; i < k; ++i) { Mv[i] = p * tn[oMv + i] + (1.0 - p) * v[i]; tn[oMv + i] = Mv[i]; }
    const double c0 = b.amm_scale * b.amm_scale / (double)k / p;
    for (int i = 0; i < k; ++i) for (int c = 0; c < k; ++c) {
      const double mvv = p * tn[oMvv + i + c * k] + (1.0 - p) * v[i] * v[c];
      tn[oMvv + i + c * k] = mvv;
      Sigma[i + c * k] = c0 * (mvv - Mv[i] * Mv[c]);
    }
    if (pivoted_chol_PL(Sigma, k, PL) == k) for (int i = 0; i < k * k; ++i) tn[oSLm + i] = PL[i];
  }
}

}  // namespace mcu
