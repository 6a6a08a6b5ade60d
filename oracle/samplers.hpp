// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// samplers.hpp — the stand-alone `sample!(v::XVariate, logf)` faces of the reference samplers
// (SURVEY.md App. A), with the random draws taken from an injectable Rng in EXACTLY the order
// the reference consumes them.
//   AMWG        : src/samplers/amwg.jl:68-115
//   Slice (uni) : src/samplers/slice.jl:66-92      Slice (multi): src/samplers/slice.jl:95-117
//   RWM         : src/samplers/rwm.jl:65-71
//   NUTS        : src/samplers/nuts.jl:63-205
//   HMC         : src/samplers/hmc.jl:72-111
//   AMM         : src/samplers/amm.jl:66-108
#pragma once
#include <cmath>
#include <functional>
#include <vector>

#include "model.hpp"
#include "rng.hpp"

namespace orc {

typedef std::vector<double> Vec;
typedef std::function<double(const Vec&)> LogF;
typedef std::function<double(const Vec&, Vec&)> LogFGrad;   // returns logf, fills grad

// Decision audit (tests/helpers.py::audit_divergence): every comparison that steers a sampler notes how close it came to its
// threshold, relative to max(1, |threshold|); the engine keeps the minimum per chain and iteration.  A device chain may part from
// the oracle's ONLY in an iteration where this minimum is at rounding level (the two sides evaluate the same density in a
// different floating-point order) — anything else is a real difference in draw order or arithmetic.
inline thread_local double* g_margin_slot = nullptr;
inline void note_margin(double lhs, double rhs, double scale = 1.0) {
  if (!g_margin_slot || !std::isfinite(lhs) || !std::isfinite(rhs)) return;
  const double m = std::fabs(lhs - rhs) / std::fmax(scale, std::fabs(rhs));
  if (m < *g_margin_slot) *g_margin_slot = m;
}
// rand() < exp(delta), the Metropolis-Hastings test of every sampler file (e.g. amwg.jl:107); noted on the log scale
inline bool mh_test(double u, double delta) {
  note_margin(std::log(u), delta);
  return u < std::exp(delta);
}

inline double dotv(const Vec& a, const Vec& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
inline double dotv(const Vec& a) { return dotv(a, a); }   // utils.jl:62

// ---------------------------------------------------------------------------- AMWG
inline void amwg_setadapt(Tune& t, bool adapt) {           // amwg.jl:88-96
  if (adapt && !t.adapt) { std::fill(t.accept.begin(), t.accept.end(), 0L); t.m = 0; }
  t.adapt = adapt;
}
inline void amwg_sub(Vec& v, Tune& t, const LogF& logf, Rng& rng) {   // amwg.jl:99-115
  double logf0 = logf(v);
  size_t n = v.size();
  Vec z(n);
  for (size_t i = 0; i < n; ++i) z[i] = t.sigma[i] * rng.normal();    // sigma .* randn(n)
  for (size_t i = 0; i < n; ++i) {
    double x = v[i];
    v[i] += z[i];
    double logfprime = logf(v);
    if (mh_test(rng.uniform(), logfprime - logf0)) {
      logf0 = logfprime;
      t.accept[i] += t.adapt ? 1 : 0;
    } else {
      v[i] = x;
    }
  }
}
inline void amwg_sample(Vec& v, Tune& t, const LogF& logf, bool adapt, Rng& rng) {   // amwg.jl:68-85
  amwg_setadapt(t, adapt);
  if (t.adapt) {
    t.m += 1;
    amwg_sub(v, t, logf, rng);
    if (t.m % t.batchsize == 0) {
      double delta = std::fmin(0.01, std::pow((double)t.m / (double)t.batchsize, -0.5));
      for (size_t i = 0; i < t.sigma.size(); ++i) {
        double epsilon = (double)t.accept[i] / (double)t.m < t.target ? -delta : delta;
        t.sigma[i] *= std::exp(epsilon);
      }
    }
  } else {
    amwg_sub(v, t, logf, rng);
  }
}

// ---------------------------------------------------------------------------- Slice
// rand(Uniform(a, b)) = a + (b - a) * rand()   (Distributions.jl)
inline double runif(double a, double b, Rng& rng) { return a + (b - a) * rng.uniform(); }

// Gamma(shape a, scale 1) by Marsaglia & Tsang (2000) — stands in for Distributions.jl's rand(Gamma(...)) inside a user-defined
// Gibbs sampler; draw order (the engine's RNG contract): for a < 1 one uniform first (boost U^(1/a)), then per attempt normals
// until 1 + c x > 0, then one uniform
inline double rgamma_mt(double a, Rng& rng) {
  double boost = 1.0;
  if (a < 1.0) { boost = std::pow(rng.uniform(), 1.0 / a); a += 1.0; }
  const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
  for (;;) {
    double x, v;
    do { x = rng.normal(); v = 1.0 + c * x; note_margin(v, 0.0); } while (v <= 0.0);
    v = v * v * v;
    const double u = rng.uniform();
    const double x2 = x * x;
    note_margin(u, 1.0 - 0.0331 * x2 * x2);
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    note_margin(std::log(u), 0.5 * x2 + d * (1.0 - v + std::log(v)));
    if (std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return boost * d * v;
  }
}

inline void slice_uni_sample(Vec& v, const Vec& width, const LogF& logf, Rng& rng) {   // slice.jl:66-92
  double logf0 = logf(v);
  size_t n = v.size();
  Vec lower(n), upper(n);
  for (size_t i = 0; i < n; ++i) lower[i] = v[i] - width[i] * rng.uniform();
  for (size_t i = 0; i < n; ++i) upper[i] = lower[i] + width[i];
  for (size_t i = 0; i < n; ++i) {
    double p0 = logf0 + std::log(rng.uniform());
    double x = v[i];
    v[i] = runif(lower[i], upper[i], rng);
    while (true) {
      logf0 = logf(v);
      note_margin(logf0, p0);
      if (!(logf0 < p0)) break;
      double value = v[i];
      if (value < x) lower[i] = value; else upper[i] = value;
      v[i] = runif(lower[i], upper[i], rng);
    }
  }
}
inline void slice_multi_sample(Vec& v, const Vec& width, const LogF& logf, Rng& rng) {  // slice.jl:95-117
  double p0 = logf(v) + std::log(rng.uniform());
  size_t n = v.size();
  Vec lower(n), upper(n), x(n);
  for (size_t i = 0; i < n; ++i) lower[i] = v[i] - width[i] * rng.uniform();
  for (size_t i = 0; i < n; ++i) upper[i] = lower[i] + width[i];
  for (size_t i = 0; i < n; ++i) x[i] = width[i] * rng.uniform() + lower[i];
  while (true) {
    const double lx = logf(x);
    note_margin(lx, p0);
    if (!(lx < p0)) break;
    for (size_t i = 0; i < n; ++i) {
      double value = x[i];
      if (value < v[i]) lower[i] = value; else upper[i] = value;
      x[i] = runif(lower[i], upper[i], rng);
    }
  }
  v = x;
}

// ---------------------------------------------------------------------------- RWM
// proposal(0,1) draws: Normal → randn; SymUniform(0,1) = Uniform(-1,1) (extensions.jl:43-46);
// SymTriangularDist(0,1): rand = mu + sigma * (rand() - rand())  (Distributions.jl)
// Cosine / Epanechnikov / Biweight / Triweight (extensions.jl:51-53): Distributions.jl's samplers are not in the tree; the engine contract
// (mamba.jl_b200/csrc/samplers.cuh) is one uniform + CDF inversion (Cosine), the middle-of-three rule (Epanechnikov), 2 Beta(3,3) - 1 and
// 2 Beta(4,4) - 1 through two Gamma draws (Biweight, Triweight).
inline double rwm_draw(int proposal, Rng& rng) {
  switch (proposal) {
    case 1: return runif(-1.0, 1.0, rng);
    case 2: { double a = rng.uniform(); double b = rng.uniform(); return a - b; }
    case 3: {
      const double u = rng.uniform();
      double lo = -1.0, hi = 1.0;
      for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (0.5 * (1.0 + mid + std::sin(M_PI * mid) * 0.31830988618379067154) < u) lo = mid; else hi = mid;
      }
      return 0.5 * (lo + hi);
    }
    case 4: {
      const double u1 = -1.0 + 2.0 * rng.uniform(), u2 = -1.0 + 2.0 * rng.uniform(), u3 = -1.0 + 2.0 * rng.uniform();
      return (std::fabs(u3) >= std::fabs(u2) && std::fabs(u3) >= std::fabs(u1)) ? u2 : u3;
    }
    case 5: case 6: {
      const double a = proposal == 5 ? 3.0 : 4.0;
      const double g1 = rgamma_mt(a, rng), g2 = rgamma_mt(a, rng);
      return 2.0 * (g1 / (g1 + g2)) - 1.0;
    }
    default: return rng.normal();
  }
}
inline void rwm_sample(Vec& v, const Vec& scale, int proposal, const LogF& logf, Rng& rng) {  // rwm.jl:65-71
  size_t n = v.size();
  Vec x(n);
  for (size_t i = 0; i < n; ++i) x[i] = v[i] + scale[i] * rwm_draw(proposal, rng);
  double u = rng.uniform();
  double lx = logf(x);          // logf(x) is evaluated first, then logf(v): the model is left at v
  double lv = logf(v);
  if (mh_test(u, lx - lv)) v = x;
}

// ---------------------------------------------------------------------------- NUTS
struct Leap { Vec x, r, grad; double logf; };
inline Leap leapfrog(const Vec& x, const Vec& r, const Vec& grad, double epsilon, const LogFGrad& f) {  // nuts.jl:129-136
  Leap o; size_t n = x.size();
  o.r.resize(n); o.x.resize(n);
  for (size_t i = 0; i < n; ++i) o.r[i] = r[i] + (0.5 * epsilon) * grad[i];
  for (size_t i = 0; i < n; ++i) o.x[i] = x[i] + epsilon * o.r[i];
  o.logf = f(o.x, o.grad);
  for (size_t i = 0; i < n; ++i) o.r[i] += (0.5 * epsilon) * o.grad[i];
  return o;
}
inline bool nouturn(const Vec& xminus, const Vec& xplus, const Vec& rminus, const Vec& rplus) {  // nuts.jl:183-187
  size_t n = xminus.size(); double a = 0, b = 0, sa = 0, sb = 0;
  for (size_t i = 0; i < n; ++i) { double d = xplus[i] - xminus[i]; a += d * rminus[i]; b += d * rplus[i]; sa += std::fabs(d * rminus[i]); sb += std::fabs(d * rplus[i]); }
  if (sa > 0) note_margin(a, 0.0, sa);
  if (sb > 0) note_margin(b, 0.0, sb);
  return a >= 0 && b >= 0;
}
struct Tree {
  Vec xminus, rminus, gradminus, xplus, rplus, gradplus, xprime;
  long nprime; bool sprime; double alphaprime; long nalphaprime;
};
inline Tree buildtree(const Vec& x, const Vec& r, const Vec& grad, int pm, int j, double epsilon,
                      const LogFGrad& f, double logp0, double logu0, Rng& rng) {   // nuts.jl:139-180
  Tree t;
  if (j == 0) {
    Leap l = leapfrog(x, r, grad, pm * epsilon, f);
    double logpprime = l.logf - 0.5 * dotv(l.r);
    note_margin(logu0, logpprime);
    t.nprime = logu0 < logpprime ? 1 : 0;
    t.sprime = logu0 < logpprime + 1000.0;
    t.xminus = t.xplus = l.x; t.rminus = t.rplus = l.r; t.gradminus = t.gradplus = l.grad;
    t.xprime = l.x;
    t.alphaprime = std::fmin(1.0, std::exp(logpprime - logp0));
    t.nalphaprime = 1;
  } else {
    t = buildtree(x, r, grad, pm, j - 1, epsilon, f, logp0, logu0, rng);
    if (t.sprime) {
      Tree t2;
      if (pm == -1) {
        t2 = buildtree(t.xminus, t.rminus, t.gradminus, pm, j - 1, epsilon, f, logp0, logu0, rng);
        t.xminus = t2.xminus; t.rminus = t2.rminus; t.gradminus = t2.gradminus;
      } else {
        t2 = buildtree(t.xplus, t.rplus, t.gradplus, pm, j - 1, epsilon, f, logp0, logu0, rng);
        t.xplus = t2.xplus; t.rplus = t2.rplus; t.gradplus = t2.gradplus;
      }
      const double um = rng.uniform(), ratio = (double)t2.nprime / (double)(t.nprime + t2.nprime);
      note_margin(um, ratio);
      if (um < ratio) t.xprime = t2.xprime;
      t.nprime += t2.nprime;
      t.sprime = t2.sprime && nouturn(t.xminus, t.xplus, t.rminus, t.rplus);
      t.alphaprime += t2.alphaprime;
      t.nalphaprime += t2.nalphaprime;
    }
  }
  return t;
}
// max_depth == 0 reproduces the reference (no cap, nuts.jl:106-124); the engine caps the number
// of doublings (documented deviation, SURVEY.md §7 hard part 4), so the oracle can too.
inline void nuts_sub(Vec& v, Tune& tune, double epsilon, const LogFGrad& f, Rng& rng, int max_depth) {  // nuts.jl:95-126
  size_t n = v.size();
  Vec r0(n), zero(n, 0.0);
  for (size_t i = 0; i < n; ++i) r0[i] = rng.normal();
  Leap l = leapfrog(v, r0, zero, 0.0, f);
  double logp0 = l.logf - 0.5 * dotv(l.r);
  double logu0 = logp0 + std::log(rng.uniform());
  Vec xminus = l.x, xplus = l.x, rminus = l.r, rplus = l.r, gradminus = l.grad, gradplus = l.grad;
  int j = 0; long nn = 1; bool s = true;
  while (s) {
    const double ud = rng.uniform();
    note_margin(ud, 0.5);
    int pm = 2 * (ud > 0.5 ? 1 : 0) - 1;
    Tree t;
    if (pm == -1) {
      t = buildtree(xminus, rminus, gradminus, pm, j, epsilon, f, logp0, logu0, rng);
      xminus = t.xminus; rminus = t.rminus; gradminus = t.gradminus;
    } else {
      t = buildtree(xplus, rplus, gradplus, pm, j, epsilon, f, logp0, logu0, rng);
      xplus = t.xplus; rplus = t.rplus; gradplus = t.gradplus;
    }
    if (t.sprime) {
      const double um = rng.uniform(), ratio = (double)t.nprime / (double)nn;
      note_margin(um, ratio);
      if (um < ratio) v = t.xprime;
    }
    j += 1;
    nn += t.nprime;
    s = t.sprime && nouturn(xminus, xplus, rminus, rplus);
    tune.alpha = t.alphaprime; tune.nalpha = t.nalphaprime;
    if (max_depth > 0 && j >= max_depth) s = false;
  }
}
inline double nutsepsilon(const Vec& x, const LogFGrad& f, Rng& rng) {   // nuts.jl:192-205
  size_t n = x.size();
  Vec r(n), zero(n, 0.0);
  for (size_t i = 0; i < n; ++i) r[i] = rng.normal();
  Leap l0 = leapfrog(x, r, zero, 0.0, f);
  double epsilon = 1.0;
  Leap l1 = leapfrog(x, l0.r, l0.grad, epsilon, f);
  double prob = std::exp(l1.logf - l0.logf - 0.5 * (dotv(l1.r) - dotv(l0.r)));
  note_margin(prob, 0.5);
  int pm = 2 * (prob > 0.5 ? 1 : 0) - 1;
  int guard = 0;
  while (std::pow(prob, pm) > std::pow(0.5, pm)) {
    epsilon *= std::pow(2.0, pm);
    l1 = leapfrog(x, l0.r, l0.grad, epsilon, f);
    prob = std::exp(l1.logf - l0.logf - 0.5 * (dotv(l1.r) - dotv(l0.r)));
    note_margin(prob, 0.5);
    if (++guard > 2000) break;   // not in the reference; guards the oracle against NaN loops
  }
  return epsilon;
}
inline void nuts_setadapt(Tune& t, bool adapt) {   // nuts.jl:84-92
  if (adapt && !t.adapt) { t.m = 0; t.mu = std::log(10.0 * t.epsilon); }
  t.adapt = adapt;
}
inline void nuts_sample(Vec& v, Tune& t, const LogFGrad& f, bool adapt, Rng& rng, int max_depth) {  // nuts.jl:63-81
  nuts_setadapt(t, adapt);
  if (t.adapt) {
    t.m += 1;
    nuts_sub(v, t, t.epsilon, f, rng, max_depth);
    double p = 1.0 / ((double)t.m + t.t0);
    t.Hbar = (1.0 - p) * t.Hbar + p * (t.target - t.alpha / (double)t.nalpha);
    t.epsilon = std::exp(t.mu - std::sqrt((double)t.m) * t.Hbar / t.gamma);
    p = std::pow((double)t.m, -t.kappa);
    t.epsilonbar = std::exp(p * std::log(t.epsilon) + (1.0 - p) * std::log(t.epsilonbar));
  } else {
    if (t.m > 0) t.epsilon = t.epsilonbar;
    nuts_sub(v, t, t.epsilon, f, rng, max_depth);
  }
}

// ---------------------------------------------------------------------------- MALA
// SigmaL: empty = identity (UniformScaling), else k×k lower-triangular Cholesky factor, column-major.
inline void mala_sample(Vec& v, double epsilon, const Vec& SigmaL, const LogFGrad& f, Rng& rng) {   // mala.jl:67-86
  const size_t n = v.size();
  const double se = std::sqrt(epsilon);
  auto Lmul = [&](const Vec& z) {            // L z, L = sqrt(epsilon) SigmaL
    Vec r(n);
    if (SigmaL.empty()) { for (size_t i = 0; i < n; ++i) r[i] = se * z[i]; return r; }
    for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t k = 0; k <= i; ++k) s += SigmaL[i + k * n] * z[k]; r[i] = se * s; }
    return r;
  };
  auto Ltmul = [&](const Vec& z) {           // L' z
    Vec r(n);
    if (SigmaL.empty()) { for (size_t i = 0; i < n; ++i) r[i] = se * z[i]; return r; }
    for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t k = i; k < n; ++k) s += SigmaL[k + i * n] * z[k]; r[i] = se * s; }
    return r;
  };
  auto Linv = [&](const Vec& w) {            // inv(L) w by forward substitution
    Vec r(n);
    if (SigmaL.empty()) { for (size_t i = 0; i < n; ++i) r[i] = w[i] / se; return r; }
    for (size_t i = 0; i < n; ++i) { double s = w[i]; for (size_t k = 0; k < i; ++k) s -= se * SigmaL[i + k * n] * r[k]; r[i] = s / (se * SigmaL[i + i * n]); }
    return r;
  };
  auto M2mul = [&](const Vec& g) { Vec t = Lmul(Ltmul(g)); for (double& x : t) x *= 0.5; return t; };   // M2 = 0.5 L L'
  Vec grad0, grad1, z(n), y(n), w(n);
  const double logf0 = f(v, grad0);
  for (size_t i = 0; i < n; ++i) z[i] = rng.normal();
  const Vec m0 = M2mul(grad0), lz = Lmul(z);
  for (size_t i = 0; i < n; ++i) y[i] = v[i] + m0[i] + lz[i];
  const double logf1 = f(y, grad1);
  const Vec m1 = M2mul(grad1);
  for (size_t i = 0; i < n; ++i) w[i] = v[i] - y[i] - m1[i];
  const double q0 = -0.5 * dotv(Linv(w));
  for (size_t i = 0; i < n; ++i) w[i] = y[i] - v[i] - m0[i];
  const double q1 = -0.5 * dotv(Linv(w));
  if (mh_test(rng.uniform(), (logf1 - q1) - (logf0 - q0))) v = y;
}

// ---------------------------------------------------------------------------- HMC
// SigmaL: empty = identity (UniformScaling), else k×k lower-triangular Cholesky factor, column-major.
inline void hmc_sample(Vec& v, double epsilon, int L, const Vec& SigmaL, const LogFGrad& f, Rng& rng) {   // hmc.jl:72-111
  size_t n = v.size();
  Vec x1 = v, grad0, grad1;
  double logf0 = f(x1, grad0); double logf1 = logf0; grad1 = grad0;
  Vec z(n), p0(n), p1(n);
  for (size_t i = 0; i < n; ++i) z[i] = rng.normal();
  if (SigmaL.empty()) p0 = z;
  else for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t k = 0; k <= i; ++k) s += SigmaL[i + k * n] * z[k]; p0[i] = s; }
  p1 = p0;
  for (size_t i = 0; i < n; ++i) p1[i] += 0.5 * epsilon * grad0[i];
  for (int l = 0; l < L; ++l) {
    for (size_t i = 0; i < n; ++i) x1[i] += epsilon * p1[i];
    logf1 = f(x1, grad1);
    for (size_t i = 0; i < n; ++i) p1[i] += epsilon * grad1[i];
  }
  for (size_t i = 0; i < n; ++i) p1[i] -= 0.5 * epsilon * grad1[i];
  for (size_t i = 0; i < n; ++i) p1[i] *= -1.0;
  auto kinetic = [&](const Vec& p) {
    if (SigmaL.empty()) return 0.5 * dotv(p);
    Vec w(n);   // forward substitution: SigmaL \ p  (== inv(SigmaL) * p)
    for (size_t i = 0; i < n; ++i) { double s = p[i]; for (size_t k = 0; k < i; ++k) s -= SigmaL[i + k * n] * w[k]; w[i] = s / SigmaL[i + i * n]; }
    return 0.5 * dotv(w);
  };
  double Kp0 = kinetic(p0), Kp1 = kinetic(p1);
  if (mh_test(rng.uniform(), (logf1 - Kp1) - (logf0 - Kp0))) v = x1;
}

// ---------------------------------------------------------------------------- AMM
// cholfact(Sigma)[:L] — unpivoted lower Cholesky (amm.jl:18-19); returns false if not p.d.
inline bool chol_lower(const Vec& A, size_t n, Vec& L) {
  L.assign(n * n, 0.0);
  for (size_t j = 0; j < n; ++j) {
    double d = A[j + j * n];
    for (size_t k = 0; k < j; ++k) d -= L[j + k * n] * L[j + k * n];
    if (!(d > 0)) return false;
    double dj = std::sqrt(d); L[j + j * n] = dj;
    for (size_t i = j + 1; i < n; ++i) {
      double s = A[i + j * n];
      for (size_t k = 0; k < j; ++k) s -= L[i + k * n] * L[j + k * n];
      L[i + j * n] = s / dj;
    }
  }
  return true;
}
// cholfact(Hermitian(Sigma), Val{true}) — LAPACK dpstrf (lower, complete pivoting, default
// tolerance n*eps*max(diag)); returns rank and writes P*L (amm.jl:88-91).
inline size_t pivoted_chol_PL(const Vec& A_in, size_t n, Vec& PL) {
  Vec A = A_in; std::vector<size_t> piv(n);
  for (size_t i = 0; i < n; ++i) piv[i] = i;
  Vec L(n * n, 0.0);
  double amax = 0; for (size_t i = 0; i < n; ++i) amax = std::fmax(amax, A[i + i * n]);
  if (!(amax > 0)) { PL.assign(n * n, 0.0); return 0; }
  double tol = (double)n * 2.220446049250313e-16 * amax;
  Vec dots(n, 0.0);   // running sum of squares of the computed row of L (dpstrf work array)
  size_t rank = n;
  for (size_t j = 0; j < n; ++j) {
    // pivot: largest remaining diagonal of the Schur complement
    size_t pvt = j; double best = -1;
    for (size_t i = j; i < n; ++i) {
      double d = A[i + i * n] - dots[i];
      if (i > j) note_margin(d, best, amax);            // pivot choice: two candidates within rounding of each other
      if (d > best) { best = d; pvt = i; }
    }
    note_margin(best, tol, amax);                        // rank decision on a (near-)singular Schur complement
    if (best <= tol || std::isnan(best)) { rank = j; break; }
    if (pvt != j) {
      // symmetric swap of rows/cols j and pvt in A, and of the computed rows of L
      for (size_t k = 0; k < n; ++k) std::swap(A[j + k * n], A[pvt + k * n]);
      for (size_t k = 0; k < n; ++k) std::swap(A[k + j * n], A[k + pvt * n]);
      for (size_t k = 0; k < j; ++k) std::swap(L[j + k * n], L[pvt + k * n]);
      std::swap(dots[j], dots[pvt]); std::swap(piv[j], piv[pvt]);
    }
    double ajj = std::sqrt(best); L[j + j * n] = ajj;
    for (size_t i = j + 1; i < n; ++i) {
      double s = A[i + j * n];
      for (size_t k = 0; k < j; ++k) s -= L[i + k * n] * L[j + k * n];
      L[i + j * n] = s / ajj;
      dots[i] += L[i + j * n] * L[i + j * n];
    }
  }
  // F[:P] * F[:L]: row piv[i] of the result is row i of L
  PL.assign(n * n, 0.0);
  for (size_t i = 0; i < n; ++i) for (size_t k = 0; k < n; ++k) PL[piv[i] + k * n] = L[i + k * n];
  return rank;
}
inline void amm_setadapt(Vec& v, Tune& t, bool adapt) {   // amm.jl:97-108
  if (adapt && !t.adapt) {
    size_t n = v.size();
    t.m = 0; t.Mv = v; t.Mvv.assign(n * n, 0.0);
    for (size_t i = 0; i < n; ++i) for (size_t k = 0; k < n; ++k) t.Mvv[i + k * n] = v[i] * v[k];
    t.SigmaLm.assign(n * n, 0.0);
  }
  t.adapt = adapt;
}
inline void amm_sample(Vec& v, Tune& t, const LogF& logf, bool adapt, Rng& rng) {   // amm.jl:66-94
  size_t n = v.size();
  amm_setadapt(v, t, adapt);
  Vec z(n), x(n);
  for (size_t i = 0; i < n; ++i) z[i] = rng.normal();
  for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t k = 0; k <= i; ++k) s += t.SigmaL[i + k * n] * z[k]; x[i] = s; }
  if (t.m > 2 * (long)n) {
    Vec z2(n);
    for (size_t i = 0; i < n; ++i) z2[i] = rng.normal();
    for (size_t i = 0; i < n; ++i) {
      double s = 0; for (size_t k = 0; k < n; ++k) s += t.SigmaLm[i + k * n] * z2[k];
      x[i] = t.beta * x[i] + (1.0 - t.beta) * s;
    }
  }
  for (size_t i = 0; i < n; ++i) x[i] += v[i];
  double u = rng.uniform();
  double lx = logf(x), lv = logf(v);
  if (mh_test(u, lx - lv)) v = x;
  if (t.adapt) {
    t.m += 1;
    double p = (double)t.m / ((double)t.m + 1.0);
    for (size_t i = 0; i < n; ++i) t.Mv[i] = p * t.Mv[i] + (1.0 - p) * v[i];
    for (size_t i = 0; i < n; ++i) for (size_t k = 0; k < n; ++k) t.Mvv[i + k * n] = p * t.Mvv[i + k * n] + (1.0 - p) * v[i] * v[k];
    Vec Sigma(n * n);
    double c = t.scale * t.scale / (double)n / p;
    for (size_t i = 0; i < n; ++i) for (size_t k = 0; k < n; ++k) Sigma[i + k * n] = c * (t.Mvv[i + k * n] - t.Mv[i] * t.Mv[k]);
    Vec PL;
    if (pivoted_chol_PL(Sigma, n, PL) == n) t.SigmaLm = PL;
  }
}

}  // namespace orc
