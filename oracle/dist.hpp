// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// dist.hpp — the distribution layer the reference's density bottoms out in.
//   * link / invlink / transformed logpdf : src/distributions/transformdistribution.jl:6-93
//   * logpdf_sub / link_sub / invlink_sub dispatch over {single univariate on an array node,
//     array of univariates, single multivariate} : src/distributions/distributionstruct.jl:84-168
//   * the log-densities themselves live in Distributions.jl (>= 0.10.0, REQUIRE:2; NOT in the
//     reference tree).  They are restated from that package's published formulas (SURVEY.md
//     App. B); the Rmath saddle-point evaluation of dbinom/dpois is replaced by the closed form,
//     which agrees to rounding.  PARITY UNPINNED at this boundary: no reference test asserts any
//     logpdf value.
#pragma once
#include <cmath>
#include <limits>
#include <vector>

namespace orc {

static const double LOG2PI = 1.8378770664093454835606594728112;
static const double NEG_INF = -std::numeric_limits<double>::infinity();

// src/utils.jl:64-68
inline double invlogit(double x) { return 1.0 / (std::exp(-x) + 1.0); }
inline double logit(double x) { return std::log(x / (1.0 - x)); }

inline double lgam(double x) { int sg; return ::lgamma_r(x, &sg); }

inline double digamma(double x) {
  // asymptotic series after upward recurrence; |err| < 1e-15 for x > 0
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
             f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + std::log(x) - 0.5 / x + t;
}

enum DKind { D_NULL = 0, D_NORMAL, D_INVGAMMA, D_GAMMA, D_EXPONENTIAL, D_BINOMIAL, D_POISSON, D_BERNOULLI, D_LAPLACE, D_UNIFORM, D_BETA, D_TRUNCNORMAL };

struct UDist {
  DKind k = D_NULL;
  double a = 0, b = 0;  // Normal(mu, sigma); InverseGamma(shape, scale); Gamma(shape, scale);
                        // Exponential(scale); Binomial(n, p); Poisson(lambda); Bernoulli(p); Laplace(location, scale);
                        // Uniform(a, b); Beta(alpha, beta); Truncated(Normal(mu = a, sigma = b), lo, hi)
  double lo = 0, hi = 0;  // truncation bounds (D_TRUNCNORMAL only; +-Inf for a one-sided truncation)
};

// @distr_support bounds of Distributions.jl (minimum(d), maximum(d))
inline double dmin(const UDist& d) {
  switch (d.k) {
    case D_NORMAL: case D_LAPLACE: return NEG_INF;
    case D_UNIFORM: return d.a;
    case D_TRUNCNORMAL: return d.lo;
    default: return 0.0;
  }
}
inline double dmax(const UDist& d) {
  switch (d.k) {
    case D_BINOMIAL: return d.a;
    case D_BERNOULLI: case D_BETA: return 1.0;
    case D_UNIFORM: return d.b;
    case D_TRUNCNORMAL: return d.hi;
    default: return -NEG_INF;
  }
}
inline bool is_discrete(const UDist& d) { return d.k == D_BINOMIAL || d.k == D_POISSON || d.k == D_BERNOULLI; }

inline bool insupport(const UDist& d, double x) {
  if (std::isnan(x)) return false;
  if (is_discrete(d)) return x == std::floor(x) && x >= dmin(d) && x <= dmax(d);
  return x >= dmin(d) && x <= dmax(d);
}

inline double logpdf(const UDist& d, double x) {
  switch (d.k) {
    case D_NORMAL: {  // StatsFuns.normlogpdf
      double z = (x - d.a) / d.b;
      return -(z * z + LOG2PI) / 2.0 - std::log(d.b);
    }
    case D_INVGAMMA:  // Distributions/univariate/continuous/inversegamma.jl
      return d.a * std::log(d.b) - lgam(d.a) - (d.a + 1.0) * std::log(x) - d.b / x;
    case D_GAMMA:     // dgamma(x, shape, scale, log)
      return -lgam(d.a) - d.a * std::log(d.b) + (d.a - 1.0) * std::log(x) - x / d.b;
    case D_EXPONENTIAL: {
      double lambda = 1.0 / d.a;
      return x < 0 ? NEG_INF : std::log(lambda) - lambda * x;
    }
    case D_BINOMIAL: {  // dbinom(x, n, p, log): closed form of the saddle-point evaluation
      double n = d.a, p = d.b, q = 1.0 - p;
      if (p == 0.0) return x == 0.0 ? 0.0 : NEG_INF;
      if (q == 0.0) return x == n ? 0.0 : NEG_INF;
      double lc = lgam(n + 1.0) - lgam(x + 1.0) - lgam(n - x + 1.0);
      double lp = lc;
      if (x > 0) lp += x * std::log(p);
      if (n - x > 0) lp += (n - x) * std::log(q);
      return lp;
    }
    case D_POISSON: {
      double lam = d.a;
      if (lam == 0.0) return x == 0.0 ? 0.0 : NEG_INF;
      return x * std::log(lam) - lam - lgam(x + 1.0);
    }
    case D_BERNOULLI:
      return x == 0.0 ? std::log(1.0 - d.a) : (x == 1.0 ? std::log(d.a) : NEG_INF);
    case D_LAPLACE:   // Distributions/univariate/continuous/laplace.jl: -(|x - mu| / theta + log(2 theta))
      return -(std::fabs(x - d.a) / d.b + std::log(2.0 * d.b));
    case D_UNIFORM:   // Distributions/univariate/continuous/uniform.jl: insupport ? -log(b - a) : -Inf
      return (x >= d.a && x <= d.b) ? -std::log(d.b - d.a) : NEG_INF;
    case D_BETA:      // (alpha - 1) log x + (beta - 1) log1p(-x) - lbeta(alpha, beta)
      return (d.a - 1.0) * std::log(x) + (d.b - 1.0) * std::log1p(-x) - (lgam(d.a) + lgam(d.b) - lgam(d.a + d.b));
    case D_TRUNCNORMAL: {  // Distributions/truncate.jl: logpdf(untruncated, x) - logtp inside [lo, hi], tp = cdf(hi) - cdf(lo)
      if (!(x >= d.lo && x <= d.hi)) return NEG_INF;
      auto Phi = [&](double t) { return std::isinf(t) ? (t > 0 ? 1.0 : 0.0) : 0.5 * std::erfc(-(t - d.a) / d.b * M_SQRT1_2); };
      const double z = (x - d.a) / d.b;
      return -(z * z + LOG2PI) / 2.0 - std::log(d.b) - std::log(Phi(d.hi) - Phi(d.lo));
    }
    default: return 0.0;
  }
}

// link / invlink / transformed logpdf: transformdistribution.jl:6-93.  The generic TransformDistribution methods (:6-48) decide from
// (minimum(d), maximum(d)): both finite -> logit((x - a) / (b - a)); lower only -> log(x - a); upper only -> log(b - x); neither -> x.
// The Real / Positive / Unit unions (:53-93) are the special cases (a, b) = (-Inf, Inf), (0, Inf), (0, 1) of the same maps.
// Discrete distributions fall to the `link(d::Distribution, x) = x` fallbacks of distributionstruct.jl:84,104,136.
enum LinkKind { LK_IDENT = 0, LK_LOG = 1, LK_BOUNDED = 2, LK_UPPER = 3 };
inline LinkKind linkkind(const UDist& d) {
  if (is_discrete(d) || d.k == D_NULL) return LK_IDENT;
  const bool lower = std::isfinite(dmin(d)), upper = std::isfinite(dmax(d));
  return lower && upper ? LK_BOUNDED : lower ? LK_LOG : upper ? LK_UPPER : LK_IDENT;
}
inline double link(const UDist& d, double x) {
  const double a = dmin(d), b = dmax(d);
  switch (linkkind(d)) {
    case LK_BOUNDED: return logit((x - a) / (b - a));
    case LK_LOG: return std::log(x - a);
    case LK_UPPER: return std::log(b - x);
    default: return x;
  }
}
inline double invlink(const UDist& d, double x) {
  const double a = dmin(d), b = dmax(d);
  switch (linkkind(d)) {
    case LK_BOUNDED: return (b - a) * invlogit(x) + a;
    case LK_LOG: return std::exp(x) + a;
    case LK_UPPER: return b - std::exp(x);
    default: return x;
  }
}
inline double logpdf(const UDist& d, double x, bool transform) {
  double lp = logpdf(d, x);
  if (transform) {
    const double a = dmin(d), b = dmax(d);
    switch (linkkind(d)) {
      case LK_BOUNDED: lp += std::log((x - a) * (b - x) / (b - a)); break;   // :39-40
      case LK_LOG: lp += std::log(x - a); break;                              // :41-42, :75-78
      case LK_UPPER: lp += std::log(b - x); break;                            // :43-44
      default: break;
    }
  }
  return lp;
}
// d(constrained value)/d(link value) and d(log-Jacobian)/d(link value) at the constrained value x: chain rule of the analytic gradient
inline void link_chain(const UDist& d, double x, double& dtheta, double& djac) {
  const double a = dmin(d), b = dmax(d);
  switch (linkkind(d)) {
    case LK_BOUNDED: dtheta = (x - a) * (b - x) / (b - a); djac = ((b - x) - (x - a)) / (b - a); break;
    case LK_LOG: dtheta = x - a; djac = 1.0; break;
    case LK_UPPER: dtheta = -(b - x); djac = 1.0; break;
    default: dtheta = 1.0; djac = 0.0; break;
  }
}
// distributionstruct.jl:138-140
inline double logpdf_sub(const UDist& d, double x, bool transform) {
  return insupport(d, x) ? logpdf(d, x, transform) : NEG_INF;
}

// A node's `distr` field: one of the three DistributionStruct shapes.
struct Distr {
  enum Form { NONE = 0, UNI, UNI_ARRAY, MVNORMAL_ISO } form = NONE;
  UDist u;                   // UNI: one univariate distribution (possibly on an array node)
  std::vector<UDist> arr;    // UNI_ARRAY: UnivariateDistribution[...]
  std::vector<double> mu;    // MVNORMAL_ISO: MvNormal(mu, sigma) / MvNormal(d, sigma)
  double sigma = 1.0;
};

// logpdf_sub(distr, value, transform): distributionstruct.jl:142-158 ; IsoNormal from Distributions.jl
inline double logpdf_sub(const Distr& D, const std::vector<double>& x, bool transform) {
  switch (D.form) {
    case Distr::UNI: {
      double lp = 0.0;
      for (double xi : x) lp += logpdf_sub(D.u, xi, transform);
      return lp;
    }
    case Distr::UNI_ARRAY: {
      double lp = 0.0;
      for (size_t i = 0; i < D.arr.size(); ++i) lp += logpdf_sub(D.arr[i], x[i], transform);
      return lp;
    }
    case Distr::MVNORMAL_ISO: {
      // insupport(d::AbstractMvNormal, x) = length match && all finite
      if (x.size() != D.mu.size()) return NEG_INF;
      for (double xi : x) if (!std::isfinite(xi)) return NEG_INF;
      double v = D.sigma * D.sigma;   // ScalMat(d, abs2(sigma))
      double d = (double)x.size();
      double sq = 0.0;
      for (size_t i = 0; i < x.size(); ++i) { double e = x[i] - D.mu[i]; sq += e * e; }
      // mvnormal_c0 - sqmahal/2 ; link is the identity for multivariate normals
      return -(d * LOG2PI + d * std::log(v)) / 2.0 - (sq / v) / 2.0;
    }
    default: return 0.0;
  }
}
inline void link_sub(const Distr& D, const std::vector<double>& x, std::vector<double>& y) {
  y.resize(x.size());
  for (size_t i = 0; i < x.size(); ++i)
    y[i] = D.form == Distr::UNI ? link(D.u, x[i]) : D.form == Distr::UNI_ARRAY ? link(D.arr[i], x[i]) : x[i];
}
inline void invlink_sub(const Distr& D, const double* x, size_t n, std::vector<double>& y) {
  y.resize(n);
  for (size_t i = 0; i < n; ++i)
    y[i] = D.form == Distr::UNI ? invlink(D.u, x[i]) : D.form == Distr::UNI_ARRAY ? invlink(D.arr[i], x[i]) : x[i];
}

}  // namespace orc
