// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// output.hpp — chain post-processing on the hot path, restated.
//   gelmandiag            : src/output/gelmandiag.jl:3-60
//   link(c) heuristics    : src/output/chains.jl:237-246 ; src/output/modelchains.jl:57-76
//   summarystats / ESS    : src/output/stats.jl:85-94
//   mcse_bm / mcse_imse   : src/output/mcse.jl:10-33
//   autocov, sem          : StatsBase.jl (>= 0.7.4, REQUIRE:6; not in tree)
//   quantile(FDist(..))   : Distributions.jl → Rmath qf (not in tree); solved here by bisection
//                           on the regularised incomplete beta function.
// Chains are Julia column-major [n × p × m] (iteration fastest).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "dist.hpp"

namespace orc {

// ---- special functions ----------------------------------------------------------------------
inline double betacf(double a, double b, double x) {   // Lentz continued fraction for I_x(a,b)
  const double FPMIN = 1e-300, EPS = 1e-16;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (std::fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d; double h = d;
  for (int m = 1; m <= 100000; ++m) {
    double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; double del = d * c; h *= del;
    if (std::fabs(del - 1.0) < EPS) break;
  }
  return h;
}
inline double ibeta(double a, double b, double x) {   // regularised incomplete beta
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  double lbt = lgam(a + b) - lgam(a) - lgam(b) + a * std::log(x) + b * std::log1p(-x);
  if (x < (a + 1.0) / (a + b + 2.0)) return std::exp(lbt) * betacf(a, b, x) / a;
  return 1.0 - std::exp(lbt) * betacf(b, a, 1.0 - x) / b;
}
inline double gammp(double a, double x) {   // regularised lower incomplete gamma
  if (x <= 0) return 0.0;
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (std::fabs(del) < std::fabs(sum) * 1e-16) break; }
    return sum * std::exp(-x + a * std::log(x) - lgam(a));
  }
  const double FPMIN = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / FPMIN, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; ++i) {
    double an = -i * (i - a); b += 2.0;
    d = an * d + b; if (std::fabs(d) < FPMIN) d = FPMIN;
    c = b + an / c; if (std::fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; double del = d * c; h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lgam(a)) * h;
}
inline double fcdf(double q, double d1, double d2) {
  if (q <= 0) return 0.0;
  if (std::isinf(d2)) return gammp(d1 / 2.0, d1 * q / 2.0);   // F(d1, inf) = chisq(d1)/d1
  double xx = d1 * q / (d1 * q + d2);
  if (xx > 0.5) return 1.0 - ibeta(d2 / 2.0, d1 / 2.0, d2 / (d1 * q + d2));
  return ibeta(d1 / 2.0, d2 / 2.0, xx);
}
inline double fquantile(double p, double d1, double d2) {   // quantile(FDist(d1, d2), p)
  if (std::isnan(d1) || std::isnan(d2) || !(d1 > 0) || !(d2 > 0)) return NAN;
  double lo = 0.0, hi = 1.0;
  int guard = 0;
  while (fcdf(hi, d1, d2) < p && ++guard < 2000) hi *= 2.0;
  for (int it = 0; it < 300; ++it) {
    double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    if (fcdf(mid, d1, d2) < p) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

// ---- basic statistics (Julia Base / StatsBase semantics) ---------------------------------------
inline double mean_of(const double* x, size_t n, size_t stride = 1) { double s = 0; for (size_t i = 0; i < n; ++i) s += x[i * stride]; return s / (double)n; }
inline double var_of(const double* x, size_t n, size_t stride = 1) {   // corrected (n-1)
  double mu = mean_of(x, n, stride), s = 0;
  for (size_t i = 0; i < n; ++i) { double d = x[i * stride] - mu; s += d * d; }
  return s / (double)(n - 1);
}
inline double cov_of(const double* x, const double* y, size_t n) {
  double mx = mean_of(x, n), my = mean_of(y, n), s = 0;
  for (size_t i = 0; i < n; ++i) s += (x[i] - mx) * (y[i] - my);
  return s / (double)(n - 1);
}
inline double sem_of(const double* x, size_t n) { return std::sqrt(var_of(x, n)) / std::sqrt((double)n); }

// StatsBase.autocov(x, lags): demeaned, divided by length(x)
inline double autocov_lag(const std::vector<double>& z /*demeaned*/, size_t k) {
  size_t n = z.size(); double s = 0;
  for (size_t t = 0; t + k < n; ++t) s += z[t] * z[t + k];
  return s / (double)n;
}

// mcse_bm(x; size=100): mcse.jl:10-19 ; returns NaN where the reference throws
inline double mcse_bm(const std::vector<double>& x, size_t size) {
  size_t n = x.size(), m = n / size;
  if (m < 2) return NAN;
  std::vector<double> mbar(m);
  for (size_t i = 0; i < m; ++i) mbar[i] = mean_of(&x[i * size], size);
  return sem_of(mbar.data(), m);
}
// mcse_imse(x): mcse.jl:21-33
inline double mcse_imse(const std::vector<double>& x) {
  size_t n = x.size(); long m = ((long)n - 2) / 2;
  double mu = mean_of(x.data(), n);
  std::vector<double> z(n); for (size_t i = 0; i < n; ++i) z[i] = x[i] - mu;
  double g0 = autocov_lag(z, 0), g1 = autocov_lag(z, 1);
  double Ghat = g0 + g1;
  double value = -g0 + 2.0 * Ghat;
  for (long i = 1; i <= m; ++i) {
    Ghat = std::fmin(Ghat, autocov_lag(z, 2 * i) + autocov_lag(z, 2 * i + 1));
    if (!(Ghat > 0)) break;
    value += 2.0 * Ghat;
  }
  return std::sqrt(value / (double)n);
}

// summarystats(c; etype): stats.jl:85-94.  out [p × 5] row-major: Mean, SD, Naive SE, MCSE, ESS.
inline void summarystats(const double* c, size_t n, size_t p, size_t m, int etype, size_t batch, double* out) {
  std::vector<double> x(n * m);
  for (size_t j = 0; j < p; ++j) {
    for (size_t k = 0; k < m; ++k) for (size_t i = 0; i < n; ++i) x[k * n + i] = c[i + n * (j + p * k)];  // vec(x): chain-major
    double mu = mean_of(x.data(), x.size());
    double sd = std::sqrt(var_of(x.data(), x.size()));
    double se = sd / std::sqrt((double)x.size());
    double mc = etype == 0 ? mcse_bm(x, batch) : mcse_imse(x);
    double ess = std::fmin((sd / mc) * (sd / mc), (double)n);
    out[j * 5 + 0] = mu; out[j * 5 + 1] = sd; out[j * 5 + 2] = se; out[j * 5 + 3] = mc; out[j * 5 + 4] = ess;
  }
}

// link(c::AbstractChains): chains.jl:237-246 — per column: log if all > 0 (logit if also all < 1).
// For ModelChains (modelchains.jl:57-76) monitored stochastic nodes use the node's own link;
// `linkcode[j]` = -1 heuristic, 0 identity, 1 log.
inline void link_chains(const double* c, size_t n, size_t p, size_t m, const int* linkcode, std::vector<double>& cc) {
  cc.assign(c, c + n * p * m);
  for (size_t j = 0; j < p; ++j) {
    int code = linkcode ? linkcode[j] : -1;
    if (code == 0) continue;
    double mn = INFINITY, mx = -INFINITY;
    for (size_t k = 0; k < m; ++k) for (size_t i = 0; i < n; ++i) { double v = c[i + n * (j + p * k)]; mn = std::fmin(mn, v); mx = std::fmax(mx, v); }
    bool dolog = code == 1, dologit = false;
    if (code == -1) { if (mn > 0.0) { if (mx < 1.0) dologit = true; else dolog = true; } }
    if (!dolog && !dologit) continue;
    for (size_t k = 0; k < m; ++k) for (size_t i = 0; i < n; ++i) {
      double& v = cc[i + n * (j + p * k)];
      v = dologit ? logit(v) : std::log(v);
    }
  }
}

// gelmandiag(c; alpha): gelmandiag.jl:5-47 (without MPSRF and without the final 3-dp rounding).
// psrf [p × 2] row-major.
inline void gelmandiag(const double* psi, size_t n, size_t p, size_t m, double alpha, double* psrf) {
  std::vector<double> psibar(m), s2(m), pb2(m);
  for (size_t j = 0; j < p; ++j) {
    for (size_t k = 0; k < m; ++k) {
      const double* col = psi + n * (j + p * k);
      psibar[k] = mean_of(col, n); s2[k] = var_of(col, n); pb2[k] = psibar[k] * psibar[k];
    }
    double w = mean_of(s2.data(), m);                         // diag(W)
    double b = (double)n * var_of(psibar.data(), m);          // diag(B)
    double psibar2 = mean_of(psibar.data(), m);
    double var_w = var_of(s2.data(), m) / (double)m;
    double var_b = (2.0 / (double)(m - 1)) * b * b;
    double var_wb = ((double)n / (double)m) * (cov_of(s2.data(), pb2.data(), m) - 2.0 * psibar2 * cov_of(s2.data(), psibar.data(), m));
    double V = ((double)(n - 1) / (double)n) * w + ((double)(m + 1) / (double)(m * n)) * b;
    double var_V = ((double)(n - 1) * (double)(n - 1) * var_w + ((double)(m + 1) / (double)m) * ((double)(m + 1) / (double)m) * var_b +
                    (2.0 * (double)(n - 1) * (double)(m + 1) / (double)m) * var_wb) / ((double)n * (double)n);
    double df = 2.0 * V * V / var_V;
    double B_df = (double)(m - 1);
    double W_df = 2.0 * w * w / var_w;
    double R_fixed = (double)(n - 1) / (double)n;
    double R_random_scale = (double)(m + 1) / (double)(m * n);
    double q = 1.0 - alpha / 2.0;
    double correction = (df + 3.0) / (df + 1.0);
    double R_random = R_random_scale * b / w;
    psrf[j * 2 + 0] = std::sqrt(correction * (R_fixed + R_random));
    if (!std::isnan(R_random)) R_random *= fquantile(q, B_df, W_df);
    psrf[j * 2 + 1] = std::sqrt(correction * (R_fixed + R_random));
  }
}

}  // namespace orc
