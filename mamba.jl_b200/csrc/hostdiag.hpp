// hostdiag.hpp — O(p) host-side finalisation of the on-device diagnostics.
//
// The cross-chain reductions run on the device (engine.cuh); what is left is arithmetic on a few
// doubles per monitored column:
//   gelman_column      PSRF + upper limit from centred cross-chain sums (src/output/gelmandiag.jl:20-47)
//   summary_column     mean / SD / naive SE / batch-means MCSE / ESS from streaming sums (src/output/stats.jl:85-94)
//   summarystats_soa   the reference's exact summarystats over materialised samples
//                      (src/output/stats.jl:85-94, src/output/mcse.jl:10-33; StatsBase autocov/sem)
//   f_quantile         quantile(FDist(d1, d2), p) (gelmandiag.jl:43) — hypergeometric series for the
//                      regularised incomplete beta / gamma functions + bisection
#pragma once
#include <cmath>
#include <vector>

namespace hostdiag {

inline double lgam(double x) { int s; return ::lgamma_r(x, &s); }

// I_x(a,b) = x^a (1-x)^b / (a B(a,b)) * [1 + sum_{n>=0} prod_{i=0..n} x (a+b+i)/(a+1+i)]   (DLMF 8.17.8 form)
inline double ibeta_series(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  const bool flip = x > (a + 1.0) / (a + b + 2.0);
  const double aa = flip ? b : a, bb = flip ? a : b, xx = flip ? 1.0 - x : x;
  double term = 1.0, sum = 1.0;
  for (int n = 0; n < 2000000; ++n) {
    term *= xx * (aa + bb + n) / (aa + 1.0 + n);
    sum += term;
    if (term < sum * 1e-17) break;
  }
  const double logpre = aa * std::log(xx) + bb * std::log1p(-xx) - std::log(aa) - (lgam(aa) + lgam(bb) - lgam(aa + bb));
  const double v = std::exp(logpre) * sum;
  return flip ? 1.0 - v : v;
}
// P(a, x) = x^a e^-x / Gamma(a+1) * sum_{n>=0} x^n / ((a+1)...(a+n))
inline double igamma_series(double a, double x) {
  if (x <= 0.0) return 0.0;
  double term = 1.0, sum = 1.0;
  for (int n = 1; n < 5000000; ++n) {
    term *= x / (a + n);
    sum += term;
    if (term < sum * 1e-17) break;
  }
  const double v = std::exp(a * std::log(x) - x - lgam(a + 1.0)) * sum;
  return v > 1.0 ? 1.0 : v;
}
inline double f_cdf(double q, double d1, double d2) {
  if (q <= 0.0) return 0.0;
  if (std::isinf(d2)) return igamma_series(0.5 * d1, 0.5 * d1 * q);
  return ibeta_series(0.5 * d1, 0.5 * d2, d1 * q / (d1 * q + d2));
}
inline double f_quantile(double p, double d1, double d2) {
  if (!(d1 > 0.0) || !(d2 > 0.0) || std::isnan(d1) || std::isnan(d2)) return NAN;
  double lo = 0.0, hi = 1.0;
  for (int g = 0; g < 1100 && f_cdf(hi, d1, d2) < p; ++g) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (f_cdf(mid, d1, d2) < p) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

// sums = { m, Σd, Σd², Σe, Σe², Σe·d, Σe·d² } with d = psibar - c1, e = s2 - c2 over chains.
inline void gelman_column(double n, double c1, double c2, const double* s, double alpha, double* out) {
  const double m = s[0], Sd = s[1], Sdd = s[2], Se = s[3], See = s[4], Sed = s[5], Sedd = s[6];
  (void)c1;
  const double w = c2 + Se / m;                                   // W = mean of the chain variances
  const double b = n * (Sdd - Sd * Sd / m) / (m - 1.0);           // B = n * var(chain means)
  const double var_w = ((See - Se * Se / m) / (m - 1.0)) / m;
  const double var_b = (2.0 / (m - 1.0)) * b * b;
  const double cov_e_d = (Sed - Se * Sd / m) / (m - 1.0);
  const double cov_e_dd = (Sedd - Se * Sdd / m) / (m - 1.0);
  // cov(s2, psibar²) - 2 mean(psibar) cov(s2, psibar), rewritten around the centre c1
  const double var_wb = (n / m) * (cov_e_dd - 2.0 * (Sd / m) * cov_e_d);
  const double V = ((n - 1.0) / n) * w + ((m + 1.0) / (m * n)) * b;
  const double var_V = ((n - 1.0) * (n - 1.0) * var_w + ((m + 1.0) / m) * ((m + 1.0) / m) * var_b +
                        (2.0 * (n - 1.0) * (m + 1.0) / m) * var_wb) / (n * n);
  const double df = 2.0 * V * V / var_V;
  const double W_df = 2.0 * w * w / var_w;
  const double R_fixed = (n - 1.0) / n;
  const double correction = (df + 3.0) / (df + 1.0);
  double R_random = ((m + 1.0) / (m * n)) * b / w;
  out[0] = std::sqrt(correction * (R_fixed + R_random));
  if (!std::isnan(R_random)) R_random *= f_quantile(1.0 - alpha / 2.0, m - 1.0, W_df);
  out[1] = std::sqrt(correction * (R_fixed + R_random));
}

// sums = { C, Σ mean_c, Σ M2_c, Σ (mean_c - c1)², Σ nb_c, Σ nb_c bmean_c, Σ bM2_c, Σ nb_c (bmean_c - c2)² }
inline void summary_column(double n, double c1, double c2, const double* s, double* out) {
  const double C = s[0], N = C * n;
  const double mean = s[1] / C;
  const double ss = s[2] + n * (s[3] - C * (mean - c1) * (mean - c1));
  const double sd = std::sqrt(ss / (N - 1.0));
  const double NB = s[4];
  double mcse = NAN;
  if (NB >= 2.0) {
    const double gb = s[5] / NB;
    const double ssb = s[6] + s[7] - NB * (gb - c2) * (gb - c2);
    mcse = std::sqrt(ssb / (NB - 1.0)) / std::sqrt(NB);
  }
  const double r = sd / mcse;
  out[0] = mean; out[1] = sd; out[2] = sd / std::sqrt(N); out[3] = mcse; out[4] = std::fmin(r * r, n);
}

inline double mean_v(const double* x, size_t n) { double s = 0; for (size_t i = 0; i < n; ++i) s += x[i]; return s / (double)n; }
inline double sd_v(const double* x, size_t n) {
  const double mu = mean_v(x, n); double s = 0;
  for (size_t i = 0; i < n; ++i) s += (x[i] - mu) * (x[i] - mu);
  return std::sqrt(s / (double)(n - 1));
}

// smp is the device layout [kept][P][C] (chain fastest).  Returns nonzero where the reference throws.
inline int summarystats_soa(const double* smp, long long kept, int P, long long C, int etype, int batch, double* out) {
  const size_t N = (size_t)kept * (size_t)C;
  std::vector<double> x(N);
  for (int j = 0; j < P; ++j) {
    for (long long k = 0; k < C; ++k)
      for (long long i = 0; i < kept; ++i) x[(size_t)k * kept + i] = smp[((size_t)i * P + j) * C + k];   // vec(x): chain-major
    const double mu = mean_v(x.data(), N), sd = sd_v(x.data(), N);
    double mc;
    if (etype == 0) {   // mcse_bm: mcse.jl:10-19
      const size_t nb = N / (size_t)batch;
      if (nb < 2) return 1;
      std::vector<double> mbar(nb);
      for (size_t q = 0; q < nb; ++q) mbar[q] = mean_v(&x[q * batch], batch);
      mc = sd_v(mbar.data(), nb) / std::sqrt((double)nb);
    } else {            // mcse_imse: mcse.jl:21-33
      std::vector<double> z(N);
      for (size_t i = 0; i < N; ++i) z[i] = x[i] - mu;
      auto acov = [&](size_t lag) { double s = 0; for (size_t t = 0; t + lag < N; ++t) s += z[t] * z[t + lag]; return s / (double)N; };
      const double g0 = acov(0);
      double Ghat = g0 + acov(1);
      double value = -g0 + 2.0 * Ghat;
      const long long mm = ((long long)N - 2) / 2;
      for (long long i = 1; i <= mm; ++i) {
        Ghat = std::fmin(Ghat, acov(2 * i) + acov(2 * i + 1));
        if (!(Ghat > 0)) break;
        value += 2.0 * Ghat;
      }
      mc = std::sqrt(value / (double)N);
    }
    const double r = sd / mc;
    out[j * 5 + 0] = mu; out[j * 5 + 1] = sd; out[j * 5 + 2] = sd / std::sqrt((double)N); out[j * 5 + 3] = mc;
    out[j * 5 + 4] = std::fmin(r * r, (double)kept);
  }
  return 0;
}

}  // namespace hostdiag
