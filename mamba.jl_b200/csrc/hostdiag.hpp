// hostdiag.hpp — O(p) host-side finalisation of the on-device diagnostics.
//
// The cross-chain reductions run on the device (engine.cuh); what is left is arithmetic on a few
// doubles per monitored column:
//   gelman_column      PSRF + upper limit from centred cross-chain sums (src/output/gelmandiag.jl:20-47)
//   summary_column     mean / SD / naive SE / batch-means MCSE / ESS from streaming sums (src/output/stats.jl:85-94)
//   summarystats_soa   the reference's exact summarystats over materialised samples
//                      (src/output/stats.jl:85-94, src/output/mcse.jl:10-33; StatsBase autocov/sem)
//   f_quantile         quantile(FDist(d1, d2), p) (gelmandiag.jl:43) — continued fraction / series for the
//                      regularised incomplete beta / gamma functions + bracketed Newton
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace hostdiag {

inline double lgam(double x) { int s; return ::lgamma_r(x, &s); }

// Regularised incomplete beta function I_x(a, b): continued fraction (DLMF 8.17.22) evaluated with the modified Lentz algorithm on the
// side where it converges in O(sqrt(max(a, b))) steps (x < (a + 1) / (a + b + 2), else 1 - I_{1-x}(b, a)).  With one chain per GPU
// thread a = (chains - 1) / 2 is 1e5 - 1e7: the power series needs up to ~1e6 terms near the mode, the fraction a few hundred.
inline double ibeta_series(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  const bool flip = x > (a + 1.0) / (a + b + 2.0);
  const double aa = flip ? b : a, bb = flip ? a : b, xx = flip ? 1.0 - x : x;
  const double tiny = 1e-300, qab = aa + bb, qap = aa + 1.0, qam = aa - 1.0;
  double c = 1.0, d = 1.0 - qab * xx / qap;
  if (std::fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m < 200000; ++m) {
    const double m2 = 2.0 * m;
    double an = m * (bb - m) * xx / ((qam + m2) * (aa + m2));
    d = 1.0 + an * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1.0 + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d; h *= d * c;
    an = -(aa + m) * (qab + m) * xx / ((aa + m2) * (qap + m2));
    d = 1.0 + an * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1.0 + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 2e-16) break;
  }
  const double logpre = aa * std::log(xx) + bb * std::log1p(-xx) - std::log(aa) - (lgam(aa) + lgam(bb) - lgam(aa + bb));
  const double v = std::exp(logpre) * h;
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
// log density of F(d1, d2) (d2 = inf: chi-square(d1) / d1)
inline double f_logpdf(double q, double d1, double d2) {
  if (std::isinf(d2)) return std::log(d1) + (0.5 * d1 - 1.0) * std::log(d1 * q) - 0.5 * d1 * q - 0.5 * d1 * std::log(2.0) - lgam(0.5 * d1);
  return 0.5 * (d1 * std::log(d1 * q) + d2 * std::log(d2) - (d1 + d2) * std::log(d1 * q + d2)) - std::log(q) -
         (lgam(0.5 * d1) + lgam(0.5 * d2) - lgam(0.5 * (d1 + d2)));
}
// quantile(FDist(d1, d2), p): bracketed Newton on the CDF (a CDF evaluation is a series of O(sqrt(d)) terms, and with one chain per GPU
// thread d1 = chains - 1 is ~1e5-1e7: plain bisection cost 10 ms per gelmandiag call), bisection whenever a Newton step leaves the bracket
inline double f_quantile(double p, double d1, double d2) {
  if (!(d1 > 0.0) || !(d2 > 0.0) || std::isnan(d1) || std::isnan(d2)) return NAN;
  double lo = 0.0, hi = 1.0;
  for (int g = 0; g < 1100 && f_cdf(hi, d1, d2) < p; ++g) { lo = hi; hi *= 2.0; }
  double q = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    const double F = f_cdf(q, d1, d2);
    if (F < p) lo = q; else hi = q;
    const double dens = std::exp(f_logpdf(q, d1, d2));
    double qn = (dens > 0.0 && std::isfinite(dens)) ? q - (F - p) / dens : NAN;
    if (!(qn > lo && qn < hi)) qn = 0.5 * (lo + hi);
    if (std::fabs(qn - q) <= 4e-16 * q || !(hi > lo)) { q = qn; break; }
    q = qn;
  }
  return q;
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


// element (iteration i, parameter j, chain k) of a ModelChains.value array [n x p x m], column-major
inline double at(const double* v, long long n, int p, long long i, int j, long long k) { return v[(size_t)i + (size_t)n * ((size_t)j + (size_t)p * (size_t)k)]; }

// ---- per-series pieces shared by the convergence diagnostics ------------------------------------------------------------
// mcse(x, method): src/output/mcse.jl:3-46; etype 0 = :bm (batch size `batch`), 1 = :imse, 2 = :ipse.  Returns nonzero where the
// reference throws ("iterations are < 2 * size ...").
inline int mcse_vec(const double* x, size_t N, int etype, int batch, double* out) {
  if (etype == 0) {
    const size_t nb = N / (size_t)batch;
    if (nb < 2) return 1;
    std::vector<double> mbar(nb);
    for (size_t q = 0; q < nb; ++q) mbar[q] = mean_v(x + q * batch, batch);
    *out = sd_v(mbar.data(), nb) / std::sqrt((double)nb);
    return 0;
  }
  const double mu = mean_v(x, N);
  std::vector<double> z(N);
  for (size_t i = 0; i < N; ++i) z[i] = x[i] - mu;
  auto acov = [&](size_t lag) { double s = 0; for (size_t t = 0; t + lag < N; ++t) s += z[t] * z[t + lag]; return s / (double)N; };   // StatsBase.autocov
  const double g0 = acov(0), g1 = acov(1);
  const long long mm = ((long long)N - 2) / 2;
  double value;
  if (etype == 1) {          // initial monotone sequence estimator
    double Ghat = g0 + g1;
    value = -g0 + 2.0 * Ghat;
    for (long long i = 1; i <= mm; ++i) {
      Ghat = std::fmin(Ghat, acov(2 * i) + acov(2 * i + 1));
      if (!(Ghat > 0)) break;
      value += 2.0 * Ghat;
    }
  } else {                   // initial positive sequence estimator
    value = g0 + 2.0 * g1;
    for (long long i = 1; i <= mm; ++i) {
      const double Ghat = acov(2 * i) + acov(2 * i + 1);
      if (!(Ghat > 0)) break;
      value += 2.0 * Ghat;
    }
  }
  *out = std::sqrt(value / (double)N);
  return 0;
}
inline double erfinv_d(double y) {   // inverse error function: Newton on erf from a logarithmic starting value
  if (!(y > -1.0 && y < 1.0)) return y == 1.0 ? INFINITY : y == -1.0 ? -INFINITY : NAN;
  const double a = 0.147, ln = std::log(1.0 - y * y), t = 2.0 / (M_PI * a) + 0.5 * ln;
  double x = (y < 0 ? -1.0 : 1.0) * std::sqrt(std::sqrt(t * t - ln / a) - t);
  for (int it = 0; it < 4; ++it) x -= (std::erf(x) - y) / (2.0 / std::sqrt(M_PI) * std::exp(-x * x));
  return x;
}
// CDF of the Cramer-von Mises statistic, 4-term series: src/utils.jl:73-81
inline double pcramer(double q) {
  double p = 0.0;
  const double fact[4] = {1.0, 1.0, 2.0, 6.0};
  for (int k = 0; k < 4; ++k) {
    const double c1 = 4.0 * k + 1.0, c2 = c1 * c1 / (16.0 * q);
    p += std::tgamma(k + 0.5) / fact[k] * std::sqrt(c1) * std::exp(-c2) * std::cyl_bessel_k(0.25, c2);
  }
  return p / (std::pow(M_PI, 1.5) * std::sqrt(q));
}
inline long long jround(double v) { return (long long)std::llround(v); }   // Julia 0.5 round(Int, x): ties away from zero
// gewekediag(x; first, last, etype): src/output/gewekediag.jl:3-19 → (z, p), NOT rounded
inline int geweke_vec(const double* x, long long n, double first, double last, int etype, int batch, double* out) {
  const long long n1 = jround(first * (double)n), s2 = jround((double)n - last * (double)n + 1.0);   // x[1:n1], x[s2:n]
  if (n1 < 1 || s2 < 1 || s2 > n) return 1;
  double m1, m2;
  if (mcse_vec(x, (size_t)n1, etype, batch, &m1) || mcse_vec(x + (s2 - 1), (size_t)(n - s2 + 1), etype, batch, &m2)) return 1;
  const double z = (mean_v(x, (size_t)n1) - mean_v(x + (s2 - 1), (size_t)(n - s2 + 1))) / std::sqrt(m1 * m1 + m2 * m2);
  out[0] = z; out[1] = 1.0 - std::erf(std::fabs(z) / std::sqrt(2.0));
  return 0;
}
// heideldiag(x; alpha, eps, etype, start): src/output/heideldiag.jl:3-27 → (burn-in, stationarity, p-value, mean, halfwidth, test)
inline int heidel_vec(const double* x, long long n, double alpha, double eps, int etype, int batch, long long start, double* out) {
  const long long delta = (long long)(0.10 * (double)n);
  const long long h0 = (long long)((double)n / 2.0);              // y = x[trunc(Int, n / 2):end]
  if (h0 < 1) return 1;
  double mc;
  if (mcse_vec(x + (h0 - 1), (size_t)(n - h0 + 1), etype, batch, &mc)) return 1;
  const double S0 = (double)(n - h0 + 1) * mc * mc;
  long long i = 1; double pvalue = 1.0, ybar = NAN; bool converged = false;
  long long ylen = n - h0 + 1, yoff = h0 - 1;                      // the series the halfwidth is computed on (last y of the loop)
  while ((double)i < (double)n / 2.0) {
    const double* y = x + (i - 1); const long long m = n - i + 1;
    yoff = i - 1; ylen = m;
    ybar = mean_v(y, (size_t)m);
    double cs = 0.0, I = 0.0;
    for (long long t = 0; t < m; ++t) { cs += y[t]; const double B = cs - ybar * (double)(t + 1); I += B * B / ((double)m * S0); }
    I /= (double)m;
    pvalue = 1.0 - pcramer(I);
    converged = pvalue > alpha;
    if (converged || delta < 1) break;
    i += delta;
  }
  if (mcse_vec(x + yoff, (size_t)ylen, etype, batch, &mc)) return 1;
  const double halfwidth = std::sqrt(2.0) * erfinv_d(1.0 - alpha) * mc;
  out[0] = (double)(i + start - 2); out[1] = converged ? 1.0 : 0.0; out[2] = pvalue; out[3] = ybar; out[4] = halfwidth;
  out[5] = (halfwidth / std::fabs(ybar) <= eps) ? 1.0 : 0.0;
  return 0;
}
// rafterydiag(x; q, r, s, eps, range): src/output/rafterydiag.jl:3-46 → (thinning, burn-in, total, nmin, dependence factor)
inline void raftery_vec(const double* x, long long nx, double q, double r, double s, double eps, long long rstart, long long rstep, double* out) {
  const double phi = std::sqrt(2.0) * erfinv_d(s);
  const double nmin = std::ceil(q * (1.0 - q) * (phi / r) * (phi / r));
  out[3] = nmin;
  if (nmin > (double)nx) { out[0] = out[1] = out[2] = out[4] = NAN; return; }
  std::vector<double> srt(x, x + nx);
  std::sort(srt.begin(), srt.end());
  const double h = (double)(nx - 1) * q; const size_t lo = (size_t)std::floor(h), hi = lo + 1 < (size_t)nx ? lo + 1 : lo;
  const double cut = srt[lo] + (h - (double)lo) * (srt[hi] - srt[lo]);
  std::vector<int> dich((size_t)nx);
  for (long long i = 0; i < nx; ++i) dich[i] = x[i] <= cut ? 1 : 0;
  long long kthin = 0; double bic = 1.0;
  std::vector<int> test;
  while (bic >= 0.0) {
    ++kthin;
    test.clear();
    for (long long i = 0; i < nx; i += kthin) test.push_back(dich[i]);
    const long long nt = (long long)test.size();
    if (nt < 3) break;
    double tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = 0; i + 2 < nt; ++i) tr[test[i] + 2 * test[i + 1] + 4 * test[i + 2]] += 1.0;   // reshape(counts, 2, 2, 2): [i1, i2, i3]
    double g2 = 0.0;
    for (int i1 = 0; i1 < 2; ++i1) for (int i2 = 0; i2 < 2; ++i2) for (int i3 = 0; i3 < 2; ++i3) {
      const double tt = tr[i1 + 2 * i2 + 4 * i3];
      if (tt > 0) {
        const double a = tr[0 + 2 * i2 + 4 * i3] + tr[1 + 2 * i2 + 4 * i3];
        const double b = tr[i1 + 2 * i2] + tr[i1 + 2 * i2 + 4];
        const double c = tr[2 * i2] + tr[1 + 2 * i2] + tr[2 * i2 + 4] + tr[1 + 2 * i2 + 4];
        g2 += 2.0 * tt * std::log(tt / (a * b / c));
      }
    }
    bic = g2 - 2.0 * std::log((double)nt - 2.0);
  }
  const long long nt = (long long)test.size();
  double tf[4] = {0, 0, 0, 0};
  for (long long i = 0; i + 1 < nt; ++i) tf[test[i] + 2 * test[i + 1]] += 1.0;
  const double alpha = tf[2] / (tf[0] + tf[2]), beta = tf[1] / (tf[1] + tf[3]);
  const double kt = (double)(kthin * rstep);
  const double m = std::log(eps * (alpha + beta) / std::fmax(alpha, beta)) / std::log(std::fabs(1.0 - alpha - beta));
  const double burnin = kt * std::ceil(m) + (double)rstart - 1.0;
  const double nn = ((2.0 - alpha - beta) * alpha * beta * phi * phi) / (r * r * std::pow(alpha + beta, 3.0));
  const double total = burnin + kt * std::ceil(nn);
  out[0] = kt; out[1] = burnin; out[2] = total; out[4] = total / nmin;
}
// the three diagnostics over every (parameter, chain) series of a ModelChains.value; out [p x K x m] column-major (K = 2, 6, 5)
// summarystats(c; etype) on a materialised array: src/output/stats.jl:85-94 — each column pooled over chains in vec(x[:, j, :]) order
// (chain-major), ESS = min((SD / MCSE)^2, iterations).  out [p x 5]: mean, SD, naive SE, MCSE, ESS.
inline int chains_summarystats(const double* v, long long n, int p, long long m, int etype, int batch, double* out) {
  const size_t N = (size_t)n * (size_t)m;
  std::vector<double> x(N);
  for (int j = 0; j < p; ++j) {
    for (long long k = 0; k < m; ++k)
      for (long long i = 0; i < n; ++i) x[(size_t)k * n + i] = at(v, n, p, i, j, k);
    const double mu = mean_v(x.data(), N), sd = sd_v(x.data(), N);
    double mc;
    if (mcse_vec(x.data(), N, etype, batch, &mc)) return 1;
    const double r = sd / mc;
    out[j * 5 + 0] = mu; out[j * 5 + 1] = sd; out[j * 5 + 2] = sd / std::sqrt((double)N); out[j * 5 + 3] = mc;
    out[j * 5 + 4] = std::fmin(r * r, (double)n);
  }
  return 0;
}

template <class F>
inline int chains_series(const double* v, long long n, int p, long long m, int K, double* out, F f) {
  std::vector<double> x((size_t)n);
  for (long long k = 0; k < m; ++k)
    for (int j = 0; j < p; ++j) {
      for (long long i = 0; i < n; ++i) x[i] = at(v, n, p, i, j, k);
      double r[8];
      if (f(x.data(), r)) return 1;
      for (int c = 0; c < K; ++c) out[(size_t)j + (size_t)p * ((size_t)c + (size_t)K * (size_t)k)] = r[c];
    }
  return 0;
}

// ---- post-processing of a materialised chain array (ModelChains.value: [n iterations x p parameters x m chains], column-major,
//      iteration fastest) — src/output/stats.jl:3-83 and the multivariate PSRF of src/output/gelmandiag.jl:49-55 -----------------

// quantile(c; q): stats.jl:74-83 — Julia's quantile(vec(x), q) (linear interpolation between order statistics, "type 7")
inline void chains_quantile(const double* v, long long n, int p, long long m, const double* q, int nq, double* out) {
  const size_t N = (size_t)n * (size_t)m;
  std::vector<double> x(N);
  for (int j = 0; j < p; ++j) {
    for (long long k = 0; k < m; ++k) for (long long i = 0; i < n; ++i) x[(size_t)k * n + i] = at(v, n, p, i, j, k);
    std::sort(x.begin(), x.end());
    for (int a = 0; a < nq; ++a) {
      const double h = (double)(N - 1) * q[a];
      const size_t lo = (size_t)std::floor(h);
      const size_t hi = lo + 1 < N ? lo + 1 : lo;
      out[(size_t)j * nq + a] = x[lo] + (h - (double)lo) * (x[hi] - x[lo]);
    }
  }
}
// hpd(c; alpha): stats.jl:52-72 — shortest of the intervals [y_i, y_(n-m+i)], m = max(1, ceil(alpha n)), first minimum
inline void chains_hpd(const double* v, long long n, int p, long long m, double alpha, double* out) {
  const size_t N = (size_t)n * (size_t)m;
  std::vector<double> x(N);
  for (int j = 0; j < p; ++j) {
    for (long long k = 0; k < m; ++k) for (long long i = 0; i < n; ++i) x[(size_t)k * n + i] = at(v, n, p, i, j, k);
    std::sort(x.begin(), x.end());
    size_t mm = (size_t)std::ceil(alpha * (double)N); if (mm < 1) mm = 1; if (mm > N) mm = N;
    size_t best = 0; double bw = x[N - mm] - x[0];
    for (size_t i = 1; i < mm; ++i) { const double w = x[N - mm + i] - x[i]; if (w < bw) { bw = w; best = i; } }
    out[j * 2 + 0] = x[best]; out[j * 2 + 1] = x[N - mm + best];
  }
}
// autocor(c; lags, relative): stats.jl:3-13 over StatsBase.autocor (demeaned, normalised by the lag-0 sum); `lags` are the
// index lags actually applied to the stored series (the caller multiplies by the thinning step when relative = true, as the
// reference does).  out [p x nlags x m], column-major like the reference's ChainSummary value.
inline void chains_autocor(const double* v, long long n, int p, long long m, const long long* lags, int nlags, double* out) {
  std::vector<double> z((size_t)n);
  for (long long k = 0; k < m; ++k)
    for (int j = 0; j < p; ++j) {
      double mu = 0; for (long long i = 0; i < n; ++i) mu += at(v, n, p, i, j, k); mu /= (double)n;
      double s0 = 0; for (long long i = 0; i < n; ++i) { z[i] = at(v, n, p, i, j, k) - mu; s0 += z[i] * z[i]; }
      for (int a = 0; a < nlags; ++a) {
        const long long lag = lags[a];
        double s = 0;
        if (lag >= 0 && lag < n) for (long long t = 0; t + lag < n; ++t) s += z[t] * z[t + lag];
        out[(size_t)j + (size_t)p * ((size_t)a + (size_t)nlags * (size_t)k)] = (lag >= 0 && lag < n) ? s / s0 : NAN;
      }
    }
}
// changerate(c): stats.jl:19-39 — per-parameter and multivariate fraction of iterations whose value changed; NOT rounded
inline void chains_changerate(const double* v, long long n, int p, long long m, double* out) {
  std::vector<double> r((size_t)p, 0.0); double rmv = 0.0;
  for (long long k = 0; k < m; ++k)
    for (long long i = 1; i < n; ++i) {
      bool any = false;
      for (int j = 0; j < p; ++j) { const bool dx = at(v, n, p, i, j, k) != at(v, n, p, i - 1, j, k); r[j] += dx; any = any || dx; }
      rmv += any;
    }
  const double den = (double)m * (double)(n - 1);
  for (int j = 0; j < p; ++j) out[j] = r[j] / den;
  out[p] = rmv / den;
}
// largest eigenvalue of a symmetric matrix (cyclic Jacobi; p is a handful of monitored parameters)
inline double sym_eigmax(std::vector<double> A, int p) {
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0; for (int a = 0; a < p; ++a) for (int b = a + 1; b < p; ++b) off += A[a * p + b] * A[a * p + b];
    if (off < 1e-300) break;
    for (int a = 0; a < p; ++a)
      for (int b = a + 1; b < p; ++b) {
        if (A[a * p + b] == 0.0) continue;
        const double th = (A[b * p + b] - A[a * p + a]) / (2.0 * A[a * p + b]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < p; ++k) { const double ka = A[k * p + a], kb = A[k * p + b]; A[k * p + a] = c * ka - s * kb; A[k * p + b] = s * ka + c * kb; }
        for (int k = 0; k < p; ++k) { const double ak = A[a * p + k], bk = A[b * p + k]; A[a * p + k] = c * ak - s * bk; A[b * p + k] = s * ak + c * bk; }
      }
  }
  double mx = A[0]; for (int a = 1; a < p; ++a) mx = std::fmax(mx, A[a * p + a]);
  return mx;
}
// multivariate PSRF (gelmandiag.jl:49-55): isposdef(W) ? R_fixed + R_random_scale * eigmax(inv(cholfact(W)) * B) : NaN, with
// W = mean within-chain covariance, B = n cov(chain means) (p x p, row-major) — W^-1 B is similar to L^-1 B L^-T (W = L L')
inline double mpsrf_from_WB(const std::vector<double>& W, const std::vector<double>& B, long long n, long long m, int p) {
  std::vector<double> L((size_t)p * p, 0.0); bool pd = true;
  for (int a = 0; a < p && pd; ++a)
    for (int b = 0; b <= a; ++b) {
      double s = W[a * p + b]; for (int k = 0; k < b; ++k) s -= L[a * p + k] * L[b * p + k];
      if (a == b) { if (!(s > 0)) { pd = false; break; } L[a * p + a] = std::sqrt(s); } else L[a * p + b] = s / L[b * p + b];
    }
  if (!pd) return NAN;
  std::vector<double> Y((size_t)p * p), S((size_t)p * p);
  for (int c = 0; c < p; ++c) for (int a = 0; a < p; ++a) { double s = B[a * p + c]; for (int k = 0; k < a; ++k) s -= L[a * p + k] * Y[k * p + c]; Y[a * p + c] = s / L[a * p + a]; }   // Y = L^-1 B
  for (int r = 0; r < p; ++r) for (int a = 0; a < p; ++a) { double s = Y[r * p + a]; for (int k = 0; k < a; ++k) s -= L[a * p + k] * S[r * p + k]; S[r * p + a] = s / L[a * p + a]; }   // S = Y L^-T
  for (int a = 0; a < p; ++a) for (int b = a + 1; b < p; ++b) { const double t = 0.5 * (S[a * p + b] + S[b * p + a]); S[a * p + b] = S[b * p + a] = t; }
  return (double)(n - 1) / (double)n + (double)(m + 1) / ((double)m * (double)n) * sym_eigmax(S, p);
}
// gelmandiag(c; alpha, mpsrf, transform) on a materialised array: gelmandiag.jl:3-60.  codes[j] = 1: log scale, 2: logit scale (link(c)).
// out [(p + mpsrf) x 2] row-major, NOT rounded; the multivariate row is (R_fixed + R_random_scale eigmax(W^-1 B), NaN), NaN when W is
// not positive definite.  Returns 1 for fewer than 2 chains.
inline int chains_gelman(const double* v, long long n, int p, long long m, double alpha, const int* codes, int mpsrf, double* out) {
  if (m < 2) return 1;
  auto val = [&](long long i, int j, long long k) { const double x = at(v, n, p, i, j, k); return (codes && codes[j] == 1) ? std::log(x) : (codes && codes[j] == 2) ? std::log(x / (1.0 - x)) : x; };
  std::vector<double> mean((size_t)m * p), W((size_t)p * p, 0.0), s2((size_t)m * p);
  for (long long k = 0; k < m; ++k) {
    for (int j = 0; j < p; ++j) { double s = 0; for (long long i = 0; i < n; ++i) s += val(i, j, k); mean[k * p + j] = s / (double)n; }
    for (int a = 0; a < p; ++a)
      for (int b = a; b < p; ++b) {
        double s = 0; for (long long i = 0; i < n; ++i) s += (val(i, a, k) - mean[k * p + a]) * (val(i, b, k) - mean[k * p + b]);
        s /= (double)(n - 1);
        W[a * p + b] += s / (double)m; if (b != a) W[b * p + a] += s / (double)m;
        if (a == b) s2[k * p + a] = s;
      }
  }
  std::vector<double> gm((size_t)p, 0.0), B((size_t)p * p, 0.0);
  for (int j = 0; j < p; ++j) { for (long long k = 0; k < m; ++k) gm[j] += mean[k * p + j]; gm[j] /= (double)m; }
  for (int a = 0; a < p; ++a) for (int b = 0; b < p; ++b) {
    double s = 0; for (long long k = 0; k < m; ++k) s += (mean[k * p + a] - gm[a]) * (mean[k * p + b] - gm[b]);
    B[a * p + b] = (double)n * s / (double)(m - 1);
  }
  for (int j = 0; j < p; ++j) {   // the univariate columns through the same centred sums as the device path
    double c2 = 0; for (long long k = 0; k < m; ++k) c2 += s2[k * p + j]; c2 /= (double)m;
    double s[7] = {(double)m, 0, 0, 0, 0, 0, 0};
    for (long long k = 0; k < m; ++k) { const double d = mean[k * p + j] - gm[j], e = s2[k * p + j] - c2; s[1] += d; s[2] += d * d; s[3] += e; s[4] += e * e; s[5] += e * d; s[6] += e * d * d; }
    gelman_column((double)n, gm[j], c2, s, alpha, out + 2 * j);
  }
  if (mpsrf) {
    const double x = mpsrf_from_WB(W, B, n, m, p);
    out[2 * p] = x; out[2 * p + 1] = NAN;
  }
  return 0;
}

}  // namespace hostdiag
