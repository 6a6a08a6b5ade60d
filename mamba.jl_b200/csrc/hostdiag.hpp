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
#include <algorithm>
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


// ---- post-processing of a materialised chain array (ModelChains.value: [n iterations x p parameters x m chains], column-major,
//      iteration fastest) — src/output/stats.jl:3-83 and the multivariate PSRF of src/output/gelmandiag.jl:49-55 -----------------
inline double at(const double* v, long long n, int p, long long i, int j, long long k) { return v[(size_t)i + (size_t)n * ((size_t)j + (size_t)p * (size_t)k)]; }

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
// gelmandiag(c; alpha, mpsrf, transform) on a materialised array: gelmandiag.jl:3-60.  codes[j] = 1: log scale (link(c)).
// out [(p + mpsrf) x 2] row-major, NOT rounded; the multivariate row is (R_fixed + R_random_scale eigmax(W^-1 B), NaN), NaN when W is
// not positive definite.  Returns 1 for fewer than 2 chains.
inline int chains_gelman(const double* v, long long n, int p, long long m, double alpha, const int* codes, int mpsrf, double* out) {
  if (m < 2) return 1;
  auto val = [&](long long i, int j, long long k) { const double x = at(v, n, p, i, j, k); return (codes && codes[j] == 1) ? std::log(x) : x; };
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
    // isposdef(W) ? R_fixed + R_random_scale * eigmax(inv(cholfact(W)) * B) : NaN — W^-1 B is similar to L^-1 B L^-T (W = L L')
    std::vector<double> L((size_t)p * p, 0.0); bool pd = true;
    for (int a = 0; a < p && pd; ++a)
      for (int b = 0; b <= a; ++b) {
        double s = W[a * p + b]; for (int k = 0; k < b; ++k) s -= L[a * p + k] * L[b * p + k];
        if (a == b) { if (!(s > 0)) { pd = false; break; } L[a * p + a] = std::sqrt(s); } else L[a * p + b] = s / L[b * p + b];
      }
    double x = NAN;
    if (pd) {
      std::vector<double> Y((size_t)p * p), S((size_t)p * p);
      for (int c = 0; c < p; ++c) for (int a = 0; a < p; ++a) { double s = B[a * p + c]; for (int k = 0; k < a; ++k) s -= L[a * p + k] * Y[k * p + c]; Y[a * p + c] = s / L[a * p + a]; }   // Y = L^-1 B
      for (int r = 0; r < p; ++r) for (int a = 0; a < p; ++a) { double s = Y[r * p + a]; for (int k = 0; k < a; ++k) s -= L[a * p + k] * S[r * p + k]; S[r * p + a] = s / L[a * p + a]; }   // S = Y L^-T
      for (int a = 0; a < p; ++a) for (int b = a + 1; b < p; ++b) { const double t = 0.5 * (S[a * p + b] + S[b * p + a]); S[a * p + b] = S[b * p + a] = t; }
      x = (double)(n - 1) / (double)n + (double)(m + 1) / ((double)m * (double)n) * sym_eigmax(S, p);
    }
    out[2 * p] = x; out[2 * p + 1] = NAN;
  }
  return 0;
}

}  // namespace hostdiag
