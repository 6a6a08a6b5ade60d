// check_fasttab.cpp — host build of the table-driven log / exp (mamba.jl_b200/csrc/fasttab_fn.cuh + fasttab.cuh) against long double.
// Build and run: g++ -O2 -std=c++17 -ffp-contract=off -mfma -I mamba.jl_b200/csrc tools/check_fasttab.cpp -o tools/bin/check_fasttab && tools/bin/check_fasttab
// Prints the largest error in ulps over random and structured arguments (the CPU suite runs it: tests/test_fastmath_tables.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#define __device__
#include "fasttab.cuh"
#include "fasttab_fn.cuh"

// the arithmetic of fast_sqrt (fastfn.cuh) with a float-precision reciprocal square root standing in for MUFU.RSQ64H
static double fast_sqrt_host(double x) {
  double y = (double)(float)(1.0 / std::sqrt(x));
  const double hx = 0.5 * x;
  y = __builtin_fma(y, __builtin_fma(-hx * y, y, 0.5), y);
  y = __builtin_fma(y, __builtin_fma(-hx * y, y, 0.5), y);
  double sq = x * y;
  sq = __builtin_fma(__builtin_fma(-sq, sq, x), 0.5 * y, sq);
  return x > 0.0 ? sq : 0.0;
}

static double ulp_of(double x) { int e; frexp(x, &e); return ldexp(1.0, e - 53); }

int main(int argc, char** argv) {
  const long n_samples = argc > 1 ? atol(argv[1]) : 20000000;
  std::mt19937_64 g(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double worst_log = 0, worst_exp = 0, arg_log = 0, arg_exp = 0, worst_sqrt = 0;
  auto chk_log = [&](double x) {
    const double got = mcu::tab::tlog(x, mcu::kLogTabG);
    const long double ref = logl((long double)x);
    const double err = ref == 0 ? fabs(got) : (double)fabsl(((long double)got - ref)) / ulp_of((double)ref);
    if (err > worst_log) { worst_log = err; arg_log = x; }
  };
  auto chk_exp = [&](double x) {
    const double got = mcu::tab::texp(x, mcu::kExpTabG);
    const long double ref = expl((long double)x);
    const double err = (double)fabsl(((long double)got - ref)) / ulp_of((double)ref);
    if (err > worst_exp) { worst_exp = err; arg_exp = x; }
  };
  for (long n = 0; n < n_samples; ++n) {
    const double u = U(g);
    chk_log(u > 0 ? u : 0.5);                       // log of a uniform draw
    chk_log(1.0 + exp(40.0 * (u - 0.5)));           // log(1 + e^eta)
    chk_log(ldexp(0.5 + u, (int)(g() % 200) - 100));
    chk_log(1.0 + (u - 0.5) * 0.03);                // around 1, across the intervals that touch it
    chk_log(1.0 + (u - 0.5) * 1e-6);
    chk_exp(1400.0 * (u - 0.5));
    chk_exp(4.0 * (u - 0.5));
    chk_exp((u - 0.5) * 1e-3);
    const double xs = -2.0 * log(1.0 - u) * (n % 3 == 0 ? 1e-12 : 1.0);   // the Box-Muller radius argument
    if (xs > 0) { const double e = fabs(fast_sqrt_host(xs) - sqrt(xs)) / ulp_of(sqrt(xs)); if (e > worst_sqrt) worst_sqrt = e; }
  }
  chk_log(1.0); chk_log(0.6875); chk_log(1.375); chk_log(2.0); chk_log(0.5); chk_exp(0.0);
  printf("tlog: max error %.3f ulp at %.17g\n", worst_log, arg_log);
  printf("texp: max error %.3f ulp at %.17g\n", worst_exp, arg_exp);
  printf("fast_sqrt arithmetic: max error %.3f ulp\n", worst_sqrt);
  return (worst_log <= 1.5 && worst_exp <= 1.0 && worst_sqrt <= 1.0) ? 0 : 1;
}
