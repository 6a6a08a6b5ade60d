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

static double ulp_of(double x) { int e; frexp(x, &e); return ldexp(1.0, e - 53); }

int main() {
  std::mt19937_64 g(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double worst_log = 0, worst_exp = 0, arg_log = 0, arg_exp = 0;
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
  for (int n = 0; n < 20000000; ++n) {
    const double u = U(g);
    chk_log(u > 0 ? u : 0.5);                       // log of a uniform draw
    chk_log(1.0 + exp(40.0 * (u - 0.5)));           // log(1 + e^eta)
    chk_log(ldexp(0.5 + u, (int)(g() % 200) - 100));
    chk_log(1.0 + (u - 0.5) * 0.03);                // around 1, across the intervals that touch it
    chk_log(1.0 + (u - 0.5) * 1e-6);
    chk_exp(1400.0 * (u - 0.5));
    chk_exp(4.0 * (u - 0.5));
    chk_exp((u - 0.5) * 1e-3);
  }
  chk_log(1.0); chk_log(0.6875); chk_log(1.375); chk_log(2.0); chk_log(0.5); chk_exp(0.0);
  printf("tlog: max error %.3f ulp at %.17g\n", worst_log, arg_log);
  printf("texp: max error %.3f ulp at %.17g\n", worst_exp, arg_exp);
  return (worst_log <= 1.5 && worst_exp <= 1.0) ? 0 : 1;
}
