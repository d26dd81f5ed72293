/* Accuracy of dragonfly_b200/csrc/exp_nonpos.h against glibc's exp (long double reference) on the host.
 * gcc -O2 -mfma -o /tmp/check_exp tools/check_exp.c -lm && /tmp/check_exp */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../dragonfly_b200/csrc/exp_nonpos.h"

static double ulp_of(double y) {
  int e;
  frexp(y, &e);
  return ldexp(1.0, e - 53);
}

int main(void) {
  double worst = 0.0, worst_x = 0.0, worst_libm = 0.0;
  srand48(12345);
  const long n = 10000000;
  for (long i = 0; i < n; i++) {
    double x;
    const int sel = (int)(i % 4);
    if (sel == 0) x = -drand48() * 1.0;
    else if (sel == 1) x = -drand48() * 40.0;
    else if (sel == 2) x = -drand48() * 700.0;
    else x = -ldexp(drand48(), -(int)(drand48() * 60));
    const long double ref = expl((long double)x);
    const double got = dfb_exp_nonpos(x);
    const double lib = exp(x);
    const double u = ulp_of((double)ref);
    const double err = fabs((double)(((long double)got - ref) / u));
    const double errl = fabs((double)(((long double)lib - ref) / u));
    if (err > worst) { worst = err; worst_x = x; }
    if (errl > worst_libm) worst_libm = errl;
  }
  printf("max error: dfb_exp_nonpos %.4f ulp (at x = %.17g), glibc exp %.4f ulp\n", worst, worst_x, worst_libm);
  printf("edge: exp(0) = %.17g, exp(-706.9) = %.6g (ref %.6g), exp(-708) = %g, exp(nan) = %g\n", dfb_exp_nonpos(0.0),
         dfb_exp_nonpos(-706.9), exp(-706.9), dfb_exp_nonpos(-708.0), dfb_exp_nonpos(NAN));
  return worst <= 1.0 ? 0 : 1;
}
