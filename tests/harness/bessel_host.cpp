// CPU build of csrc/bessel.cuh so the Matern/Bessel device function can be checked against
// scipy / mpmath without a GPU (tests/test_bessel_cpu.py compiles this with g++).
#include "../../magi_v2_b200/csrc/bessel.cuh"
extern "C" int matern_lag_host(double nu, double phi1, double phi2, const double* l, int n, double* kap,
                               double* dkap, double* d2kap) {
  MaternConsts mc;
  if (matern_consts_init(nu, &mc) != 0) return -1;
  for (int i = 0; i < n; ++i) matern_lag(mc, phi1, phi2, l[i], kap[i], dkap[i], d2kap[i]);
  return 0;
}
