// FP64 modified Bessel function of the second kind, non-integer order, for the Matern blocks.
//
// Replaces scipy.special.kvp(v, u, n=0/1/2) at magi_v2.py:787 (AMOS zbesk on the CPU).  The three
// derivative orders the reference combines reduce analytically to three consecutive orders,
//   kappa(l)            =  c   u^nu     K_nu(u)
//   d kappa / d l       = -c a u^nu     K_{nu-1}(u)
//   d^2 kappa / d l^2   =  c a^2 u^(nu-1) [u K_{nu-2}(u) - K_{nu-1}(u)]
// with u = a l, a = sqrt(2 nu)/phi2, c = phi1 2^(1-nu)/Gamma(nu)  (from d/du[u^nu K_nu] = -u^nu K_{nu-1}).
// K_mu and K_{mu+1} (|mu| <= 1/2) come from Temme's series for u <= 2 and Steed's continued
// fraction CF2 for u > 2 (Temme 1975; Press et al., "Numerical Recipes" sec. 6.7 describes the
// method), then the upward recurrence K_{k+1} = K_{k-1} + (2k/u) K_k.  The Gamma-function
// constants depend only on nu and are computed once on the host in extended precision.
//
// Compiles as plain C++ too (tests/ builds it with g++ to check it against scipy/mpmath on the CPU).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MAGI_HD __host__ __device__ __forceinline__
#else
#define MAGI_HD inline
#endif

struct MaternConsts {
  double nu;      // smoothness
  double mu;      // nu - nl, |mu| <= 1/2
  int nl;         // number of upward recurrences from mu to nu
  double gam1;    // (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
  double gam2;    // (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2
  double gampl;   // 1/Gamma(1+mu)
  double gammi;   // 1/Gamma(1-mu)
  double cnorm;   // 2^(1-nu) / Gamma(nu)
  double sq2nu;   // sqrt(2 nu)
};

// Host-only: fill the constants for a given nu (> 1).  Extended precision so that the
// cancellation in gam1 costs nothing in double.
inline int matern_consts_init(double nu, MaternConsts* mc) {
  if (!(nu > 1.0) || !(nu < 170.0)) return -1;
  const int nl = (int)(nu + 0.5);
  const long double mu = (long double)nu - nl;
  const long double rp = 1.0L / tgammal(1.0L + mu), rm = 1.0L / tgammal(1.0L - mu);
  long double g1;
  if (fabsl(mu) < 1e-4L) {
    // (rm - rp)/(2 mu) = -(gamma_E + c4 mu^2 + ...):  1/Gamma(1+z) = 1 + gamma z + c3 z^2 + c4 z^3 ...
    const long double gE = 0.577215664901532860606512090082402431L, c4 = -0.04200263503409523552900393487542981871L;
    g1 = -(gE + c4 * mu * mu);
  } else {
    g1 = (rm - rp) / (2.0L * mu);
  }
  mc->nu = nu;
  mc->mu = (double)mu;
  mc->nl = nl;
  mc->gam1 = (double)g1;
  mc->gam2 = (double)((rm + rp) / 2.0L);
  mc->gampl = (double)rp;
  mc->gammi = (double)rm;
  mc->cnorm = (double)(powl(2.0L, 1.0L - (long double)nu) / tgammal((long double)nu));
  mc->sq2nu = (double)sqrtl(2.0L * (long double)nu);
  return 0;
}

// K_mu(x) and K_{mu+1}(x), scaled by exp(x) when x > 2 (scaled = true on return), |mu| <= 1/2.
MAGI_HD void bessel_k_pair(const MaternConsts& mc, double x, double& kmu, double& kmu1, bool& scaled) {
  const double EPS = 1.0e-16;
  const int MAXIT = 10000;
  const double xmu = mc.mu, xmu2 = xmu * xmu;
  if (x < 2.0) {
    const double x2 = 0.5 * x;
    const double pimu = M_PI * xmu;
    const double fact = fabs(pimu) < EPS ? 1.0 : pimu / sin(pimu);
    double d = -log(x2);
    double e = xmu * d;
    const double fact2 = fabs(e) < EPS ? 1.0 : sinh(e) / e;
    double ff = fact * (mc.gam1 * cosh(e) + mc.gam2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / mc.gampl;
    double q = 0.5 / (e * mc.gammi);
    double c = 1.0;
    d = x2 * x2;
    double sum1 = p;
    for (int i = 1; i <= MAXIT; ++i) {
      ff = (i * ff + p + q) / (i * (double)i - xmu2);
      c *= d / i;
      p /= (i - xmu);
      q /= (i + xmu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    kmu = sum;
    kmu1 = sum1 * (2.0 / x);
    scaled = false;
  } else {
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - xmu2;
    double q = a1, c = a1;
    double a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= MAXIT; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels / s) < EPS) break;
    }
    h = a1 * h;
    kmu = sqrt(M_PI / (2.0 * x)) / s;  // times exp(-x)
    kmu1 = kmu * (xmu + x + 0.5 - h) / x;
    scaled = true;
  }
}

// The three Matern quantities at lag l > 0 for hyper-parameters (phi1, phi2):
//   kap = kappa(l), dkap = d kappa/d l (< 0), d2kap = d^2 kappa / d l^2.
MAGI_HD void matern_lag(const MaternConsts& mc, double phi1, double phi2, double l, double& kap, double& dkap,
                        double& d2kap) {
  const double a = mc.sq2nu / phi2;
  const double u = a * l;
  double k0, k1;
  bool scaled;
  bessel_k_pair(mc, u, k0, k1, scaled);  // orders mu, mu+1
  // walk to orders nu-2, nu-1, nu
  double km2, km1, kn;
  const double two_over_u = 2.0 / u;
  if (mc.nl >= 2) {
    double lo = k0, hi = k1;  // orders mu + t, mu + t + 1
    for (int t = 0; t < mc.nl - 2; ++t) {
      const double nx = lo + (mc.mu + t + 1) * two_over_u * hi;
      lo = hi;
      hi = nx;
    }
    km2 = lo;
    km1 = hi;
    kn = lo + (mc.nu - 1.0) * two_over_u * hi;
  } else {  // nl == 1: nu - 2 = mu - 1, K_{mu-1} = K_{mu+1} - (2 mu/u) K_mu
    km2 = k1 - mc.mu * two_over_u * k0;
    km1 = k0;
    kn = k1;
  }
  const double sc = scaled ? exp(-u) : 1.0;
  const double c = phi1 * mc.cnorm;
  const double un = pow(u, mc.nu);  // u^nu
  const double unm1 = un / u;
  kap = c * un * kn * sc;
  dkap = -c * a * un * km1 * sc;
  d2kap = c * a * a * unm1 * (u * km2 - km1) * sc;
}
