// FP64 modified Bessel function of the second kind, non-integer order, for the Matern blocks.
//
// Replaces scipy.special.kvp(v, u, n=0/1/2) at magi_v2.py:787 (AMOS zbesk on the CPU).  The three
// derivative orders the reference combines reduce analytically to three consecutive orders,
//   kappa(l)            =  c   u^nu     K_nu(u)
//   d kappa / d l       = -c a u^nu     K_{nu-1}(u)
//   d^2 kappa / d l^2   =  c a^2 u^(nu-1) [u K_{nu-2}(u) - K_{nu-1}(u)]
// with u = a l, a = sqrt(2 nu)/phi2, c = phi1 2^(1-nu)/Gamma(nu)  (from d/du[u^nu K_nu] = -u^nu K_{nu-1}).
// K_mu and K_{mu+1} (|mu| <= 1/2) come from Temme's series for u <= 2 and the continued fraction CF2
// (Steed's algorithm, Thompson-Barnett normalisation) for u > 2 -- both written from the papers' formulas, which
// are restated above the functions -- then the upward recurrence K_{k+1} = K_{k-1} + (2k/u) K_k.  The Gamma-function
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

// ---- K_mu(x), K_{mu+1}(x) for |mu| <= 1/2 ------------------------------------------------------------------------
// Written from the published formulas (no third-party code):
//
// x <= 2 -- Temme's series (N. M. Temme, J. Comput. Phys. 19 (1975) 324, eqs. (1.3)-(1.9)).  With L = ln(2/x),
//   sigma = mu L, G1 = [1/Gamma(1-mu) - 1/Gamma(1+mu)] / (2 mu), G2 = [1/Gamma(1-mu) + 1/Gamma(1+mu)] / 2:
//     f_0 = (pi mu / sin pi mu) [G1 cosh sigma + G2 L sinh(sigma)/sigma]
//     p_0 = e^sigma Gamma(1+mu) / 2,   q_0 = e^-sigma Gamma(1-mu) / 2,   c_0 = 1
//     f_k = (k f_{k-1} + p_{k-1} + q_{k-1}) / (k^2 - mu^2),  p_k = p_{k-1}/(k - mu),  q_k = q_{k-1}/(k + mu),
//     c_k = c_{k-1} (x^2/4) / k
//     K_mu = sum_k c_k f_k ,   K_{mu+1} = (2/x) sum_k c_k (p_k - k f_k)
//
// x > 2 -- the second continued fraction of the Bessel recurrences evaluated by Steed's forward algorithm with the
//   normalising sum of I. J. Thompson and A. R. Barnett, Comput. Phys. Commun. 47 (1987) 245, sec. 3:
//     CF2 = 1 / (b_1 + a_2 / (b_2 + a_3 / (b_3 + ...))),  b_n = 2 (x + n),  a_{n+1} = -[(n + 1/2)^2 - mu^2]
//     K_mu = sqrt(pi / 2x) e^-x / S,   S = 1 + sum_n Q_n dh_n,   Q_n = sum_{k<=n} C_k q_k
//     (C_1 = a_1 = 1/4 - mu^2, C_{k+1} = -a_{k+1} C_k / (k+1);  q_0 = 0, q_1 = 1, q_{k+1} = (q_{k-1} - b_k q_k)/a_{k+1};
//      dh_n the n-th increment of Steed's evaluation of CF2)
//     K_{mu+1} = K_mu (mu + x + 1/2 - a_1 CF2) / x
// The pair is returned scaled by e^x in the second case (scaled = true): the caller multiplies e^-x back in where
// it combines it with u^nu, so nothing underflows early.
struct TemmeTerms {
  double f, p, q, c;
};

MAGI_HD void bessel_k_temme_series(const MaternConsts& mc, double x, double& kmu, double& kmu1) {
  const double kTiny = 1.0e-16;
  const double mu = mc.mu, half_x = 0.5 * x, quarter_x2 = half_x * half_x;
  const double L = -log(half_x), sigma = mu * L;
  const double pi_mu = M_PI * mu;
  const double sinc_ratio = fabs(pi_mu) < kTiny ? 1.0 : pi_mu / sin(pi_mu);        // pi mu / sin(pi mu)
  const double sinhc = fabs(sigma) < kTiny ? 1.0 : sinh(sigma) / sigma;             // sinh(sigma) / sigma
  const double es = exp(sigma);
  TemmeTerms t;
  t.f = sinc_ratio * (mc.gam1 * cosh(sigma) + mc.gam2 * sinhc * L);
  t.p = 0.5 * es / mc.gampl;        // gampl = 1 / Gamma(1 + mu)
  t.q = 0.5 / (es * mc.gammi);      // gammi = 1 / Gamma(1 - mu)
  t.c = 1.0;
  double s0 = t.f, s1 = t.p;        // k = 0 terms of the two sums
  for (int k = 1; k <= 500; ++k) {
    const double kk = (double)k;
    t.f = (kk * t.f + t.p + t.q) / (kk * kk - mu * mu);
    t.p /= kk - mu;
    t.q /= kk + mu;
    t.c *= quarter_x2 / kk;
    const double term0 = t.c * t.f;
    s0 += term0;
    s1 += t.c * (t.p - kk * t.f);
    if (fabs(term0) < fabs(s0) * kTiny) break;
  }
  kmu = s0;
  kmu1 = s1 * (2.0 / x);
}

MAGI_HD void bessel_k_steed_cf2(const MaternConsts& mc, double x, double& kmu_scaled, double& kmu1_scaled) {
  const double kTiny = 1.0e-16;
  const double mu = mc.mu, a1 = 0.25 - mu * mu;
  // Steed's forward evaluation: D_n = 1 / (b_n + a_n D_{n-1}),  dh_n = (b_n D_n - 1) dh_{n-1},  h = sum dh_n
  double b = 2.0 * (1.0 + x);
  double D = 1.0 / b;
  double dh = D, h = D;
  // Thompson-Barnett sum
  double q_prev = 0.0, q_cur = 1.0;   // q_{k-1}, q_k
  double Ck = a1;                     // C_k
  double Q = a1;                      // Q_n
  double S = 1.0 + Q * dh;
  double an = -a1;                    // a_n (n = 1), becomes a_{n+1} below
  for (int n = 2; n <= 10000; ++n) {
    an -= 2.0 * (n - 1);              // a_n = -[(n - 1/2)^2 - mu^2]
    Ck = -an * Ck / n;
    const double q_next = (q_prev - b * q_cur) / an;
    q_prev = q_cur;
    q_cur = q_next;
    Q += Ck * q_next;
    b += 2.0;
    D = 1.0 / (b + an * D);
    dh = (b * D - 1.0) * dh;
    h += dh;
    const double dS = Q * dh;
    S += dS;
    if (fabs(dS / S) < kTiny) break;
  }
  kmu_scaled = sqrt(M_PI / (2.0 * x)) / S;
  kmu1_scaled = kmu_scaled * (mu + x + 0.5 - a1 * h) / x;
}

// K_mu(x) and K_{mu+1}(x), scaled by exp(x) when x > 2 (scaled = true on return), |mu| <= 1/2.
MAGI_HD void bessel_k_pair(const MaternConsts& mc, double x, double& kmu, double& kmu1, bool& scaled) {
  scaled = !(x < 2.0);
  if (scaled) bessel_k_steed_cf2(mc, x, kmu, kmu1);
  else bessel_k_temme_series(mc, x, kmu, kmu1);
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
