// Compiled-in ODE right-hand sides: f(x, theta) and the vector-Jacobian products
//   vx[d]  = sum_d' g[d'] * d f_d'/d x_d        vth[k] = sum_d' g[d'] * d f_d'/d theta_k
// which is exactly what TF reverse-mode autodiff of the user's f_vec contributes to the gradient
// of the log-posterior (magi_v2.py:335).  One struct per model; D and P are compile-time.
#pragma once
#include "common.cuh"

struct Seir3 {  // vignette.ipynb:68-79: x = (E, I, R), S = 1 - E - I - R, theta = (beta, gamma, sigma)
  static constexpr int D = 3, P = 3, ID = MAGI_MODEL_SEIR3;
  __device__ static __forceinline__ void f(const double* x, const double* th, double* out) {
    const double S = 1.0 - (x[0] + x[1] + x[2]);
    out[0] = (th[0] * S * x[1]) - (th[2] * x[0]);
    out[1] = (th[2] * x[0]) - (th[1] * x[1]);
    out[2] = th[1] * x[1];
  }
  __device__ static __forceinline__ void vjp(const double* x, const double* th, const double* g, double* vx,
                                             double* vth) {
    const double E = x[0], I = x[1];
    const double S = 1.0 - (x[0] + x[1] + x[2]);
    const double b = th[0], gm = th[1], s = th[2];
    vx[0] = g[0] * (-b * I - s) + g[1] * s;
    vx[1] = g[0] * (b * S - b * I) - g[1] * gm + g[2] * gm;
    vx[2] = -g[0] * b * I;
    vth[0] = g[0] * S * I;
    vth[1] = (g[2] - g[1]) * I;
    vth[2] = (g[1] - g[0]) * E;
  }
};

struct Seir4 {  // S explicit: dS=-bSI, dE=bSI-sE, dI=sE-gI, dR=gI, theta = (beta, gamma, sigma)
  static constexpr int D = 4, P = 3, ID = MAGI_MODEL_SEIR4;
  __device__ static __forceinline__ void f(const double* x, const double* th, double* out) {
    const double bSI = th[0] * x[0] * x[2];
    out[0] = -bSI;
    out[1] = bSI - th[2] * x[1];
    out[2] = th[2] * x[1] - th[1] * x[2];
    out[3] = th[1] * x[2];
  }
  __device__ static __forceinline__ void vjp(const double* x, const double* th, const double* g, double* vx,
                                             double* vth) {
    const double S = x[0], E = x[1], I = x[2];
    const double b = th[0], gm = th[1], s = th[2];
    vx[0] = (g[1] - g[0]) * b * I;
    vx[1] = (g[2] - g[1]) * s;
    vx[2] = (g[1] - g[0]) * b * S + (g[3] - g[2]) * gm;
    vx[3] = 0.0;
    vth[0] = (g[1] - g[0]) * S * I;
    vth[1] = (g[3] - g[2]) * I;
    vth[2] = (g[2] - g[1]) * E;
  }
};

struct Sirw {  // test_magi_script.py:19-45: x = (S, I, R, W), theta = (beta, phi, xi, chi, kappa)
  static constexpr int D = 4, P = 5, ID = MAGI_MODEL_SIRW;
  __device__ static __forceinline__ void f(const double* x, const double* th, double* out) {
    const double S = x[0], I = x[1], R = x[2], W = x[3];
    const double beta = th[0], phi = th[1], xi = th[2], chi = th[3], kappa = th[4];
    out[0] = -beta * S * I + kappa * W;
    out[1] = beta * S * I - phi * I;
    out[2] = phi * I - xi * R + chi * I * W;
    out[3] = xi * R - chi * I * W - kappa * W;
  }
  __device__ static __forceinline__ void vjp(const double* x, const double* th, const double* g, double* vx,
                                             double* vth) {
    const double S = x[0], I = x[1], R = x[2], W = x[3];
    const double beta = th[0], phi = th[1], xi = th[2], chi = th[3], kappa = th[4];
    vx[0] = (g[1] - g[0]) * beta * I;
    vx[1] = (g[1] - g[0]) * beta * S + (g[2] - g[1]) * phi + (g[2] - g[3]) * chi * W;
    vx[2] = (g[3] - g[2]) * xi;
    vx[3] = (g[0] - g[3]) * kappa + (g[2] - g[3]) * chi * I;
    vth[0] = (g[1] - g[0]) * S * I;
    vth[1] = (g[2] - g[1]) * I;
    vth[2] = (g[3] - g[2]) * R;
    vth[3] = (g[2] - g[3]) * I * W;
    vth[4] = (g[0] - g[3]) * W;
  }
};

struct Lorenz96 {  // dx_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F, D = 10, theta = (F)
  static constexpr int D = 10, P = 1, ID = MAGI_MODEL_LORENZ96;
  __device__ static __forceinline__ void f(const double* x, const double* th, double* out) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int ip1 = (i + 1) % D, im1 = (i + D - 1) % D, im2 = (i + D - 2) % D;
      out[i] = (x[ip1] - x[im2]) * x[im1] - x[i] + th[0];
    }
  }
  __device__ static __forceinline__ void vjp(const double* x, const double* th, const double* g, double* vx,
                                             double* vth) {
    // f_i depends on x_{i+1}, x_{i-2}, x_{i-1}, x_i  =>  x_d appears in f_{d-1}, f_{d+2}, f_{d+1}, f_d
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int dm1 = (d + D - 1) % D, dm2 = (d + D - 2) % D, dp1 = (d + 1) % D, dp2 = (d + 2) % D;
      vx[d] = g[dm1] * x[dm2]                 // f_{d-1}: x_d is its x_{i+1}, factor x_{i-1} = x_{d-2}
              - g[dp2] * x[dp1]               // f_{d+2}: x_d is its x_{i-2}, factor x_{i-1} = x_{d+1}
              + g[dp1] * (x[dp2] - x[dm1])    // f_{d+1}: x_d is its x_{i-1}, factor x_{i+1} - x_{i-2}
              - g[d];
      s += g[d];
    }
    vth[0] = s;
  }
};
