/* C restatement of the reference's posterior-evaluation path: log-posterior + analytic gradient, leapfrog,
 * fixed-length HMC, NUTS, dual averaging, the annealing schedule.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a third, compiled implementation beside the numpy /
 * torch-autograd ones in oracle/magi_oracle.py.  It exists so that (i) whole NUTS runs at the vignette's size
 * (n = 161, tree depth 10, thousands of transitions) can be replayed on the CPU in seconds and compared with
 * the CUDA sampler, and (ii) bench.py's cpu_baseline has a compiled, chain-blocked CPU arm that is not dominated
 * by interpreter overhead.  Nothing under magi_v2_b200/ links or loads it.
 *
 * Follows /root/reference/magi_v2.py:
 *   log-posterior            :308-348  (formula: SURVEY.md A.2; gradient: A.3 -- (A + A^T) x for the quadratic
 *                                       forms, because the reference's C^-1 / K^-1 are not exactly symmetric)
 *   temperature schedule     :833-835, :855-856
 *   sampler stack            :357-371, :862-876 -> tensorflow-probability==0.24.0 (requirements.txt:8; not on
 *                            disk): SimpleLeapfrogIntegrator, NoUTurnSampler (multinomial sampling, generalised
 *                            U-turn criterion, max_tree_depth 10, max_energy_diff 1000),
 *                            DualAveragingStepSizeAdaptation (target 0.75, shrinkage 0.05, smoothing 10, decay 0.75)
 *   "parity unpinned" for everything that restates TFP.
 * Random numbers: Philox4x32-10 with the counters of oracle/magi_oracle.py::rng_normals / rng_uniform_pair.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp-simd -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MODEL_SEIR3 0
#define MODEL_SEIR4 1
#define MODEL_SIRW 2
#define MODEL_LORENZ96 3
#define MAXD 16
#define MAXP 8

typedef struct {
  int n, D, P, model, band;
  double beta;
  double *SC;   /* [D,n,n]  C^-1 + C^-T */
  double *SK;   /* [D,n,n]  K^-1 + K^-T */
  double *M;    /* [D,n,n]  m */
  double *mu, *y, *N_ds, *LB;
  unsigned char *mask;
} mo_problem;

static double *dup_d(const double *a, size_t k) {
  double *r = (double *)malloc(k * sizeof(double));
  memcpy(r, a, k * sizeof(double));
  return r;
}

/* band < 0: dense.  The matrices are taken as given (already banded by the caller, magi_v2.py:271-274);
 * `band` only lets the loops skip the zeros. */
mo_problem *mo_create(int n, int D, int P, int model, int band, const double *Cinv, const double *m,
                      const double *Kinv, const double *mu, const double *y, const unsigned char *mask,
                      const double *N_ds, double beta, const double *LB) {
  if (D > MAXD || P > MAXP) return NULL;
  mo_problem *h = (mo_problem *)calloc(1, sizeof(mo_problem));
  size_t nn = (size_t)n * n, tot = nn * D;
  h->n = n; h->D = D; h->P = P; h->model = model; h->band = band; h->beta = beta;
  h->M = dup_d(m, tot);
  h->SC = (double *)malloc(tot * sizeof(double));
  h->SK = (double *)malloc(tot * sizeof(double));
  for (int d = 0; d < D; ++d)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        size_t ij = d * nn + (size_t)i * n + j, ji = d * nn + (size_t)j * n + i;
        h->SC[ij] = Cinv[ij] + Cinv[ji];
        h->SK[ij] = Kinv[ij] + Kinv[ji];
      }
  h->mu = dup_d(mu, D); h->N_ds = dup_d(N_ds, D); h->LB = dup_d(LB, D);
  h->y = dup_d(y, (size_t)n * D);
  h->mask = (unsigned char *)malloc((size_t)n * D);
  memcpy(h->mask, mask, (size_t)n * D);
  return h;
}

void mo_destroy(mo_problem *h) {
  if (!h) return;
  free(h->SC); free(h->SK); free(h->M);
  free(h->mu); free(h->y); free(h->N_ds); free(h->LB); free(h->mask); free(h);
}

int mo_state_size(const mo_problem *h) { return h->n * h->D + h->D + h->P; }

/* ---- user ODE right-hand sides (vignette.ipynb:68-79; test_magi_script.py:19-45) ---- */
static void ode_f(int model, int D, const double *x, const double *th, double *f) {
  switch (model) {
    case MODEL_SEIR3: {
      double S = 1.0 - (x[0] + x[1] + x[2]);
      f[0] = th[0] * S * x[1] - th[2] * x[0];
      f[1] = th[2] * x[0] - th[1] * x[1];
      f[2] = th[1] * x[1];
    } break;
    case MODEL_SEIR4: {
      double inf = th[0] * x[0] * x[2];
      f[0] = -inf; f[1] = inf - th[2] * x[1]; f[2] = th[2] * x[1] - th[1] * x[2]; f[3] = th[1] * x[2];
    } break;
    case MODEL_SIRW: {
      double S = x[0], I = x[1], R = x[2], W = x[3];
      f[0] = -th[0] * S * I + th[4] * W;
      f[1] = th[0] * S * I - th[1] * I;
      f[2] = th[1] * I - th[2] * R + th[3] * I * W;
      f[3] = th[2] * R - th[3] * I * W - th[4] * W;
    } break;
    default:
      for (int i = 0; i < D; ++i)
        f[i] = (x[(i + 1) % D] - x[(i + D - 2) % D]) * x[(i + D - 1) % D] - x[i] + th[0];
  }
}

/* vx[d] += sum_d' g[d'] df_d'/dx_d ; vth[k] += sum_d' g[d'] df_d'/dtheta_k  (what reverse-mode autodiff of
 * f_vec contributes, magi_v2.py:335) */
static void ode_vjp(int model, int D, const double *x, const double *th, const double *g, double *vx, double *vth) {
  switch (model) {
    case MODEL_SEIR3: {
      double E = x[0], I = x[1], S = 1.0 - (x[0] + x[1] + x[2]);
      double b = th[0], gm = th[1], sg = th[2];
      /* f0 = b S I - sg E ; dS/dx_d = -1 */
      vx[0] = g[0] * (-b * I - sg) + g[1] * sg;
      vx[1] = g[0] * (b * S - b * I) + g[1] * (-gm) + g[2] * gm;
      vx[2] = g[0] * (-b * I);
      vth[0] += g[0] * S * I;
      vth[1] += -g[1] * I + g[2] * I;
      vth[2] += -g[0] * E + g[1] * E;
    } break;
    case MODEL_SEIR4: {
      double S = x[0], E = x[1], I = x[2], b = th[0], gm = th[1], sg = th[2];
      vx[0] = -b * I * g[0] + b * I * g[1];
      vx[1] = -sg * g[1] + sg * g[2];
      vx[2] = -b * S * g[0] + b * S * g[1] - gm * g[2] + gm * g[3];
      vx[3] = 0.0;
      vth[0] += S * I * (g[1] - g[0]);
      vth[1] += I * (g[3] - g[2]);
      vth[2] += E * (g[2] - g[1]);
    } break;
    case MODEL_SIRW: {
      double S = x[0], I = x[1], R = x[2], W = x[3];
      double be = th[0], ph = th[1], xi = th[2], ch = th[3], ka = th[4];
      vx[0] = -be * I * g[0] + be * I * g[1];
      vx[1] = -be * S * g[0] + (be * S - ph) * g[1] + (ph + ch * W) * g[2] - ch * W * g[3];
      vx[2] = -xi * g[2] + xi * g[3];
      vx[3] = ka * g[0] + ch * I * g[2] - (ch * I + ka) * g[3];
      vth[0] += S * I * (g[1] - g[0]);
      vth[1] += I * (g[2] - g[1]);
      vth[2] += R * (g[3] - g[2]);
      vth[3] += I * W * (g[2] - g[3]);
      vth[4] += W * (g[0] - g[3]);
    } break;
    default: {
      for (int j = 0; j < D; ++j) vx[j] = 0.0;
      for (int i = 0; i < D; ++i) {
        int ip1 = (i + 1) % D, im1 = (i + D - 1) % D, im2 = (i + D - 2) % D;
        vx[ip1] += g[i] * x[im1];
        vx[im2] -= g[i] * x[im1];
        vx[im1] += g[i] * (x[ip1] - x[im2]);
        vx[i] -= g[i];
        vth[0] += g[i];
      }
    }
  }
}

static double softplus(double z) { return z > 0 ? z + log1p(exp(-z)) : log1p(exp(z)); }
static double sigmoid(double z) { return z >= 0 ? 1.0 / (1.0 + exp(-z)) : exp(z) / (1.0 + exp(z)); }

static inline void band_range(int n, int band, int i, int *j0, int *j1) {
  if (band < 0) { *j0 = 0; *j1 = n; return; }
  *j0 = i - band < 0 ? 0 : i - band;
  *j1 = i + band + 1 > n ? n : i + band + 1;
}

/* One evaluation: z = [X (n*D, row-major [n,D]), sigma_sqs_pre (D), thetas_pre (P)] -> lp, g (same packing).
 * work: 6*n*D doubles. */
static double logpost_grad(const mo_problem *h, const double *z, double bt, double *g, double *work) {
  const int n = h->n, D = h->D, P = h->P;
  const size_t nn = (size_t)n * n;
  const double *X = z, *s = z + n * D, *tau = z + n * D + D;
  double *xc = work, *u = xc + n * D, *w = u + n * D, *r = w + n * D, *q = r + n * D, *gx = q + n * D; /* [D][n] */
  double sig2[MAXD], th[MAXP], SSE[MAXD], vth[MAXP];
  for (int d = 0; d < D; ++d) sig2[d] = softplus(s[d]) + h->LB[d];                       /* :318 */
  for (int k = 0; k < P; ++k) { th[k] = softplus(tau[k]); vth[k] = 0.0; }                /* :319 */
  for (int d = 0; d < D; ++d)
    for (int i = 0; i < n; ++i) xc[d * n + i] = X[i * D + d] - h->mu[d];                 /* :329 */
  /* x^T A x = 1/2 x^T (A + A^T) x: one pass over SC = C^-1 + C^-T gives both t1 and its gradient */
  double t1 = 0.0, t2 = 0.0;
  for (int d = 0; d < D; ++d) {                                                          /* :332, :336 */
    const double *SC = h->SC + d * nn, *M = h->M + d * nn, *x = xc + d * n;
    for (int i = 0; i < n; ++i) {
      int j0, j1; band_range(n, h->band, i, &j0, &j1);
      double b = 0.0, c = 0.0;
      const double *sc = SC + (size_t)i * n, *mi = M + (size_t)i * n;
#pragma omp simd reduction(+ : b, c)
      for (int j = j0; j < j1; ++j) { b += sc[j] * x[j]; c += mi[j] * x[j]; }
      t1 += 0.5 * x[i] * b;
      u[d * n + i] = b;
      w[d * n + i] = c;
    }
  }
  for (int i = 0; i < n; ++i) {                                                          /* :335-336 */
    double f[MAXD];
    ode_f(h->model, D, X + i * D, th, f);
    for (int d = 0; d < D; ++d) r[d * n + i] = f[d] - w[d * n + i];
  }
  for (int d = 0; d < D; ++d) {                                                          /* :337 */
    const double *SK = h->SK + d * nn, *rr = r + d * n;
    for (int i = 0; i < n; ++i) {
      int j0, j1; band_range(n, h->band, i, &j0, &j1);
      double b = 0.0;
      const double *sk = SK + (size_t)i * n;
#pragma omp simd reduction(+ : b)
      for (int j = j0; j < j1; ++j) b += sk[j] * rr[j];
      t2 += 0.5 * rr[i] * b;
      q[d * n + i] = b;                                                                  /* d t2 / d r */
    }
  }
  for (int d = 0; d < D; ++d) {                                                          /* u - m^T q, row by row */
    const double *M = h->M + d * nn, *qq = q + d * n;
    double *gd = gx + d * n;
    for (int i = 0; i < n; ++i) gd[i] = u[d * n + i];
    for (int i = 0; i < n; ++i) {
      int j0, j1; band_range(n, h->band, i, &j0, &j1);
      const double *mi = M + (size_t)i * n, qi = qq[i];
#pragma omp simd
      for (int j = j0; j < j1; ++j) gd[j] -= mi[j] * qi;
    }
  }
  double t3 = 0.0, t4 = 0.0;
  for (int d = 0; d < D; ++d) { SSE[d] = 0.0; t3 += h->N_ds[d] * log(2.0 * M_PI * sig2[d]); }   /* :340 */
  for (int i = 0; i < n; ++i) {
    double qi[MAXD], vx[MAXD];
    for (int d = 0; d < D; ++d) qi[d] = q[d * n + i];
    ode_vjp(h->model, D, X + i * D, th, qi, vx, vth);
    for (int d = 0; d < D; ++d) {
      double e = h->mask[i * D + d] ? X[i * D + d] - h->y[i * D + d] : 0.0;               /* :343-345 */
      SSE[d] += e * e;
      g[i * D + d] = bt * -0.5 * ((gx[d * n + i] + vx[d]) / h->beta + 2.0 * e / sig2[d]);
    }
  }
  double logJ = 0.0;
  for (int d = 0; d < D; ++d) {
    t4 += SSE[d] / sig2[d];
    logJ += s[d] - softplus(s[d]);                                                       /* :322 */
    double sg = sigmoid(s[d]);
    g[n * D + d] = bt * (-0.5 * (h->N_ds[d] / sig2[d] - SSE[d] / (sig2[d] * sig2[d])) * sg + (1.0 - sg));
  }
  for (int k = 0; k < P; ++k) {
    logJ += tau[k] - softplus(tau[k]);                                                   /* :323 */
    double sg = sigmoid(tau[k]);
    g[n * D + D + k] = bt * (-0.5 / h->beta * vth[k] * sg + (1.0 - sg));
  }
  return bt * (-0.5 * ((t1 + t2) / h->beta + t3 + t4) + logJ);                           /* :348 */
}

double mo_logpost_grad(const mo_problem *h, const double *z, double beta_temp, double *g) {
  double *work = (double *)malloc(sizeof(double) * 6 * h->n * h->D);
  double lp = logpost_grad(h, z, beta_temp, g, work);
  free(work);
  return lp;
}

/* R chains of ONE dataset, one after the other (bench.py cpu_baseline).  The image has
 * no libgomp, so host threads come from the caller: ctypes releases the GIL, bench.py runs one call per thread. */
void mo_logpost_grad_batch(const mo_problem *h, int R, const double *Z, const double *bt, double *lp, double *G) {
  const int S = mo_state_size(h);
  double *work = (double *)malloc(sizeof(double) * 6 * h->n * h->D);
  for (int c = 0; c < R; ++c) lp[c] = logpost_grad(h, Z + (size_t)c * S, bt[c], G + (size_t)c * S, work);
  free(work);
}

/* ---- Philox4x32-10 + Box-Muller, counters as oracle/magi_oracle.py ---- */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
             n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
static double u53(uint32_t hi, uint32_t lo) {
  uint64_t k = ((uint64_t)hi << 21) ^ ((uint64_t)lo >> 11);
  return ((double)k + 0.5) * 1.1102230246251565e-16;
}
static void rng_normals(uint64_t seed, uint32_t chain, uint32_t iter, int count, double *out) {
  for (int j = 0; 2 * j < count; ++j) {
    uint32_t c[4] = {(uint32_t)j, chain, iter, 0u};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
    double rad = sqrt(-2.0 * log(u1));
    out[2 * j] = rad * cos(2.0 * M_PI * u2);
    if (2 * j + 1 < count) out[2 * j + 1] = rad * sin(2.0 * M_PI * u2);
  }
}
static void rng_pair(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t purpose, uint32_t index, double *a,
                     double *b) {
  uint32_t c[4] = {index, chain, iter, purpose};
  philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  *a = u53(c[0], c[1]);
  *b = u53(c[2], c[3]);
}

void mo_rng_normals(uint64_t seed, uint32_t chain, uint32_t iter, int count, double *out) {
  rng_normals(seed, chain, iter, count, out);
}

static double schedule(double step, double min_temp) {                                    /* :833-835 */
  double v = 1.0 / log(step + 2.0);
  return v > min_temp ? v : min_temp;
}

/* ---- leapfrog (TFP SimpleLeapfrogIntegrator, identity mass) ---- */
void mo_leapfrog(const mo_problem *h, double *z, double *p, double eps, int n_steps, double beta_temp,
                 double *traj /* [n_steps, S] or NULL */) {
  const int S = mo_state_size(h);
  double *g = (double *)malloc(sizeof(double) * S), *work = (double *)malloc(sizeof(double) * 6 * h->n * h->D);
  logpost_grad(h, z, beta_temp, g, work);
  for (int t = 0; t < n_steps; ++t) {
    for (int i = 0; i < S; ++i) { p[i] += 0.5 * eps * g[i]; z[i] += eps * p[i]; }
    logpost_grad(h, z, beta_temp, g, work);
    for (int i = 0; i < S; ++i) p[i] += 0.5 * eps * g[i];
    if (traj) memcpy(traj + (size_t)t * S, z, sizeof(double) * S);
  }
  free(g); free(work);
}

/* ---- dual averaging ---- */
typedef struct { double eps, log_target, err, log_avg; int step; } da_state;
static void da_update(da_state *s, double acc, int num_adapt) {
  const double target = 0.75, shrink = 0.05, smooth = 10.0, decay = 0.75;
  if (s->step < num_adapt) {
    s->err += target - acc;
    double t = s->step + 1.0;
    double log_x = s->log_target - sqrt(t) * s->err / (shrink * (t + smooth));
    double eta = pow(t, -decay);
    s->log_avg = eta * log_x + (1.0 - eta) * s->log_avg;
    s->eps = exp(log_x);
    if (s->step + 1 == num_adapt) s->eps = exp(s->log_avg);
  }
  s->step += 1;
}

/* ---- NUTS ---- */
typedef struct { double *z, *p, *g, *rho, *pf; double logw, lp; int ok; } subtree;

typedef struct {
  const mo_problem *h;
  int S, dir;
  double eps, bt, H0, max_dE;
  uint64_t seed; uint32_t chain, iter;
  double sum_acc; int n_leaf, leaf_index, diverged;
  double run_w;          /* running log weight of the top-level subtree being built */
  double *sel_z, *sel_g; double sel_lp; int sel_valid;
  subtree *tmp;          /* one scratch subtree per level */
  double *work;
} nuts_ctx;

static double logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}
static double dot(const double *a, const double *b, int k) { double s = 0; for (int i = 0; i < k; ++i) s += a[i] * b[i]; return s; }
static int no_uturn(const double *rho, const double *pf, const double *pl, int S) {
  return dot(rho, pf, S) > 0.0 && dot(rho, pl, S) > 0.0;
}

/* 2^depth leaves continuing from (zc, pc, gc) in direction ctx->dir; result in *out */
static void build(nuts_ctx *c, const double *zc, const double *pc, const double *gc, int depth, subtree *out) {
  const int S = c->S;
  if (depth == 0) {
    const double e = c->dir * c->eps;
    for (int i = 0; i < S; ++i) { double ph = pc[i] + 0.5 * e * gc[i]; out->p[i] = ph; out->z[i] = zc[i] + e * ph; }
    double lpn = logpost_grad(c->h, out->z, c->bt, out->g, c->work);
    for (int i = 0; i < S; ++i) out->p[i] += 0.5 * e * out->g[i];
    double H = -lpn + 0.5 * dot(out->p, out->p, S);
    double dE = H - c->H0;
    if (!isfinite(dE)) dE = INFINITY;
    double a = dE > 0.0 ? exp(-dE) : 1.0;
    c->sum_acc += a; c->n_leaf += 1;
    double u_leaf, dummy;
    rng_pair(c->seed, c->chain, c->iter, 3u, (uint32_t)c->leaf_index, &u_leaf, &dummy);
    c->leaf_index += 1;
    int div = dE > c->max_dE;
    c->diverged |= div;
    memcpy(out->rho, out->p, sizeof(double) * S);
    memcpy(out->pf, out->p, sizeof(double) * S);
    out->logw = -dE; out->lp = lpn; out->ok = !div;
    /* uniform progressive selection inside the top-level subtree: leaf i replaces the running proposal with
       probability w_i / (w_0 + ... + w_i) */
    c->run_w = logaddexp(c->run_w, out->logw);
    if (log(u_leaf) < out->logw - c->run_w) {
      memcpy(c->sel_z, out->z, sizeof(double) * S);
      memcpy(c->sel_g, out->g, sizeof(double) * S);
      c->sel_lp = lpn; c->sel_valid = 1;
    }
    return;
  }
  build(c, zc, pc, gc, depth - 1, out);
  if (!out->ok) return;
  subtree *b = &c->tmp[depth - 1];
  build(c, out->z, out->p, out->g, depth - 1, b);
  for (int i = 0; i < S; ++i) out->rho[i] += b->rho[i];
  memcpy(out->z, b->z, sizeof(double) * S);
  memcpy(out->p, b->p, sizeof(double) * S);
  memcpy(out->g, b->g, sizeof(double) * S);
  out->logw = logaddexp(out->logw, b->logw);
  out->lp = b->lp;
  out->ok = b->ok && no_uturn(out->rho, out->pf, out->p, S);
}

static void subtree_alloc(subtree *t, int S) {
  t->z = (double *)malloc(sizeof(double) * 5 * S);
  t->p = t->z + S; t->g = t->p + S; t->rho = t->g + S; t->pf = t->rho + S;
}

/* The reference's sampler stack, one chain.  cached_lp = 0: the log-posterior / gradient at the current point are
 * re-evaluated at the new temperature at the start of a transition (what magi_v2_b200 does); 1: TFP's kernel
 * results carry the previous transition's target_log_prob / gradients, computed at the PREVIOUS temperature, into
 * the next one_step (SURVEY.md section 7 hard part 8).
 * fixed_bt: NaN -> schedule(step0 + it).  out_z may be NULL; out_tail [n_iter, D + P] (pre-activations). */
void mo_nuts_chain(const mo_problem *h, const double *z0, int n_iter, double eps0, uint64_t seed, uint32_t chain,
                   int num_adapt, double min_temp, int step0, double fixed_bt, int max_depth, int cached_lp,
                   double *out_z, double *out_tail, double *out_acc, double *out_eps, int *out_nleap,
                   int *out_depth, double *out_lp) {
  const int S = mo_state_size(h), nD = h->n * h->D, T = h->D + h->P;
  nuts_ctx c; memset(&c, 0, sizeof(c));
  c.h = h; c.S = S; c.seed = seed; c.chain = chain; c.max_dE = 1000.0;
  c.work = (double *)malloc(sizeof(double) * 6 * nD);
  c.tmp = (subtree *)malloc(sizeof(subtree) * (max_depth + 1));
  for (int d = 0; d <= max_depth; ++d) subtree_alloc(&c.tmp[d], S);
  subtree top; subtree_alloc(&top, S);
  double *buf = (double *)malloc(sizeof(double) * 12 * S);
  double *z = buf, *g0 = z + S, *p0 = g0 + S, *zl = p0 + S, *pl = zl + S, *gl = pl + S, *zr = gl + S, *pr = zr + S,
         *gr = pr + S, *rho = gr + S, *prop_z = rho + S, *prop_g = prop_z + S;
  c.sel_z = (double *)malloc(sizeof(double) * 2 * S); c.sel_g = c.sel_z + S;
  memcpy(z, z0, sizeof(double) * S);
  da_state da = {eps0, log(10.0 * eps0), 0.0, 0.0, 0};
  double lp_cache = 0.0; int have_cache = 0;
  for (int it = 0; it < n_iter; ++it) {
    double bt = isnan(fixed_bt) ? schedule((double)(step0 + it), min_temp) : fixed_bt;
    c.bt = bt; c.eps = da.eps; c.iter = (uint32_t)(step0 + it);
    double lp0;
    if (cached_lp && have_cache) lp0 = lp_cache;            /* g0 already holds the cached gradient */
    else {
      /* no carried values yet: TFP's bootstrap_results evaluates at schedule(0) (magi_v2.py:357-364); a run that
         continues at step0 > 0 starts from values of the previous step's temperature */
      double bt0 = bt;
      if (cached_lp && isnan(fixed_bt) && step0 + it > 0) bt0 = schedule((double)(step0 + it - 1), min_temp);
      lp0 = logpost_grad(h, z, bt0, g0, c.work);
    }
    rng_normals(seed, chain, c.iter, S, p0);
    c.H0 = -lp0 + 0.5 * dot(p0, p0, S);
    c.sum_acc = 0.0; c.n_leaf = 0; c.leaf_index = 0; c.diverged = 0;
    memcpy(zl, z, sizeof(double) * S); memcpy(zr, z, sizeof(double) * S);
    memcpy(pl, p0, sizeof(double) * S); memcpy(pr, p0, sizeof(double) * S);
    memcpy(gl, g0, sizeof(double) * S); memcpy(gr, g0, sizeof(double) * S);
    memcpy(rho, p0, sizeof(double) * S);
    memcpy(prop_z, z, sizeof(double) * S); memcpy(prop_g, g0, sizeof(double) * S);
    double prop_lp = lp0, logw = 0.0;
    int depth = 0;
    while (depth < max_depth) {
      double u_dir, u_acc;
      rng_pair(seed, chain, c.iter, 2u, (uint32_t)depth, &u_dir, &u_acc);
      c.dir = u_dir < 0.5 ? 1 : -1;
      c.run_w = -INFINITY; c.sel_valid = 0;
      if (c.dir > 0) build(&c, zr, pr, gr, depth, &top);
      else build(&c, zl, pl, gl, depth, &top);
      depth += 1;
      if (!top.ok) break;
      if (log(u_acc) < top.logw - logw && c.sel_valid) {      /* biased progressive sampling between trees */
        memcpy(prop_z, c.sel_z, sizeof(double) * S); memcpy(prop_g, c.sel_g, sizeof(double) * S);
        prop_lp = c.sel_lp;
      }
      logw = logaddexp(logw, top.logw);
      for (int i = 0; i < S; ++i) rho[i] += top.rho[i];
      if (c.dir > 0) { memcpy(zr, top.z, sizeof(double) * S); memcpy(pr, top.p, sizeof(double) * S); memcpy(gr, top.g, sizeof(double) * S); }
      else { memcpy(zl, top.z, sizeof(double) * S); memcpy(pl, top.p, sizeof(double) * S); memcpy(gl, top.g, sizeof(double) * S); }
      if (!no_uturn(rho, pl, pr, S)) break;
    }
    double acc = c.sum_acc / (c.n_leaf > 0 ? c.n_leaf : 1);
    memcpy(z, prop_z, sizeof(double) * S); memcpy(g0, prop_g, sizeof(double) * S);
    lp_cache = prop_lp; have_cache = 1;
    if (out_eps) out_eps[it] = da.eps;
    da_update(&da, acc, num_adapt);
    if (out_z) memcpy(out_z + (size_t)it * S, z, sizeof(double) * S);
    if (out_tail) memcpy(out_tail + (size_t)it * T, z + nD, sizeof(double) * T);
    if (out_acc) out_acc[it] = acc;
    if (out_nleap) out_nleap[it] = c.n_leaf;
    if (out_depth) out_depth[it] = depth;
    if (out_lp) out_lp[it] = prop_lp;
  }
  for (int d = 0; d <= max_depth; ++d) free(c.tmp[d].z);
  free(c.tmp); free(top.z); free(buf); free(c.sel_z); free(c.work);
}

/* Fixed-length HMC with the same glue (oracle/magi_oracle.py::hmc_chain). */
void mo_hmc_chain(const mo_problem *h, const double *z0, int n_iter, int n_leapfrog, double eps0, uint64_t seed,
                  uint32_t chain, int num_adapt, double min_temp, int step0, double fixed_bt, double *out_z,
                  double *out_tail, double *out_acc, double *out_eps) {
  const int S = mo_state_size(h), nD = h->n * h->D, T = h->D + h->P;
  double *buf = (double *)malloc(sizeof(double) * 5 * S), *work = (double *)malloc(sizeof(double) * 6 * nD);
  double *z = buf, *z1 = z + S, *p = z1 + S, *g = p + S, *p0 = g + S;
  memcpy(z, z0, sizeof(double) * S);
  da_state da = {eps0, log(10.0 * eps0), 0.0, 0.0, 0};
  for (int it = 0; it < n_iter; ++it) {
    double bt = isnan(fixed_bt) ? schedule((double)(step0 + it), min_temp) : fixed_bt;
    double eps = da.eps;
    rng_normals(seed, chain, (uint32_t)(step0 + it), S, p0);
    memcpy(p, p0, sizeof(double) * S); memcpy(z1, z, sizeof(double) * S);
    double lp0 = logpost_grad(h, z1, bt, g, work), lp1 = lp0;
    for (int t = 0; t < n_leapfrog; ++t) {
      for (int i = 0; i < S; ++i) { p[i] += 0.5 * eps * g[i]; z1[i] += eps * p[i]; }
      lp1 = logpost_grad(h, z1, bt, g, work);
      for (int i = 0; i < S; ++i) p[i] += 0.5 * eps * g[i];
    }
    double dH = (-lp1 + 0.5 * dot(p, p, S)) - (-lp0 + 0.5 * dot(p0, p0, S));
    double acc = !isfinite(dH) ? 0.0 : (dH > 0.0 ? exp(-dH) : 1.0);
    double u, dummy;
    rng_pair(seed, chain, (uint32_t)(step0 + it), 1u, 0u, &u, &dummy);
    if (u < acc) memcpy(z, z1, sizeof(double) * S);
    if (out_eps) out_eps[it] = eps;
    da_update(&da, acc, num_adapt);
    if (out_z) memcpy(out_z + (size_t)it * S, z, sizeof(double) * S);
    if (out_tail) memcpy(out_tail + (size_t)it * T, z + nD, sizeof(double) * T);
    if (out_acc) out_acc[it] = acc;
  }
  free(buf); free(work);
}
