/* cd_oracle_c.c -- plain-C (OpenMP) restatement of the reference's CD-EKF and CD-KF filters.
 *
 * TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load it.  PARITY STATUS: parity unpinned at rel 1e-9 against real JAX/diffrax (see cd_oracle.py header);
 * this file is validated against cd_oracle.py and the reference-generated golden vectors in tests/.
 *
 * It follows the reference's arithmetic naively (dense full-matrix products, no symmetry or sparsity shortcuts), one
 * trajectory per OpenMP task -- the cost model of SURVEY.md 8(d):
 *   extended_kalman_filter   src/continuous_discrete_nonlinear_gaussian_ssm/inference_ekf.py:202-326
 *     _predict :46-148 (first / second order; second-order term is identically 0 for the registry drifts)
 *     _condition_on :153-199, psd_solve dynamax/utils/utils.py:202-207, TFP MVN.log_prob
 *   cdlgssm_filter           src/continuous_discrete_linear_gaussian_ssm/inference.py:555-632
 *   diffeqsolve              src/utils/diffrax_utils.py:150-163 (diffrax ConstantStepSize stepping, restated)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 64
#define MAXS 6

typedef struct {
  int S;
  double a[MAXS][MAXS];
  double b[MAXS];
} tableau;

static int get_tableau(int solver, tableau* t) {
  memset(t, 0, sizeof(*t));
  switch (solver) {
    case 0: t->S = 1; t->b[0] = 1.0; return 0;                                             /* euler */
    case 1: t->S = 2; t->a[1][0] = 1.0; t->b[0] = 0.5; t->b[1] = 0.5; return 0;            /* heun */
    case 2: t->S = 2; t->a[1][0] = 0.5; t->b[1] = 1.0; return 0;                           /* midpoint */
    case 3: t->S = 2; t->a[1][0] = 0.75; t->b[0] = 1.0 / 3.0; t->b[1] = 2.0 / 3.0; return 0; /* ralston */
    case 4: t->S = 3; t->a[1][0] = 0.5; t->a[2][1] = 0.75; t->b[0] = 2.0 / 9.0; t->b[1] = 1.0 / 3.0; t->b[2] = 4.0 / 9.0; return 0;
    case 5: t->S = 4; t->a[1][0] = 0.5; t->a[2][1] = 0.5; t->a[3][2] = 1.0;
            t->b[0] = 1.0 / 6.0; t->b[1] = 1.0 / 3.0; t->b[2] = 1.0 / 3.0; t->b[3] = 1.0 / 6.0; return 0; /* rk4 */
    case 6: t->S = 6;                                                                      /* dopri5 */
      t->a[1][0] = 1.0 / 5.0;
      t->a[2][0] = 3.0 / 40.0; t->a[2][1] = 9.0 / 40.0;
      t->a[3][0] = 44.0 / 45.0; t->a[3][1] = -56.0 / 15.0; t->a[3][2] = 32.0 / 9.0;
      t->a[4][0] = 19372.0 / 6561.0; t->a[4][1] = -25360.0 / 2187.0; t->a[4][2] = 64448.0 / 6561.0; t->a[4][3] = -212.0 / 729.0;
      t->a[5][0] = 9017.0 / 3168.0; t->a[5][1] = -355.0 / 33.0; t->a[5][2] = 46732.0 / 5247.0; t->a[5][3] = 49.0 / 176.0; t->a[5][4] = -5103.0 / 18656.0;
      t->b[0] = 35.0 / 384.0; t->b[2] = 500.0 / 1113.0; t->b[3] = 125.0 / 192.0; t->b[4] = -2187.0 / 6784.0; t->b[5] = 11.0 / 84.0;
      return 0;
  }
  return -1;
}

/* ---- drift registry (same ids / theta layout as include/cdk.h) ---- */
static void drift_f(int id, const double* th, int n, const double* x, double* f) {
  if (id == 0) {
    for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += th[i * n + k] * x[k]; f[i] = s + th[n * n + i]; }
  } else if (id == 1) {
    f[0] = th[0] * (x[1] - x[0]); f[1] = x[0] * (th[1] - x[2]) - x[1]; f[2] = x[0] * x[1] - th[2] * x[2];
  } else {
    for (int i = 0; i < n; ++i) {
      int ip = (i + 1) % n, im1 = (i + n - 1) % n, im2 = (i + n - 2) % n;
      f[i] = (x[ip] - x[im2]) * x[im1] - x[i] + th[0];
    }
  }
}
static void drift_jac(int id, const double* th, int n, const double* x, double* J) {
  if (id == 0) { memcpy(J, th, sizeof(double) * n * n); return; }
  memset(J, 0, sizeof(double) * n * n);
  if (id == 1) {
    J[0] = -th[0]; J[1] = th[0]; J[3] = th[1] - x[2]; J[4] = -1.0; J[5] = -x[0]; J[6] = x[1]; J[7] = x[0]; J[8] = -th[2];
  } else {
    for (int i = 0; i < n; ++i) {
      int ip = (i + 1) % n, im1 = (i + n - 1) % n, im2 = (i + n - 2) % n;
      J[i * n + ip] += x[im1]; J[i * n + im2] -= x[im1]; J[i * n + im1] += x[ip] - x[im2]; J[i * n + i] -= 1.0;
    }
  }
}

static void matmul(const double* A, const double* B, double* C, int r, int k, int c) { /* C[r,c] = A[r,k] B[k,c] */
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) { double s = 0; for (int q = 0; q < k; ++q) s += A[i * k + q] * B[q * c + j]; C[i * c + j] = s; }
}
static void matmul_nt(const double* A, const double* B, double* C, int r, int k, int c) { /* C[r,c] = A[r,k] B[c,k]^T */
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) { double s = 0; for (int q = 0; q < k; ++q) s += A[i * k + q] * B[j * k + q]; C[i * c + j] = s; }
}
static void cholesky(const double* A, double* L, int n, double boost) {
  memset(L, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j] + boost;
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    d = sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * n + j];
      for (int k = 0; k < j; ++k) v -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = v / d;
    }
  }
}
static void cho_solve(const double* L, int n, double* B, int c) { /* in place, B[n,c] */
  for (int col = 0; col < c; ++col) {
    for (int i = 0; i < n; ++i) { double v = B[i * c + col]; for (int k = 0; k < i; ++k) v -= L[i * n + k] * B[k * c + col]; B[i * c + col] = v / L[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double v = B[i * c + col]; for (int k = i + 1; k < n; ++k) v -= L[k * n + i] * B[k * c + col]; B[i * c + col] = v / L[i * n + i]; }
  }
}

typedef struct {
  int kind; /* 0 = EKF moments (m,P); 1 = KF pushforward (A,Q) */
  int n, drift_id;
  const double* th;  /* drift params or F */
  const double* lql;
} ode_ctx;

static void ode_rhs(const ode_ctx* c, const double* y, double* dy, double* work) {
  const int n = c->n, nn = n * n;
  if (c->kind == 0) {
    const double* m = y; const double* P = y + n;
    double* J = work; double* JP = work + nn; double* PJt = work + 2 * nn;
    drift_f(c->drift_id, c->th, n, m, dy);
    drift_jac(c->drift_id, c->th, n, m, J);
    matmul(J, P, JP, n, n, n);
    matmul_nt(P, J, PJt, n, n, n);
    for (int e = 0; e < nn; ++e) dy[n + e] = JP[e] + PJt[e] + c->lql[e];
  } else {
    const double* A = y; const double* Q = y + nn; const double* F = c->th;
    double* FQ = work; double* QFt = work + nn;
    matmul(F, A, dy, n, n, n);
    matmul(F, Q, FQ, n, n, n);
    matmul_nt(Q, F, QFt, n, n, n);
    for (int e = 0; e < nn; ++e) dy[nn + e] = FQ[e] + QFt[e] + c->lql[e];
  }
}

/* diffrax ConstantStepSize stepping (restated): returns 1 if max_steps was exceeded */
static int rk_solve(const ode_ctx* c, const tableau* tb, double* y, int S, double t0, double t1, double dt0, int max_steps,
                    double* ks, double* ys, double* acc, double* work, long long* nsub) {
  double tprev = t0, tnext = fmin(t0 + dt0, t1);
  int steps = 0;
  while (tprev < t1) {
    if (steps >= max_steps) { for (int e = 0; e < S; ++e) y[e] = NAN; return 1; }
    const double dt = tnext - tprev;
    memcpy(acc, y, sizeof(double) * S);
    for (int i = 0; i < tb->S; ++i) {
      memcpy(ys, y, sizeof(double) * S);
      for (int j = 0; j < i; ++j)
        if (tb->a[i][j] != 0.0) for (int e = 0; e < S; ++e) ys[e] += tb->a[i][j] * ks[j * S + e];
      ode_rhs(c, ys, ks + i * S, work);
      for (int e = 0; e < S; ++e) ks[i * S + e] *= dt;
      if (tb->b[i] != 0.0) for (int e = 0; e < S; ++e) acc[e] += tb->b[i] * ks[i * S + e];
    }
    memcpy(y, acc, sizeof(double) * S);
    ++steps;
    tprev = tnext;
    const double cand = tprev + dt0;
    tnext = cand > t1 - 1e-10 ? t1 : cand;
  }
  if (nsub) *nsub += steps;
  return 0;
}

/* update shared by KF / EKF with linear emission; returns the log-likelihood increment */
static double condition_on(int n, int m, const double* H, const double* d, const double* R, const double* yk,
                           double* mu, double* P, int num_iter, double* w) {
  double* HP = w; double* S = HP + m * n; double* Sl = S + m * m; double* Kt = Sl + m * m; double* SK = Kt + m * n;
  double* r = SK + m * n; double* Sb = r + m; double* z = Sb + m * m;
  double ll = 0.0;
  for (int it = 0; it < num_iter; ++it) {
    matmul(H, P, HP, m, n, n);
    matmul_nt(HP, H, S, m, n, m);
    for (int e = 0; e < m * m; ++e) S[e] += R[e];
    for (int a = 0; a < m; ++a) { double s = d[a]; for (int q = 0; q < n; ++q) s += H[a * n + q] * mu[q]; r[a] = yk[a] - s; }
    if (it == 0) {
      cholesky(S, Sl, m, 0.0);
      double quad = 0, logdet = 0;
      for (int i = 0; i < m; ++i) { double v = r[i]; for (int q = 0; q < i; ++q) v -= Sl[i * m + q] * z[q]; v /= Sl[i * m + i]; z[i] = v; quad += v * v; logdet += log(Sl[i * m + i]); }
      ll = -0.5 * quad - logdet - 0.5 * m * log(2.0 * M_PI);
    }
    for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) Sb[a * m + b] = 0.5 * (S[a * m + b] + S[b * m + a]);
    cholesky(Sb, Sl, m, 1e-9);
    memcpy(Kt, HP, sizeof(double) * m * n);
    cho_solve(Sl, m, Kt, n);
    matmul(S, Kt, SK, m, m, n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int a = 0; a < m; ++a) s += Kt[a * n + i] * SK[a * n + j]; P[i * n + j] -= s; }
    for (int i = 0; i < n; ++i) { double s = 0; for (int a = 0; a < m; ++a) s += Kt[a * n + i] * r[a]; mu[i] += s; }
  }
  for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) { double v = 0.5 * (P[i * n + j] + P[j * n + i]); P[i * n + j] = v; P[j * n + i] = v; }
  return ll;
}

/* algo: 0 = EKF (theta = drift params), 1 = KF (theta = F[n,n], bias b added after the pushforward).
 * Shared parameters, data batched [N,K,m], [N,K].  Outputs may be NULL.  Returns total substeps (for flop accounting). */
long long cdo_filter(int algo, long long N, int K, int n, int m, int drift_id, int solver, double dt0, double dt_final,
                     int max_steps, int num_iter, const double* Y, const double* T, const double* m0, const double* P0,
                     const double* theta, const double* bias, const double* Lm, const double* Qc, const double* H,
                     const double* d, const double* R, double* LL, double* FM, double* FP, double* PM, double* PP,
                     int num_threads) {
  tableau tb;
  if (get_tableau(solver, &tb) != 0 || n > MAXN || m > MAXN) return -1;
  const int nn = n * n;
  double* lql = (double*)malloc(sizeof(double) * nn);
  double* tmp = (double*)malloc(sizeof(double) * nn);
  matmul(Lm, Qc, tmp, n, n, n);
  matmul_nt(tmp, Lm, lql, n, n, n);
  free(tmp);
  long long total_sub = 0;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel reduction(+ : total_sub)
  {
    const int S = algo == 0 ? n + nn : 2 * nn;
    double* buf = (double*)malloc(sizeof(double) * (size_t)(S * (MAXS + 3) + 3 * nn + n + nn + 4 * m * n + 4 * m * m + 4 * m + nn + n));
    double* y = buf; double* ks = y + S; double* ys = ks + MAXS * S; double* acc = ys + S; double* work = acc + S;
    double* mu = work + 3 * nn; double* P = mu + n; double* w = P + nn;
    double* AP = w + 4 * m * n + 4 * m * m + 4 * m; double* mnew = AP + nn;
    ode_ctx c; c.kind = algo; c.n = n; c.drift_id = drift_id; c.th = theta; c.lql = lql;
#pragma omp for schedule(dynamic, 4)
    for (long long tr = 0; tr < N; ++tr) {
      memcpy(mu, m0, sizeof(double) * n);
      memcpy(P, P0, sizeof(double) * nn);
      double ll = 0.0;
      const double* Yt = Y + tr * (long long)K * m;
      const double* Tt = T + tr * (long long)K;
      for (int k = 0; k < K; ++k) {
        ll += condition_on(n, m, H, d, R, Yt + (long long)k * m, mu, P, algo == 0 ? num_iter : 1, w);
        const long long row = tr * (long long)K + k;
        if (FM) memcpy(FM + row * n, mu, sizeof(double) * n);
        if (FP) memcpy(FP + row * nn, P, sizeof(double) * nn);
        const double t0 = Tt[k], t1 = k + 1 < K ? Tt[k + 1] : Tt[k] + dt_final;
        if (algo == 0) {
          memcpy(y, mu, sizeof(double) * n);
          memcpy(y + n, P, sizeof(double) * nn);
          rk_solve(&c, &tb, y, S, t0, t1, dt0, max_steps, ks, ys, acc, work, &total_sub);
          memcpy(mu, y, sizeof(double) * n);
          memcpy(P, y + n, sizeof(double) * nn);
        } else {
          memset(y, 0, sizeof(double) * S);
          for (int i = 0; i < n; ++i) y[i * n + i] = 1.0;
          rk_solve(&c, &tb, y, S, t0, t1, dt0, max_steps, ks, ys, acc, work, &total_sub);
          const double* A = y; const double* Q = y + nn;
          matmul(A, P, AP, n, n, n);
          for (int i = 0; i < n; ++i) { double s = 0; for (int q = 0; q < n; ++q) s += A[i * n + q] * mu[q]; mnew[i] = s + (bias ? bias[i] : 0.0); }
          matmul_nt(AP, A, P, n, n, n);
          for (int e = 0; e < nn; ++e) P[e] += Q[e];
          memcpy(mu, mnew, sizeof(double) * n);
        }
        if (PM) memcpy(PM + row * n, mu, sizeof(double) * n);
        if (PP) memcpy(PP + row * nn, P, sizeof(double) * nn);
      }
      if (LL) LL[tr] = ll;
    }
    free(buf);
  }
  free(lql);
  return total_sub;
}

int cdo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
