// cdk_small.cu -- register-resident CD-EKF for tiny state dimensions: ONE THREAD PER TRAJECTORY.
//
// Replaces extended_kalman_filter (src/continuous_discrete_nonlinear_gaussian_ssm/inference_ekf.py:202-326) with its
// _predict (:46-148, moment ODE dm = f(m), dP = F P + P F^T + L Qc L^T) and _condition_on (:153-199) for the
// registry drifts whose whole filter state fits in registers (Lorenz-63: m[3] + symmetric P[6]).
//
// B200 mapping (BASELINE config 3: N = 65,536, K = 1,000, n = 3, m = 1):
//  * the state never leaves registers; the per-step outputs stream straight to HBM, observations/time stamps are
//    prefetched one gap ahead so the dependent global-load latency is hidden behind ~4.5 RK substeps of FP64 math;
//  * irregular gaps give every lane its own substep count q_k.  Instead of running each gap to the warp-wide maximum
//    (what jax.vmap does to diffrax's while_loop), lanes run a FLATTENED substep stream: every loop iteration is one
//    RK substep for the whole warp, and lanes whose gap just ended do the (cheap) measurement update under a
//    predicate.  Warp efficiency is ~(q*c_step)/(q*c_step + c_update) instead of mean(q)/max(q);
//  * 64-thread CTAs so that 65,536 trajectories = 1,024 CTAs fit in ONE wave at 7 CTAs/SM (148*7 = 1,036), which
//    needs <= 144 registers/thread: model constants live in shared memory, not registers.
#include "cdk_common.cuh"

namespace cdk {
namespace {

// ---- symmetric packed storage: upper triangle, row-major ----------------------------------------------------------
template <int NX>
__host__ __device__ constexpr int pidx(int i, int j) {
  return i <= j ? i * NX - (i * (i - 1)) / 2 + (j - i) : j * NX - (j * (j - 1)) / 2 + (i - j);
}

template <typename T, int NX>
struct St {
  static constexpr int NP = NX * (NX + 1) / 2;
  T m[NX];
  T P[NP];
};

// ---- drift policies: f(x), G = J(x) P (P symmetric packed) ---------------------------------------------------------
struct DriftL63 {
  static constexpr int NX = 3;
  static constexpr int NTHETA = 3;
  template <typename T>
  __device__ __forceinline__ static void f(const T* th, const T (&x)[3], T (&o)[3]) {
    // LearnableLorenz63.f, cdnlgssm_utils.py:77-83
    o[0] = th[0] * (x[1] - x[0]);
    o[1] = x[0] * (th[1] - x[2]) - x[1];
    o[2] = x[0] * x[1] - th[2] * x[2];
  }
  template <typename T>
  __device__ __forceinline__ static void jp(const T* th, const T (&x)[3], const T (&P)[6], T (&G)[3][3]) {
    // J = [[-s, s, 0], [r - z, -1, -x], [y, x, -b]]  (jacfwd(f), inference_ekf.py:95)
    const T rz = th[1] - x[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const T p0 = P[pidx<3>(0, j)], p1 = P[pidx<3>(1, j)], p2 = P[pidx<3>(2, j)];
      G[0][j] = th[0] * (p1 - p0);
      G[1][j] = rz * p0 - p1 - x[0] * p2;
      G[2][j] = x[1] * p0 + x[0] * p1 - th[2] * p2;
    }
  }
};

template <int N_>
struct DriftLinear {
  static constexpr int NX = N_;
  static constexpr int NTHETA = N_ * N_ + N_;
  template <typename T>
  __device__ __forceinline__ static void f(const T* th, const T (&x)[N_], T (&o)[N_]) {
    // LearnableLinear.f, cdnlgssm_utils.py:60-61
#pragma unroll
    for (int i = 0; i < N_; ++i) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < N_; ++k) s += th[i * N_ + k] * x[k];
      o[i] = s + th[N_ * N_ + i];
    }
  }
  template <typename T>
  __device__ __forceinline__ static void jp(const T* th, const T (&x)[N_], const T (&P)[N_ * (N_ + 1) / 2],
                                            T (&G)[N_][N_]) {
#pragma unroll
    for (int i = 0; i < N_; ++i)
#pragma unroll
      for (int j = 0; j < N_; ++j) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < N_; ++k) s += th[i * N_ + k] * P[pidx<N_>(k, j)];
        G[i][j] = s;
      }
  }
};

// dt * rhs of the EKF moment ODE (orders 'first' and 'second'; the reference's second-order term
// 0.5*einsum('iik,kl->l', Hess, P) is identically zero for every drift handled here -- SURVEY F8).
template <typename T, class Drift>
__device__ __forceinline__ void ekf_rhs(const T* th, const T* lql, const St<T, Drift::NX>& y, T dt,
                                        St<T, Drift::NX>& k) {
  constexpr int NX = Drift::NX;
  T f[NX];
  Drift::f(th, y.m, f);
  T G[NX][NX];
  Drift::jp(th, y.m, y.P, G);
#pragma unroll
  for (int i = 0; i < NX; ++i) k.m[i] = dt * f[i];
#pragma unroll
  for (int i = 0; i < NX; ++i)
#pragma unroll
    for (int j = i; j < NX; ++j) k.P[pidx<NX>(i, j)] = dt * ((G[i][j] + G[j][i]) + lql[pidx<NX>(i, j)]);
}

template <typename T, class Drift, int SOLVER>
__device__ __forceinline__ void rk_step(const T* th, const T* lql, St<T, Drift::NX>& y, T dt) {
  using TB = Tab<SOLVER>;
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  St<T, NX> k[TB::S];
  St<T, NX> acc = y;
#pragma unroll
  for (int i = 0; i < TB::S; ++i) {
    St<T, NX> yi = y;
#pragma unroll
    for (int j = 0; j < i; ++j) {
      if (TB::a(i, j) != 0.0) {
        const T aij = T(TB::a(i, j));
#pragma unroll
        for (int e = 0; e < NX; ++e) yi.m[e] += aij * k[j].m[e];
#pragma unroll
        for (int e = 0; e < NP; ++e) yi.P[e] += aij * k[j].P[e];
      }
    }
    ekf_rhs<T, Drift>(th, lql, yi, dt, k[i]);
    if (TB::b(i) != 0.0) {
      const T bi = T(TB::b(i));
#pragma unroll
      for (int e = 0; e < NX; ++e) acc.m[e] += bi * k[i].m[e];
#pragma unroll
      for (int e = 0; e < NP; ++e) acc.P[e] += bi * k[i].P[e];
    }
  }
  y = acc;
}

// Measurement update + log-likelihood increment (inference_ekf.py:285-289, :153-199; psd_solve utils.py:202-207).
// sh: H[NY*NX], d[NY], R[NY*NY] in shared memory.
template <typename T, int NX, int NY>
__device__ __forceinline__ T ekf_update(const T* H, const T* dvec, const T* R, St<T, NX>& s, const T (&y)[NY],
                                        int num_iter) {
  T ll = T(0);
  for (int it = 0; it < num_iter; ++it) {
    T HP[NY][NX];
#pragma unroll
    for (int a = 0; a < NY; ++a)
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < NX; ++k) acc += H[a * NX + k] * s.P[pidx<NX>(k, j)];
        HP[a][j] = acc;
      }
    T S[NY][NY];
#pragma unroll
    for (int a = 0; a < NY; ++a)
#pragma unroll
      for (int b = 0; b < NY; ++b) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < NX; ++k) acc += HP[a][k] * H[b * NX + k];
        S[a][b] = R[a * NY + b] + acc;
      }
    T r[NY];
#pragma unroll
    for (int a = 0; a < NY; ++a) {
      T acc = dvec[a];
#pragma unroll
      for (int k = 0; k < NX; ++k) acc += H[a * NX + k] * s.m[k];
      r[a] = y[a] - acc;
    }
    if (it == 0) {
      // MVN(h(m), H P H^T + R).log_prob(y): Cholesky of S without jitter (TFP)
      T Lc[NY][NY];
      T z[NY];
      T logdet = T(0), quad = T(0);
#pragma unroll
      for (int j = 0; j < NY; ++j) {
        T dsum = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) dsum -= Lc[j][k] * Lc[j][k];
        const T ljj = sqrt(dsum);
        Lc[j][j] = ljj;
        const T inv = T(1) / ljj;
#pragma unroll
        for (int i = j + 1; i < NY; ++i) {
          T v = S[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
          Lc[i][j] = v * inv;
        }
        T zz = r[j];
#pragma unroll
        for (int k = 0; k < j; ++k) zz -= Lc[j][k] * z[k];
        z[j] = zz * inv;
        quad += z[j] * z[j];
        logdet += log(ljj);
      }
      ll = T(-0.5) * quad - logdet - T(NY) * half_log_2pi<T>();
    }
    // K = psd_solve(S, H P)^T : Cholesky of sym(S) + 1e-9 I
    T Lb[NY][NY];
    T inv_d[NY];
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      T dsum = S[j][j] + T(1e-9);
#pragma unroll
      for (int k = 0; k < j; ++k) dsum -= Lb[j][k] * Lb[j][k];
      const T ljj = sqrt(dsum);
      Lb[j][j] = ljj;
      inv_d[j] = T(1) / ljj;
#pragma unroll
      for (int i = j + 1; i < NY; ++i) {
        T v = T(0.5) * (S[i][j] + S[j][i]);
#pragma unroll
        for (int k = 0; k < j; ++k) v -= Lb[i][k] * Lb[j][k];
        Lb[i][j] = v * inv_d[j];
      }
    }
    T Kt[NY][NX];  // Kt = (S + boost)^-1 H P, K = Kt^T
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      T w[NY];
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        T v = HP[i][c];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= Lb[i][k] * w[k];
        w[i] = v * inv_d[i];
      }
#pragma unroll
      for (int i = NY - 1; i >= 0; --i) {
        T v = w[i];
#pragma unroll
        for (int k = i + 1; k < NY; ++k) v -= Lb[k][i] * Kt[k][c];
        Kt[i][c] = v * inv_d[i];
      }
    }
    // P <- P - K S K^T (un-boosted S), m <- m + K (y - h(m))
    T KS[NX][NY];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int b = 0; b < NY; ++b) {
        T acc = T(0);
#pragma unroll
        for (int a = 0; a < NY; ++a) acc += Kt[a][i] * S[a][b];
        KS[i][b] = acc;
      }
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) {
        T acc = T(0);
#pragma unroll
        for (int b = 0; b < NY; ++b) acc += KS[i][b] * Kt[b][j];
        s.P[pidx<NX>(i, j)] -= acc;
      }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T acc = T(0);
#pragma unroll
      for (int a = 0; a < NY; ++a) acc += Kt[a][i] * r[a];
      s.m[i] += acc;
    }
  }
  return ll;
}

template <typename T, int NX>
__device__ __forceinline__ void store_moments(T* __restrict__ M, T* __restrict__ C, long long row,
                                              const St<T, NX>& s) {
  if (M) {
#pragma unroll
    for (int i = 0; i < NX; ++i) M[row * NX + i] = s.m[i];
  }
  if (C) {
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) C[row * NX * NX + i * NX + j] = s.P[pidx<NX>(i, j)];
  }
}

constexpr int SMALL_TPB = 64;

template <typename T, class Drift, int NY, int SOLVER>
__global__ void __launch_bounds__(SMALL_TPB, 7) ekf_small_kernel(const KArgs<T> a) {
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  // shared model constants: theta | lql (packed) | H | d | R.  One copy per CTA, or one per thread when batched.
  constexpr int NPAR = NTH + NP + NY * NX + NY + NY * NY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sh = reinterpret_cast<T*>(smem_raw);

  const long long N = a.d.N;
  const int K = a.d.K;
  const long long traj = (long long)blockIdx.x * SMALL_TPB + threadIdx.x;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  T* par = par_batched ? sh + threadIdx.x * NPAR : sh;
  const long long tj = traj < N ? traj : (N - 1);
  if (par_batched || threadIdx.x == 0) {
    const T* th = a.in[CDK_IN_F] + tj * a.in_stride[CDK_IN_F];
    const T* Lm = a.in[CDK_IN_L] + tj * a.in_stride[CDK_IN_L];
    const T* Qc = a.in[CDK_IN_QC] + tj * a.in_stride[CDK_IN_QC];
    const T* H = a.in[CDK_IN_H] + tj * a.in_stride[CDK_IN_H];
    const T* dv = a.in[CDK_IN_D] + tj * a.in_stride[CDK_IN_D];
    const T* R = a.in[CDK_IN_R] + tj * a.in_stride[CDK_IN_R];
    for (int i = 0; i < NTH; ++i) par[i] = th[i];
    // L Qc L^T (inference_ekf.py:86-87,105; loop-invariant, hoisted)
    for (int i = 0; i < NX; ++i)
      for (int j = i; j < NX; ++j) {
        T acc = T(0);
        for (int p = 0; p < NX; ++p) {
          T lq = T(0);
          for (int q = 0; q < NX; ++q) lq += Lm[i * NX + q] * Qc[q * NX + p];
          acc += lq * Lm[j * NX + p];
        }
        par[NTH + pidx<NX>(i, j)] = acc;
      }
    for (int i = 0; i < NY * NX; ++i) par[NTH + NP + i] = H[i];
    for (int i = 0; i < NY; ++i) par[NTH + NP + NY * NX + i] = dv[i];
    for (int i = 0; i < NY * NY; ++i) par[NTH + NP + NY * NX + NY + i] = R[i];
  }
  __syncthreads();
  if (traj >= N) return;
  const T* th = par;
  const T* lql = par + NTH;
  const T* Hs = par + NTH + NP;
  const T* ds = Hs + NY * NX;
  const T* Rs = ds + NY;

  const T* __restrict__ Y = a.in[CDK_IN_Y] + traj * a.in_stride[CDK_IN_Y];
  const T* __restrict__ Tm = a.in[CDK_IN_T] + traj * a.in_stride[CDK_IN_T];
  T* __restrict__ FM = static_cast<T*>(a.out[CDK_OUT_FM]);
  T* __restrict__ FP = static_cast<T*>(a.out[CDK_OUT_FP]);
  T* __restrict__ PM = static_cast<T*>(a.out[CDK_OUT_PM]);
  T* __restrict__ PP = static_cast<T*>(a.out[CDK_OUT_PP]);
  T* __restrict__ LLC = static_cast<T*>(a.out[CDK_OUT_LLCUM]);

  St<T, NX> s;
  {
    const T* m0 = a.in[CDK_IN_M0] + traj * a.in_stride[CDK_IN_M0];
    const T* P0 = a.in[CDK_IN_P0] + traj * a.in_stride[CDK_IN_P0];
#pragma unroll
    for (int i = 0; i < NX; ++i) s.m[i] = m0[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) s.P[pidx<NX>(i, j)] = P0[i * NX + j];
  }

  const T dt0 = T(a.d.dt0);
  const T dtf = T(a.d.dt_final);
  const T tol = clip_tol<T>();
  const int max_steps = a.d.max_steps;
  const int num_iter = a.d.num_iter;

  // prefetched observation for the next update and the end time of the gap that follows it
  T y_nx[NY];
#pragma unroll
  for (int i = 0; i < NY; ++i) y_nx[i] = Y[i];
  T t_cur = Tm[0];
  T t_nxt = K > 1 ? Tm[1] : t_cur + dtf;

  T ll = T(0);
  int status = 0;
  int k = 0;
  int nsteps = 0;
  T tprev = T(0), tnext = T(0), t1 = T(0);
  bool need_update = true;
  const long long row0 = traj * (long long)K;

  while (true) {
    if (need_update) {
      ll += ekf_update<T, NX, NY>(Hs, ds, Rs, s, y_nx, num_iter);
      if (LLC) LLC[row0 + k] = ll;
      store_moments<T, NX>(FM, FP, row0 + k, s);
      tprev = t_cur;
      t1 = t_nxt;
      if (k + 1 < K) {
#pragma unroll
        for (int i = 0; i < NY; ++i) y_nx[i] = Y[(long long)(k + 1) * NY + i];
        t_cur = t1;
        t_nxt = (k + 2 < K) ? Tm[k + 2] : t1 + dtf;
      }
      tnext = fmin(tprev + dt0, t1);
      nsteps = 0;
      need_update = false;
    }
    if (tprev < t1) {
      if (nsteps >= max_steps) {  // diffrax max_steps exceeded: poison this trajectory, abandon the gap
        status = 2;
#pragma unroll
        for (int i = 0; i < NX; ++i) s.m[i] = T(NAN);
#pragma unroll
        for (int i = 0; i < NP; ++i) s.P[i] = T(NAN);
        tprev = t1;
      } else {
        rk_step<T, Drift, SOLVER>(th, lql, s, tnext - tprev);
        ++nsteps;
        tprev = tnext;
        const T cand = tprev + dt0;
        tnext = cand > t1 - tol ? t1 : cand;
      }
    }
    if (!(tprev < t1)) {
      store_moments<T, NX>(PM, PP, row0 + k, s);
      ++k;
      if (k == K) break;
      need_update = true;
    }
  }
  if (status == 0 && !isfinite(ll)) status = 1;
  if (a.out[CDK_OUT_LL]) static_cast<T*>(a.out[CDK_OUT_LL])[traj] = ll;
  if (a.out[CDK_OUT_STATUS]) static_cast<int*>(a.out[CDK_OUT_STATUS])[traj] = status;
}

template <typename T, class Drift, int NY, int SOLVER>
int launch_one(const KArgs<T>& a, cudaStream_t s) {
  constexpr int NX = Drift::NX;
  constexpr int NPAR = Drift::NTHETA + NX * (NX + 1) / 2 + NY * NX + NY + NY * NY;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  const size_t smem = sizeof(T) * NPAR * (par_batched ? SMALL_TPB : 1);
  const long long blocks = (a.d.N + SMALL_TPB - 1) / SMALL_TPB;
  if (blocks == 0) return CDK_OK;
  ekf_small_kernel<T, Drift, NY, SOLVER><<<(unsigned)blocks, SMALL_TPB, smem, s>>>(a);
  note_launch();
  return check_launch("ekf_small_kernel");
}

template <typename T, class Drift, int NY>
int launch_solver(const KArgs<T>& a, cudaStream_t s) {
  switch (a.d.solver) {
    case CDK_RK4: return launch_one<T, Drift, NY, CDK_RK4>(a, s);
    case CDK_DOPRI5: return launch_one<T, Drift, NY, CDK_DOPRI5>(a, s);
    case CDK_EULER: return launch_one<T, Drift, NY, CDK_EULER>(a, s);
    case CDK_HEUN: return launch_one<T, Drift, NY, CDK_HEUN>(a, s);
  }
  return CDK_E_UNSUPPORTED;
}

template <typename T, class Drift>
int launch_ny(const KArgs<T>& a, cudaStream_t s) {
  switch (a.d.m) {
    case 1: return launch_solver<T, Drift, 1>(a, s);
    case 2: return launch_solver<T, Drift, 2>(a, s);
    case 3: return launch_solver<T, Drift, 3>(a, s);
  }
  return CDK_E_UNSUPPORTED;
}

}  // namespace

// Fast path coverage: EKF filter, state_order first/second, Lorenz-63 (n = 3), m <= 3, solvers rk4/dopri5/euler/heun.
template <typename T>
int launch_ekf_small(const KArgs<T>& a, cudaStream_t s) {
  const cdk_desc& d = a.d;
  if (d.state_order == CDK_ORDER_ZEROTH) return CDK_E_UNSUPPORTED;
  if (d.drift_id == CDK_DRIFT_LORENZ63 && d.n == 3) return launch_ny<T, DriftL63>(a, s);
  return CDK_E_UNSUPPORTED;
}

template int launch_ekf_small<double>(const KArgs<double>&, cudaStream_t);
template int launch_ekf_small<float>(const KArgs<float>&, cudaStream_t);

}  // namespace cdk
