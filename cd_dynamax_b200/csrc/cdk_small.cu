// cdk_small.cu -- register-arithmetic CD-EKF for tiny state dimensions (Lorenz-63: m[3] + symmetric P[6]).
//
// Replaces extended_kalman_filter (src/continuous_discrete_nonlinear_gaussian_ssm/inference_ekf.py:202-326) with its
// _predict (:46-148, moment ODE dm = f(m), dP = F P + P F^T + L Qc L^T) and _condition_on (:153-199) for the
// registry drifts whose whole filter state fits in registers.
//
// B200 mapping (BASELINE config 3: N = 65,536, K = 1,000, n = 3, m = 1) -- kernel `ekf_small_v5`:
//  * One thread integrates one trajectory-gap at a time, all arithmetic in registers (FP64 FMA pipe bound).
//  * Irregular gaps give every trajectory its own substep count q_k in {3..6}.  A warp that keeps a fixed set of 32
//    trajectories runs every gap to the warp-wide maximum (what jax.vmap does to diffrax's while_loop: 4.5/6 = 75 %
//    lane efficiency).  Instead the CANONICAL per-trajectory state lives in shared memory (SoA by slot) and, every
//    observation step, the CTA re-assigns trajectories to threads with a counting sort on ceil(gap / dt0), so that the
//    32 lanes of a warp integrate gaps with (nearly) the same number of substeps.  The sort key only depends on the
//    time stamps, so it is computed one step ahead, off the critical path; it costs one extra barrier per step.
//    Which thread integrates a trajectory never changes its arithmetic, so results are bit-reproducible.
//  * Observations and time stamps stream HBM -> shared memory through a 4-deep cp.async ring (issued three steps ahead).
//  * Warp specialisation: 7 worker warps (224 trajectory slots) + 1 I/O warp per CTA.  While the workers integrate step
//    k, the I/O warp (a) issues the cp.async loads of step k+3, (b) computes the assignment of step k+1, and (c) flushes
//    the outputs of step k-1 -- which sit in the same shared-memory slots that hold the canonical state (double-buffered
//    by step parity) -- with coalesced stores, instead of 24 lane-scattered 8-byte stores per step (32 L1 wavefronts
//    each).  The worker critical path is update + substeps + ONE barrier per step.
//  * 256-thread CTAs, 2 per SM: 148 * 2 * 224 = 66,304 >= 65,536 trajectories in ONE wave (128 registers/thread).
#include <type_traits>

#include "cdk_common.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

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
  __device__ __forceinline__ static void jac(const T* th, const T (&x)[3], T (&J)[3][3]) {
    J[0][0] = -th[0]; J[0][1] = th[0]; J[0][2] = T(0);
    J[1][0] = th[1] - x[2]; J[1][1] = T(-1); J[1][2] = -x[0];
    J[2][0] = x[1]; J[2][1] = x[0]; J[2][2] = -th[2];
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

// rhs of the EKF moment ODE, WITHOUT the dt factor (orders 'first' and 'second'; the reference's second-order term
// 0.5*einsum('iik,kl->l', Hess, P) is identically zero for every drift handled here -- SURVEY F8).
template <typename T, class Drift>
__device__ __forceinline__ void ekf_rhs(const T* th, const T* lql, const St<T, Drift::NX>& y, St<T, Drift::NX>& k) {
  constexpr int NX = Drift::NX;
  Drift::f(th, y.m, k.m);
  T G[NX][NX];
  Drift::jp(th, y.m, y.P, G);
#pragma unroll
  for (int i = 0; i < NX; ++i)
#pragma unroll
    for (int j = i; j < NX; ++j)
      k.P[pidx<NX>(i, j)] = (i == j) ? fma(T(2), G[i][i], lql[pidx<NX>(i, i)]) : (G[i][j] + G[j][i]) + lql[pidx<NX>(i, j)];
}

// One explicit RK step y <- y + dt * sum_i b_i f(y_i), y_i = y + dt * sum_j a_ij f(y_j)  (diffrax stores k_i = dt f;
// folding dt into the coefficients is the same arithmetic up to rounding and saves one multiply per state element).
template <typename T, class Drift, int SOLVER>
__device__ __forceinline__ void rk_step(const T* th, const T* lql, St<T, Drift::NX>& y, T dt) {
  using TB = Tab<SOLVER>;
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  // The weighted stage sum is accumulated relative to the first non-zero weight, ksum = sum_i (b_i / b_i0) k_i, and added
  // with ONE in-place fma y += (b_i0 dt) ksum: the same operation count as accumulating y + dt b_i k_i stage by stage, but
  // the new state is written straight into the registers of the old one (no register copies at the loop back-edge).
  constexpr int I0 = TB::b(0) != 0.0 ? 0 : (TB::S > 1 && TB::b(1) != 0.0 ? 1 : (TB::S > 2 && TB::b(2) != 0.0 ? 2 : 3));
  St<T, NX> k[TB::S];
  St<T, NX> ksum;
#pragma unroll
  for (int i = 0; i < TB::S; ++i) {
    St<T, NX> yi = y;
#pragma unroll
    for (int j = 0; j < i; ++j) {
      if (TB::a(i, j) != 0.0) {
        const T c = T(TB::a(i, j)) * dt;
#pragma unroll
        for (int e = 0; e < NX; ++e) yi.m[e] = fma(c, k[j].m[e], yi.m[e]);
#pragma unroll
        for (int e = 0; e < NP; ++e) yi.P[e] = fma(c, k[j].P[e], yi.P[e]);
      }
    }
    ekf_rhs<T, Drift>(th, lql, yi, k[i]);
    if (i == I0) {
      ksum = k[i];
    } else if (TB::b(i) != 0.0) {
      const T c = T(TB::b(i) / TB::b(I0));
      if (TB::b(i) == TB::b(I0)) {
#pragma unroll
        for (int e = 0; e < NX; ++e) ksum.m[e] += k[i].m[e];
#pragma unroll
        for (int e = 0; e < NP; ++e) ksum.P[e] += k[i].P[e];
      } else {
#pragma unroll
        for (int e = 0; e < NX; ++e) ksum.m[e] = fma(c, k[i].m[e], ksum.m[e]);
#pragma unroll
        for (int e = 0; e < NP; ++e) ksum.P[e] = fma(c, k[i].P[e], ksum.P[e]);
      }
    }
  }
  const T w = T(TB::b(I0)) * dt;
#pragma unroll
  for (int e = 0; e < NX; ++e) y.m[e] = fma(w, ksum.m[e], y.m[e]);
#pragma unroll
  for (int e = 0; e < NP; ++e) y.P[e] = fma(w, ksum.P[e], y.P[e]);
}

// Measurement update + log-likelihood increment (inference_ekf.py:285-289, :153-199; psd_solve utils.py:202-207).
// sh: H[NY*NX], d[NY], R[NY*NY] in shared memory.
template <typename T, int NX, int NY>
__device__ __forceinline__ T ekf_update(const T* H, const T* dvec, const T* R, St<T, NX>& s, const T (&y)[NY],
                                        int num_iter, T* s_defer = nullptr) {
  // s_defer (scalar emission only): the caller accumulates log S itself (as the log of a running product, one log per
  // 8 steps instead of one per step); the returned increment then omits the -log(S)/2 term and *s_defer = S.
  T ll = T(0);
  if constexpr (NY == 1) {
    // Scalar emission: the 1x1 Cholesky / triangular solves collapse to two independent reciprocals and one log, which
    // the scheduler can overlap (the generic path below is one long sqrt -> rcp -> log dependency chain).
    //   log N(y; h(m), S) = -r^2 / (2 S) - log(S) / 2 - log(2 pi) / 2;   K = P H^T / (S + 1e-9);   P -= K S K^T
    for (int it = 0; it < num_iter; ++it) {
      T HP[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < NX; ++k) acc += H[k] * s.P[pidx<NX>(k, j)];
        HP[j] = acc;
      }
      T S = R[0], hm = dvec[0];
#pragma unroll
      for (int k = 0; k < NX; ++k) {
        S += HP[k] * H[k];
        hm += H[k] * s.m[k];
      }
      const T r = y[0] - hm;
      const T rb = T(1) / (S + T(1e-9));
      if (it == 0) {
        if (s_defer) {
          *s_defer = S;
          ll = T(-0.5) * (r * r / S) - half_log_2pi<T>();
        } else {
          ll = T(-0.5) * (r * r / S) - T(0.5) * log(S) - half_log_2pi<T>();
        }
      }
      T Kt[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) Kt[j] = HP[j] * rb;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const T ks = Kt[i] * S;
#pragma unroll
        for (int j = i; j < NX; ++j) s.P[pidx<NX>(i, j)] -= ks * Kt[j];
        s.m[i] += Kt[i] * r;
      }
    }
    return ll;
  } else {
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
}


// ---- cp.async (LDGSTS) element copies: sizeof(T) in {4, 8} is always naturally aligned -----------------------------
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef CDK_EKF_DEFAULT_MODE
#define CDK_EKF_DEFAULT_MODE 2
#endif
constexpr int V5_W = 224;         // worker threads (7 warps) = trajectory slots per CTA
constexpr int V5_TPB = 256;       // + 1 helper warp; 2 CTAs / SM (128 registers / thread)
constexpr int V5_LD = V5_W + 1;   // SoA row stride of the input rings (odd: conflict-free)
constexpr int V5_TRING = 6;       // time-stamp ring depth (steps k .. k+3 are read, k+5 is in flight)
constexpr int V5_TAHEAD = 5;
constexpr int V5_YRING = 4;       // emission ring depth (step k is read, k+2 is in flight)
constexpr int V5_YAHEAD = 2;
constexpr int V5_NB = 32;         // counting-sort buckets (substep count clamped to 1..31; 0 = dead slot)
constexpr int V5_SPL = V5_W / 32; // slots per helper-warp lane

// TMA descriptors of the four per-step output arrays viewed as 2-D tensors [N][K*len] (len = NX or NX*NX): one box is
// {2 steps x len elements, 224 trajectories}, so ONE cp.async.bulk.tensor store per array writes a whole CTA's two steps.
struct alignas(64) V5Maps {
  CUtensorMap m[4];  // FM, FP, PM, PP
  int use_tma;
};

template <typename T, int NX, int NY>
struct alignas(128) V5Smem {
  // Output staging = canonical state, in the exact global row layout [slot][2 steps][len] (dense: it is a TMA box).
  T fm[V5_W][2][NX];
  alignas(128) T fp[V5_W][2][NX * NX];
  alignas(128) T pm[V5_W][2][NX];
  alignas(128) T pp[V5_W][2][NX * NX];
  T inY[V5_YRING][NY][V5_LD];
  T inT[V5_TRING][V5_LD];
  T ll[V5_LD];
  int status[V5_W];
  int perm[2][V5_W];   // thread -> slot assignment, by step parity
  int hist[3][V5_NB];  // counting-sort histograms, by step mod 3 (counted 2 steps ahead, scanned 1 step ahead)
  // followed by the model constants: NPAR values (shared) or V5_W * NPAR (one block per slot when batched)
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_src));
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(s)
               : "memory");
}

// Fallback flush (fp32, odd K, unaligned outputs): all threads copy rows [k0, k0 + nrow) of every live slot.
template <typename T, int NX, int NY>
__device__ __forceinline__ void v5_flush_generic(const V5Smem<T, NX, NY>& sm, void* const* out, int k0, int nrow, int K,
                                                 long long traj0, int nlive) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    T* __restrict__ G = static_cast<T*>(out[a == 0 ? CDK_OUT_FM : a == 1 ? CDK_OUT_FP : a == 2 ? CDK_OUT_PM : CDK_OUT_PP]);
    if (!G) continue;
    const int len = (a & 1) ? NX * NX : NX;
    const T* src = a == 0 ? &sm.fm[0][0][0] : a == 1 ? &sm.fp[0][0][0] : a == 2 ? &sm.pm[0][0][0] : &sm.pp[0][0][0];
    const int per = nrow * len;
    const int total = nlive * per;
    const int r0 = k0 & 1;
    for (int u = threadIdx.x; u < total; u += V5_TPB) {
      const int slot = u / per;
      const int e = u - slot * per;
      G[((traj0 + slot) * (long long)K + k0) * len + e] = src[slot * 2 * len + r0 * len + e];
    }
  }
}

template <typename T, class Drift, int NY, int SOLVER, bool REGROUP>
__global__ void __launch_bounds__(V5_TPB, 2) ekf_small_v5(const KArgs<T> a, const __grid_constant__ V5Maps maps) {
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  constexpr int NPAR = NTH + NP + NY * NX + NY + NY * NY;  // theta | lql (packed) | H | d | R
  using S = V5Smem<T, NX, NY>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  T* parbase = reinterpret_cast<T*>(smem_raw + sizeof(S));

  const long long N = a.d.N;
  const int K = a.d.K;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool helper = tid >= V5_W;
  const long long traj0 = (long long)blockIdx.x * V5_W;
  const int nlive = (int)((N - traj0) < V5_W ? (N - traj0) : V5_W);
  const bool home_live = tid < nlive;  // worker thread whose home slot holds a trajectory
  const long long traj = traj0 + tid;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  const bool use_tma = sizeof(T) == 8 && maps.use_tma != 0;
  const T* __restrict__ Ybase = a.in[CDK_IN_Y] + traj0 * a.in_stride[CDK_IN_Y];
  const T* __restrict__ Tbase = a.in[CDK_IN_T] + traj0 * a.in_stride[CDK_IN_T];
  const long long ystride = a.in_stride[CDK_IN_Y], tstride = a.in_stride[CDK_IN_T];
  const T dt0 = T(a.d.dt0);
  const T dtf = T(a.d.dt_final);
  const T inv_dt0 = T(1) / dt0;

  // ---- prologue: model constants, initial moments, first input steps ----
  if ((par_batched && home_live) || (!par_batched && tid == 0)) {
    T* par = par_batched ? parbase + tid * NPAR : parbase;
    const long long tj = par_batched ? traj : 0;
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
  if (!helper) {
    sm.status[tid] = 0;
    sm.ll[tid] = T(0);
    sm.perm[0][tid] = tid;
    sm.perm[1][tid] = tid;
  } else {
    for (int i = lane; i < 3 * V5_NB; i += 32) (&sm.hist[0][0])[i] = 0;
  }
  if (home_live) {
    const T* m0 = a.in[CDK_IN_M0] + traj * a.in_stride[CDK_IN_M0];
    const T* P0 = a.in[CDK_IN_P0] + traj * a.in_stride[CDK_IN_P0];
    // the prior is the prediction for t_0 (inference_ekf.py:320): it sits where step -1 (row 1) would have left it
#pragma unroll
    for (int i = 0; i < NX; ++i) sm.pm[tid][1][i] = m0[i];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) sm.pp[tid][1][i] = P0[i];
    for (int kk = 0; kk < V5_TAHEAD && kk < K; ++kk) {
      if (kk < V5_YAHEAD) {
#pragma unroll
        for (int c = 0; c < NY; ++c) cp_async_elem(&sm.inY[kk][c][tid], Ybase + tid * ystride + (long long)kk * NY + c);
      }
      cp_async_elem(&sm.inT[kk][tid], Tbase + tid * tstride + kk);
    }
    cp_async_commit();
    cp_async_wait_all();
  }
  __syncthreads();

  // ---- per-step regrouping (worker threads): counting sort of the home slots by the substep count of the gap after
  //      observation kk.  count(kk) runs during step kk-2, scatter(kk) during step kk-1 -> perm[kk & 1] for step kk.
  int my_bucket = 0, my_rank = 0;
  auto sort_count = [&](int kk) {
    int b = 0;
    if (home_live) {
      const T t0 = sm.inT[kk % V5_TRING][tid];
      const T t1 = kk + 1 < K ? sm.inT[(kk + 1) % V5_TRING][tid] : t0 + dtf;
      const T q = ceil((t1 - t0) * inv_dt0);
      b = q > T(1) ? (q < T(V5_NB - 1) ? (int)q : V5_NB - 1) : 1;
    }
    const unsigned grp = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&sm.hist[kk % 3][b], __popc(grp));
    base = __shfl_sync(0xffffffffu, base, leader);
    my_bucket = b;
    my_rank = base + __popc(grp & ((1u << lane) - 1u));
  };
  auto sort_scatter = [&](int kk) {
    const int cnt = sm.hist[kk % 3][lane];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int off = __shfl_sync(0xffffffffu, incl - cnt, my_bucket);
    sm.perm[kk & 1][off + my_rank] = tid;
  };
  if (REGROUP) {
    if (!helper) sort_count(0);
    __syncthreads();
    if (!helper) {
      sort_scatter(0);
      if (K > 1) sort_count(1);
    }
    __syncthreads();
  }

  const T tol = clip_tol<T>();
  const int max_steps = a.d.max_steps;
  const int num_iter = a.d.num_iter;
  T* __restrict__ LLC = static_cast<T*>(a.out[CDK_OUT_LLCUM]);

  for (int k = 0; k < K; ++k) {
    const int row = k & 1;
    if (!helper) {
      if (REGROUP && k + 1 < K) sort_scatter(k + 1);
      // ---- worker: update at t_k, then integrate the gap t_k -> t_{k+1}, for the slot assigned to this thread ----
      // Sorted chunk c (32 consecutive ranks, c = 6 has the longest gaps) -> warp: the heavy and light chunks are paired on
      // the warps that share an SM sub-partition (warp w issues on sub-partition w % 4), the heaviest chunk goes to the
      // sub-partition that only hosts one worker warp (+ the helper warp), so the four FP64 pipes carry equal work.
      const int chunk = (0x2106345 >> (4 * (tid >> 5))) & 7;  // warps 0..6 -> chunks 5,4,3,6,0,1,2
      const int p = REGROUP ? sm.perm[k & 1][chunk * 32 + lane] : tid;
      const bool live = p < nlive;
      const T* par = par_batched ? parbase + (live ? p : 0) * NPAR : parbase;
      const T* th = par;
      const T* lql = par + NTH;
      St<T, NX> s;
      T tprev = T(0), t1 = T(0), ll = T(0);
      if (live) {
        const T* Hs = par + NTH + NP;
        const T* ds = Hs + NY * NX;
        const T* Rs = ds + NY;
#pragma unroll
        for (int i = 0; i < NX; ++i) s.m[i] = sm.pm[p][row ^ 1][i];
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = i; j < NX; ++j) s.P[pidx<NX>(i, j)] = sm.pp[p][row ^ 1][i * NX + j];
        T y[NY];
#pragma unroll
        for (int c = 0; c < NY; ++c) y[c] = sm.inY[k % V5_YRING][c][p];
        tprev = sm.inT[k % V5_TRING][p];
        t1 = k + 1 < K ? sm.inT[(k + 1) % V5_TRING][p] : tprev + dtf;
        ll = sm.ll[p] + ekf_update<T, NX, NY>(Hs, ds, Rs, s, y, num_iter);
      }
      // the TMA store of the previous 2-step block must have finished READING the staging rows before row 0 is rewritten
      if (use_tma && row == 0) asm volatile("bar.sync 1, %0;" ::"n"(V5_TPB) : "memory");
      if (live) {
        sm.ll[p] = ll;
        if (LLC) LLC[(traj0 + p) * (long long)K + k] = ll;
#pragma unroll
        for (int i = 0; i < NX; ++i) sm.fm[p][row][i] = s.m[i];
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = 0; j < NX; ++j) sm.fp[p][row][i * NX + j] = s.P[pidx<NX>(i, j)];
        // diffrax ConstantStepSize stepping (diffrax_utils.py:150-163)
        T tnext = fmin(tprev + dt0, t1);
        int nsteps = 0;
        while (tprev < t1) {
          if (nsteps >= max_steps) {  // diffrax max_steps exceeded: poison this trajectory, abandon the gap
            sm.status[p] = 2;
#pragma unroll
            for (int i = 0; i < NX; ++i) s.m[i] = T(NAN);
#pragma unroll
            for (int i = 0; i < NP; ++i) s.P[i] = T(NAN);
            break;
          }
          rk_step<T, Drift, SOLVER>(th, lql, s, tnext - tprev);
          ++nsteps;
          tprev = tnext;
          const T cand = tprev + dt0;
          tnext = cand > t1 - tol ? t1 : cand;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) sm.pm[p][row][i] = s.m[i];
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = 0; j < NX; ++j) sm.pp[p][row][i * NX + j] = s.P[pidx<NX>(i, j)];
      }
      if (REGROUP && k + 2 < K) sort_count(k + 2);
    } else {
      // ---- helper warp: output store of the previous block, input loads, sort housekeeping ----
      if (use_tma && row == 0) {
        if (k > 0 && lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const int c1 = (int)traj0;
          if (a.out[CDK_OUT_FM]) tma_store_2d(&maps.m[0], &sm.fm[0][0][0], (k - 2) * NX, c1);
          if (a.out[CDK_OUT_FP]) tma_store_2d(&maps.m[1], &sm.fp[0][0][0], (k - 2) * NX * NX, c1);
          if (a.out[CDK_OUT_PM]) tma_store_2d(&maps.m[2], &sm.pm[0][0][0], (k - 2) * NX, c1);
          if (a.out[CDK_OUT_PP]) tma_store_2d(&maps.m[3], &sm.pp[0][0][0], (k - 2) * NX * NX, c1);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(V5_TPB) : "memory");
      }
      const int kt = k + V5_TAHEAD, ky = k + V5_YAHEAD;
#pragma unroll
      for (int r = 0; r < V5_SPL; ++r) {
        const int slot = lane + 32 * r;
        if (slot < nlive) {
          if (ky < K) {
#pragma unroll
            for (int c = 0; c < NY; ++c)
              cp_async_elem(&sm.inY[ky % V5_YRING][c][slot], Ybase + slot * ystride + (long long)ky * NY + c);
          }
          if (kt < K) cp_async_elem(&sm.inT[kt % V5_TRING][slot], Tbase + slot * tstride + kt);
        }
      }
      cp_async_commit();
      sm.hist[k % 3][lane] = 0;  // scanned during step k-1, counted again during step k+1
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // the loads issued during step k-1 have landed
    }
    __syncthreads();
    if (!use_tma && (row == 1 || k == K - 1)) {
      v5_flush_generic<T, NX, NY>(sm, a.out, k - row, row + 1, K, traj0, nlive);
      __syncthreads();
    }
  }
  if (use_tma && helper && lane == 0) {  // K is even on this path: the last block is rows K-2, K-1
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int c1 = (int)traj0;
    if (a.out[CDK_OUT_FM]) tma_store_2d(&maps.m[0], &sm.fm[0][0][0], (K - 2) * NX, c1);
    if (a.out[CDK_OUT_FP]) tma_store_2d(&maps.m[1], &sm.fp[0][0][0], (K - 2) * NX * NX, c1);
    if (a.out[CDK_OUT_PM]) tma_store_2d(&maps.m[2], &sm.pm[0][0][0], (K - 2) * NX, c1);
    if (a.out[CDK_OUT_PP]) tma_store_2d(&maps.m[3], &sm.pp[0][0][0], (K - 2) * NX * NX, c1);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (home_live) {
    const T ll = sm.ll[tid];
    int status = sm.status[tid];
    if (status == 0 && !isfinite(ll)) status = 1;
    if (a.out[CDK_OUT_LL]) static_cast<T*>(a.out[CDK_OUT_LL])[traj] = ll;
    if (a.out[CDK_OUT_STATUS]) static_cast<int*>(a.out[CDK_OUT_STATUS])[traj] = status;
  }
}

// ======================================================================================================================
// Variant B: INDEPENDENT WARPS.  One warp = 32 trajectories kept for the whole kernel (7 such warps share a CTA only to
// get an even 2-CTAs-per-SM placement; they never synchronise with each other); the filter state never
// leaves registers, every gap runs to the warp-wide maximum substep count (predicated lanes, 4.5/6 = 75 % lane efficiency
// on the benchmark grid) -- but there is no CTA-wide barrier at all, so the ~15 resident warps of an SM drift out of phase
// and the latency-bound measurement update of one warp hides behind the FP64-bound substeps of the others.  Inputs come
// through a private 4-deep cp.async ring; each warp stores its own 2-step output block with four TMA tensor stores.
// ======================================================================================================================
constexpr int LW_RING = 4;

// Diagnostics (scripts/trace_lw.py): when a device buffer is registered with cdk_debug_set_trace(), lane 0 of every warp
// records {globaltimer at entry, at exit, %smid, %warpid} -- used to see how warp run times spread over SM sub-partitions.
__device__ unsigned long long* g_lw_trace = nullptr;
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <typename T, int NX, int NY>
struct alignas(128) LWSmem {
  T fm[32][2][NX];
  alignas(128) T fp[32][2][NX * NX];
  alignas(128) T pm[32][2][NX];
  alignas(128) T pp[32][2][NX * NX];
  T inY[LW_RING][NY][32];
  T inT[LW_RING][32];
  // followed by the model constants: NPAR values (shared) or 32 * NPAR (one block per lane when batched)
};

// Write one staging row (LEN elements of step parity ROW) of this lane's [2][LEN] block.  fp64: the block starts 16-byte
// aligned (lane stride 16 LEN bytes), so all but at most one element go out as 128-bit stores (conflict-free per quarter
// warp) -- 14 shared-memory stores per step instead of 24.
template <int ROW, int LEN, typename T>
__device__ __forceinline__ void stage_row(T* lane_block, const T (&v)[LEN]) {
  T* p = lane_block + ROW * LEN;
  if constexpr (sizeof(T) == 8) {
    constexpr int FIRST = (ROW * LEN) & 1;
    if (FIRST) p[0] = v[0];
#pragma unroll
    for (int e = FIRST; e + 1 < LEN; e += 2) *reinterpret_cast<double2*>(p + e) = make_double2(v[e], v[e + 1]);
    if ((LEN - FIRST) & 1) p[LEN - 1] = v[LEN - 1];
  } else {
#pragma unroll
    for (int e = 0; e < LEN; ++e) p[e] = v[e];
  }
}

// Warps per CTA: 7 (two CTAs per SM, <= 128 registers) or 14 (ONE CTA per SM, <= 144 registers: the drift parameters and
// L Qc L^T then live in registers instead of being re-read from shared memory every substep).  Either way 65,536
// trajectories are one wave of 2,048 warps over 148 SMs.  CDK_LW_WPC selects; CDK_LW_SYNC=p adds a CTA-wide barrier every p
// steps (keeps the warps of an SM sub-partition progressing together instead of two of them finishing early).
template <typename T, class Drift, int NY, int SOLVER, int WPC>
__global__ void __launch_bounds__(32 * WPC, WPC == 7 ? 2 : 1)
    ekf_small_lw(const KArgs<T> a, const __grid_constant__ V5Maps maps, const int warp_bytes, const int sync_period,
                 const int use_token) {
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  constexpr int NPAR = NTH + NP + NY * NX + NY + NY * NY;  // theta | lql (packed) | H | d | R
  using S = LWSmem<T, NX, NY>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  S& sm = *reinterpret_cast<S*>(smem_raw + (size_t)warp * warp_bytes);  // every warp owns a private slice
  T* parbase = reinterpret_cast<T*>(smem_raw + (size_t)warp * warp_bytes + sizeof(S));

  const long long N = a.d.N;
  const int K = a.d.K;
  const int lane = threadIdx.x & 31;
  const long long cta0 = (long long)blockIdx.x * WPC * 32;
  const long long traj0 = cta0 + warp * 32;
  // Optional FP64-pipe semaphore (CDK_LW_TOKEN = permits, default off): at most `permits` warps of an SM sub-partition are
  // inside the RK substep loop at a time (FIFO tickets), the others update / stage / store.  Built to break the convoy
  // that forms because warps sharing a pipe equally re-synchronise their phases; measured no gain (7.2-7.5 ms with 2-3
  // permits, 8.1 ms with 1, 7.3 ms without), because the loop is bound by register-file bandwidth, not latency: a DFMA
  // with three distinct register operands issues every 3 cycles, not 2 (scripts/micro/fp64_regs.cu), and the update phase
  // of one warp already hides behind the substeps of the others.  Kept for experiments.
  __shared__ unsigned lw_token[4][2];  // [sub-partition][next ticket, tenures completed]
  if (threadIdx.x < 8) (&lw_token[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  if (traj0 >= N) return;  // whole warp out of range (warps never wait for it: see live_threads)
  unsigned hw_warp;
  asm volatile("mov.u32 %0, %warpid;" : "=r"(hw_warp));
  volatile unsigned* const tok = &lw_token[hw_warp & 3][0];
  const long long rem_cta = N - cta0;
  const int live_threads = 32 * (int)(rem_cta >= 32 * WPC ? WPC : (rem_cta + 31) / 32);
  const long long traj = traj0 + lane;
  const bool live = traj < N;
  const int nlive = (int)((N - traj0) < 32 ? (N - traj0) : 32);
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  // log-likelihood-only calls (output_fields = []) skip the staging stores and the output flush altogether
  const bool any_out = a.out[CDK_OUT_FM] || a.out[CDK_OUT_FP] || a.out[CDK_OUT_PM] || a.out[CDK_OUT_PP];
  const bool use_tma = sizeof(T) == 8 && maps.use_tma != 0 && any_out;
  unsigned long long* const trace = g_lw_trace;
  const unsigned long long t_entry = trace ? globaltimer() : 0ull;
  if ((par_batched && live) || (!par_batched && lane == 0)) {
    T* par = par_batched ? parbase + lane * NPAR : parbase;
    const long long tj = par_batched ? traj : 0;
    const T* th = a.in[CDK_IN_F] + tj * a.in_stride[CDK_IN_F];
    const T* Lm = a.in[CDK_IN_L] + tj * a.in_stride[CDK_IN_L];
    const T* Qc = a.in[CDK_IN_QC] + tj * a.in_stride[CDK_IN_QC];
    const T* H = a.in[CDK_IN_H] + tj * a.in_stride[CDK_IN_H];
    const T* dv = a.in[CDK_IN_D] + tj * a.in_stride[CDK_IN_D];
    const T* R = a.in[CDK_IN_R] + tj * a.in_stride[CDK_IN_R];
    for (int i = 0; i < NTH; ++i) par[i] = th[i];
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
  const T* __restrict__ Yg = a.in[CDK_IN_Y] + (live ? traj : 0) * a.in_stride[CDK_IN_Y];
  const T* __restrict__ Tg = a.in[CDK_IN_T] + (live ? traj : 0) * a.in_stride[CDK_IN_T];
  auto prefetch = [&](int kk) {
    if (live && kk < K) {
#pragma unroll
      for (int c = 0; c < NY; ++c) cp_async_elem(&sm.inY[kk & (LW_RING - 1)][c][lane], Yg + (long long)kk * NY + c);
      cp_async_elem(&sm.inT[kk & (LW_RING - 1)][lane], Tg + kk);
    }
    cp_async_commit();
  };
  prefetch(0);
  prefetch(1);
  prefetch(2);
  St<T, NX> s;
#pragma unroll
  for (int i = 0; i < NX; ++i) s.m[i] = T(0);
#pragma unroll
  for (int i = 0; i < NP; ++i) s.P[i] = T(0);
  if (live) {
    const T* m0 = a.in[CDK_IN_M0] + traj * a.in_stride[CDK_IN_M0];
    const T* P0 = a.in[CDK_IN_P0] + traj * a.in_stride[CDK_IN_P0];
#pragma unroll
    for (int i = 0; i < NX; ++i) s.m[i] = m0[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) s.P[pidx<NX>(i, j)] = P0[i * NX + j];
  }
  __syncwarp();
  const T* par = par_batched ? parbase + lane * NPAR : parbase;
  // WPC == 14: theta and L Qc L^T stay in registers for the whole kernel; WPC == 7: re-read from shared memory
  T th_r[NTH], lql_r[NP];
#pragma unroll
  for (int i = 0; i < NTH; ++i) th_r[i] = par[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) lql_r[i] = par[NTH + i];
  const T* th = WPC == 7 ? par : th_r;
  const T* lql = WPC == 7 ? par + NTH : lql_r;
  const T* Hs = par + NTH + NP;
  const T* ds = Hs + NY * NX;
  const T* Rs = ds + NY;
  const T dt0 = T(a.d.dt0);
  const T dtf = T(a.d.dt_final);
  const T tol = clip_tol<T>();
  const int max_steps = a.d.max_steps;
  const int num_iter = a.d.num_iter;
  T* __restrict__ LLC = static_cast<T*>(a.out[CDK_OUT_LLCUM]);
  T ll = T(0), sprod = T(1);
  bool sbad = false;
  int status = 0;
  auto tma_pair = [&](int first, int k0) {  // lane 0: store arrays {first, first + 1} of the block starting at step k0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int c1 = (int)traj0;
    if (first == 0) {
      if (a.out[CDK_OUT_FM]) tma_store_2d(&maps.m[0], &sm.fm[0][0][0], k0 * NX, c1);
      if (a.out[CDK_OUT_FP]) tma_store_2d(&maps.m[1], &sm.fp[0][0][0], k0 * NX * NX, c1);
    } else {
      if (a.out[CDK_OUT_PM]) tma_store_2d(&maps.m[2], &sm.pm[0][0][0], k0 * NX, c1);
      if (a.out[CDK_OUT_PP]) tma_store_2d(&maps.m[3], &sm.pp[0][0][0], k0 * NX * NX, c1);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  };
  // one RK step with the diffrax stepping rule (tnext = tprev + dt0, clipped to t1 within tol)
  auto substep = [&](T& tprev, T& tnext, const T t1) {
    rk_step<T, Drift, SOLVER>(th, lql, s, tnext - tprev);
    tprev = tnext;
    const T cand = tprev + dt0;
    tnext = cand > t1 - tol ? t1 : cand;
  };
  auto poison = [&]() {
    status = 2;
#pragma unroll
    for (int i = 0; i < NX; ++i) s.m[i] = T(NAN);
#pragma unroll
    for (int i = 0; i < NP; ++i) s.P[i] = T(NAN);
  };

  // current mean / full covariance -> this lane's staging rows of step parity `row`
  auto stage = [&](auto rowc, T* mrow, T* prow) {
    constexpr int row = decltype(rowc)::value;
    T full[NX * NX];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) full[i * NX + j] = s.P[pidx<NX>(i, j)];
    stage_row<row, NX>(mrow, s.m);
    stage_row<row, NX * NX>(prow, full);
  };
  // one observation step; `row` (= k & 1, the staging row) is a compile-time constant: the k loop is unrolled by two
  auto step = [&](auto rowc, const int k) {
    constexpr int row = decltype(rowc)::value;
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // own loads of steps <= k+1 have landed
    T tprev = T(0), t1 = T(0);
    if (live) {
      T y[NY];
#pragma unroll
      for (int c = 0; c < NY; ++c) y[c] = sm.inY[k & (LW_RING - 1)][c][lane];
      tprev = sm.inT[k & (LW_RING - 1)][lane];
      t1 = k + 1 < K ? sm.inT[(k + 1) & (LW_RING - 1)][lane] : tprev + dtf;
      if (NY == 1 && !LLC) {
        // sum_k log S_k = log prod_k S_k, folded every 8 steps (factors outside [1e-8, 1e8] take the direct log); a
        // non-positive S (non-PD covariance) must still poison the log-likelihood as log() would.
        T Sk;
        ll += ekf_update<T, NX, NY>(Hs, ds, Rs, s, y, num_iter, &Sk);
        if (!(Sk > T(0))) sbad = true;
        if (Sk > T(1e-8) && Sk < T(1e8)) {
          sprod *= Sk;  // at most 8 (fp32: 4) factors in [1e-8, 1e8]: no overflow / underflow
        } else {
          ll -= T(0.5) * log(Sk);
        }
        if ((k & (sizeof(T) == 8 ? 7 : 3)) == (sizeof(T) == 8 ? 7 : 3)) {
          ll -= T(0.5) * log(sprod);
          sprod = T(1);
        }
      } else {
        ll += ekf_update<T, NX, NY>(Hs, ds, Rs, s, y, num_iter);
        if (LLC) LLC[traj * (long long)K + k] = ll;
      }
    }
    prefetch(k + 3);
    if (use_tma && row == 0 && k > 0) {  // the FM/FP store of the previous block was issued ~6 substeps ago
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
    }
    if (live && any_out) stage(rowc, &sm.fm[lane][0][0], &sm.fp[lane][0][0]);
    if (use_tma && row == 1) {  // the filtered rows of this block are complete: store them while the gap is integrated
      __syncwarp();
      if (lane == 0) tma_pair(0, k - 1);
    }
    unsigned ticket = 0;
    if (use_token) {  // warp-uniform spin: every lane polls the same shared-memory word (a broadcast read)
      if (lane == 0) ticket = atomicAdd(const_cast<unsigned*>(tok), 1u);
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      while ((int)(ticket - tok[1]) >= use_token) {  // use_token = permits: warps inside the substep loop at a time
      }
    }
    if (live) {
      T tnext = fmin(tprev + dt0, t1);
      int nsteps = 0;
      while (tprev < t1 && nsteps < max_steps) {
        substep(tprev, tnext, t1);
        ++nsteps;
      }
      if (tprev < t1) poison();  // diffrax max_steps exceeded: the reference result is NaN
    }
    if (use_token) {
      __syncwarp();
      if (lane == 0) atomicAdd(const_cast<unsigned*>(tok + 1), 1u);
    }
    if (use_tma && row == 0 && k > 0) {  // the PM/PP store of the previous block was issued one whole step ago
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    if (live && any_out) stage(rowc, &sm.pm[lane][0][0], &sm.pp[lane][0][0]);
    if (any_out && (row == 1 || k == K - 1)) {
      __syncwarp();
      if (use_tma) {
        if (lane == 0) tma_pair(2, k - 1);
      } else {
        const int k0 = k - row, nrow = row + 1;
#pragma unroll
        for (int arr = 0; arr < 4; ++arr) {
          T* __restrict__ G = static_cast<T*>(
              a.out[arr == 0 ? CDK_OUT_FM : arr == 1 ? CDK_OUT_FP : arr == 2 ? CDK_OUT_PM : CDK_OUT_PP]);
          if (!G) continue;
          const int len = (arr & 1) ? NX * NX : NX;
          const T* src = arr == 0 ? &sm.fm[0][0][0] : arr == 1 ? &sm.fp[0][0][0] : arr == 2 ? &sm.pm[0][0][0] : &sm.pp[0][0][0];
          const int per = nrow * len;
          for (int u = lane; u < nlive * per; u += 32) {
            const int slot = u / per, e = u - slot * per;
            G[((traj0 + slot) * (long long)K + k0) * len + e] = src[slot * 2 * len + e];
          }
        }
        __syncwarp();
      }
    }
  };

  for (int k = 0; k < K; k += 2) {
    step(std::integral_constant<int, 0>{}, k);
    if (k + 1 < K) step(std::integral_constant<int, 1>{}, k + 1);
    if (sync_period > 0 && ((k >> 1) + 1) % sync_period == 0)
      asm volatile("bar.sync 1, %0;" ::"r"(live_threads) : "memory");
  }
  if (use_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (trace && lane == 0) {
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %warpid;" : "=r"(wid));
    unsigned long long* r = trace + 4 * (traj0 >> 5);
    r[0] = t_entry;
    r[1] = globaltimer();
    r[2] = smid;
    r[3] = wid;
  }
  if (live) {
    ll -= T(0.5) * log(sprod);
    if (sbad) ll = T(NAN);
    if (status == 0 && !isfinite(ll)) status = 1;
    if (a.out[CDK_OUT_LL]) static_cast<T*>(a.out[CDK_OUT_LL])[traj] = ll;
    if (a.out[CDK_OUT_STATUS]) static_cast<int*>(a.out[CDK_OUT_STATUS])[traj] = status;
  }
}

// ======================================================================================================================
// Variant C: TRAJECTORY POOL PER WARP.  The two structural losses of variant B are (i) lock-step gaps -- every gap of a
// warp's 32 trajectories runs to the warp-wide maximum substep count (6 on the benchmark grid against a mean of 4.5) --
// and (ii) 2,048 warps over 592 sub-partitions: 272 of them hold four warps, the rest three, and the kernel ends with
// the slowest.  Here a CTA is 12 warps (three per sub-partition, one CTA per SM) and a warp owns a POOL of up to 40
// trajectories (37 on the benchmark: 148 x 12 x 37 >= 65,536) whose canonical state lives in its private shared memory.
// The warp is a tiny scheduler: a lane holds one trajectory for the length of one gap, integrates it one RK substep per
// loop iteration and hands it back; trajectories whose gap is complete queue up (FIFO, so nobody falls behind) and are
// processed -- `thresh` or more at a time -- by an UPDATE PASS in which lane i takes the i-th waiting trajectory at
// whatever observation index it has reached (measurement update, log-likelihood, output rows), after which free lanes
// pick them up again.  Lanes never idle through somebody else's longer gap, nothing ever synchronises across warps, and
// the per-trajectory arithmetic is exactly that of variant B (bit-identical results).
// Outputs go straight from registers to HBM in the update pass: a lane writes the prediction row of step k-1 and the
// filtered row of step k of ITS trajectory with 128-bit stores (rows are 24 / 72 bytes, so at most one 8-byte store per
// row); the 126 MB L2 merges the rows of consecutive steps into full lines before they reach DRAM.
// Observations: y_{k+1} and t_{k+2} are prefetched with cp.async into a two-deep per-trajectory ring during pass k.
// ======================================================================================================================
constexpr int PL_WPC = 12;  // warps per CTA
constexpr int PL_PP = 40;   // largest pool per warp

template <int NX, int NY>
struct alignas(16) PoolSmem {
  double cm[NX][PL_PP];                                   // canonical mean, structure-of-arrays
  double cP[NX * (NX + 1) / 2][PL_PP];                    // canonical covariance (packed upper triangle)
  double ct0[PL_PP], ct1[PL_PP], cll[PL_PP], csp[PL_PP];  // gap start / end, log-likelihood, running product of S
  double py[2][NY][PL_PP], pt[2][PL_PP];                  // prefetched y_k and t_{k+1}, ring slot k & 1
  int ck[PL_PP];                                          // next observation index
  int cflag[PL_PP];                                       // bits 0..1 status, bit 2 "S was not positive"
  unsigned char rq[64], nq[64];                           // FIFO rings: ready for a lane / waiting for an update pass
};

// One output row (LEN doubles at row index `row` of a [rows][LEN] array whose base is 16-byte aligned): 128-bit stores
// wherever the address allows (row * LEN * 8 is 16-byte aligned iff row is even for odd LEN).
template <int LEN>
__device__ __forceinline__ void store_row_f64(double* __restrict__ base, long long row, const double (&v)[LEN]) {
  double* p = base + row * LEN;
  if ((LEN & 1) == 0 || (row & 1) == 0) {
#pragma unroll
    for (int e = 0; e + 1 < LEN; e += 2) *reinterpret_cast<double2*>(p + e) = make_double2(v[e], v[e + 1]);
    if (LEN & 1) p[LEN - 1] = v[LEN - 1];
  } else {
    p[0] = v[0];
#pragma unroll
    for (int e = 1; e + 1 < LEN; e += 2) *reinterpret_cast<double2*>(p + e) = make_double2(v[e], v[e + 1]);
  }
}

template <class Drift, int NY, int SOLVER>
__global__ void __launch_bounds__(32 * PL_WPC, 1)
    ekf_small_pool(const KArgs<double> a, const int pool, const int thresh, const int vec_out) {
  using T = double;
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  using S = PoolSmem<NX, NY>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  S& sm = *reinterpret_cast<S*>(smem_raw + (size_t)warp * sizeof(S));
  const long long N = a.d.N;
  const int K = a.d.K;
  const long long gw = (long long)blockIdx.x * PL_WPC + warp;
  const long long traj0 = gw * pool;
  if (traj0 >= N) return;
  const int np = (int)((N - traj0) < pool ? (N - traj0) : pool);  // trajectories in this warp's pool
  unsigned long long* const trace = g_lw_trace;
  const unsigned long long t_entry = trace ? globaltimer() : 0ull;

  // model constants: shared by all trajectories (per-trajectory parameter blocks take variant B), kept in registers
  T th[NTH], lql[NP], Hs[NY * NX], ds[NY], Rs[NY * NY];
  {
    const T* thg = a.in[CDK_IN_F];
    const T* Lm = a.in[CDK_IN_L];
    const T* Qc = a.in[CDK_IN_QC];
#pragma unroll
    for (int i = 0; i < NTH; ++i) th[i] = thg[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) {
        T acc = T(0);
        for (int p = 0; p < NX; ++p) {
          T lq = T(0);
          for (int q = 0; q < NX; ++q) lq += Lm[i * NX + q] * Qc[q * NX + p];
          acc += lq * Lm[j * NX + p];
        }
        lql[pidx<NX>(i, j)] = acc;
      }
#pragma unroll
    for (int i = 0; i < NY * NX; ++i) Hs[i] = a.in[CDK_IN_H][i];
#pragma unroll
    for (int i = 0; i < NY; ++i) ds[i] = a.in[CDK_IN_D][i];
#pragma unroll
    for (int i = 0; i < NY * NY; ++i) Rs[i] = a.in[CDK_IN_R][i];
  }
  const T* __restrict__ Yg = a.in[CDK_IN_Y] + traj0 * a.in_stride[CDK_IN_Y];
  const T* __restrict__ Tg = a.in[CDK_IN_T] + traj0 * a.in_stride[CDK_IN_T];
  const long long ystride = a.in_stride[CDK_IN_Y], tstride = a.in_stride[CDK_IN_T];
  // canonical state of every pool member = the prior; all of them queue for their first update
  for (int u = lane; u < np; u += 32) {
    const long long tj = traj0 + u;
    const T* m0 = a.in[CDK_IN_M0] + tj * a.in_stride[CDK_IN_M0];
    const T* P0 = a.in[CDK_IN_P0] + tj * a.in_stride[CDK_IN_P0];
#pragma unroll
    for (int i = 0; i < NX; ++i) sm.cm[i][u] = m0[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) sm.cP[pidx<NX>(i, j)][u] = P0[i * NX + j];
    sm.ct0[u] = T(0);
    sm.ct1[u] = Tg[u * tstride];  // "end of the previous gap" of step 0 = t_0
    sm.cll[u] = T(0);
    sm.csp[u] = T(1);
#pragma unroll
    for (int c = 0; c < NY; ++c) sm.py[0][c][u] = Yg[u * ystride + c];
    sm.pt[0][u] = K > 1 ? Tg[u * tstride + 1] : T(0);
    sm.ck[u] = 0;
    sm.cflag[u] = 0;
    sm.nq[u] = (unsigned char)u;
  }
  __syncwarp();
  T* const FM = static_cast<T*>(a.out[CDK_OUT_FM]);
  T* const FP = static_cast<T*>(a.out[CDK_OUT_FP]);
  T* const PM = static_cast<T*>(a.out[CDK_OUT_PM]);
  T* const PP = static_cast<T*>(a.out[CDK_OUT_PP]);
  T* const LLC = static_cast<T*>(a.out[CDK_OUT_LLCUM]);
  const T dt0 = T(a.d.dt0), dtf = T(a.d.dt_final), tol = clip_tol<T>();
  const int max_steps = a.d.max_steps, num_iter = a.d.num_iter;

  // FIFO rings (indices are warp-uniform registers, entries live in shared memory)
  unsigned nq_head = 0, nq_cnt = (unsigned)np;  // waiting for an update pass
  unsigned rq_head = 0, rq_cnt = 0;             // updated, waiting for a lane
  int slot = -1;                                // pool member this lane integrates, or -1
  int nsteps = 0;
  St<T, NX> s;
  T tprev = T(0), tnext = T(0), t1 = T(0);
#pragma unroll
  for (int i = 0; i < NX; ++i) s.m[i] = T(0);
#pragma unroll
  for (int i = 0; i < NP; ++i) s.P[i] = T(0);
  const unsigned lt_mask = (1u << lane) - 1u;

  for (;;) {
    // ---- (A) free lanes pick up ready trajectories, oldest first ----
    if (rq_cnt != 0u) {
      const unsigned free_m = __ballot_sync(0xffffffffu, slot < 0);
      if (free_m != 0u) {
        const unsigned nfree = __popc(free_m);
        const unsigned ntake = nfree < rq_cnt ? nfree : rq_cnt;
        const unsigned myrank = __popc(free_m & lt_mask);
        if (slot < 0 && myrank < ntake) {
          const int mine = sm.rq[(rq_head + myrank) & 63u];
          slot = mine;
#pragma unroll
          for (int i = 0; i < NX; ++i) s.m[i] = sm.cm[i][mine];
#pragma unroll
          for (int i = 0; i < NP; ++i) s.P[i] = sm.cP[i][mine];
          tprev = sm.ct0[mine];
          t1 = sm.ct1[mine];
          tnext = fmin(tprev + dt0, t1);
          nsteps = 0;
        }
        rq_head += ntake;
        rq_cnt -= ntake;
      }
    }
    const unsigned act_m = __ballot_sync(0xffffffffu, slot >= 0);
    if (nq_cnt >= (unsigned)thresh || (act_m == 0u && nq_cnt != 0u)) {
      // ---- (C) update pass: lane i takes the i-th waiting trajectory ----
      const unsigned ntake = nq_cnt < 32u ? nq_cnt : 32u;
      asm volatile("cp.async.wait_group 0;" ::: "memory");  // observation prefetches issued by earlier passes have landed
      __syncwarp();
      const bool has = (unsigned)lane < ntake;
      bool again = false;  // the trajectory goes back to the ready ring (it has observations left)
      int us = -1;
      if (has) {
        us = sm.nq[(nq_head + lane) & 63u];
        const long long tj = traj0 + us;
        const int k = sm.ck[us];
        St<T, NX> u;
#pragma unroll
        for (int i = 0; i < NX; ++i) u.m[i] = sm.cm[i][us];
#pragma unroll
        for (int i = 0; i < NP; ++i) u.P[i] = sm.cP[i][us];
        const long long row = tj * (long long)K + k;
        if (vec_out && k > 0) {  // the state before the update is the prediction made at step k - 1
          T full[NX * NX];
#pragma unroll
          for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) full[i * NX + j] = u.P[pidx<NX>(i, j)];
          if (PM) store_row_f64<NX>(PM, row - 1, u.m);
          if (PP) store_row_f64<NX * NX>(PP, row - 1, full);
        }
        T ll = sm.cll[us];
        int flag = sm.cflag[us];
        if (k < K) {
          T y[NY];
#pragma unroll
          for (int c = 0; c < NY; ++c) y[c] = sm.py[k & 1][c][us];
          const T tk = sm.ct1[us];  // t_k is the end of the previous gap, bit for bit
          const T tk1 = k + 1 < K ? sm.pt[k & 1][us] : tk + dtf;
          if (k + 1 < K) {  // prefetch y_{k+1}, t_{k+2} into the other ring slot
#pragma unroll
            for (int c = 0; c < NY; ++c)
              cp_async_elem(&sm.py[(k + 1) & 1][c][us], Yg + us * ystride + (long long)(k + 1) * NY + c);
            if (k + 2 < K) cp_async_elem(&sm.pt[(k + 1) & 1][us], Tg + us * tstride + k + 2);
          }
          T sprod = sm.csp[us];
          if (NY == 1 && !LLC) {
            T Sk;
            ll += ekf_update<T, NX, NY>(Hs, ds, Rs, u, y, num_iter, &Sk);
            if (!(Sk > T(0))) flag |= 4;
            if (Sk > T(1e-8) && Sk < T(1e8)) {
              sprod *= Sk;
            } else {
              ll -= T(0.5) * log(Sk);
            }
            if ((k & 7) == 7) {
              ll -= T(0.5) * log(sprod);
              sprod = T(1);
            }
          } else {
            ll += ekf_update<T, NX, NY>(Hs, ds, Rs, u, y, num_iter);
            if (LLC) LLC[row] = ll;
          }
          if (vec_out) {
            T full[NX * NX];
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
              for (int j = 0; j < NX; ++j) full[i * NX + j] = u.P[pidx<NX>(i, j)];
            if (FM) store_row_f64<NX>(FM, row, u.m);
            if (FP) store_row_f64<NX * NX>(FP, row, full);
          }
#pragma unroll
          for (int i = 0; i < NX; ++i) sm.cm[i][us] = u.m[i];
#pragma unroll
          for (int i = 0; i < NP; ++i) sm.cP[i][us] = u.P[i];
          sm.ct0[us] = tk;
          sm.ct1[us] = tk1;
          sm.cll[us] = ll;
          sm.csp[us] = sprod;
          sm.ck[us] = k + 1;
          sm.cflag[us] = flag;
        } else {
          // k == K: the last prediction has been written above; finish the trajectory
          ll -= T(0.5) * log(sm.csp[us]);
          if (flag & 4) ll = T(NAN);
          int status = flag & 3;
          if (status == 0 && !isfinite(ll)) status = 1;
          if (a.out[CDK_OUT_LL]) static_cast<T*>(a.out[CDK_OUT_LL])[tj] = ll;
          if (a.out[CDK_OUT_STATUS]) static_cast<int*>(a.out[CDK_OUT_STATUS])[tj] = status;
        }
        again = k < K;
      }
      cp_async_commit();
      const unsigned again_m = __ballot_sync(0xffffffffu, again);  // push onto the ready ring in pass order
      if (again) sm.rq[(rq_head + rq_cnt + __popc(again_m & lt_mask)) & 63u] = (unsigned char)us;
      rq_cnt += __popc(again_m);
      nq_head += ntake;
      nq_cnt -= ntake;
      __syncwarp();
      continue;
    }
    if (act_m == 0u) break;  // nothing integrating, nothing waiting: the pool is finished
    // ---- (B) one RK substep for every lane that holds a trajectory ----
    bool fin = false;
    if (slot >= 0) {
      if (tprev < t1 && nsteps < max_steps) {
        rk_step<T, Drift, SOLVER>(th, lql, s, tnext - tprev);
        ++nsteps;
        tprev = tnext;
        const T cand = tprev + dt0;
        tnext = cand > t1 - tol ? t1 : cand;
      }
      fin = !(tprev < t1) || nsteps >= max_steps;
    }
    const unsigned fin_m = __ballot_sync(0xffffffffu, fin);
    if (fin_m != 0u) {
      if (fin) {
        if (tprev < t1) {  // diffrax max_steps exceeded: the reference result is NaN
          sm.cflag[slot] = (sm.cflag[slot] & ~3) | 2;
#pragma unroll
          for (int i = 0; i < NX; ++i) s.m[i] = T(NAN);
#pragma unroll
          for (int i = 0; i < NP; ++i) s.P[i] = T(NAN);
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) sm.cm[i][slot] = s.m[i];
#pragma unroll
        for (int i = 0; i < NP; ++i) sm.cP[i][slot] = s.P[i];
        sm.nq[(nq_head + nq_cnt + __popc(fin_m & lt_mask)) & 63u] = (unsigned char)slot;
        slot = -1;
      }
      nq_cnt += __popc(fin_m);
      __syncwarp();
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (trace && lane == 0) {
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %warpid;" : "=r"(wid));
    unsigned long long* r = trace + 4 * gw;
    r[0] = t_entry;
    r[1] = globaltimer();
    r[2] = smid;
    r[3] = wid;
  }
}

// ======================================================================================================================
// EKS backward pass for the register-sized drifts (extended_kalman_smoother, inference_ekf.py:450-539 with _smooth
// :363-448): for k = K-2 .. 0, with the Jacobian and the drift frozen at the filtered mean m_f,
//   aux = psd_solve(P_f, L Qc L^T)^T,  G = J(m_f) + aux,
//   d/ds (m_s, P_s) = -( f(m_f) + G (m_s - m_f),  G P_s + P_s G^T - L Qc L^T ),  s in [0, t_{k+1} - t_k]  (reverse_rhs,
//   diffrax_utils.py:13-25), integrated with the caller's fixed-step solver from the smoothed moments of step k + 1.
// Same mapping as ekf_small_lw: one warp = 32 trajectories for the whole kernel, smoothed state, G and the frozen terms in
// registers, gaps run to the warp-wide maximum substep count, warps never synchronise with each other.  The filtered
// moments are read back from HBM through a private 3-deep cp.async ring (12 + 1 values per lane and step, issued two
// steps ahead); the smoothed rows leave through the same two-step staging blocks and TMA tensor stores as the filter's.
// ======================================================================================================================
constexpr int EK_RING = 3;

template <typename T, int NX>
struct alignas(128) EKSmem {
  T sm[32][2][NX];
  alignas(128) T sp[32][2][NX * NX];
  T inM[EK_RING][NX][32];
  T inP[EK_RING][NX * NX][32];
  T inT[EK_RING][32];
};

// generic explicit RK step y <- y + dt * sum_i b_i rhs(y_i) (same accumulation scheme as rk_step)
template <typename T, int NX, int SOLVER, class RHS>
__device__ __forceinline__ void rk_step_fn(St<T, NX>& y, T dt, RHS rhs) {
  using TB = Tab<SOLVER>;
  constexpr int NP = St<T, NX>::NP;
  constexpr int I0 = TB::b(0) != 0.0 ? 0 : (TB::S > 1 && TB::b(1) != 0.0 ? 1 : (TB::S > 2 && TB::b(2) != 0.0 ? 2 : 3));
  St<T, NX> k[TB::S];
  St<T, NX> ksum;
#pragma unroll
  for (int i = 0; i < TB::S; ++i) {
    St<T, NX> yi = y;
#pragma unroll
    for (int j = 0; j < i; ++j) {
      if (TB::a(i, j) != 0.0) {
        const T c = T(TB::a(i, j)) * dt;
#pragma unroll
        for (int e = 0; e < NX; ++e) yi.m[e] = fma(c, k[j].m[e], yi.m[e]);
#pragma unroll
        for (int e = 0; e < NP; ++e) yi.P[e] = fma(c, k[j].P[e], yi.P[e]);
      }
    }
    rhs(yi, k[i]);
    if (i == I0) {
      ksum = k[i];
    } else if (TB::b(i) != 0.0) {
      const T c = T(TB::b(i) / TB::b(I0));
      if (TB::b(i) == TB::b(I0)) {
#pragma unroll
        for (int e = 0; e < NX; ++e) ksum.m[e] += k[i].m[e];
#pragma unroll
        for (int e = 0; e < NP; ++e) ksum.P[e] += k[i].P[e];
      } else {
#pragma unroll
        for (int e = 0; e < NX; ++e) ksum.m[e] = fma(c, k[i].m[e], ksum.m[e]);
#pragma unroll
        for (int e = 0; e < NP; ++e) ksum.P[e] = fma(c, k[i].P[e], ksum.P[e]);
      }
    }
  }
  const T w = T(TB::b(I0)) * dt;
#pragma unroll
  for (int e = 0; e < NX; ++e) y.m[e] = fma(w, ksum.m[e], y.m[e]);
#pragma unroll
  for (int e = 0; e < NP; ++e) y.P[e] = fma(w, ksum.P[e], y.P[e]);
}

template <typename T, class Drift, int SOLVER>
__global__ void __launch_bounds__(32 * 14, 1) eks_small_lw(const KArgs<T> a, const __grid_constant__ V5Maps maps) {
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  constexpr int WPC = 14;
  using S = EKSmem<T, NX>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  S& sm = *reinterpret_cast<S*>(smem_raw + (size_t)warp * sizeof(S));
  const long long N = a.d.N;
  const int K = a.d.K;
  const long long traj0 = ((long long)blockIdx.x * WPC + warp) * 32;
  if (traj0 >= N) return;
  const long long traj = traj0 + lane;
  const bool live = traj < N;
  const int nlive = (int)((N - traj0) < 32 ? (N - traj0) : 32);
  const bool use_tma = sizeof(T) == 8 && maps.use_tma != 0;
  const long long tl = live ? traj : 0;

  // model constants, per lane (a leading N on any parameter is simply this lane's block)
  T th[NTH], lql[NP];
  {
    const T* thg = a.in[CDK_IN_F] + tl * a.in_stride[CDK_IN_F];
    const T* Lm = a.in[CDK_IN_L] + tl * a.in_stride[CDK_IN_L];
    const T* Qc = a.in[CDK_IN_QC] + tl * a.in_stride[CDK_IN_QC];
#pragma unroll
    for (int i = 0; i < NTH; ++i) th[i] = thg[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) {
        T acc = T(0);
        for (int p = 0; p < NX; ++p) {
          T lq = T(0);
          for (int q = 0; q < NX; ++q) lq += Lm[i * NX + q] * Qc[q * NX + p];
          acc += lq * Lm[j * NX + p];
        }
        lql[pidx<NX>(i, j)] = acc;
      }
  }
  const T* __restrict__ FMg = a.in[CDK_IN_FM] + tl * a.in_stride[CDK_IN_FM];
  const T* __restrict__ FPg = a.in[CDK_IN_FP] + tl * a.in_stride[CDK_IN_FP];
  const T* __restrict__ Tg = a.in[CDK_IN_T] + tl * a.in_stride[CDK_IN_T];
  T* const SMg = static_cast<T*>(a.out[CDK_OUT_SM]);
  T* const SPg = static_cast<T*>(a.out[CDK_OUT_SP]);
  auto prefetch = [&](int kk) {  // filtered moments and time stamp of step kk -> ring slot kk % EK_RING
    if (live && kk >= 0) {
      const int r = kk % EK_RING;
#pragma unroll
      for (int i = 0; i < NX; ++i) cp_async_elem(&sm.inM[r][i][lane], FMg + (long long)kk * NX + i);
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) cp_async_elem(&sm.inP[r][i][lane], FPg + (long long)kk * NX * NX + i);
      cp_async_elem(&sm.inT[r][lane], Tg + kk);
    }
    cp_async_commit();
  };
  prefetch(K - 1);
  prefetch(K - 2);
  prefetch(K - 3);
  const T dt0 = T(a.d.dt0), tol = clip_tol<T>();
  const int max_steps = a.d.max_steps;
  int status = 0;

  auto tma_store = [&](int k0) {  // lane 0: store the two-step block starting at the (even) step k0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_store_2d(&maps.m[0], &sm.sm[0][0][0], k0 * NX, (int)traj0);
    tma_store_2d(&maps.m[1], &sm.sp[0][0][0], k0 * NX * NX, (int)traj0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  };
  auto flush_generic = [&](int k0, int nrow) {  // fp32 / odd K / unaligned outputs: cooperative copy of rows [k0, k0 + nrow)
#pragma unroll
    for (int arr = 0; arr < 2; ++arr) {
      T* __restrict__ G = arr == 0 ? SMg : SPg;
      const int len = arr ? NX * NX : NX;
      const T* src = arr ? &sm.sp[0][0][0] : &sm.sm[0][0][0];
      const int per = nrow * len, off = (k0 & 1) * len;
      for (int u = lane; u < nlive * per; u += 32) {
        const int slot = u / per, e = u - slot * per;
        G[((traj0 + slot) * (long long)K + k0) * len + e] = src[slot * 2 * len + off + e];
      }
    }
    __syncwarp();
  };
  auto stage = [&](int row, const St<T, NX>& v) {
    T full[NX * NX];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) full[i * NX + j] = v.P[pidx<NX>(i, j)];
    if (row == 0) {
      stage_row<0, NX>(&sm.sm[lane][0][0], v.m);
      stage_row<0, NX * NX>(&sm.sp[lane][0][0], full);
    } else {
      stage_row<1, NX>(&sm.sm[lane][0][0], v.m);
      stage_row<1, NX * NX>(&sm.sp[lane][0][0], full);
    }
  };

  // step K-1: smoothed = filtered (:466-470 / :813-814 in the linear twin)
  St<T, NX> s;
  T t1 = T(0);
  asm volatile("cp.async.wait_group 2;" ::: "memory");
  __syncwarp();
  if (live) {
    const int r = (K - 1) % EK_RING;
#pragma unroll
    for (int i = 0; i < NX; ++i) s.m[i] = sm.inM[r][i][lane];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) s.P[pidx<NX>(i, j)] = sm.inP[r][i * NX + j][lane];
    t1 = sm.inT[r][lane];
    // the reference copies the filtered covariance verbatim: keep its lower triangle too
    T full[NX * NX];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) full[i] = sm.inP[r][i][lane];
    if (((K - 1) & 1) == 0) {
      stage_row<0, NX>(&sm.sm[lane][0][0], s.m);
      stage_row<0, NX * NX>(&sm.sp[lane][0][0], full);
    } else {
      stage_row<1, NX>(&sm.sm[lane][0][0], s.m);
      stage_row<1, NX * NX>(&sm.sp[lane][0][0], full);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NX; ++i) s.m[i] = T(0);
#pragma unroll
    for (int i = 0; i < NP; ++i) s.P[i] = T(0);
  }
  __syncwarp();
  if (((K - 1) & 1) == 0) {  // odd K: the last row is a block of its own (never the TMA path)
    flush_generic(K - 1, 1);
  }

  for (int k = K - 2; k >= 0; --k) {
    const int row = k & 1;
    prefetch(k - 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");  // own loads of step k have landed
    __syncwarp();
    if (live) {
      const int r = k % EK_RING;
      T mf[NX], Pf[NX][NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) mf[i] = sm.inM[r][i][lane];
#pragma unroll
      for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = 0; j < NX; ++j) Pf[i][j] = sm.inP[r][i * NX + j][lane];
      const T t0 = sm.inT[r][lane];
      // psd_solve(P_f, L Qc L^T): Cholesky of sym(P_f) + 1e-9 I (utils.py:202-207), three right-hand sides
      T Lc[NX][NX], inv[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        T sjj = T(0.5) * (Pf[j][j] + Pf[j][j]) + T(1e-9);
#pragma unroll
        for (int q = 0; q < j; ++q) sjj -= Lc[j][q] * Lc[j][q];
        const T dj = sqrt(sjj);
        Lc[j][j] = dj;
        inv[j] = T(1) / dj;
#pragma unroll
        for (int i = j + 1; i < NX; ++i) {
          T v = T(0.5) * (Pf[i][j] + Pf[j][i]);
#pragma unroll
          for (int q = 0; q < j; ++q) v -= Lc[i][q] * Lc[j][q];
          Lc[i][j] = v * inv[j];
        }
      }
      T G[NX][NX];
      Drift::jac(th, mf, G);
#pragma unroll
      for (int c = 0; c < NX; ++c) {  // column c of X = (P_f + eps I)^-1 L Qc L^T;  aux = X^T
        T w[NX], x[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          T v = lql[pidx<NX>(i, c)];
#pragma unroll
          for (int q = 0; q < i; ++q) v -= Lc[i][q] * w[q];
          w[i] = v * inv[i];
        }
#pragma unroll
        for (int i = NX - 1; i >= 0; --i) {
          T v = w[i];
#pragma unroll
          for (int q = i + 1; q < NX; ++q) v -= Lc[q][i] * x[q];
          x[i] = v * inv[i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) G[c][i] += x[i];
      }
      T c0[NX];
      Drift::f(th, mf, c0);
      auto rhs = [&](const St<T, NX>& y, St<T, NX>& kk) {
        T GP[NX][NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          T acc = c0[i];
#pragma unroll
          for (int q = 0; q < NX; ++q) acc = fma(G[i][q], y.m[q] - mf[q], acc);
          kk.m[i] = -acc;
#pragma unroll
          for (int j = 0; j < NX; ++j) {
            T g = T(0);
#pragma unroll
            for (int q = 0; q < NX; ++q) g = fma(G[i][q], y.P[pidx<NX>(q, j)], g);
            GP[i][j] = g;
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
          for (int j = i; j < NX; ++j) kk.P[pidx<NX>(i, j)] = lql[pidx<NX>(i, j)] - (GP[i][j] + GP[j][i]);
      };
      const T span = t1 - t0;  // integrate s from 0 to t_{k+1} - t_k
      T tprev = T(0), tnext = fmin(dt0, span);
      int nsteps = 0;
      while (tprev < span && nsteps < max_steps) {
        rk_step_fn<T, NX, SOLVER>(s, tnext - tprev, rhs);
        ++nsteps;
        tprev = tnext;
        const T cand = tprev + dt0;
        tnext = cand > span - tol ? span : cand;
      }
      if (tprev < span) {  // diffrax max_steps exceeded: NaN, as in the reference
        status = 2;
#pragma unroll
        for (int i = 0; i < NX; ++i) s.m[i] = T(NAN);
#pragma unroll
        for (int i = 0; i < NP; ++i) s.P[i] = T(NAN);
      }
      t1 = t0;
    }
    if (use_tma && row == 1) {  // the block (k-1, k) is about to be rewritten: its predecessor (k+1, k+2) must have been read
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    if (live) stage(row, s);
    __syncwarp();
    if (use_tma) {
      if (row == 0 && lane == 0) tma_store(k);
    } else if (row == 0 || k == 0) {
      const int nrow = (k + 1 < K && row == 0) ? 2 : 1;
      flush_generic(k, nrow);
    }
  }
  if (use_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  cp_async_wait_all();
  if (live && a.out[CDK_OUT_STATUS]) {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NX; ++i) bad |= !isfinite(s.m[i]);
    if (status == 0 && bad) status = 1;
    if (status != 0) static_cast<int*>(a.out[CDK_OUT_STATUS])[traj] = status;  // keep the filter's status otherwise
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = []() -> encode_tiled_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<encode_tiled_fn>(p);
  }();
  return fn;
}

// Build the four output tensor maps.  Returns false when the TMA path does not apply (fp32, odd K, unaligned pointers,
// CDK_EKF_TMA=0): the kernel then uses the cooperative-copy flush.
template <typename T>
bool make_maps(const KArgs<T>& a, int NX, int box_rows, V5Maps& maps, const int* slot_list = nullptr) {
  memset(&maps, 0, sizeof(maps));
  static const bool disabled = []() {
    const char* e = getenv("CDK_EKF_TMA");
    return e && e[0] == '0';
  }();
  if (disabled || sizeof(T) != 8 || (a.d.K & 1) || a.d.N > 0x7fffffffLL) return false;
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return false;
  const int default_slots[4] = {CDK_OUT_FM, CDK_OUT_FP, CDK_OUT_PM, CDK_OUT_PP};
  const int* slots = slot_list ? slot_list : default_slots;  // even entries: [N][K][NX] arrays, odd: [N][K][NX*NX]
  for (int i = 0; i < 4; ++i) {
    if (slots[i] < 0) continue;
    void* ptr = a.out[slots[i]];
    if (!ptr) continue;
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return false;
    const cuuint64_t len = (i & 1) ? NX * NX : NX;
    const cuuint64_t gdim[2] = {len * (cuuint64_t)a.d.K, (cuuint64_t)a.d.N};
    const cuuint64_t gstr[1] = {len * (cuuint64_t)a.d.K * sizeof(T)};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * len), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  maps.use_tma = 1;
  return true;
}

template <typename T, class Drift, int NY, int SOLVER>
int launch_one(const KArgs<T>& a, cudaStream_t s) {
  constexpr int NX = Drift::NX;
  constexpr int NPAR = Drift::NTHETA + NX * (NX + 1) / 2 + NY * NX + NY + NY * NY;
  using S = V5Smem<T, NX, NY>;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  const size_t smem = sizeof(S) + sizeof(T) * NPAR * (par_batched ? V5_W : 1);
  const long long blocks = (a.d.N + V5_W - 1) / V5_W;
  if (blocks == 0) return CDK_OK;
  if (blocks > 2147483647LL) return CDK_E_SIZE;
  // CDK_EKF_MODE=regroup : CTA-wide per-step regrouping (ekf_small_v5);  lockstep : same kernel, fixed assignment;
  //              warp    : independent warps (ekf_small_lw);  pool : trajectory pool per warp (ekf_small_pool).
  static const int mode = []() {
    const char* e = getenv("CDK_EKF_MODE");
    if (e && e[0] == 'r') return 0;
    if (e && e[0] == 'l') return 1;
    if (e && e[0] == 'w') return 2;
    if (e && e[0] == 'p') return 3;
    return CDK_EKF_DEFAULT_MODE;
  }();
  V5Maps maps;
  if constexpr (sizeof(T) == 8) {
    // Variant C (trajectory pool per warp): fp64, parameters shared by all trajectories, 16-byte aligned output bases
    // (128-bit row stores).
    const bool any_out = a.out[CDK_OUT_FM] || a.out[CDK_OUT_FP] || a.out[CDK_OUT_PM] || a.out[CDK_OUT_PP];
    bool aligned = true;
    for (int sl : {CDK_OUT_FM, CDK_OUT_FP, CDK_OUT_PM, CDK_OUT_PP})
      if (a.out[sl] && (reinterpret_cast<uintptr_t>(a.out[sl]) & 15)) aligned = false;
    if (mode == 3 && !par_batched && aligned) {
      using SP = PoolSmem<NX, NY>;
      static const int thresh = []() {
        const char* e = getenv("CDK_POOL_T");
        const int v = e ? atoi(e) : 16;
        return v < 1 ? 1 : (v > 32 ? 32 : v);
      }();
      int dev = 0, nsm = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      long long pool = (a.d.N + (long long)nsm * PL_WPC - 1) / ((long long)nsm * PL_WPC);  // one wave when it fits
      pool = pool < 32 ? 32 : (pool > PL_PP ? PL_PP : pool);
      const long long pblocks = (a.d.N + pool * PL_WPC - 1) / (pool * PL_WPC);
      if (pblocks > 2147483647LL) return CDK_E_SIZE;
      const size_t smp = sizeof(SP) * PL_WPC;
      auto kp = ekf_small_pool<Drift, NY, SOLVER>;
      if (smp > 48 * 1024 &&
          cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smp) != cudaSuccess)
        return check_launch("cudaFuncSetAttribute(ekf_small_pool)");
      kp<<<(unsigned)pblocks, 32 * PL_WPC, smp, s>>>(a, (int)pool, thresh, any_out ? 1 : 0);
      note_launch();
      return check_launch("ekf_small_pool");
    }
  }
  if (mode == 2 || mode == 3) {
    using SW = LWSmem<T, NX, NY>;
    static const int wpc_env = []() {
      const char* e = getenv("CDK_LW_WPC");
      return e && atoi(e) == 7 ? 7 : 14;
    }();
    static const int sync_period = []() {
      const char* e = getenv("CDK_LW_SYNC");
      return e ? atoi(e) : 0;
    }();
    const int warp_bytes = (int)((sizeof(SW) + sizeof(T) * NPAR * (par_batched ? 32 : 1) + 127) & ~size_t(127));
    const int wpc = (size_t)warp_bytes * wpc_env > 227 * 1024 ? 7 : wpc_env;  // per-lane parameter blocks: 2 CTAs of 7 warps
    const size_t smw = (size_t)warp_bytes * wpc;
    const long long wblocks = (a.d.N + 32 * wpc - 1) / (32 * wpc);
    if (wblocks > 2147483647LL) return CDK_E_SIZE;
    make_maps<T>(a, NX, 32, maps);
    auto kw = wpc == 7 ? ekf_small_lw<T, Drift, NY, SOLVER, 7> : ekf_small_lw<T, Drift, NY, SOLVER, 14>;
    if (smw > 48 * 1024) {
      if (cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw) != cudaSuccess)
        return check_launch("cudaFuncSetAttribute(ekf_small_lw)");
    }
    static const int token_env = []() {
      const char* e = getenv("CDK_LW_TOKEN");
      return e ? atoi(e) : 0;
    }();
    const int use_token = wpc == 14 ? token_env : 0;  // with two CTAs per SM the lock would have to span CTAs
    kw<<<(unsigned)wblocks, 32 * wpc, smw, s>>>(a, maps, warp_bytes, sync_period, use_token);
    note_launch();
    return check_launch("ekf_small_lw");
  }
  make_maps<T>(a, NX, V5_W, maps);
  auto kern = mode == 0 ? ekf_small_v5<T, Drift, NY, SOLVER, true> : ekf_small_v5<T, Drift, NY, SOLVER, false>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(ekf_small_v5)");
  }
  kern<<<(unsigned)blocks, V5_TPB, smem, s>>>(a, maps);
  note_launch();
  return check_launch("ekf_small_v5");
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

template <typename T, class Drift, int SOLVER>
int launch_eks_one(const KArgs<T>& a, cudaStream_t s) {
  constexpr int NX = Drift::NX;
  using S = EKSmem<T, NX>;
  const long long blocks = (a.d.N + 32 * 14 - 1) / (32 * 14);
  if (blocks == 0) return CDK_OK;
  if (blocks > 2147483647LL) return CDK_E_SIZE;
  V5Maps maps;
  const int slots[4] = {CDK_OUT_SM, CDK_OUT_SP, -1, -1};
  make_maps<T>(a, NX, 32, maps, slots);
  const size_t smem = sizeof(S) * 14;
  auto kern = eks_small_lw<T, Drift, SOLVER>;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("cudaFuncSetAttribute(eks_small_lw)");
  kern<<<(unsigned)blocks, 32 * 14, smem, s>>>(a, maps);
  note_launch();
  return check_launch("eks_small_lw");
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

int set_lw_trace(void* devbuf) {
  unsigned long long* p = static_cast<unsigned long long*>(devbuf);
  return cudaMemcpyToSymbol(g_lw_trace, &p, sizeof(p)) == cudaSuccess ? CDK_OK : CDK_E_CUDA;
}

// EKS backward pass fast path: Lorenz-63, solvers rk4 / dopri5 / euler / heun (anything else: generic_smooth_kernel).
template <typename T>
int launch_eks_small(const KArgs<T>& a, cudaStream_t s) {
  const cdk_desc& d = a.d;
  static const bool disabled = []() {
    const char* e = getenv("CDK_EKS_FAST");
    return e && e[0] == '0';
  }();
  if (disabled || d.drift_id != CDK_DRIFT_LORENZ63 || d.n != 3) return CDK_E_UNSUPPORTED;
  switch (d.solver) {
    case CDK_RK4: return launch_eks_one<T, DriftL63, CDK_RK4>(a, s);
    case CDK_DOPRI5: return launch_eks_one<T, DriftL63, CDK_DOPRI5>(a, s);
    case CDK_EULER: return launch_eks_one<T, DriftL63, CDK_EULER>(a, s);
    case CDK_HEUN: return launch_eks_one<T, DriftL63, CDK_HEUN>(a, s);
  }
  return CDK_E_UNSUPPORTED;
}

template int launch_eks_small<double>(const KArgs<double>&, cudaStream_t);
template int launch_eks_small<float>(const KArgs<float>&, cudaStream_t);
template int launch_ekf_small<double>(const KArgs<double>&, cudaStream_t);
template int launch_ekf_small<float>(const KArgs<float>&, cudaStream_t);

}  // namespace cdk
