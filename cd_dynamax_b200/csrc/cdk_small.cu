// cdk_small.cu -- register-arithmetic CD-EKF for tiny state dimensions (Lorenz-63: m[3] + symmetric P[6]).
//
// Replaces extended_kalman_filter (src/continuous_discrete_nonlinear_gaussian_ssm/inference_ekf.py:202-326) with its
// _predict (:46-148, moment ODE dm = f(m), dP = F P + P F^T + L Qc L^T) and _condition_on (:153-199) for the
// registry drifts whose whole filter state fits in registers.
//
// B200 mapping (BASELINE config 3: N = 65,536, K = 1,000, n = 3, m = 1) -- kernel `ekf_small_v3`:
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
#include "cdk_common.cuh"

#include <stdlib.h>

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
  St<T, NX> k[TB::S];
  St<T, NX> acc = y;
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
    if (TB::b(i) != 0.0) {
      const T c = T(TB::b(i)) * dt;
#pragma unroll
      for (int e = 0; e < NX; ++e) acc.m[e] = fma(c, k[i].m[e], acc.m[e]);
#pragma unroll
      for (int e = 0; e < NP; ++e) acc.P[e] = fma(c, k[i].P[e], acc.P[e]);
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
      if (it == 0) ll = T(-0.5) * (r * r / S) - T(0.5) * log(S) - half_log_2pi<T>();
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
  }
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


// ---- cp.async (LDGSTS) element copies: sizeof(T) in {4, 8} is always naturally aligned -----------------------------
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int V3_W = 224;         // worker threads (7 warps) = trajectory slots per CTA
constexpr int V3_TPB = 256;       // + 1 I/O warp; 2 CTAs / SM (128 registers / thread)
constexpr int V3_LD = V3_W + 1;   // SoA row stride: odd, so the fields of one slot fall into distinct banks
constexpr int V3_RING = 4;        // input ring depth in observation steps
constexpr int V3_NB = 32;         // counting-sort buckets (substep count clamped to 1..31; 0 = dead slot)
constexpr int V3_SPL = V3_W / 32; // slots per I/O-warp lane

template <typename T, int NX, int NY>
struct V3Smem {
  static constexpr int NP = NX * (NX + 1) / 2;
  static constexpr int NF = 2 * (NX + NP);  // FM | FP | PM | PP (P packed)
  static constexpr int OFF_FM = 0, OFF_FP = NX, OFF_PM = NX + NP, OFF_PP = 2 * NX + NP;
  T stg[2][NF][V3_LD];  // canonical state + output staging, by step parity
  T inY[V3_RING][NY][V3_LD];
  T inT[V3_RING][V3_LD];
  T ll[V3_LD];
  int status[V3_W];
  int perm[2][V3_W];  // thread -> slot assignment, by step parity
  int hist[V3_NB];
  // followed by the model constants: NPAR values (shared) or V3_W * NPAR (one block per slot when batched)
};

// I/O warp: coalesced write of the staged outputs of step kk for every live slot.  Lane l owns a FIXED element of the
// row (l % len) and walks over the slots with a constant stride (32 / len), so each pass is LDS + STG + pointer bumps;
// consecutive lanes write consecutive addresses inside one trajectory's row.
template <typename T, int NX, int NY>
__device__ __forceinline__ void v3_flush(const V3Smem<T, NX, NY>& sm, void* const* out, int kk, int K, long long traj0,
                                         int nlive, int lane) {
  using S = V3Smem<T, NX, NY>;
  const T* base = &sm.stg[kk & 1][0][0];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    T* __restrict__ G = static_cast<T*>(out[a == 0 ? CDK_OUT_FM : a == 1 ? CDK_OUT_FP : a == 2 ? CDK_OUT_PM : CDK_OUT_PP]);
    if (!G) continue;
    const bool mat = (a & 1) != 0;
    const int len = mat ? NX * NX : NX;
    const int off = a == 0 ? S::OFF_FM : a == 1 ? S::OFF_FP : a == 2 ? S::OFF_PM : S::OFF_PP;
    const int stride = 32 / len;  // slots per pass
    if (lane >= stride * len) continue;
    const int slot0 = lane / len, el = lane - slot0 * len;
    const int field = off + (mat ? pidx<NX>(el / NX, el % NX) : el);
    const T* src = base + field * V3_LD + slot0;
    T* dst = G + ((traj0 + slot0) * (long long)K + kk) * len + el;
    const long long dstep = (long long)stride * K * len;
#pragma unroll 4
    for (int slot = slot0; slot < nlive; slot += stride) {
      *dst = *src;
      src += stride;
      dst += dstep;
    }
  }
}

template <typename T, class Drift, int NY, int SOLVER, bool REGROUP>
__global__ void __launch_bounds__(V3_TPB, 2) ekf_small_v3(const KArgs<T> a) {
  constexpr int NX = Drift::NX;
  constexpr int NP = St<T, NX>::NP;
  constexpr int NTH = Drift::NTHETA;
  constexpr int NPAR = NTH + NP + NY * NX + NY + NY * NY;  // theta | lql (packed) | H | d | R
  using S = V3Smem<T, NX, NY>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  T* parbase = reinterpret_cast<T*>(smem_raw + ((sizeof(S) + 15) & ~size_t(15)));

  const long long N = a.d.N;
  const int K = a.d.K;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool io_warp = tid >= V3_W;
  const long long traj0 = (long long)blockIdx.x * V3_W;
  const int nlive = (int)((N - traj0) < V3_W ? (N - traj0) : V3_W);
  const bool home_live = tid < nlive;  // worker thread whose home slot holds a trajectory
  const long long traj = traj0 + tid;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  const T* __restrict__ Ybase = a.in[CDK_IN_Y] + traj0 * a.in_stride[CDK_IN_Y];
  const T* __restrict__ Tbase = a.in[CDK_IN_T] + traj0 * a.in_stride[CDK_IN_T];
  const long long ystride = a.in_stride[CDK_IN_Y], tstride = a.in_stride[CDK_IN_T];
  const T dt0 = T(a.d.dt0);
  const T dtf = T(a.d.dt_final);
  const T inv_dt0 = T(1) / dt0;

  // ---- prologue: model constants, initial moments, first three input steps ----
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
  if (!io_warp) {
    sm.status[tid] = 0;
    sm.ll[tid] = T(0);
    sm.perm[0][tid] = tid;
    sm.perm[1][tid] = tid;
  }
  if (home_live) {
    const T* m0 = a.in[CDK_IN_M0] + traj * a.in_stride[CDK_IN_M0];
    const T* P0 = a.in[CDK_IN_P0] + traj * a.in_stride[CDK_IN_P0];
    // the prior is the prediction for t_0 (inference_ekf.py:320): it sits where step -1 would have left it
#pragma unroll
    for (int i = 0; i < NX; ++i) sm.stg[1][S::OFF_PM + i][tid] = m0[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) sm.stg[1][S::OFF_PP + pidx<NX>(i, j)][tid] = P0[i * NX + j];
    for (int kk = 0; kk < 3 && kk < K; ++kk) {
#pragma unroll
      for (int c = 0; c < NY; ++c) cp_async_elem(&sm.inY[kk][c][tid], Ybase + tid * ystride + (long long)kk * NY + c);
      cp_async_elem(&sm.inT[kk][tid], Tbase + tid * tstride + kk);
    }
    cp_async_commit();
    cp_async_wait_all();
  }
  __syncthreads();

  // ---- I/O warp helpers (one warp serves all V3_W slots: slot = lane + 32 r) ----
  auto io_prefetch = [&](int kk) {
    if (kk < K) {
#pragma unroll
      for (int r = 0; r < V3_SPL; ++r) {
        const int slot = lane + 32 * r;
        if (slot < nlive) {
#pragma unroll
          for (int c = 0; c < NY; ++c)
            cp_async_elem(&sm.inY[kk & (V3_RING - 1)][c][slot], Ybase + slot * ystride + (long long)kk * NY + c);
          cp_async_elem(&sm.inT[kk & (V3_RING - 1)][slot], Tbase + slot * tstride + kk);
        }
      }
    }
    cp_async_commit();
  };
  // counting sort of the slots by the substep count of the gap after observation kk -> perm[kk & 1]
  auto io_sort = [&](int kk) {
    sm.hist[lane] = 0;
    __syncwarp();
    int bkt[V3_SPL], pos[V3_SPL];
#pragma unroll
    for (int r = 0; r < V3_SPL; ++r) {
      const int slot = lane + 32 * r;
      int b = 0;
      if (slot < nlive) {
        const T t0 = sm.inT[kk & (V3_RING - 1)][slot];
        const T t1 = kk + 1 < K ? sm.inT[(kk + 1) & (V3_RING - 1)][slot] : t0 + dtf;
        const T q = ceil((t1 - t0) * inv_dt0);
        b = q > T(1) ? (q < T(V3_NB - 1) ? (int)q : V3_NB - 1) : 1;
      }
      const unsigned grp = __match_any_sync(0xffffffffu, b);
      const int leader = __ffs(grp) - 1;
      int base = 0;
      if (lane == leader) {
        base = sm.hist[b];
        sm.hist[b] = base + __popc(grp);
      }
      __syncwarp();
      base = __shfl_sync(0xffffffffu, base, leader);
      bkt[r] = b;
      pos[r] = base + __popc(grp & ((1u << lane) - 1u));
    }
    const int cnt = sm.hist[lane];
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - cnt;
#pragma unroll
    for (int r = 0; r < V3_SPL; ++r) {
      const int off = __shfl_sync(0xffffffffu, excl, bkt[r]);
      sm.perm[kk & 1][off + pos[r]] = lane + 32 * r;
    }
  };
  if (REGROUP && io_warp) io_sort(0);
  __syncthreads();

  const T tol = clip_tol<T>();
  const int max_steps = a.d.max_steps;
  const int num_iter = a.d.num_iter;
  T* __restrict__ LLC = static_cast<T*>(a.out[CDK_OUT_LLCUM]);

  for (int k = 0; k < K; ++k) {
    if (!io_warp) {
      // ---- worker: update at t_k, then integrate the gap t_k -> t_{k+1}, for the slot assigned to this thread ----
      const int p = REGROUP ? sm.perm[k & 1][tid] : tid;
      if (p < nlive) {
        const T* par = par_batched ? parbase + p * NPAR : parbase;
        const T* th = par;
        const T* lql = par + NTH;
        const T* Hs = par + NTH + NP;
        const T* ds = Hs + NY * NX;
        const T* Rs = ds + NY;
        const int prv = (k + 1) & 1, cur = k & 1, ring = k & (V3_RING - 1);
        St<T, NX> s;
#pragma unroll
        for (int i = 0; i < NX; ++i) s.m[i] = sm.stg[prv][S::OFF_PM + i][p];
#pragma unroll
        for (int i = 0; i < NP; ++i) s.P[i] = sm.stg[prv][S::OFF_PP + i][p];
        T y[NY];
#pragma unroll
        for (int c = 0; c < NY; ++c) y[c] = sm.inY[ring][c][p];
        T tprev = sm.inT[ring][p];
        const T t1 = k + 1 < K ? sm.inT[(k + 1) & (V3_RING - 1)][p] : tprev + dtf;
        const T ll = sm.ll[p] + ekf_update<T, NX, NY>(Hs, ds, Rs, s, y, num_iter);
        sm.ll[p] = ll;
        if (LLC) LLC[(traj0 + p) * (long long)K + k] = ll;
#pragma unroll
        for (int i = 0; i < NX; ++i) sm.stg[cur][S::OFF_FM + i][p] = s.m[i];
#pragma unroll
        for (int i = 0; i < NP; ++i) sm.stg[cur][S::OFF_FP + i][p] = s.P[i];
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
        for (int i = 0; i < NX; ++i) sm.stg[cur][S::OFF_PM + i][p] = s.m[i];
#pragma unroll
        for (int i = 0; i < NP; ++i) sm.stg[cur][S::OFF_PP + i][p] = s.P[i];
      }
    } else {
      // ---- I/O warp, overlapped with the workers: inputs for step k+3, assignment for step k+1, outputs of step k-1 ----
      cp_async_wait_all();  // the group of step k+2 was issued a full step ago
      io_prefetch(k + 3);   // ring slot of step k-1, no longer read by anyone
      if (REGROUP && k + 1 < K) io_sort(k + 1);
      if (k > 0) v3_flush<T, NX, NY>(sm, a.out, k - 1, K, traj0, nlive, lane);
    }
    __syncthreads();
  }
  if (io_warp) {
    v3_flush<T, NX, NY>(sm, a.out, K - 1, K, traj0, nlive, lane);
  } else if (home_live) {
    const T ll = sm.ll[tid];
    int status = sm.status[tid];
    if (status == 0 && !isfinite(ll)) status = 1;
    if (a.out[CDK_OUT_LL]) static_cast<T*>(a.out[CDK_OUT_LL])[traj] = ll;
    if (a.out[CDK_OUT_STATUS]) static_cast<int*>(a.out[CDK_OUT_STATUS])[traj] = status;
  }
}

template <typename T, class Drift, int NY, int SOLVER>
int launch_one(const KArgs<T>& a, cudaStream_t s) {
  constexpr int NX = Drift::NX;
  constexpr int NPAR = Drift::NTHETA + NX * (NX + 1) / 2 + NY * NX + NY + NY * NY;
  using S = V3Smem<T, NX, NY>;
  const uint32_t par_mask = (1u << CDK_IN_F) | (1u << CDK_IN_L) | (1u << CDK_IN_QC) | (1u << CDK_IN_H) |
                            (1u << CDK_IN_D) | (1u << CDK_IN_R);
  const bool par_batched = (a.d.batched_mask & par_mask) != 0;
  const size_t smem = ((sizeof(S) + 15) & ~size_t(15)) + sizeof(T) * NPAR * (par_batched ? V3_W : 1);
  const long long blocks = (a.d.N + V3_W - 1) / V3_W;
  if (blocks == 0) return CDK_OK;
  if (blocks > 2147483647LL) return CDK_E_SIZE;
  // CDK_EKF_REGROUP=0 keeps a fixed thread <-> trajectory assignment (profiling A/B); default is per-step regrouping
  static const bool regroup = []() {
    const char* e = getenv("CDK_EKF_REGROUP");
    return !(e && e[0] == '0');
  }();
  auto kern = regroup ? ekf_small_v3<T, Drift, NY, SOLVER, true> : ekf_small_v3<T, Drift, NY, SOLVER, false>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(ekf_small_v3)");
  }
  kern<<<(unsigned)blocks, V3_TPB, smem, s>>>(a);
  note_launch();
  return check_launch("ekf_small_v3");
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
