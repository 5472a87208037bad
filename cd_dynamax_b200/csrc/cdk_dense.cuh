// cdk_dense.cuh -- block-cooperative small dense helpers and the drift registry, shared by the shared-memory kernels.
#pragma once
#include "cdk_common.cuh"

namespace cdk {

#define FOR_T(i, cnt) for (int i = threadIdx.x; i < (cnt); i += blockDim.x)

__host__ __device__ inline int ldp(int c) { return c | 1; }


// ---- drift registry, element-wise, on an arbitrary accessor x(j) ---------------------------------------------------
template <typename T, class XF>
__device__ __forceinline__ T drift_f(int id, const T* th, int n, int i, XF x) {
  switch (id) {
    case CDK_DRIFT_LINEAR: {
      T s = T(0);
      for (int k = 0; k < n; ++k) s += th[i * n + k] * x(k);
      return s + th[n * n + i];
    }
    case CDK_DRIFT_LORENZ63:
      if (i == 0) return th[0] * (x(1) - x(0));
      if (i == 1) return x(0) * (th[1] - x(2)) - x(1);
      return x(0) * x(1) - th[2] * x(2);
    case CDK_DRIFT_LORENZ96: {
      const int ip = i + 1 == n ? 0 : i + 1, im1 = i == 0 ? n - 1 : i - 1, im2 = im1 == 0 ? n - 1 : im1 - 1;
      return (x(ip) - x(im2)) * x(im1) - x(i) + th[0];
    }
    default: {  // quadratic
      const T* B = th + n;
      const T* C = th + n + n * n;
      T s = th[i];
      for (int j = 0; j < n; ++j) {
        const T xj = x(j);
        T c = B[i * n + j];
        for (int k = 0; k < n; ++k) c += C[(i * n + j) * n + k] * x(k);
        s += c * xj;
      }
      return s;
    }
  }
}

template <typename T>
__device__ __forceinline__ T drift_jac(int id, const T* th, int n, int i, int j, const T* x) {
  switch (id) {
    case CDK_DRIFT_LINEAR: return th[i * n + j];
    case CDK_DRIFT_LORENZ63: {
      if (i == 0) return j == 0 ? -th[0] : (j == 1 ? th[0] : T(0));
      if (i == 1) return j == 0 ? th[1] - x[2] : (j == 1 ? T(-1) : -x[0]);
      return j == 0 ? x[1] : (j == 1 ? x[0] : -th[2]);
    }
    case CDK_DRIFT_LORENZ96: {
      const int ip = i + 1 == n ? 0 : i + 1, im1 = i == 0 ? n - 1 : i - 1, im2 = im1 == 0 ? n - 1 : im1 - 1;
      T v = T(0);
      if (j == ip) v += x[im1];
      if (j == im2) v -= x[im1];
      if (j == im1) v += x[ip] - x[im2];
      if (j == i) v -= T(1);
      return v;
    }
    default: {
      const T* B = th + n;
      const T* C = th + n + n * n;
      T s = B[i * n + j];
      for (int k = 0; k < n; ++k) s += (C[(i * n + j) * n + k] + C[(i * n + k) * n + j]) * x[k];
      return s;
    }
  }
}

// g_k = sum_i d2 f_i / dx_i dx_k: the only Hessian contraction the reference's 'second' order uses
// (0.5*jnp.trace(H_t @ P) traces axes (0,1): inference_ekf.py:111-114).  Zero unless quadratic.
template <typename T>
__device__ __forceinline__ T drift_graddiv(int id, const T* th, int n, int k) {
  if (id != CDK_DRIFT_QUADRATIC) return T(0);
  const T* C = th + n + n * n;
  T s = T(0);
  for (int i = 0; i < n; ++i) s += C[(i * n + i) * n + k] + C[(i * n + k) * n + i];
  return s;
}

// ---- block-cooperative dense helpers (all end WITHOUT a barrier unless noted) ---------------------------------------
// L = chol(A + boost I) (lower; reads the lower triangle of A; A != L). Upper triangle of L zeroed. Ends with a barrier.
// Left-looking and SINGLE-WARP: lane r owns rows r and r + 32 (n <= 64), columns are separated by __syncwarp() only, and
// the rest of the CTA waits at the one barrier at the end.  The critical path of a Cholesky is the dependent chain
// dot product -> sqrt -> divide of every column; the first version (one thread per row, two CTA barriers per column, a
// scalar k-loop whose every iteration waited for its own shared-memory load) spent ~1,400 cycles per column on it -- 62 %
// of the UKF's time, which factors P at every RK stage (profiles/r01_generic_ukf_n40.txt).  Here the k-loop is unrolled
// by four with all loads of a block issued before its FMAs, both rows of a lane and the pivot share the loads of row j,
// and every sum runs on two interleaved accumulators.  A non-positive pivot yields NaN (the reference's behaviour for
// non-PD input).
template <typename T>
__device__ void chol(const T* A, T* L, int n, int ld, T boost) {
  FOR_T(e, n * ld) {
    const int i = e / ld, j = e - i * ld;
    if (j > i) L[e] = T(0);
  }
  __syncthreads();  // A may have been produced by other threads just before the call
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int r0 = lane, r1 = lane + 32;
    const T* L0 = L + (r0 < n ? r0 : 0) * ld;  // rows outside the matrix alias row 0 and are discarded
    const T* L1 = L + (r1 < n ? r1 : 0) * ld;
    for (int j = 0; j < n; ++j) {
      const T* Lj = L + j * ld;
      T s0 = A[j * ld + j] + boost, s1 = T(0);
      T a0 = (r0 > j && r0 < n) ? A[r0 * ld + j] : T(0), a1 = T(0);
      T b0 = (r1 > j && r1 < n) ? A[r1 * ld + j] : T(0), b1 = T(0);
      int k = 0;
      for (; k + 3 < j; k += 4) {
        const T p0 = Lj[k], p1 = Lj[k + 1], p2 = Lj[k + 2], p3 = Lj[k + 3];
        const T x0 = L0[k], x1 = L0[k + 1], x2 = L0[k + 2], x3 = L0[k + 3];
        const T y0 = L1[k], y1 = L1[k + 1], y2 = L1[k + 2], y3 = L1[k + 3];
        s0 -= p0 * p0; s1 -= p1 * p1; s0 -= p2 * p2; s1 -= p3 * p3;
        a0 -= x0 * p0; a1 -= x1 * p1; a0 -= x2 * p2; a1 -= x3 * p3;
        b0 -= y0 * p0; b1 -= y1 * p1; b0 -= y2 * p2; b1 -= y3 * p3;
      }
      for (; k < j; ++k) {
        const T p0 = Lj[k];
        s0 -= p0 * p0;
        a0 -= L0[k] * p0;
        b0 -= L1[k] * p0;
      }
      // one rsqrt instead of sqrt + divide on the critical path (<= 2 ulp apart; NaN / inf propagate the same way)
      const T sjj = s0 + s1;
      const T rinv = rsqrt(sjj);
      if (r0 == j || r1 == j) L[j * ld + j] = sjj * rinv;
      if (r0 > j && r0 < n) L[r0 * ld + j] = (a0 + a1) * rinv;
      if (r1 > j && r1 < n) L[r1 * ld + j] = (b0 + b1) * rinv;
      __syncwarp();
    }
  }
  __syncthreads();
}

// Solve (L L^T) X = B in place, B is [n x c] with leading dimension ldb; one thread per column. Ends with a barrier.
template <typename T>
__device__ void chol_solve(const T* L, int n, int ld, T* B, int c, int ldb) {
  FOR_T(col, c) {
    for (int i = 0; i < n; ++i) {
      T v = B[i * ldb + col];
      for (int k = 0; k < i; ++k) v -= L[i * ld + k] * B[k * ldb + col];
      B[i * ldb + col] = v / L[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      T v = B[i * ldb + col];
      for (int k = i + 1; k < n; ++k) v -= L[k * ld + i] * B[k * ldb + col];
      B[i * ldb + col] = v / L[i * ld + i];
    }
  }
  __syncthreads();
}


// ---- block-cooperative small GEMM on the FP64 tensor cores ----------------------------------------------------------
// C(i, j) = sum_k opA(i, k) * opB(k, j), i < M, j < N, k < Kd, operands in shared memory (any T, converted to double in
// the fragment loads), result handed element-wise to `epi(i, j, double value)`.
//   TRANSA = false: opA(i, k) = A[i * lda + k];   true: opA(i, k) = A[k * lda + i]
//   TRANSB = false: opB(k, j) = B[k * ldb + j];   true: opB(k, j) = B[j * ldb + k]
// mma.sync.m8n8k4.f64 (SASS DMMA): each warp owns whole 8x8 output tiles and sweeps k in steps of 4 with two interleaved
// accumulator sets.  Fragment layout: A (8x4) lane -> (row lane/4, col lane%4); B (4x8) lane -> (row lane%4, col lane/4);
// C (8x8) lane -> (row lane/4, cols 2*(lane%4), +1).  Compared with one output element per thread (two shared-memory
// operand loads per FMA) this needs 2 loads per 256 FMAs and 1/16 of the instructions.  Every thread of the CTA must
// call it; no barrier inside (callers synchronise before the operands are read and after the epilogue writes).
template <typename T, bool TRANSA, bool TRANSB, class Epi>
__device__ __forceinline__ void mm_dmma(const T* __restrict__ A, int lda, const T* __restrict__ B, int ldb, int M, int N,
                                        int Kd, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int tm = (M + 7) >> 3, tn = (N + 7) >> 3;
  for (int t = warp; t < tm * tn; t += nwarp) {
    const int r0 = (t / tn) << 3, c0 = (t % tn) << 3;
    const int ia = r0 + gid, jb = c0 + gid;
    const bool va = ia < M, vb = jb < N;
    double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
    for (int k0 = 0; k0 < Kd; k0 += 8) {
      {
        const int k = k0 + tig;
        const bool vk = k < Kd;
        const double av = (va && vk) ? (double)(TRANSA ? A[k * lda + ia] : A[ia * lda + k]) : 0.0;
        const double bv = (vb && vk) ? (double)(TRANSB ? B[jb * ldb + k] : B[k * ldb + jb]) : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(d0), "+d"(d1)
                     : "d"(av), "d"(bv));
      }
      {
        const int k = k0 + 4 + tig;
        const bool vk = k < Kd;
        const double av = (va && vk) ? (double)(TRANSA ? A[k * lda + ia] : A[ia * lda + k]) : 0.0;
        const double bv = (vb && vk) ? (double)(TRANSB ? B[jb * ldb + k] : B[k * ldb + jb]) : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(e0), "+d"(e1)
                     : "d"(av), "d"(bv));
      }
    }
    const int i = r0 + gid, j = c0 + 2 * tig;
    if (i < M) {
      if (j < N) epi(i, j, d0 + e0);
      if (j + 1 < N) epi(i, j + 1, d1 + e1);
    }
  }
}

}  // namespace cdk
