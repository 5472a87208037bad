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
// L = chol(A) (lower; reads the lower triangle of A; A != L). Upper triangle of L zeroed. Ends with a barrier.
template <typename T>
__device__ void chol(const T* A, T* L, int n, int ld, T boost) {
  FOR_T(e, n * ld) L[e] = T(0);
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) {
      T sjj = A[j * ld + j] + boost;
      for (int k = 0; k < j; ++k) sjj -= L[j * ld + k] * L[j * ld + k];
      const T dj = sqrt(sjj);
      if (i == j) {
        L[j * ld + j] = dj;
      } else {
        T v = A[i * ld + j];
        for (int k = 0; k < j; ++k) v -= L[i * ld + k] * L[j * ld + k];
        L[i * ld + j] = v / dj;
      }
    }
    __syncthreads();
  }
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

}  // namespace cdk
