// Example user-defined drift (cd_dynamax_b200.LearnableUserDrift / build.build_user_drift): Van der Pol oscillator with a
// cubic restoring force.  Used by tests/test_gpu_user_drift.py; __graft_entry__.build() pre-builds its variant library.
namespace cdk_user {
// f0 = x1,  f1 = mu (1 - x0^2) x1 - x0 - eps x0^3;   theta = (mu, eps)
template <typename T, class XF> __device__ __forceinline__ T f(const T* th, int n, int i, XF x) {
  if (i == 0) return x(1);
  return th[0] * (T(1) - x(0) * x(0)) * x(1) - x(0) - th[1] * x(0) * x(0) * x(0);
}
template <typename T> __device__ __forceinline__ T jac(const T* th, int n, int i, int j, const T* x) {
  if (i == 0) return j == 1 ? T(1) : T(0);
  if (j == 0) return T(-2) * th[0] * x[0] * x[1] - T(1) - T(3) * th[1] * x[0] * x[0];
  return th[0] * (T(1) - x[0] * x[0]);
}
template <typename T> __device__ __forceinline__ T graddiv(const T* th, int n, int k, const T* x) {
  return k == 0 ? T(-2) * th[0] * x[0] : T(0);  // sum_i d2 f_i / dx_i dx_k: only d2 f1 / dx1 dx0 = -2 mu x0 is non-zero
}
}
