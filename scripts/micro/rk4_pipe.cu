// How fast does the FP64 pipe run the REAL Lorenz-63 EKF RK4 substep (203 FP64 instructions, 3 distinct register operands
// each) with 1..4 warps per SM sub-partition and nothing else in the way?  Separates a structural instruction-mix ceiling
// from the latency of the measurement-update phase in ekf_small_lw.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -I cd_dynamax_b200/csrc \
//        -o scripts/micro/rk4_pipe scripts/micro/rk4_pipe.cu -lcuda
#include <cstdio>

#include "../../cd_dynamax_b200/csrc/cdk_small.cu"

namespace cdk {
__global__ void rk4_probe(double* sink, long long* cyc, int iters, double dt, unsigned lane_mask = 0xffffffffu) {
  St<double, 3> s;
  s.m[0] = 1.0 + 1e-3 * threadIdx.x; s.m[1] = 1.0; s.m[2] = 20.0;
  for (int i = 0; i < 6; ++i) s.P[i] = (i == 0 || i == 3 || i == 5) ? 1.0 : 0.0;
  double th[3] = {10.0, 28.0, 8.0 / 3.0};
  double lql[6] = {1, 0, 0, 1, 0, 1};
  __syncthreads();
  long long t0 = clock64();
  // lane_mask: does the FP64 pipe spend less time on a warp instruction when only half (a quarter) of the lanes are
  // active?  (It does not -- profiles/r02_micro_rk4_lanes.jsonl -- so spreading a small batch over more, emptier warps
  // buys nothing: the cost of this arithmetic is per WARP instruction.)
  if ((lane_mask >> (threadIdx.x & 31)) & 1u) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) rk_step<double, DriftL63, CDK_RK4>(th, lql, s, dt);
  }
  long long t1 = clock64();
  double acc = 0;
  for (int i = 0; i < 3; ++i) acc += s.m[i];
  for (int i = 0; i < 6; ++i) acc += s.P[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
}  // namespace cdk

int main() {
  double* sink;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4000;
  for (int w : {1, 4, 8, 12, 14, 16}) {
    for (int rep = 0; rep < 2; ++rep) cdk::rk4_probe<<<148, 32 * w>>>(sink, cyc, iters, 1e-4);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += h[i];
    m /= 148;
    const double per_smsp = w < 4 ? 1 : (w + 3) / 4;  // warps on the fullest sub-partition
    printf("{\"warps_per_sm\": %d, \"cycles_per_substep_per_warp\": %.1f, \"cycles_per_substep_fullest_smsp\": %.1f}\n", w,
           m / iters, m / iters / per_smsp);
  }
  struct { const char* name; unsigned mask; } masks[] = {{"32 lanes", 0xffffffffu}, {"lanes 0-15", 0x0000ffffu},
                                                         {"even lanes", 0x55555555u}, {"lanes 0-7", 0x000000ffu}};
  for (auto& mk : masks) {
    for (int rep = 0; rep < 2; ++rep) cdk::rk4_probe<<<148, 32 * 4>>>(sink, cyc, iters, 1e-4, mk.mask);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += h[i];
    printf("{\"active\": \"%s\", \"warps_per_sm\": 4, \"cycles_per_substep_per_warp\": %.1f}\n", mk.name, m / 148 / iters);
  }
  return 0;
}
// link stubs for the two host helpers cdk_small.cu expects from cdk_api.cu
namespace cdk {
int check_launch(const char*) { return 0; }
void note_launch() {}
}  // namespace cdk
