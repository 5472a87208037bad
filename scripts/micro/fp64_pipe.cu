// FP64 pipe micro-benchmark for sm_100a: cycles per warp-level DFMA as a function of warps per SM sub-partition and
// independent chains per thread (ILP), with and without interleaved integer instructions.  Informs the latency /
// occupancy model of the FP64-bound filter kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp64_pipe scripts/micro/fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIX>
__global__ void k(double* sink, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  int z = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
      if (MIX) {
#pragma unroll
        for (int i = 0; i < MIX; ++i) z = z * 3 + it;  // integer work that cannot be folded
      }
    }
  }
  long long t1 = clock64();
  double s = z;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int MIX>
void run(int warps_per_sm, double* sink, long long* cyc) {
  const int iters = 2000, blocks = 148;
  k<ILP, MIX><<<blocks, 32 * warps_per_sm>>>(sink, cyc, iters, 1.0000001, 1e-9);
  k<ILP, MIX><<<blocks, 32 * warps_per_sm>>>(sink, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  const double dfma_per_warp = (double)iters * 8 * ILP;
  const double per_smsp_warps = warps_per_sm / 4.0;
  printf("{\"ilp\": %d, \"mix\": %d, \"warps_per_sm\": %d, \"cycles_per_dfma_per_warp\": %.3f, \"smsp_cycles_per_dfma\": %.3f}\n",
         ILP, MIX, warps_per_sm, m / dfma_per_warp, m / (dfma_per_warp * (per_smsp_warps < 1 ? 1 : per_smsp_warps)));
}

int main() {
  double* sink;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int ws[] = {1, 4, 8, 12, 16, 32};
  for (int w : ws) {
    run<1, 0>(w, sink, cyc);
    run<2, 0>(w, sink, cyc);
    run<4, 0>(w, sink, cyc);
    run<8, 0>(w, sink, cyc);
    run<16, 0>(w, sink, cyc);
    run<8, 2>(w, sink, cyc);
    run<8, 4>(w, sink, cyc);
  }
  return 0;
}
