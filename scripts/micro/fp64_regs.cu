// Does a DFMA with three DISTINCT vector-register operands issue at the same rate as one with uniform operands?
// (register-file bandwidth: 3 x 64-bit x 32 lanes per instruction).  Variants: fma(x, U, U), fma(x, y, U), fma(x, y, z),
// and dadd(x, y); 8 independent chains per thread, 1..4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp64_regs scripts/micro/fp64_regs.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* sink, long long* cyc, int iters, double a, double b, const double* __restrict__ init) {
  double x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = init[threadIdx.x + 32 * i];
    y[i] = init[threadIdx.x + 32 * i + 1] * 1e-9 + 1.0;
    z[i] = init[threadIdx.x + 32 * i + 2] * 1e-9;
  }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = fma(x[i], a, b);
        if (MODE == 1) x[i] = fma(x[i], y[i], b);
        if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);
        if (MODE == 3) x[i] = fma(y[(i + u) & 7], z[(i + 3 + u) & 7], x[i]);  // operands from other chains, no reuse pattern
        if (MODE == 4) x[i] = x[i] + y[i];
        if (MODE == 5) x[i] = x[i] * y[(i + u) & 7];
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int w, double* sink, long long* cyc, const double* init) {
  const int iters = 2000;
  for (int r = 0; r < 2; ++r) k<MODE><<<148, 32 * w>>>(sink, cyc, iters, 1.0000001, 1e-9, init);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  const double per = (double)iters * 32;
  const double ws = w < 4 ? 1 : w / 4.0;
  printf("{\"mode\": %d, \"warps_per_sm\": %d, \"smsp_cycles_per_fp64_instr\": %.3f}\n", MODE, w, m / per / ws);
}

int main() {
  double *sink, *init;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  cudaMalloc(&init, 4096 * sizeof(double));
  cudaMemset(init, 0, 4096 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int w : {4, 8, 16}) {
    run<0>(w, sink, cyc, init);
    run<1>(w, sink, cyc, init);
    run<2>(w, sink, cyc, init);
    run<3>(w, sink, cyc, init);
    run<4>(w, sink, cyc, init);
    run<5>(w, sink, cyc, init);
  }
  return 0;
}
