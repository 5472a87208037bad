// FP64 tensor-core (mma.sync.m8n8k4.f64 = SASS DMMA) issue rate with shared versus distinct operand registers:
// MODE 0: every DMMA of a sweep reuses one (a, b) pair (what cdk_dmma_probe_f64 does); MODE 1: 8 accumulator tiles x
// distinct a and b registers per DMMA (a 2 x 4 outer-product block: 2 a-fragments x 4 b-fragments, the shape a register
// -blocked GEMM issues); MODE 2: fully distinct a and b per DMMA.  16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_regs scripts/micro/dmma_regs.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void k(double* sink, long long* cyc, int iters, const double* __restrict__ init) {
  double c[8][2], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = c[i][1] = 0.0;
    a[i] = init[threadIdx.x + i] + 1e-3 * i;
    b[i] = init[threadIdx.x + 8 + i] + 2e-3 * i;
  }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) dmma(c[i][0], c[i][1], a[0], b[0]);
      if (MODE == 1) dmma(c[i][0], c[i][1], a[i >> 2], b[i & 3]);
      if (MODE == 2) dmma(c[i][0], c[i][1], a[i], b[(i + 3) & 7]);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + a[i] + b[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int w, double* sink, long long* cyc, const double* init) {
  const int iters = 4000;
  for (int r = 0; r < 2; ++r) k<MODE><<<148, 32 * w>>>(sink, cyc, iters, init);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  const double per_smsp = w < 4 ? 1 : w / 4.0;
  printf("{\"mode\": %d, \"warps_per_sm\": %d, \"smsp_cycles_per_dmma\": %.2f}\n", MODE, w, m / (iters * 8.0) / per_smsp);
}

int main() {
  double *sink, *init;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * sizeof(double));
  cudaMalloc(&init, 4096 * sizeof(double));
  cudaMemset(init, 0, 4096 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int w : {4, 12, 16}) {
    run<0>(w, sink, cyc, init);
    run<1>(w, sink, cyc, init);
    run<2>(w, sink, cyc, init);
  }
  return 0;
}
