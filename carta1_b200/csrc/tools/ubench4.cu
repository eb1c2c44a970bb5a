// ubench4.cu -- dependent-DFMA latency on B200: how many independent accumulator chains per warp, and how many
// warps per SM sub-partition, keep the FP64 pipe busy (sizes the FIR loops of the QMF kernels: 8 chains in stage 1,
// 4 in stage 2, 4 warps per sub-partition).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench4 ubench4.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int ILP>
__global__ void k(double *out, double a, double b, int iters) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = a + threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8 / ILP; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(double *d_out, int sms, int warps_per_sm) {
  const int threads = 32 * warps_per_sm;  // one CTA per SM
  k<ILP><<<sms, threads>>>(d_out, 1.0000001, 1e-9, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP><<<sms, threads>>>(d_out, 1.0000001, 1e-9, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-DFMAs per sub-partition: warps_per_sm / 4 warps x ITERS x 8
  const double per_smsp = (double)warps_per_sm / 4.0 * ITERS * 8;
  printf("chains %d  warps/SMSP %4.1f  %7.3f ms  %6.2f clk per warp-DFMA per SMSP (@1.965 GHz)\n", ILP, warps_per_sm / 4.0, ms,
         ms * 1e-3 * 1.965e9 / per_smsp);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d_out;
  cudaMalloc(&d_out, sizeof(double) * sms * 1024);
  for (int w : {4, 8, 12, 16, 24, 32}) {
    run<1>(d_out, sms, w);
    run<2>(d_out, sms, w);
    run<4>(d_out, sms, w);
    run<8>(d_out, sms, w);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
