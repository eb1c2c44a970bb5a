// ubench.cu -- instruction-throughput microbenchmarks that size the FP64 / conversion
// ceiling of the bit-exact ATRAC1 kernels on B200 (DESIGN.md "FP64-pipe bound").
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) k(double *out, double a, double b, int iters) {
  double v[ILP];
  float f[ILP];
  unsigned u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { v[i] = a + threadIdx.x * 1e-3 + i; f[i] = (float)v[i]; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) v[i] = fma(v[i], a, b);                     // DFMA
      if (OP == 1) v[i] = v[i] + b;                             // DADD
      if (OP == 2) v[i] = v[i] * a;                             // DMUL
      if (OP == 3) { f[i] = (float)v[i]; v[i] = __hiloint2double(__double2hiint(v[i]) ^ __float_as_int(f[i]) & 1, __double2loint(v[i])); }  // F2F.F32.F64 + xor
      if (OP == 4) { v[i] = (double)f[i]; f[i] = __int_as_float(__float_as_int(f[i]) ^ (__double2loint(v[i]) & 1)); }  // F2F.F64.F32
      if (OP == 5) { v[i] = (double)(float)(v[i] * a); }        // DMUL + round trip f32
      if (OP == 6) {                                            // integer RNE rounding of a double to f32 precision
        unsigned lo = __double2loint(v[i]), hi = __double2hiint(v[i]);
        unsigned rb = (lo >> 29) & 1u;
        unsigned add = 0x0FFFFFFFu + rb;
        unsigned nlo = lo + add;
        hi += (nlo < lo);
        nlo &= 0xE0000000u;
        v[i] = __hiloint2double(hi, nlo) * a;
      }
      if (OP == 7) { u[i] = u[i] * 1664525u + 1013904223u; }   // IMAD
      if (OP == 8) { f[i] = fmaf(f[i], 1.0001f, 0.5f); }       // FFMA
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i] + f[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, double *d_out, int sms, double ops_per_iter) {
  const int blocks = sms * 8;
  k<OP><<<blocks, 256>>>(d_out, 1.0000001, 1e-9, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, 256>>>(d_out, 1.0000001, 1e-9, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * 256 * ITERS * ILP * ops_per_iter;
  printf("%-28s %8.3f ms  %8.2f Gop/s  %7.2f op/clk/SM @1.9GHz\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / sms / 1.9e9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d; cudaMalloc(&d, sizeof(double) * sms * 8 * 256);
  printf("SMs: %d\n", sms);
  run<0>("DFMA", d, sms, 1);
  run<1>("DADD", d, sms, 1);
  run<2>("DMUL", d, sms, 1);
  run<3>("F2F.F32.F64 (+2 int)", d, sms, 1);
  run<4>("F2F.F64.F32 (+2 int)", d, sms, 1);
  run<5>("DMUL + f32 round trip", d, sms, 1);
  run<6>("DMUL + int RNE rounding", d, sms, 1);
  run<7>("IMAD", d, sms, 1);
  run<8>("FFMA", d, sms, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
