// ubench2.cu -- what one FFT butterfly with its four binary32 roundings costs on B200, per
// rounding scheme, occupancy and per-thread ILP.  Prints SM clocks per warp-level butterfly per
// SM sub-partition (the FP64 pipe alone needs 18 * 2 = 36 for the magic-constant rounding,
// 22 * 2 = 44 for Veltkamp splitting, 10 * 2 = 20 with no rounding).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench2 ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Magic {
  int clamp;
  __device__ Magic() { asm("mov.u32 %0, 0x39E00000;" : "=r"(clamp)); }
  __device__ __forceinline__ double operator()(double v) const {
    const int h = __double2hiint(v);
    const int ce = max((h & 0x7FF00000) + 0x01D00000, clamp);
    const double c = __hiloint2double(ce, 0);
    return copysign((fabs(v) + c) - c, v);
  }
};
struct MagicNoClamp {
  __device__ __forceinline__ double operator()(double v) const {
    const int h = __double2hiint(v);
    const double c = __hiloint2double((h & 0xFFF00000) + 0x01D00000, 0);
    return (v + c) - c;
  }
};
struct Veltkamp {
  __device__ __forceinline__ double operator()(double v) const {
    const double g = v * 536870913.0;
    return g + (v - g);
  }
};
struct IntRne {
  __device__ __forceinline__ double operator()(double v) const {
    unsigned lo = __double2loint(v), hi = __double2hiint(v);
    const unsigned nlo = lo + 0x0FFFFFFFu + ((lo >> 29) & 1u);
    hi += (nlo < lo);
    return __hiloint2double(hi, nlo & 0xE0000000u);
  }
};
struct Cvt {
  __device__ __forceinline__ double operator()(double v) const { return (double)(float)v; }
};
struct None {
  __device__ __forceinline__ double operator()(double v) const { return v; }
};
// two roundings by the magic constant, one integer, one Veltkamp: spreads the work over pipes
struct Mixed {
  Magic m; IntRne i; Veltkamp w;
};

template <typename R, int ILP>
__global__ void __launch_bounds__(256) k(const double *in, double *out, const double2 *tw, int iters) {
  R rnd;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double a[ILP], b[ILP], c[ILP], d[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = in[(t + i) & 1023]; b[i] = in[(t + 3 * i + 1) & 1023]; c[i] = in[(t + 5 * i + 2) & 1023]; d[i] = in[(t + 7 * i + 3) & 1023]; }
  const double2 w = tw[threadIdx.x & 7];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const double tr = c[i] * w.x - d[i] * w.y, ti = c[i] * w.y + d[i] * w.x;
      const double na = rnd(a[i] + tr), nb = rnd(b[i] + ti), nc = rnd(a[i] - tr), nd = rnd(b[i] - ti);
      a[i] = nc; b[i] = nd; c[i] = na; d[i] = nb;   // keeps magnitudes bounded (|w| = 1, sum/difference alternate)
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i] + b[i] + c[i] + d[i];
  out[t] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) kmixed(const double *in, double *out, const double2 *tw, int iters) {
  Magic m; IntRne ir; Veltkamp v;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double a[ILP], b[ILP], c[ILP], d[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = in[(t + i) & 1023]; b[i] = in[(t + 3 * i + 1) & 1023]; c[i] = in[(t + 5 * i + 2) & 1023]; d[i] = in[(t + 7 * i + 3) & 1023]; }
  const double2 w = tw[threadIdx.x & 7];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const double tr = c[i] * w.x - d[i] * w.y, ti = c[i] * w.y + d[i] * w.x;
      const double na = m(a[i] + tr), nb = ir(b[i] + ti), nc = ir(a[i] - tr), nd = v(b[i] - ti);
      a[i] = nc; b[i] = nd; c[i] = na; d[i] = nb;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i] + b[i] + c[i] + d[i];
  out[t] = s;
}

static double *d_in, *d_out;
static double2 *d_tw;
static int sms;

template <typename K>
void time_it(const char *name, K kernel, int ilp, int ctas_per_sm) {
  const int iters = 512, blocks = sms * ctas_per_sm;
  kernel<<<blocks, 256>>>(d_in, d_out, d_tw, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kernel<<<blocks, 256>>>(d_in, d_out, d_tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-level butterflies per SM sub-partition
  const double wb = (double)ctas_per_sm * 8 /*warps*/ * iters * ilp / 4.0;
  printf("%-14s ILP %d  %2d warps/SM  %7.3f ms  %6.1f clk per warp-butterfly per SMSP (@1.965 GHz)\n", name, ilp,
         ctas_per_sm * 8, ms, ms * 1e-3 * 1.965e9 / wb);
}

#define RUN(R, ILP) do { for (int c : {1, 2, 3, 4, 8}) time_it(#R, k<R, ILP>, ILP, c); } while (0)

int main() {
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double h_in[1024]; double2 h_tw[8];
  for (int i = 0; i < 1024; i++) h_in[i] = (float)(0.001 * ((i * 7919) % 1000 - 500));
  for (int i = 0; i < 8; i++) { h_tw[i].x = cos(-0.3 - 0.7 * i); h_tw[i].y = sin(-0.3 - 0.7 * i); }
  cudaMalloc(&d_in, sizeof h_in); cudaMalloc(&d_tw, sizeof h_tw); cudaMalloc(&d_out, sizeof(double) * sms * 8 * 256);
  cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
  cudaMemcpy(d_tw, h_tw, sizeof h_tw, cudaMemcpyHostToDevice);
  printf("SMs: %d\n", sms);
  RUN(None, 4); RUN(Magic, 4); RUN(Magic, 8); RUN(MagicNoClamp, 4); RUN(Veltkamp, 4); RUN(Veltkamp, 8); RUN(IntRne, 4); RUN(Cvt, 4);
  for (int c : {1, 2, 3, 4, 8}) time_it("Mixed(m,i,i,v)", kmixed<4>, 4, c);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
