// ubench3.cu -- FP64 tensor-core (DMMA m8n8k4) probes for the QMF FIR:
//   (1) is the k-accumulation of one DMMA the sequential chain fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c))))?
//       (the reference adds the taps one by one in binary64; any other order changes the bits)
//   (2) how many clocks does a DMMA hold a sub-partition, alone and interleaved with integer work?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench3 ubench3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// order probe: A[8][4], B[4][8], C[8][8] per trial, results compared on the host
__global__ void order_probe(const double *A, const double *B, const double *C, double *D, int trials) {
  const int lane = threadIdx.x & 31;
  for (int t = blockIdx.x; t < trials; t += gridDim.x) {
    const double *a = A + t * 32, *b = B + t * 32, *c = C + t * 64;
    const int row = lane >> 2, k = lane & 3;
    double d0, d1;
    dmma(d0, d1, a[row * 4 + k], b[k * 8 + row], c[row * 8 + 2 * k], c[row * 8 + 2 * k + 1]);
    D[t * 64 + row * 8 + 2 * k] = d0;
    D[t * 64 + row * 8 + 2 * k + 1] = d1;
  }
}

template <int INT_PER_DMMA>
__global__ void __launch_bounds__(256) tput(double *out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = -c[i][0]; }
  unsigned x = threadIdx.x * 2654435761u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      dmma(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
#pragma unroll
      for (int j = 0; j < INT_PER_DMMA; j++) x = (x ^ (x >> 7)) + 0x9E3779B9u;
    }
  }
  double s = x;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void time_it(const char *name, K kernel, int sms, int ctas, double *d_out, int ints) {
  const int iters = 2048, blocks = sms * ctas;
  kernel<<<blocks, 256>>>(d_out, 8, 1.0000001, 0.999999);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kernel<<<blocks, 256>>>(d_out, iters, 1.0000001, 0.999999);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double dm = (double)ctas * 8 * iters * 8 / 4.0;  // DMMAs per sub-partition
  printf("%-24s %2d warps/SM  %8.3f ms  %6.2f clk per DMMA per SMSP  (%.1f TFLOP/s; %d dependent int ops between DMMAs)\n", name,
         ctas * 8, ms, ms * 1e-3 * 1.965e9 / dm, (double)blocks * 8 * iters * 8 * 512.0 / (ms * 1e-3) / 1e12, ints * 2);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // ---- (1) accumulation order
  const int trials = 200000;
  double *hA = (double *)malloc(trials * 32 * 8), *hB = (double *)malloc(trials * 32 * 8), *hC = (double *)malloc(trials * 64 * 8),
         *hD = (double *)malloc(trials * 64 * 8);
  srand(7);
  auto rnd = [](int spread) { return (double)(float)(((rand() / (double)RAND_MAX) - 0.5) * exp2((double)(rand() % spread - spread / 2))); };
  for (int i = 0; i < trials * 32; i++) { hA[i] = rnd(40); hB[i] = rnd(40); }
  for (int i = 0; i < trials * 64; i++) hC[i] = (rand() % 5 == 0) ? 0.0 : rnd(60) * 1.0000001192092896;
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, trials * 32 * 8); cudaMalloc(&dB, trials * 32 * 8); cudaMalloc(&dC, trials * 64 * 8); cudaMalloc(&dD, trials * 64 * 8);
  cudaMemcpy(dA, hA, trials * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, trials * 32 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, hC, trials * 64 * 8, cudaMemcpyHostToDevice);
  order_probe<<<sms * 4, 32>>>(dA, dB, dC, dD, trials);
  cudaMemcpy(hD, dD, trials * 64 * 8, cudaMemcpyDeviceToHost);
  long bad_fwd = 0, bad_rev = 0, bad_pair = 0, differ = 0;
  for (int t = 0; t < trials; t++)
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) {
        const double *a = hA + t * 32 + r * 4, *b = hB + t * 32;
        const double c0 = hC[t * 64 + r * 8 + c];
        double f = c0, rv = c0;
        for (int k = 0; k < 4; k++) f = __builtin_fma(a[k], b[k * 8 + c], f);
        for (int k = 3; k >= 0; k--) rv = __builtin_fma(a[k], b[k * 8 + c], rv);
        const double pr = (c0 + a[0] * b[c]) + (a[1] * b[8 + c]) + ((a[2] * b[16 + c]) + (a[3] * b[24 + c]));
        const double got = hD[t * 64 + r * 8 + c];
        bad_fwd += got != f; bad_rev += got != rv; bad_pair += got != pr; differ += f != rv;
      }
  printf("DMMA m8n8k4 vs host chains over %ld outputs: mismatches forward-chain %ld, reverse-chain %ld, pairwise %ld  "
         "(forward and reverse chains themselves differ on %ld)\n", (long)trials * 64, bad_fwd, bad_rev, bad_pair, differ);
  // ---- (2) throughput / issue behaviour
  double *d_out; cudaMalloc(&d_out, sizeof(double) * sms * 8 * 256);
  for (int c : {1, 2, 4, 8}) time_it("DMMA only", tput<0>, sms, c, d_out, 0);
  for (int c : {2, 8}) time_it("DMMA + 4 int", tput<2>, sms, c, d_out, 2);
  for (int c : {2, 8}) time_it("DMMA + 8 int", tput<4>, sms, c, d_out, 4);
  for (int c : {2, 8}) time_it("DMMA + 16 int", tput<8>, sms, c, d_out, 8);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
