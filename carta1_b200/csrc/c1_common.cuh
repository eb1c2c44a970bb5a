// c1_common.cuh -- shared device/host declarations for the carta1_b200 kernels.
//
// Numerical contract: "binary64 compute, binary32 store, no FMA".
// The whole library is compiled with -fmad=false so nvcc never contracts a*b+c; the only
// fused operations are the explicit fma() calls in the QMF kernels, where both factors are
// widened binary32 values and the product is therefore exact (Appendix A.2).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace c1 {

constexpr int kFrame = 512;
constexpr int kSuBytes = 212;
constexpr int kSuWords = 53;
constexpr int kNumBfu = 52;
constexpr int kFrameBits = 1696;

// Format tables (codec/core/constants.js:29-52,141-143), part of DevTables.
struct FormatTables {
  uint8_t specs[52];        // SPECS_PER_BFU
  uint16_t start_long[52];  // BFU_START_LONG
  uint16_t start_short[52]; // BFU_START_SHORT
  uint8_t bfu_of_long[512]; // inverse maps: coefficient position -> BFU index
  uint8_t bfu_of_short[512];
  uint8_t size_class[52];   // index of SPECS_PER_BFU[b] in {4,6,7,8,9,10,12,20}
  uint16_t bj_long[512];    // coefficient position -> (BFU << 5) | index inside the BFU
  uint16_t bj_short[512];
};

// libm-derived tables, uploaded once per context (global memory, read through L1).
struct DevTables {
  double qmf_even[24];  // QMF_EVEN / QMF_ODD widened to binary64 (exact)
  double qmf_odd[24];
  double win[32];       // WINDOW_SHORT
  double sf[64];        // SCALE_FACTORS
  double mdct_fwd64[32], mdct_fwd256[128], mdct_fwd512[256];
  double mdct_inv64[32], mdct_inv256[128], mdct_inv512[256];
  // FFT twiddles by the reference's recurrence (codec/transforms/fft.js:42-65): for half
  // stride h the k-th twiddle sits at index (h - 1 + k); (re, im) pairs.
  double2 fft_tw[255];
  float sf_thr[64];     // 63 thresholds of the exact findScaleFactor table (+1 pad)
  double log1p10;       // fdlibm log1p(10), the constant divisor of transient.js:211
  // quantize()'s normFactor = quantRange / scaleFactor (quantization.js:43-45) for word-length
  // index wl (bits = wl + 1, quantRange = 2^wl - 1) and scale-factor index: the IEEE division is
  // done once on the host
  double norm[16][64];
  double rcp_range[16];  // 1 / (2^wl - 1): dequantize()'s division by quantRange via div_by_range
  // dequantize() itself for the short word lengths (wl 1..6, i.e. 2..7 bits): deq_tab[deq_off(wl) +
  // (sfi << bits) + code] = f32((q * SF[sfi]) / (2^wl - 1)), code = q as a `bits`-bit two's-complement
  // field, every one of the 2^bits patterns (quantization.js:65-78; the row of sfi == 0 is +0).  The
  // IEEE multiplication and division are done on the host, once per context.
  float deq_tab[64 * 252];
  FormatTables fmt;
};

// first entry of word-length index wl (1..6) in DevTables::deq_tab: 64 * (4 + 8 + ... + 2^wl)
__host__ __device__ inline int deq_off(int wl) { return 64 * ((4 << (wl - 1)) - 4); }
constexpr int kDeqMaxWl = 6;

// Per-encoder parameters (EncoderOptions + tables derived from allocationBias).
struct DevEncParams {
  double threshold;
  double bsf[64];        // biased scale factors
  // zero_bit[sfi*8 + size_class] = f32((bsf[sfi] * 2.0) * size)   (bitallocation.js:86-87)
  float zero_bit[64 * 8];
  // 15-bit order-isomorphic images of the f32 heap priorities (bitallocation.js:226-230,
  // 266-269): key0 for word length 0, key1 for word length 1; each further step is -128.
  uint16_t key0[64], key1[64];
  int32_t use_fixed;
  int32_t fixed[3];
};

// ECMAScript ToInt32 (`| 0`, codec/coding/quantization.js:51).
__device__ __forceinline__ int js_to_int32(double x) {
  if (fabs(x) < 2147483648.0) return __double2int_rz(x);  // also false for NaN
  if (!isfinite(x)) return 0;
  double t = trunc(x);
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int)(unsigned int)(unsigned long long)m;
}

__device__ __forceinline__ int band_of_bfu(int b) { return b < 20 ? 0 : (b < 36 ? 1 : 2); }
__device__ __forceinline__ int band_of_coef(int c) { return c < 128 ? 0 : (c < 256 ? 1 : 2); }
__device__ __forceinline__ int wl_bits(int wl) { return wl == 0 ? 0 : wl + 1; }  // WORD_LENGTH_BITS


// Hint: bring the 128-byte line holding p into L2 (no register, no shared memory).  The persistent
// warps issue it for the inputs of their NEXT work item while they work on the current one, so that
// item's first loads pay an L2 hit instead of a DRAM round trip.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous, 16-byte aligned block from global into shared memory,
// completion signalled on an mbarrier: one lane arms the barrier with the byte count and issues one instruction for
// the whole block; the data goes through the TMA unit, not through the warp's LSU / L1 path.  The streaming QMF
// kernels stage the next frame (K1: 2 KB of PCM, K7: the 2 KB band record) this way while the current one is
// filtered.  A warp owns its barrier (arrival count 1); `parity` is the phase bit the caller flips after every wait.
__device__ __forceinline__ void mbar_init(uint32_t mbar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy (the TMA unit)
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t mbar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(mbar),
      "r"(parity)
      : "memory");
}

// Rings of the streaming QMF kernels (K1 / K7): a ring whose lanes produce kR outputs each keeps element e at index
// e + 2 (e / kR), i.e. groups of kR doubles with 2 doubles of padding after each.
template <int kR>
__host__ __device__ constexpr int ring_at(int e) { return e + 2 * (e / kR); }

// Experiment build only (tools/onchip_experiment.sh, -DC1_EXPERIMENT_ONCHIP_INTERMEDIATES; never the shipped library): the
// band rows between K1 and K3 and the band records between K6 and K7 are aliased onto 64 rows that stay in L1 / L2.
// The results are garbage; the kernel times bound from above what keeping those intermediates on chip (a K1->K3 /
// K6->K7 fusion) could save.
#ifdef C1_EXPERIMENT_ONCHIP_INTERMEDIATES
__device__ __forceinline__ size_t onchip_row(size_t u) { return 64 + (u & 63); }
#else
__device__ __forceinline__ size_t onchip_row(size_t u) { return u; }
#endif

}  // namespace c1
