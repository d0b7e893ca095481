// common.cuh -- shared device/host helpers for libcapdec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/capdec.h"

namespace capdec {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// error plumbing (no exceptions across the ABI)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();
void count_launch();          // process-wide kernel-launch counter (capdec_launch_count)

#define CAPDEC_CUDA_OK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      capdec::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                        cudaGetErrorString(_e));                                    \
      return CAPDEC_ERR_CUDA;                                                       \
    }                                                                               \
  } while (0)

#define CAPDEC_LAUNCH_OK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      capdec::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,           \
                        cudaGetErrorString(_e));                                    \
      return CAPDEC_ERR_CUDA;                                                       \
    }                                                                               \
    capdec::count_launch();                                                         \
  } while (0)

#define CAPDEC_TRY(expr)                                                            \
  do {                                                                              \
    int _r = (expr);                                                                \
    if (_r != CAPDEC_OK) return _r;                                                 \
  } while (0)

#define CAPDEC_REQUIRE(cond, code, ...)                                             \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      capdec::set_error(__VA_ARGS__);                                               \
      return (code);                                                                \
    }                                                                               \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// feature-type helpers (FT = float in fp32 mode, bf16 in bf16 mode)
// ---------------------------------------------------------------------------
template <typename T> struct FTraits;
template <> struct FTraits<float> {
  static constexpr int VEC = 4;            // elements per 16-byte vector
  static constexpr int PREC = CAPDEC_FP32;
};
template <> struct FTraits<bf16> {
  static constexpr int VEC = 8;
  static constexpr int PREC = CAPDEC_BF16;
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte streaming load (read-only path, no L1 allocation): features that are
// re-read every decode step live in L2, not L1.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// unpack one 16-byte vector into FTraits<T>::VEC floats
__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[4], float) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
  f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[8], bf16) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack16(const float (&f)[4], float) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                    __float_as_uint(f[3]));
}
__device__ __forceinline__ uint4 pack16(const float (&f)[8], bf16) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// counter-based dropout keep decision: splitmix64-style hash of (seed, index)
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  float u = (float)(z >> 40) * (1.0f / 16777216.0f);   // [0,1)
  return u < p ? 0.0f : 1.0f / (1.0f - p);
}

// ---------------------------------------------------------------------------
// GEMM engine interface (both engines): out[r, n] = sum_k X[r,k] * W[n,k]
// ---------------------------------------------------------------------------
struct GemmArgs {
  const void* X = nullptr;  int64_t ldx = 0;     // activations, rows x K, K contiguous, feature type
  const void* W = nullptr;  int64_t ldw = 0;     // weights,     N x K,    K contiguous, feature type
  void* out = nullptr;      int64_t ldo = 0;     // rows x N, N contiguous
  int out_ft = 0;                                // 1: write feature type, 0: fp32
  const float* bias = nullptr;                   // [N] or null
  const float* addm = nullptr; int64_t ldadd = 0;  // fp32 [rows, N] added in the epilogue (may alias out)
  int rows = 0, N = 0, K = 0;
  int rows_alloc = 0;                            // rows physically present in X (>= rows); 0 -> rows
  int splitk = 0;                                // tcgen05 engine: 0 = off, -1 = auto, n = n slices.  The
                                                 // output must then be PRE-INITIALISED (zero or the in-place
                                                 // addend); the SIMT engine ignores this and overwrites.
  int batch = 1;                                 // grid.z
  int64_t sX = 0, sW = 0, sO = 0, sBias = 0, sAdd = 0;   // element strides per batch index
};

int gemm_simt(const GemmArgs& a, cudaStream_t st);   // fp32 FFMA engine   (gemm_simt.cu)
int gemm_tc(const GemmArgs& a, cudaStream_t st);     // tcgen05 + TMA bf16 (gemm_tc.cu)
int gemm_tc_init();

static inline int gemm(int precision, const GemmArgs& a, cudaStream_t st) {
  return precision == CAPDEC_BF16 ? gemm_tc(a, st) : gemm_simt(a, st);
}

}  // namespace capdec
