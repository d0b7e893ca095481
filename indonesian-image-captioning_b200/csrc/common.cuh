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

bool pdl_enabled();           // CAPDEC_PDL=0 switches programmatic dependent launch off

// Launch `kernel` on `st` with the PDL attribute (and an optional cluster x-dimension).  Every
// kernel launched through here MUST call pdl_wait() (or pdl_prologue()) before its first global
// memory access.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while the
// previous kernel of the stream is still draining.  pdl_wait() blocks until that kernel has
// completed and its writes are visible; nothing produced by it may be touched before.
// pdl_launch_dependents() lets the NEXT kernel begin its own launch early.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}

// ---------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers shared by gemm_tc.cu, recur.cu and attention.cu
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared memory (one contiguous run), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// global data written by other CTAs with ordinary stores is about to be read through the async proxy
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
// Fused epilogues of the tcgen05 engine (gemm_tc.cu).  Every mode first forms
// val = acc (+ bias[n]) (+ addm[r, n]) and then:
enum GemmEpi {
  EPI_PLAIN = 0,   // out[r, n] = val
  EPI_G1 = 1,      // out[r, n] = val (fp32); columns n >= col0 are the SCN recurrent factor p = h W_ha:
                   //   m[(g*mB + r)*2F + F + f] = ft(val * q[r, n-col0]),  g = (n-col0)/F, f = (n-col0)%F
  EPI_P3 = 2,      // out[r, n] = val (fp32, u = x W_ia);  m[(g*mB + r)*2F + f] = ft(val * v[r, n])
  EPI_CELL = 3,    // 4 accumulators (grid batch = gate i,f,o,c of one d tile): LSTM pointwise ->
                   //   c_new, gates, h (+ dropout copy)                      (scn_cell.py:146-152)
  EPI_WR = 4,      // backward: acc = [w | r] = dpre_g [W_ic_g | W_hc_g]; grid batch = gate g.
                   //   n <  F: du[r, gF+n] = ft(w v), dv_acc[r, gF+n] += w u
                   //   n >= F: dp[r, gF+f] = ft(r q), dq_acc[r, gF+f] += r p
  EPI_DHCELL = 5   // backward: acc = recurrent gradient dh of step t; fused LSTM pointwise backward
                   //   (cell_bwd) -> dpre_t (ft), dc updated in place; rows >= rows take acc = 0
};

// schedule of the fused vocabulary kernel (gemm_tc.cu), recomputed by the selection kernel (beam.cu)
struct VocabTopkPlan { int tiles_r, num_tiles, grid, slots; };
VocabTopkPlan vocab_topk_plan(int rows, int V);

struct EpiArgs {
  // G1 / P3 / WR
  const float* fa = nullptr;        // G1: q   P3: v   WR: v
  const float* fb = nullptr;        //                  WR: q
  const float* fc = nullptr; int64_t ldc = 0;   //      WR: u (ld)
  const float* fd = nullptr; int64_t ldd = 0;   //      WR: p (ld)
  void* m = nullptr;                // G1 / P3: m operand of the P4 GEMM (feature type)
  void* du = nullptr; int64_t lddu = 0;         // WR
  void* dp = nullptr; int64_t lddp = 0;         // WR
  float* dv_acc = nullptr; float* dq_acc = nullptr;   // WR  [rows][4F]
  int mB = 0, F = 0, col0 = 0;
  int vB = 0;                       // P3: > 0 -> row r uses v[r % vB] (all (t,b) rows in one GEMM)
  // CELL / DHCELL
  const float* b1 = nullptr; const float* b2 = nullptr;
  const float* c_prev = nullptr; float* c_new = nullptr; float* gates = nullptr;
  void* h_out = nullptr; int64_t ldh = 0; void* hd_out = nullptr;
  float dropout_p = 0.f; const uint64_t* seed = nullptr; int t = 0, T = 1, D = 0;
  const float* dh_fc = nullptr; int64_t ld_dhfc = 0; float* dc = nullptr; void* dpre = nullptr;
  const float* c_new_r = nullptr;   // DHCELL: c_t (read)
  int rows_epi = 0;                 // DHCELL: rows the fused pointwise covers (>= rows)
  int lstm_order = 0;
  int topk_slots = 0;               // fused vocabulary kernel: partial slots per batch row
};

struct GemmArgs {
  const void* X = nullptr;  int64_t ldx = 0;     // activations, rows x K, K contiguous, feature type
  const void* W = nullptr;  int64_t ldw = 0;     // weights,     N x K,    K contiguous, feature type
  void* out = nullptr;      int64_t ldo = 0;     // rows x N, N contiguous
  int out_ft = 0;                                // 1: write feature type, 0: fp32
  const float* bias = nullptr;                   // [N] or null
  const float* addm = nullptr; int64_t ldadd = 0;  // fp32 [rows, N] added in the epilogue (may alias out)
  int rows = 0, N = 0, K = 0;
  int rows_alloc = 0;                            // rows physically present in X (>= rows); 0 -> rows
  int splitk = 0;                                // tcgen05 engine: 0 = off, -1 = auto, n = n K-slices that
                                                 // add into the output (EPI_PLAIN) or `abuf` (fused modes)
                                                 // with fp32 atomics: PRE-INITIALISE it (zeros, or the
                                                 // in-place addend).  Ignored by the SIMT engine.
  int batch = 1;                                 // grid.z
  int64_t sX = 0, sW = 0, sO = 0, sBias = 0, sAdd = 0;   // element strides per batch index
  int tn = 0;                                    // bit 0: W is given transposed, W^T [K][N] (pitch ldw); bit 1: X is,
                                                 // X^T [K][rows] (pitch ldx).  3: out[r, n] = sum_k XT[k, r] WT[k, n]
                                                 // (weight-gradient products; tcgen05 engine, EPI_PLAIN only)
  int epi = EPI_PLAIN;                           // tcgen05 engine only
  // fused modes: fp32 accumulation buffer, element (acc, z, r, n) at abuf[z*a_sz + acc*a_sa + r*a_ld + n]
  // (default: `out`), and the per-tile ticket counters (>= #tiles ints, zero on entry, left zero)
  float* abuf = nullptr; int64_t a_ld = 0, a_sz = 0, a_sa = 0;
  int* counters = nullptr;
  EpiArgs e;
};

int gemm_simt(const GemmArgs& a, cudaStream_t st);   // fp32 FFMA engine   (gemm_simt.cu)
int gemm_tc(const GemmArgs& a, cudaStream_t st);     // tcgen05 + TMA bf16 (gemm_tc.cu)
constexpr int GEMM_TC_MAX_TILE_COUNTERS = 4096;
int gemm_tc_init();
bool gemm_tc_cell_fusable(int rows);   // EPI_CELL keeps 4 accumulators per tile: small row counts only

static inline int gemm(int precision, const GemmArgs& a, cudaStream_t st) {
  return precision == CAPDEC_BF16 ? gemm_tc(a, st) : gemm_simt(a, st);
}

}  // namespace capdec
