// gemm_tc.cu -- bf16 tensor-core GEMM engine for sm_100a: TMA -> shared memory ->
// tcgen05.mma (accumulator in TMEM) -> tcgen05.ld epilogue.
//
//   out[r, n] = sum_k X[r, k] * W[n, k]  (+ bias[n]) (+ addm[r, n])      fp32 accumulate
//
// "Swap-AB" orientation, chosen for this decoder: the MMA M dimension (128 TMEM lanes)
// runs over the OUTPUT FEATURES n (rows of the packed weight matrix W) and the MMA N
// dimension over the activation rows r.  The recurrent GEMMs of the decoder have only
// B = 32..128 activation rows but 512..4608 output features, so the
// weights fill the 128-wide M side and the small batch is the N side (32/64/128).  A
// side effect: in the epilogue each thread owns one feature n (one TMEM lane) and the 32
// lanes of a warp store 32 consecutive n of one row r -> fully coalesced stores.
//
// One CTA computes one 128 x BNR output tile: warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer (single elected thread), warps 2..5 = epilogue (TMEM lane
// quarter = warp_id % 4).  Operand tiles are 64 bf16 (=128 B) wide in K, 128B-swizzled,
// K-major, multi-stage ring with full/empty mbarriers.
#include <cuda.h>   // CUtensorMap types only; the encode function is fetched at run time

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace capdec {

namespace {

constexpr int BM = 128;        // output features per tile (MMA M, TMEM lanes)
constexpr int BK = 64;         // bf16 elements per k-block (128 bytes, one swizzle atom)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

// ------------------------------ PTX wrappers ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address   bits [0,14)
  d |= (uint64_t)1 << 16;                           // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // SBO = 1024 B    bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
  return d;
}

template <int BNR>
struct Cfg {
  static constexpr int STAGES = (BNR == 128) ? 4 : 6;
  static constexpr int W_BYTES = BM * BK * 2;
  static constexpr int X_BYTES = BNR * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BNR < 32 ? 32 : BNR;
};

template <int BNR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
               void* out, int64_t ldo, int out_ft, const float* __restrict__ bias,
               const float* addm, int64_t ldadd, int rows, int N, int K,
               int64_t sO, int64_t sBias, int64_t sAdd, int splits) {
  using C = Cfg<BNR>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full ; then tmem base slot
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * C::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BM;
  const int r0 = blockIdx.y * BNR;
  // grid.z = batch * splits: split-K slices of one output tile are reduced with fp32 atomics
  // into an output the caller has pre-initialised (zero, or the in-place addend)
  const int z = blockIdx.z / splits;
  const int split = blockIdx.z - z * splits;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = (int)(((int64_t)nkb_total * split) / splits);
  const int kb_end = (int)(((int64_t)nkb_total * (split + 1)) / splits);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[C::STAGES + s]), 1);
    }
    mbar_init(smem_u32(&bars[2 * C::STAGES]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------- TMA producer -------------------------
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(smem_u32(&bars[C::STAGES + s]), ph ^ 1);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_expect_tx(full, C::STAGE_BYTES);
        const uint32_t ws = smem_u32(smem + s * C::STAGE_BYTES);
        tma_load_3d(ws, &mapW, full, (kb_begin + kb) * BK, n0, z);
        tma_load_3d(ws + C::W_BYTES, &mapX, full, (kb_begin + kb) * BK, r0, z);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------- MMA issuer -------------------------
      // instruction descriptor: D=F32, A=B=BF16, both K-major, N=BNR, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNR >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        tc_fence_after();
        const uint32_t ws = smem_u32(smem + s * C::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(ws);
        const uint64_t bdesc = make_smem_desc(ws + C::W_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the
          // (addr >> 4) start-address field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(smem_u32(&bars[C::STAGES + s]));     // frees the smem stage when MMAs retire
      }
      umma_commit(smem_u32(&bars[2 * C::STAGES]));       // accumulator complete
    }
  } else {
    // ------------------------- epilogue warps 2..5 -------------------------
    mbar_wait(smem_u32(&bars[2 * C::STAGES]), 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int n = n0 + q * 32 + lane;       // output feature owned by this thread
    const bool n_ok = n < N;
    float bv = 0.f;
    if (bias != nullptr && n_ok && split == 0) bv = bias[(int64_t)z * sBias + n];
    float* outf = (float*)out + (int64_t)z * sO;
    bf16* outh = (bf16*)out + (int64_t)z * sO;
    const float* add = addm ? addm + (int64_t)z * sAdd : nullptr;
    // split-K: the addend is applied once (split 0) unless it IS the output (in-place accumulate)
    if (splits > 1 && (split != 0 || (const void*)add == (const void*)outf)) add = nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BNR; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (n_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int r = r0 + c0 + j;
          if (r < rows) {
            float val = __uint_as_float(v[j]) + bv;
            if (add) val += add[(int64_t)r * ldadd + n];
            if (splits > 1) atomicAdd(&outf[(int64_t)r * ldo + n], val);
            else if (out_ft) outh[(int64_t)r * ldo + n] = __float2bfloat16_rn(val);
            else outf[(int64_t)r * ldo + n] = val;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------ host side ------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
int g_init_rc = CAPDEC_OK;

struct MapKey {
  const void* p; int64_t ld, sb; int rows, K, batch, box;
  bool operator==(const MapKey& o) const {
    return p == o.p && ld == o.ld && sb == o.sb && rows == o.rows && K == o.K && batch == o.batch &&
           box == o.box;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.p;
    auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.ld); mix((size_t)k.sb); mix((size_t)k.rows); mix((size_t)k.K);
    mix((size_t)k.batch); mix((size_t)k.box);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

int get_map(const void* p, int64_t ld, int rows, int K, int batch, int64_t sb, int box,
            CUtensorMap* out) {
  MapKey key{p, ld, sb, rows, K, batch, box};
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return CAPDEC_OK; }
  }
  CAPDEC_REQUIRE(((uintptr_t)p % 16) == 0 && (ld % 8) == 0 && (batch == 1 || (sb % 8) == 0),
                 CAPDEC_ERR_BAD_SHAPE,
                 "gemm_tc: operand needs 16-byte aligned base/pitch (ptr=%p ld=%lld sb=%lld)", p,
                 (long long)ld, (long long)sb);
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch == 1 ? (int64_t)rows * ld : sb) * 2};
  cuuint32_t box3[3] = {(cuuint32_t)BK, (cuuint32_t)box, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p), gdim, gstr,
                        box3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CAPDEC_REQUIRE(r == CUDA_SUCCESS, CAPDEC_ERR_CUDA,
                 "cuTensorMapEncodeTiled failed (%d) ptr=%p K=%d rows=%d ld=%lld batch=%d", (int)r, p,
                 K, rows, (long long)ld, batch);
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return CAPDEC_OK;
}

template <int BNR>
int launch(const GemmArgs& a, cudaStream_t st) {
  using C = Cfg<BNR>;
  CUtensorMap mW, mX;
  CAPDEC_TRY(get_map(a.W, a.ldw, a.N, a.K, a.batch, a.sW, BM, &mW));
  const int xrows = a.rows_alloc > a.rows ? a.rows_alloc : a.rows;
  CAPDEC_TRY(get_map(a.X, a.ldx, xrows, a.K, a.batch, a.sX, BNR, &mX));
  // split-K only for fp32 outputs that the caller pre-initialised (GemmArgs::splitk)
  const int nkb = ceil_div(a.K, BK);
  int splits = 1;
  if (a.splitk != 0 && !a.out_ft) {
    const int tiles = ceil_div(a.N, BM) * ceil_div(a.rows, BNR) * a.batch;
    splits = a.splitk > 0 ? a.splitk : (148 + tiles - 1) / tiles;     // auto: about one CTA per SM
    if (splits > nkb / 2) splits = nkb / 2;                            // >= 2 k-blocks per CTA
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  dim3 grid(ceil_div(a.N, BM), ceil_div(a.rows, BNR), a.batch * splits);
  gemm_tc_kernel<BNR><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(
      mW, mX, a.out, a.ldo, a.out_ft, a.bias, a.addm, a.ldadd, a.rows, a.N, a.K, a.sO, a.sBias,
      a.sAdd, splits);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace

int gemm_tc_init() {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                cudaGetErrorString(e));
      g_init_rc = CAPDEC_ERR_CUDA;
      return;
    }
    g_encode = (EncodeTiledFn)fn;
    cudaError_t e1 = cudaFuncSetAttribute(gemm_tc_kernel<32>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg<32>::SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(gemm_tc_kernel<64>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg<64>::SMEM_BYTES);
    cudaError_t e3 = cudaFuncSetAttribute(gemm_tc_kernel<128>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg<128>::SMEM_BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel) failed: %s",
                cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
      g_init_rc = CAPDEC_ERR_CUDA;
    }
  });
  return g_init_rc;
}

int gemm_tc(const GemmArgs& a, cudaStream_t st) {
  if (a.rows <= 0 || a.N <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(a.X && a.W && a.out && a.K > 0, CAPDEC_ERR_BAD_ARG, "gemm_tc: null operand");
  CAPDEC_TRY(gemm_tc_init());
  if (a.rows <= 32) return launch<32>(a, st);
  if (a.rows <= 64) return launch<64>(a, st);
  return launch<128>(a, st);
}

}  // namespace capdec
