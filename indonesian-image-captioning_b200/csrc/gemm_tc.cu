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

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace capdec {

namespace {

constexpr int BM = 128;        // output features per tile (MMA M, TMEM lanes)
constexpr int BK = 64;         // bf16 elements per k-block (128 bytes, one swizzle atom)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

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

// MN-major, 128B-swizzled operand tile (the contraction index k is the SLOW dimension in memory):
// 64 MN-elements (128 B) contiguous, 8 k-rows per swizzle atom (1024 B), k-groups SBO = 1024 B apart,
// 64-wide MN blocks LBO = 8192 B apart (one TMA box {64 mn, 64 k} each).  Canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (CUTLASS make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;                 // LBO: next 64-wide MN block
  d |= (uint64_t)(1024 >> 4) << 32;                 // SBO: next group of 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

// Epilogue stores of one 32-row chunk held in registers (thread = output feature n, v[j] = row r0 + j):
// the row pointer advances by the pitch, nothing is recomputed per element (the first version spent ~40
// instructions per element on 64-bit index arithmetic and parameter reloads: 10 us per 128 x 128 tile).
template <bool OUT_FT>
__device__ __forceinline__ void store_chunk(void* out, int64_t ldo, int r0, int rows, float bv, const float* add,
                                            int64_t ldadd, const uint32_t (&v)[32]) {
  const int nr = min(32, rows - r0);
  if (nr <= 0) return;
  if (OUT_FT) {
    bf16* po = (bf16*)out + (int64_t)r0 * ldo;
    if (add == nullptr && nr == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { *po = __float2bfloat16_rn(__uint_as_float(v[j]) + bv); po += ldo; }
    } else {
      const float* pa = add ? add + (int64_t)r0 * ldadd : nullptr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nr) {
          float val = __uint_as_float(v[j]) + bv;
          if (pa) val += *pa;
          *po = __float2bfloat16_rn(val);
        }
        po += ldo;
        if (pa) pa += ldadd;
      }
    }
  } else {
    float* po = (float*)out + (int64_t)r0 * ldo;
    if (add == nullptr && nr == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { *po = __uint_as_float(v[j]) + bv; po += ldo; }
    } else {
      const float* pa = add ? add + (int64_t)r0 * ldadd : nullptr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nr) {
          float val = __uint_as_float(v[j]) + bv;
          if (pa) val += *pa;
          *po = val;
        }
        po += ldo;
        if (pa) pa += ldadd;
      }
    }
  }
}

template <int BNR, int NACC>
struct Cfg {
  // 3 stages for the 128-row tile: 96 KB per CTA, so TWO CTAs share an SM and one tile's prologue /
  // epilogue overlaps the other's main loop (with 4 stages = 128 KB only one CTA fits)
  static constexpr int STAGES = ((BNR == 64 && NACC == 4) || BNR == 128) ? 3 : 4;
  static constexpr int W_BYTES = BM * BK * 2;
  static constexpr int X_BYTES = BNR * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int TMEM_COLS = (NACC * BNR) <= 32 ? 32 : (NACC * BNR) <= 64 ? 64 :
                                   (NACC * BNR) <= 128 ? 128 : (NACC * BNR) <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct KArgs {
  void* out; int64_t ldo; int out_ft;
  const float* bias; const float* addm; int64_t ldadd;
  int rows, N, K;
  int64_t sO, sBias, sAdd;
  int splits;            // K-slices per output tile (grid.z = batch * splits)
  // accumulation buffer of the fused epilogues: element (acc, z, r, n) lives at
  // abuf[z*a_sz + acc*a_sa + r*a_ld + n]; K-slices add into it with fp32 atomics (pre-zeroed or
  // pre-filled with the addend by the caller) and the LAST CTA of a tile (ticket counter) reads the
  // complete sums back and applies the fused epilogue
  float* abuf; int64_t a_ld, a_sz, a_sa;
  int* counters;         // one ticket counter per output tile, zero on entry, left zero
  EpiArgs e;
};

// one output element (r, n) of batch z with its NACC accumulated values.  Read-only operands are
// fetched with __ldg so that the unrolled caller can issue the loads of several rows together;
// `extra` is the value the caller pre-loaded for the mode (DHCELL: dc[r, n]).
template <int NACC, int EPI>
__device__ __forceinline__ void epilogue_elem(const KArgs& a, int z, int r, int n, const float (&acc)[NACC],
                                              float extra) {
  const EpiArgs& e = a.e;
  float val = acc[0];
  if (EPI != EPI_CELL && EPI != EPI_DHCELL) {
    if (a.bias) val += __ldg(a.bias + (int64_t)z * a.sBias + n);
    if (a.addm) val += __ldg(a.addm + (int64_t)z * a.sAdd + (int64_t)r * a.ldadd + n);
  }
  if (EPI == EPI_G1) {
    ((float*)a.out)[(int64_t)r * a.ldo + n] = val;
    if (n >= e.col0) {
      const int nn = n - e.col0;
      const int g = nn / e.F, f = nn - g * e.F;
      ((bf16*)e.m)[((int64_t)g * e.mB + r) * 2 * e.F + e.F + f] =
          __float2bfloat16_rn(val * __ldg(e.fa + (int64_t)r * 4 * e.F + nn));
    }
  } else if (EPI == EPI_P3) {
    ((float*)a.out)[(int64_t)r * a.ldo + n] = val;
    const int g = n / e.F, f = n - g * e.F;
    const int rv = e.vB > 0 ? r % e.vB : r;
    ((bf16*)e.m)[((int64_t)g * e.mB + r) * 2 * e.F + f] =
        __float2bfloat16_rn(val * __ldg(e.fa + (int64_t)rv * 4 * e.F + n));
  } else if (EPI == EPI_CELL) {
    const int D = e.D, d = n;
    float pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x = acc[g < NACC ? g : 0];
      if (e.b1) x += __ldg(e.b1 + g * D + d);
      if (e.b2) x += __ldg(e.b2 + g * D + d);
      pre[g] = x;
    }
    const int so = e.lstm_order ? 3 : 2, sg = e.lstm_order ? 2 : 3;
    const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), og = sigmoidf_(pre[so]);
    const float gg = tanhf(pre[sg]);
    const int64_t i = (int64_t)r * D + d;
    const float c = fg * __ldg(e.c_prev + i) + ig * gg;
    const float h = og * tanhf(c);
    e.c_new[i] = c;
    if (e.gates) {
      float* gp = e.gates + (int64_t)r * 4 * D + d;
      gp[0] = ig; gp[D] = fg; gp[2 * D] = og; gp[3 * D] = gg;
    }
    ((bf16*)e.h_out)[(int64_t)r * e.ldh + d] = __float2bfloat16_rn(h);
    if (e.hd_out) {
      const float sc = dropout_scale(__ldg(e.seed), ((uint64_t)r * e.T + e.t) * D + d, e.dropout_p);
      ((bf16*)e.hd_out)[(int64_t)r * e.ldh + d] = __float2bfloat16_rn(h * sc);
    }
  } else if (EPI == EPI_WR) {
    const int F = e.F, g = z;
    if (n < F) {
      const int64_t k = (int64_t)r * 4 * F + g * F + n;
      ((bf16*)e.du)[(int64_t)r * e.lddu + g * F + n] = __float2bfloat16_rn(val * __ldg(e.fa + k));
      atomicAdd(e.dv_acc + k, val * __ldg(e.fc + (int64_t)r * e.ldc + g * F + n));   // sole writer: RED, no stall
    } else {
      const int f = n - F;
      const int64_t k = (int64_t)r * 4 * F + g * F + f;
      ((bf16*)e.dp)[(int64_t)r * e.lddp + g * F + f] = __float2bfloat16_rn(val * __ldg(e.fb + k));
      atomicAdd(e.dq_acc + k, val * __ldg(e.fd + (int64_t)r * e.ldd + g * F + f));
    }
  } else if (EPI == EPI_DHCELL) {
    const int D = e.D, d = n;
    const int64_t i = (int64_t)r * D + d;
    float dh = r < a.rows ? val : 0.f;          // rows that ended at this step start from dh = 0
    {
      float g = __ldg(e.dh_fc + (int64_t)r * e.ld_dhfc + d);
      if (e.dropout_p > 0.f) g *= dropout_scale(__ldg(e.seed), ((uint64_t)r * e.T + e.t) * D + d, e.dropout_p);
      dh += g;
    }
    const float* gp = e.gates + (int64_t)r * 4 * D + d;
    const float ig = __ldg(gp), fg = __ldg(gp + D), og = __ldg(gp + 2 * D), gg = __ldg(gp + 3 * D);
    const float tc = tanhf(__ldg(e.c_new_r + i));
    const float dcn = extra + dh * og * (1.f - tc * tc);
    const float dpo = dh * tc * og * (1.f - og);
    const float dpi = dcn * gg * ig * (1.f - ig);
    const float dpf = dcn * __ldg(e.c_prev + i) * fg * (1.f - fg);
    const float dpg = dcn * ig * (1.f - gg * gg);
    e.dc[i] = dcn * fg;
    const int so = e.lstm_order ? 3 : 2, sg = e.lstm_order ? 2 : 3;
    bf16* dp = (bf16*)e.dpre + (int64_t)r * 4 * D + d;
    dp[0] = __float2bfloat16_rn(dpi);
    dp[D] = __float2bfloat16_rn(dpf);
    dp[(int64_t)so * D] = __float2bfloat16_rn(dpo);
    dp[(int64_t)sg * D] = __float2bfloat16_rn(dpg);
  }
}

// TN: bit 0 = the W operand is given TRANSPOSED (W^T [K][N]), bit 1 = the X operand is (X^T [K][rows]) --
// the layout of a weight-gradient product dW = dY^T X, whose contraction runs over the sample rows:
// MN-major UMMA operands, no explicit transposition pass.
template <int BNR, int NACC, int EPI, int TN = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
               const __grid_constant__ KArgs a) {
  using C = Cfg<BNR, NACC>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full ; then tmem base slot, then the "last CTA" flag
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * C::STAGES + 1);
  volatile uint32_t* last_flag = tmem_slot + 1;

  pdl_launch_dependents();
  const int splits = a.splits;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BM;
  const int r0 = blockIdx.y * BNR;
  // grid.z = batch * splits
  const int z = blockIdx.z / splits;
  const int split = blockIdx.z - z * splits;
  const int nkb_total = (a.K + BK - 1) / BK;
  const int kb_begin = (int)(((int64_t)nkb_total * split) / splits);
  const int kb_end = (int)(((int64_t)nkb_total * (split + 1)) / splits);
  const int nunits = (kb_end - kb_begin) * NACC;     // pipeline units: (k-block, accumulator)

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
  pdl_wait();          // everything below may read what the previous kernel of the stream wrote

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------- TMA producer -------------------------
      for (int u = 0; u < nunits; ++u) {
        const int s = u % C::STAGES;
        const uint32_t ph = (u / C::STAGES) & 1;
        const int kb = kb_begin + u / NACC;
        const int zz = NACC > 1 ? (u % NACC) : z;
        mbar_wait(smem_u32(&bars[C::STAGES + s]), ph ^ 1);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_expect_tx(full, C::STAGE_BYTES);
        const uint32_t ws = smem_u32(smem + s * C::STAGE_BYTES);
        if (TN & 1) {
#pragma unroll
          for (int h = 0; h < BM / 64; ++h) tma_load_3d(ws + h * 8192, &mapW, full, n0 + h * 64, kb * BK, zz);
        } else {
          tma_load_3d(ws, &mapW, full, kb * BK, n0, zz);
        }
        if (TN & 2) {
#pragma unroll
          for (int h = 0; h < BNR / 64; ++h) tma_load_3d(ws + C::W_BYTES + h * 8192, &mapX, full, r0 + h * 64, kb * BK, zz);
        } else {
          tma_load_3d(ws + C::W_BYTES, &mapX, full, kb * BK, r0, zz);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------- MMA issuer -------------------------
      // instruction descriptor: D=F32, A=B=BF16, both K-major, N=BNR, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNR >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24) | ((TN & 1) ? (1u << 15) : 0u) |          // a major = MN
                             ((TN & 2) ? (1u << 16) : 0u);                                        // b major = MN
      for (int u = 0; u < nunits; ++u) {
        const int s = u % C::STAGES;
        const uint32_t ph = (u / C::STAGES) & 1;
        const int acc = NACC > 1 ? (u % NACC) : 0;
        const int kb_rel = u / NACC;
        mbar_wait(smem_u32(&bars[s]), ph);
        tc_fence_after();
        const uint32_t ws = smem_u32(smem + s * C::STAGE_BYTES);
        const uint64_t adesc = (TN & 1) ? make_smem_desc_mn(ws) : make_smem_desc(ws);
        const uint64_t bdesc = (TN & 2) ? make_smem_desc_mn(ws + C::W_BYTES) : make_smem_desc(ws + C::W_BYTES);
        // K-major: 16 elements = 32 bytes along K inside the swizzle atom (+2 in the (addr >> 4) field);
        // MN-major: 16 k-rows = two 1024-byte atoms (+128)
        constexpr int ASTEP = (TN & 1) ? 128 : 2, BSTEP = (TN & 2) ? 128 : 2;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          umma_bf16(tmem_base + (uint32_t)(acc * BNR), adesc + ASTEP * k, bdesc + BSTEP * k, idesc,
                    (kb_rel | k) != 0);
        }
        umma_commit(smem_u32(&bars[C::STAGES + s]));     // frees the smem stage when MMAs retire
      }
      umma_commit(smem_u32(&bars[2 * C::STAGES]));       // accumulators complete
    }
  } else {
    // ------------------------- epilogue warps 2..5 -------------------------
    mbar_wait(smem_u32(&bars[2 * C::STAGES]), 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int nl = q * 32 + lane;           // tile-local output feature owned by this thread
    const int n = n0 + nl;
    const bool n_ok = n < a.N;
    if (EPI == EPI_PLAIN) {
      // plain epilogue: TMEM -> registers -> global; K-slices add with fp32 atomics into an output
      // the caller pre-initialised (zero, or the in-place addend)
      float bv = 0.f;
      if (a.bias != nullptr && n_ok && split == 0) bv = a.bias[(int64_t)z * a.sBias + n];
      float* outf = (float*)a.out + (int64_t)z * a.sO;
      bf16* outh = (bf16*)a.out + (int64_t)z * a.sO;
      const float* add = a.addm ? a.addm + (int64_t)z * a.sAdd : nullptr;
      if (splits > 1 && (split != 0 || (const void*)add == (const void*)outf)) add = nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BNR; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (n_ok) {
          if (splits > 1) {
            float* po = outf + (int64_t)(r0 + c0) * a.ldo + n;
            const float* pa = add ? add + (int64_t)(r0 + c0) * a.ldadd + n : nullptr;
            const int nr = min(32, a.rows - (r0 + c0));
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nr) atomicAdd(po, __uint_as_float(v[j]) + bv + (pa ? *pa : 0.f));
              po += a.ldo;
              if (pa) pa += a.ldadd;
            }
          } else if (a.out_ft) {
            store_chunk<true>(outh + n, a.ldo, r0 + c0, a.rows, bv, add ? add + n : nullptr, a.ldadd, v);
          } else {
            store_chunk<false>(outf + n, a.ldo, r0 + c0, a.rows, bv, add ? add + n : nullptr, a.ldadd, v);
          }
        }
      }
    } else {
      // fused epilogues, pass 1: this CTA's partial sums are ADDED to the accumulation buffer (the
      // caller pre-initialised it: zeros, or the in-place addend)
      float* ab = a.abuf + (int64_t)z * a.a_sz + n;
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll 1
        for (int c0 = 0; c0 < BNR; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BNR + c0), v);
          if (n_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int r = r0 + c0 + j;
              if (r < a.rows) {
                atomicAdd(ab + (int64_t)acc * a.a_sa + (int64_t)r * a.a_ld, __uint_as_float(v[j]));
              }
            }
          }
        }
      }
      bool last = true;
      if (splits > 1) {
        // ticket: the CTA that finishes a tile last owns its fused epilogue
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          int* ctr = a.counters + (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * z));
          const int ticket = atomicAdd(ctr, 1);
          const bool l = ticket == splits - 1;
          if (l) *ctr = 0;                       // every K-slice has taken its ticket: rearm
          *last_flag = l ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        last = *last_flag != 0u;
        if (last) __threadfence();
      }
      if (last && n_ok) {
        // pass 2: complete sums -> fused epilogue (thread = feature n, loop over the tile's rows)
        const int rows_lim = (EPI == EPI_DHCELL) ? a.e.rows_epi : a.rows;
        const int r_end = min(r0 + BNR, rows_lim);
        constexpr int RB = 8;                    // rows in flight per thread
        for (int rb = r0; rb < r_end; rb += RB) {
          float acc[RB][NACC], extra[RB];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int r = min(rb + i, r_end - 1);
#pragma unroll
            for (int k = 0; k < NACC; ++k)
              acc[i][k] = __ldcg(ab + (int64_t)k * a.a_sa + (int64_t)r * a.a_ld);
            extra[i] = (EPI == EPI_DHCELL) ? __ldcg(a.e.dc + (int64_t)r * a.e.D + n) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < RB; ++i)
            if (rb + i < r_end) epilogue_elem<NACC, EPI>(a, z, rb + i, n, acc[i], extra[i]);
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

// ---------------------------------------------------------------------------------------
// Persistent variant for the large plain GEMMs (vocabulary projection, att1, weight gradients, the
// beam-search GEMMs): one CTA per SM walks over its 128 x 128 output tiles; the TMA ring (6 stages)
// keeps running across tile boundaries and the accumulator is DOUBLE BUFFERED in TMEM (2 x 128
// columns), so the epilogue of tile i (TMEM -> registers -> global) overlaps the main loop of tile
// i + 1 and the per-tile prologue (barrier init, TMEM allocation, first-load latency) is paid once.
// ---------------------------------------------------------------------------------------
// KL mode: CTA b owns the CONSECUTIVE tiles [ceil(b T / G), ceil((b+1) T / G)) of the (row tile, vocabulary tile)
// sequence with the vocabulary tile running fastest, so that the epilogue's running statistics of a batch row stay in
// registers across the vocabulary tiles of a run; plain mode: round robin.
__device__ __forceinline__ int ps_run_begin(int b, int num_tiles, int grid) {
  return (int)(((int64_t)b * num_tiles + grid - 1) / grid);
}
constexpr int PS_THREADS = 320;     // TMA warp, MMA warp, EIGHT epilogue warps: two per TMEM lane quarter, 64 columns each
constexpr int PS_STAGES = 6;
constexpr int PS_BNR = 128;
constexpr int PS_STAGE_BYTES = (BM + PS_BNR) * BK * 2;
constexpr int PS_SMEM_BYTES = PS_STAGES * PS_STAGE_BYTES + 1024 + 256;

//
// KL > 0: the fused vocabulary projection + log-softmax statistics + top-k candidates of beam search
// (reference: `scores = F.log_softmax(self.fc(h), dim=1)` + `topk`, attention_scn.py:235-253).  The operand roles
// are EXCHANGED by the caller: the M side ("W", TMEM lanes) holds the batch rows h, the N side ("X", TMEM columns)
// the vocabulary rows of fc.weight, so that one epilogue thread owns ONE batch row and walks over the tile's 128
// vocabulary entries in its registers: running maximum, sum of exponentials and the KL largest logits (value,
// vocabulary index) need no cross-thread traffic.  Per (batch row, vocabulary tile) it writes 2 + 2 KL floats
// instead of 128 logits: the (rows x V) fp32 logits tensor never exists; beam_select_kernel<.., 2> merges the tiles.
template <int TN, int KL = 0>
__global__ void __launch_bounds__(PS_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
                       const __grid_constant__ KArgs a, int tiles_n, int tiles_r, int num_tiles) {
  constexpr int W_BYTES = BM * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + PS_STAGES * PS_STAGE_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * PS_STAGES + 4);
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int s = 0; s < PS_STAGES; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[PS_STAGES + s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars[2 * PS_STAGES + b]), 1);          // tmem_full: one tcgen05.commit
      mbar_init(smem_u32(&bars[2 * PS_STAGES + 2 + b]), 8);      // tmem_empty: the eight epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int nkb = (a.K + BK - 1) / BK;
  const int tiles_nr = tiles_n * tiles_r;
  const int t_first = KL > 0 ? ps_run_begin(blockIdx.x, num_tiles, gridDim.x) : (int)blockIdx.x;
  const int t_end = KL > 0 ? ps_run_begin(blockIdx.x + 1, num_tiles, gridDim.x) : num_tiles;
  const int t_step = KL > 0 ? 1 : (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------- TMA producer -------------------------
      uint32_t u = 0;
      for (int t = t_first; t < t_end; t += t_step) {
        const int z = t / tiles_nr, rem = t - z * tiles_nr;
        const int rt = KL > 0 ? rem % tiles_r : rem / tiles_n, nt = KL > 0 ? rem / tiles_r : rem - rt * tiles_n;
        const int n0 = nt * BM, r0 = rt * PS_BNR;
        for (int kb = 0; kb < nkb; ++kb, ++u) {
          const int s = u % PS_STAGES;
          const uint32_t ph = (u / PS_STAGES) & 1;
          mbar_wait(smem_u32(&bars[PS_STAGES + s]), ph ^ 1);
          const uint32_t full = smem_u32(&bars[s]);
          mbar_expect_tx(full, PS_STAGE_BYTES);
          const uint32_t ws = smem_u32(smem + s * PS_STAGE_BYTES);
          if (TN & 1) {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h) tma_load_3d(ws + h * 8192, &mapW, full, n0 + h * 64, kb * BK, z);
          } else {
            tma_load_3d(ws, &mapW, full, kb * BK, n0, z);
          }
          if (TN & 2) {
#pragma unroll
            for (int h = 0; h < PS_BNR / 64; ++h) tma_load_3d(ws + W_BYTES + h * 8192, &mapX, full, r0 + h * 64, kb * BK, z);
          } else {
            tma_load_3d(ws + W_BYTES, &mapX, full, kb * BK, r0, z);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------- MMA issuer -------------------------
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PS_BNR >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24) | ((TN & 1) ? (1u << 15) : 0u) |
                             ((TN & 2) ? (1u << 16) : 0u);
      constexpr int ASTEP = (TN & 1) ? 128 : 2, BSTEP = (TN & 2) ? 128 : 2;
      uint32_t u = 0;
      int i = 0;
      for (int t = t_first; t < t_end; t += t_step, ++i) {
        const int buf = i & 1;
        mbar_wait(smem_u32(&bars[2 * PS_STAGES + 2 + buf]), (uint32_t)((i >> 1) & 1) ^ 1u);   // epilogue drained it
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++u) {
          const int s = u % PS_STAGES;
          const uint32_t ph = (u / PS_STAGES) & 1;
          mbar_wait(smem_u32(&bars[s]), ph);
          tc_fence_after();
          const uint32_t ws = smem_u32(smem + s * PS_STAGE_BYTES);
          const uint64_t adesc = (TN & 1) ? make_smem_desc_mn(ws) : make_smem_desc(ws);
          const uint64_t bdesc = (TN & 2) ? make_smem_desc_mn(ws + W_BYTES) : make_smem_desc(ws + W_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_base + (uint32_t)(buf * PS_BNR), adesc + ASTEP * k, bdesc + BSTEP * k, idesc, (kb | k) != 0);
          umma_commit(smem_u32(&bars[PS_STAGES + s]));
        }
        umma_commit(smem_u32(&bars[2 * PS_STAGES + buf]));
      }
    }
  } else {
    // ------------------------- epilogue warps 2..9 -------------------------
    // warp w may touch TMEM lanes 32 (w & 3) ..; warps w and w + 4 share a lane quarter and take 64 columns each
    // (one warp per scheduler and quarter left the epilogue latency-exposed: it, not the main loop, set the pace)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cbeg = half * (PS_BNR / 2), cend = cbeg + PS_BNR / 2;
    int i = 0;
    // KL mode: running statistics of this thread's batch row over the vocabulary tiles of the CTA's run
    constexpr int KQ = KL > 0 ? KL : 1;
    float m = -INFINITY, ssum = 0.f;
    float lv[KQ];
    int li[KQ];
#pragma unroll
    for (int e = 0; e < KQ; ++e) { lv[e] = -INFINITY; li[e] = 0x7fffffff; }
    for (int t = t_first; t < t_end; t += t_step, ++i) {
      const int buf = i & 1;
      const int z = t / tiles_nr, rem = t - z * tiles_nr;
      const int rt = KL > 0 ? rem % tiles_r : rem / tiles_n, nt = KL > 0 ? rem / tiles_r : rem - rt * tiles_n;
      const int n = nt * BM + q * 32 + lane, r0 = rt * PS_BNR;
      const bool n_ok = n < a.N;
      float bv = 0.f;
      if (KL == 0 && a.bias != nullptr && n_ok) bv = a.bias[(int64_t)z * a.sBias + n];
      float* outf = (float*)a.out + (int64_t)z * a.sO;
      bf16* outh = (bf16*)a.out + (int64_t)z * a.sO;
      const float* add = a.addm ? a.addm + (int64_t)z * a.sAdd : nullptr;
      mbar_wait(smem_u32(&bars[2 * PS_STAGES + buf]), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      if (KL > 0) {
        // thread = batch row n, TMEM columns = vocabulary entries r0 .. r0 + 127 of this tile
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * PS_BNR + c0), v);
          const int vb = r0 + c0;
          if (vb < a.rows) {
            float x[32];
            if (vb + 32 <= a.rows) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __ldg(a.bias + vb + j);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = (vb + j < a.rows) ? __uint_as_float(v[j]) + __ldg(a.bias + vb + j) : -INFINITY;
            }
            float c4[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
            for (int j = 4; j < 32; ++j) c4[j & 3] = fmaxf(c4[j & 3], x[j]);
            const float cm = fmaxf(fmaxf(c4[0], c4[1]), fmaxf(c4[2], c4[3]));
            const float mn = fmaxf(m, cm);
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) a4[j & 3] += __expf(x[j] - mn);
            ssum = ssum * __expf(m - mn) + ((a4[0] + a4[1]) + (a4[2] + a4[3]));
            m = mn;
            // candidates: pop the chunk's maxima while they beat the list tail (usually zero or one round once the
            // run has seen a few hundred entries); the FIRST maximum wins a tie, and an equal value never displaces
            // an earlier (smaller-index) entry of the list
            float top = cm;
#pragma unroll 1
            while (top > lv[KQ - 1]) {
              int jm = 31;
#pragma unroll
              for (int j = 30; j >= 0; --j) jm = (x[j] == top) ? j : jm;
              lv[KQ - 1] = top; li[KQ - 1] = vb + jm;
#pragma unroll
              for (int e = KQ - 1; e > 0; --e) {
                if (lv[e] > lv[e - 1]) {
                  const float tv = lv[e]; lv[e] = lv[e - 1]; lv[e - 1] = tv;
                  const int ti = li[e]; li[e] = li[e - 1]; li[e - 1] = ti;
                }
              }
              top = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                x[j] = (j == jm) ? -INFINITY : x[j];
                top = fmaxf(top, x[j]);
              }
            }
          }
        }
        // the run leaves this row tile (or ends): flush the row's statistics into slot (run - first run of the row tile)
        if (t + 1 == t_end || (t + 1) / tiles_r != nt) {
          if (n_ok) {
            const int run_lo = (int)(((int64_t)nt * tiles_r * gridDim.x) / num_tiles);
            float* po = (float*)a.out + ((int64_t)n * a.e.topk_slots + ((int)blockIdx.x - run_lo) * 2 + half) * (2 + 2 * KQ);
            po[0] = m; po[1] = ssum;
#pragma unroll
            for (int e = 0; e < KQ; ++e) { po[2 + e] = lv[e]; po[2 + KQ + e] = __int_as_float(li[e]); }
          }
          m = -INFINITY; ssum = 0.f;
#pragma unroll
          for (int e = 0; e < KQ; ++e) { lv[e] = -INFINITY; li[e] = 0x7fffffff; }
        }
      } else {
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * PS_BNR + c0), v);
          if (n_ok) {
            if (a.out_ft) store_chunk<true>(outh + n, a.ldo, r0 + c0, a.rows, bv, add ? add + n : nullptr, a.ldadd, v);
            else store_chunk<false>(outf + n, a.ldo, r0 + c0, a.rows, bv, add ? add + n : nullptr, a.ldadd, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[2 * PS_STAGES + 2 + buf])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
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
  const void* p; int64_t ld, sb; int rows, K, batch, box;     // box < 0: transposed operand [K][rows]
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
  // box > 0: operand [rows][K], K contiguous, tile {64 k, box rows}.  box < 0: transposed operand
  // [K][rows], rows contiguous, tile {64 rows, 64 k}
  const bool tn = box < 0;
  cuuint64_t gdim[3] = {(cuuint64_t)(tn ? rows : K), (cuuint64_t)(tn ? K : rows), (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2,
                        (cuuint64_t)(batch == 1 ? (int64_t)(tn ? K : rows) * ld : sb) * 2};
  cuuint32_t box3[3] = {(cuuint32_t)BK, (cuuint32_t)(tn ? 64 : box), 1};
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

template <int BNR, int NACC, int EPI, int TN = 0>
int launch(const GemmArgs& a, cudaStream_t st) {
  using C = Cfg<BNR, NACC>;
  auto kernel = gemm_tc_kernel<BNR, NACC, EPI, TN>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [&] {
    attr_rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  CAPDEC_REQUIRE(attr_rc == cudaSuccess, CAPDEC_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc_kernel) failed: %s",
                 cudaGetErrorString(attr_rc));
  CUtensorMap mW, mX;
  CAPDEC_TRY(get_map(a.W, a.ldw, a.N, a.K, a.batch, a.sW, (TN & 1) ? -1 : BM, &mW));
  const int rows_epi = (EPI == EPI_DHCELL && a.e.rows_epi > a.rows) ? a.e.rows_epi : a.rows;
  int xrows = a.rows_alloc > a.rows ? a.rows_alloc : a.rows;
  if (xrows < rows_epi) xrows = rows_epi;
  CAPDEC_TRY(get_map(a.X, a.ldx, xrows, a.K, a.batch, a.sX, (TN & 2) ? -1 : BNR, &mX));
  // split-K: grid.z = batch * splits; the K-slices of a tile add into one fp32 buffer
  const int nkb = ceil_div(a.K, BK);
  const int zb = NACC > 1 ? 1 : a.batch;           // NACC > 1: the batch (gate) index is looped inside
  const int row_tiles = ceil_div(rows_epi, BNR);
  const int tiles = ceil_div(a.N, BM) * row_tiles * zb;
  int splits = 1;
  if (a.splitk != 0 && !(EPI == EPI_PLAIN && a.out_ft)) {
    int target = 148;                                                  // auto: about one CTA per SM
    if (const char* e = getenv("CAPDEC_GEMM_TARGET")) target = atoi(e);   // experiments
    splits = a.splitk > 0 ? a.splitk : (target + tiles - 1) / tiles;
    if (a.splitk < 0) {
      if (const char* e = getenv("CAPDEC_GEMM_SPLITK")) splits = atoi(e);   // experiments: fixed split count
    }
    if (splits > nkb / 2) splits = nkb / 2;                            // >= 2 k-blocks per CTA
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  KArgs k;
  k.out = a.out; k.ldo = a.ldo; k.out_ft = a.out_ft; k.bias = a.bias; k.addm = a.addm; k.ldadd = a.ldadd;
  k.rows = a.rows; k.N = a.N; k.K = a.K; k.sO = a.sO; k.sBias = a.sBias; k.sAdd = a.sAdd;
  k.splits = splits; k.e = a.e;
  k.abuf = a.abuf; k.a_ld = a.a_ld; k.a_sz = a.a_sz; k.a_sa = a.a_sa; k.counters = a.counters;
  if (EPI != EPI_PLAIN) {
    if (!k.abuf) { k.abuf = (float*)a.out; k.a_ld = a.ldo; k.a_sz = a.sO; k.a_sa = 0; }
    CAPDEC_REQUIRE(k.abuf != nullptr && !a.out_ft, CAPDEC_ERR_BAD_ARG,
                   "gemm_tc: fused epilogues need an fp32 accumulation buffer");
    CAPDEC_REQUIRE(splits == 1 || (k.counters != nullptr && tiles <= GEMM_TC_MAX_TILE_COUNTERS),
                   CAPDEC_ERR_BAD_ARG, "gemm_tc: split-K with a fused epilogue needs ticket counters (%d tiles)",
                   tiles);
    if ((const void*)k.addm == (const void*)k.abuf) k.addm = nullptr;   // addend already sits in the buffer
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(ceil_div(a.N, BM), row_tiles, zb * splits);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CAPDEC_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, mW, mX, k));
  count_launch();
  return CAPDEC_OK;
}

int g_sm_count = 0;

template <int TN, int KL = 0>
int launch_persist(const GemmArgs& a, cudaStream_t st) {
  auto kernel = gemm_tc_persist_kernel<TN, KL>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [&] {
    attr_rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_BYTES);
  });
  CAPDEC_REQUIRE(attr_rc == cudaSuccess, CAPDEC_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc_persist_kernel) failed: %s",
                 cudaGetErrorString(attr_rc));
  if (g_sm_count == 0) {
    int dev = 0;
    CAPDEC_CUDA_OK(cudaGetDevice(&dev));
    CAPDEC_CUDA_OK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap mW, mX;
  CAPDEC_TRY(get_map(a.W, a.ldw, a.N, a.K, a.batch, a.sW, (TN & 1) ? -1 : BM, &mW));
  const int xrows = a.rows_alloc > a.rows ? a.rows_alloc : a.rows;
  CAPDEC_TRY(get_map(a.X, a.ldx, xrows, a.K, a.batch, a.sX, (TN & 2) ? -1 : PS_BNR, &mX));
  const int tiles_n = ceil_div(a.N, BM), tiles_r = ceil_div(a.rows, PS_BNR);
  const int num_tiles = tiles_n * tiles_r * a.batch;
  KArgs k;
  k.out = a.out; k.ldo = a.ldo; k.out_ft = a.out_ft; k.bias = a.bias; k.addm = a.addm; k.ldadd = a.ldadd;
  k.rows = a.rows; k.N = a.N; k.K = a.K; k.sO = a.sO; k.sBias = a.sBias; k.sAdd = a.sAdd;
  k.splits = 1; k.e = a.e;
  k.abuf = nullptr; k.a_ld = 0; k.a_sz = 0; k.a_sa = 0; k.counters = nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(num_tiles < g_sm_count ? num_tiles : g_sm_count, 1, 1);
  if (KL > 0) cfg.gridDim = dim3(vocab_topk_plan(a.N, a.rows).grid, 1, 1);   // the schedule beam_select recomputes
  cfg.blockDim = dim3(PS_THREADS, 1, 1);
  cfg.dynamicSmemBytes = PS_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CAPDEC_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, mW, mX, k, tiles_n, tiles_r, num_tiles));
  count_launch();
  return CAPDEC_OK;
}

bool persist_enabled() {
  const char* s = getenv("CAPDEC_GEMM_PERSIST");        // read per call: tests flip it
  return !(s && s[0] == '0');
}

template <int NACC, int EPI>
int launch_rows(const GemmArgs& a, cudaStream_t st) {
  const int rows = (EPI == EPI_DHCELL && a.e.rows_epi > a.rows) ? a.e.rows_epi : a.rows;
  // batch-side tile: with few output tiles (one 128-row batch tile x N/128 weight tiles < one wave of CTAs) the K loop
  // had to be cut in 3..16 slices that meet in fp32 atomics -- at 128 captions per GPU and D = F = 1024 every in-loop GEMM
  // of the per-step chains took ~15 us whatever its size.  Narrower batch tiles make the wave out of TILES first (the
  // weight tile is re-read from L2 once per batch tile) and leave split-K for the long-K / few-tile products:
  // 13.2 -> 12.4 ms per step at the config-5 shape (profiles/r2b_scaled_gemm_sweep.txt).
  int cap = 128;
  if (a.splitk < 0 && rows > 32) {
    const int64_t wt = (int64_t)ceil_div(a.N, BM) * (NACC > 1 ? 1 : a.batch);
    if (wt * ceil_div(rows, 128) < 148) cap = wt * ceil_div(rows, 64) >= 148 ? 64 : 32;
  }
  if (const char* e = getenv("CAPDEC_GEMM_ROWTILE")) cap = atoi(e);        // experiments: forced batch-side tile
  if (rows <= 32 || cap <= 32) return launch<32, NACC, EPI>(a, st);
  if (rows <= 64 || NACC > 1 || cap <= 64) return launch<64, NACC, EPI>(a, st);
  return launch<128, NACC, EPI>(a, st);
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
  });
  return g_init_rc;
}

// fused CELL epilogue keeps 4 accumulators of BNR columns each: rows <= 64 per tile row
bool gemm_tc_cell_fusable(int rows) { return rows <= 64; }

int gemm_tc(const GemmArgs& a, cudaStream_t st) {
  if (a.rows <= 0 && !(a.epi == EPI_DHCELL && a.e.rows_epi > 0)) return CAPDEC_OK;
  if (a.N <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(a.X && a.W && a.K > 0, CAPDEC_ERR_BAD_ARG, "gemm_tc: null operand");
  CAPDEC_TRY(gemm_tc_init());
  switch (a.epi) {
    case EPI_PLAIN:
      CAPDEC_REQUIRE(a.out, CAPDEC_ERR_BAD_ARG, "gemm_tc: null output");
      // more than one wave of 128 x 128 tiles and no split-K: the persistent multi-tile kernel
      // (measured: it wins for many short-K tiles -- vocabulary projection 22.7 vs 27.1 us -- and loses to
      // two co-resident single-tile CTAs per SM when K is long or the tiles are few)
      if (persist_enabled() && (a.splitk == 0 || a.out_ft) && a.rows > 128 && a.K <= 1024 &&
          (int64_t)ceil_div(a.N, BM) * ceil_div(a.rows, PS_BNR) * a.batch > 4 * 148) {
        switch (a.tn & 3) {
          case 0: return launch_persist<0>(a, st);
          case 1: return launch_persist<1>(a, st);
          case 2: return launch_persist<2>(a, st);
          default: return launch_persist<3>(a, st);
        }
      }
      if (a.tn) {
        // transposed operand(s): 64- or 128-wide row tiles (one or two 64-column TMA boxes)
        const bool small = a.rows <= 64;
        switch (a.tn & 3) {
          case 1: return small ? launch<64, 1, EPI_PLAIN, 1>(a, st) : launch<128, 1, EPI_PLAIN, 1>(a, st);
          case 2: return small ? launch<64, 1, EPI_PLAIN, 2>(a, st) : launch<128, 1, EPI_PLAIN, 2>(a, st);
          default: return small ? launch<64, 1, EPI_PLAIN, 3>(a, st) : launch<128, 1, EPI_PLAIN, 3>(a, st);
        }
      }
      return launch_rows<1, EPI_PLAIN>(a, st);
    case EPI_G1: return launch_rows<1, EPI_G1>(a, st);
    case EPI_P3: return launch_rows<1, EPI_P3>(a, st);
    case EPI_WR: return launch_rows<1, EPI_WR>(a, st);
    case EPI_DHCELL: return launch_rows<1, EPI_DHCELL>(a, st);
    case EPI_CELL:
      CAPDEC_REQUIRE(a.batch == 4 && a.rows <= 64, CAPDEC_ERR_BAD_SHAPE,
                     "gemm_tc: fused cell epilogue needs batch == 4 gates and rows <= 64 (rows=%d)", a.rows);
      return launch_rows<4, EPI_CELL>(a, st);
  }
  set_error("gemm_tc: bad epilogue %d", a.epi);
  return CAPDEC_ERR_BAD_ARG;
}

// schedule of the fused vocabulary kernel: grid, tiles and partial slots per batch row (a row tile's vocabulary
// tiles are spread over at most `slots` consecutive CTA runs)
VocabTopkPlan vocab_topk_plan(int rows, int V) {
  VocabTopkPlan p;
  p.tiles_r = ceil_div(V, PS_BNR);
  p.num_tiles = ceil_div(rows, BM) * p.tiles_r;
  p.grid = p.num_tiles < 148 ? p.num_tiles : 148;
  p.slots = 2 * ((int)(((int64_t)p.tiles_r * p.grid + p.num_tiles - 1) / p.num_tiles) + 1);   // two column halves per run
  return p;
}
size_t vocab_topk_part_floats(int rows, int V, int kl) {
  return (size_t)rows * (size_t)vocab_topk_plan(rows, V).slots * (size_t)(2 + 2 * kl);
}

// Fused vocabulary projection + log-softmax statistics + top-kl candidates (gemm_tc_persist_kernel<0, KL>):
// part[row][tile][2 + 2 kl] = {max, sum exp(x - max), kl largest logits, their vocabulary indices} per 128-entry
// vocabulary tile.  H [rows][K] and Wfc [V][K] are bf16, K contiguous.
int gemm_tc_vocab_topk(const void* H, int64_t ldh, int rows, const void* Wfc, int64_t ldw, int V, int K,
                       const float* bias, float* part, int kl, cudaStream_t st) {
  if (rows <= 0 || V <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(H && Wfc && part && K > 0 && (kl == 4 || kl == 8), CAPDEC_ERR_BAD_ARG, "gemm_tc_vocab_topk: bad argument");
  CAPDEC_TRY(gemm_tc_init());
  CAPDEC_REQUIRE(bias != nullptr, CAPDEC_ERR_BAD_ARG, "gemm_tc_vocab_topk: bias is NULL");
  GemmArgs a;
  a.W = H; a.ldw = ldh; a.N = rows;          // M side: batch rows
  a.X = Wfc; a.ldx = ldw; a.rows = V;        // N side: vocabulary entries
  a.K = K; a.out = part; a.bias = bias;
  a.e.topk_slots = vocab_topk_plan(rows, V).slots;
  return kl == 4 ? launch_persist<0, 4>(a, st) : launch_persist<0, 8>(a, st);
}

}  // namespace capdec
