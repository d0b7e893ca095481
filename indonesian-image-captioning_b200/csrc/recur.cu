// recur.cu -- the persistent decoder kernels: the whole teacher-forced time loop of the SCN-LSTM / attention
// decoders, and its reverse-time gradient, in ONE cooperative launch each (bf16 feature mode).
//
// Reference math: models/decoders/attention_scn.py:139-156 (the `for t` loop),
// models/scn_cell.py:52-154, models/attention.py:26-44; restated in SURVEY.md App. A.1 / A.2.
//
// Why: at 32 captions per GPU a decode step is ~18 MFLOP per row and 1 MB of features per
// row -- every stage finishes in 1-3 us, so a chain of 7 dependent kernel launches per step
// (>= 3 us each even with programmatic dependent launch inside a CUDA graph) is pure latency.
// Here one CTA per SM stays resident for all T steps:
//   * the recurrent and factor weights (W_d | W_beta | W_ha, W_ia[M:], W_ic | W_hc; 17.3 MB in
//     bf16 at the 512-dim config) are sliced BY OUTPUT FEATURE over the CTAs and loaded into shared
//     memory ONCE; each CTA owns 16-32 output features of every GEMM with the full K, so no
//     split-K, no atomics, no zero-filled accumulators, bit-reproducible sums;
//   * TWO ROW GROUPS PER CTA: batch rows are independent through the whole recurrence, so the 512
//     threads of a CTA are two groups of 256 (group g owns rows [16 g, 16 g + 16) = one m16 tile)
//     that run the time loop INDEPENDENTLY of each other -- own staging buffers, own mbarriers, own
//     named barrier (bar.sync 1 + g) -- and share only the resident weight slices;
//   * DATAFLOW, NOT BARRIERS: the phases of a step hand their results to the other CTAs through per-step
//     buffers pre-filled with a NaN pattern; a consumer polls the data itself and multiplies straight out of
//     the loaded registers (see "Dataflow exchange" below).  The first version separated the phases by grid
//     barriers (6 per step, >= 1.3 us each, half of the step) and staged every operand through shared memory;
//   * a GEMM phase of a group = warp w takes 1/8 of K, reads its 16 x K/8 activation slice from the exchange
//     buffer with 16-byte relaxed loads straight into mma fragments (K permuted identically for both
//     operands), mma.sync m16n8k16 bf16 -> fp32, 8-way reduction through shared memory, fused epilogue
//     (bias, factor products u*v / p*q written as the next GEMM's operand);
//   * attention = scores per (row, pixel) item, softmax + weighted sum + gate per (row, 256-channel
//     chunk) item, features streamed through the staging buffers (one TMA bulk copy per 32 pixels, the first
//     copies in flight while the phase still waits for its inputs).
// tcgen05 is not used here on purpose: its 128-lane M granularity would force 8-way split-K
// with atomics (and a seventh phase to consume the sums) for GEMMs whose whole tensor work is
// 0.3 us per step; the batched GEMMs outside the loop (vocabulary, att1, embedding side,
// weight gradients) stay on the tcgen05 engine (gemm_tc.cu).
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int RT = 512;             // threads per CTA = two row groups
constexpr int GT = 256;             // threads per row group
constexpr int GW = GT / 32;         // 8 warps per group: warp w of a GEMM phase = K slice w
constexpr int GR = 16;              // rows per group = one m16 tile
constexpr int KSL = GW;             // K slices of a GEMM phase
constexpr int KC = 512;             // K elements per staged activation chunk (1 KB per row)
constexpr int WPAD = 64;            // bytes of padding per resident weight row (bank spread)
constexpr int STAGE = GR * KC * 2;  // one staging buffer: 16 rows of KC bf16 = 16 KB; two per group
constexpr int REDLD = 16 + 8;       // floats per row of the cross-warp reduction scratch (2 n-tiles)
constexpr int REDF = KSL * GR * REDLD;      // floats of one group's reduction scratch (12 KB)
constexpr int CHUNK = 256;          // channels per weighted-sum work item (512 B per pixel)
constexpr int NCOL = CHUNK / 8;     // 16-byte columns per chunk = one warp
constexpr int GROUPS = GT / NCOL;   // pixel groups of the weighted sum (8 = the warps of the row group)
constexpr int WPXS = STAGE / (CHUNK * 2);   // pixels per weighted-sum stage (32)
constexpr int SCI = 2 * STAGE / 1024;       // score items (<= 1 KB each) the two stages hold (32)
// cycles row group 1 starts after group 0 (attention decoders; CAPDEC_RECUR_SKEW overrides both).  Measured at the
// config-3 shape: forward 1 249 / 1 205 / 1 225 / 1 234 us and backward 1 315 / 1 306 / 1 280 / 1 279 us per launch
// for 0 / 8 000 / 16 000 / 24 000 cycles; separately (CAPDEC_RECUR_SKEW_FWD / _BWD): forward 1 240 / 1 202 / 1 202 us
// for 4 000 / 8 000 / 12 000, backward 1 292 / 1 277 / 1 270 us for 12 000 / 20 000 / 30 000.
constexpr int RECUR_SKEW_FWD = 8000, RECUR_SKEW_BWD = 20000;
constexpr int QW = 64;              // attention channels per item of backward phase B
constexpr int BPX = STAGE / (QW * 2);       // pixels per fill of phase B (128)
static_assert(NCOL == 32, "weighted-sum mapping: one warp per pixel group");
static_assert(REDF >= 2 * GW + (GROUPS - 1) * NCOL * 8, "reduction scratch too small for the weighted sum");
static_assert(2 * STAGE >= 2 * GW * 2 * QW * 4, "staging buffers too small for the phase-B slab");
static_assert(REDF >= CHUNK / 2 + 4 * GT, "reduction scratch too small for phase A (dawe + 4 x pad4(P) partials, P <= GT)");
static_assert(WPXS == 32 && GW == 8 && CHUNK == 256, "phase A: 2 pixel tiles x 4 channel quarters per fill");

__host__ __device__ __forceinline__ int pad4i(int x) { return (x + 3) & ~3; }

// Debug build only (-DCAPDEC_RECUR_FINE): tagged clock stamps of thread 0 of CTA 0 (row group 0) inside the phases,
// printed by the launchers when CAPDEC_RECUR_PROF=1.
#ifdef CAPDEC_RECUR_FINE
constexpr int FINE_N = 1 << 16;
__device__ unsigned long long d_fine[FINE_N];
__device__ unsigned d_fine_idx;
#define FSTAMP(tag)                                                                                   \
  do {                                                                                                \
    if (blockIdx.x == 0 && threadIdx.x == 0 && G.fidx < FINE_N)                                       \
      d_fine[G.fidx++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); \
  } while (0)
#define FSTAMP_FLUSH() do { if (blockIdx.x == 0 && threadIdx.x == 0) d_fine_idx = G.fidx; } while (0)
#else
#define FSTAMP(tag) do { } while (0)
#define FSTAMP_FLUSH() do { } while (0)
#endif

__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// MUFU-based transcendentals (abs. error ~1e-6; the results are rounded to bf16 operands anyway)
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.0f * fsigmoid(2.0f * x) - 1.0f; }

// Shared-memory pipeline state of one row group: two 16 KB staging buffers (adjacent), each with a
// "full" mbarrier.
struct Pipe {
  uint8_t* stg;        // generic pointer to stage 0; stage 1 = stg + STAGE
  uint32_t stg_a;      // shared-memory address of stage 0
  uint32_t full0;      // mbarrier address of stage 0; stage 1's is 8 bytes further
  uint32_t phase;      // bit s = parity of the next wait on stage s
};
// no dynamically indexed members: the whole group state stays in registers
__device__ __forceinline__ uint32_t pipe_full(const Pipe& pp, int s) { return pp.full0 + 8u * (uint32_t)s; }
__device__ __forceinline__ void pipe_wait(Pipe& pp, int s) {
  mbar_wait(pipe_full(pp, s), (pp.phase >> s) & 1u);
  pp.phase ^= 1u << s;
}

// Everything one row group owns.
struct Grp {
  int g;               // 0 / 1
  int tid, warp;       // thread / warp index inside the group
  int row0;            // first batch row of the group (16 g)
  int vcta;            // CTA index used to deal out (row, ...) work items: reversed for group 1, so that the two
                       // groups' items of a partially filled wave land on different SMs
  Pipe pp;
  float* red;          // [REDF]
  float* al;           // [pad4(P)]
  unsigned* abortp;    // dataflow polls: abort flag (a poll timed out)
  long long t_end;     // ... and the clock64 value after which a waiting poll raises it
#ifdef CAPDEC_RECUR_FINE
  unsigned fidx;
#endif
};

// barrier over the 256 threads of a row group
__device__ __forceinline__ void gsync(const Grp& G) {
  asm volatile("bar.sync %0, %1;" ::"r"(G.g + 1), "n"(GT) : "memory");
}

// the group's leader copies `bytes` contiguous bytes into stage s
__device__ __forceinline__ void stage_fill(const Grp& G, int s, const void* src, uint32_t bytes) {
  if (G.tid == 0) {
    mbar_expect_tx(pipe_full(G.pp, s), bytes);
    bulk_g2s(G.pp.stg_a + s * STAGE, src, bytes, pipe_full(G.pp, s));
  }
}

// =====================================================================================
// Dataflow exchange between CTAs: NO grid barriers in the forward time loop.
// Whatever one CTA hands to the others inside the loop (h_t, att2 | beta_pre, the attention scores, z, the factor
// products m, the gate pre-activations) lives in a PER-STEP buffer that the launcher fills with an all-ones bit
// pattern (0xFFFF per bf16 / 0xFFFFFFFF per float: a NaN encoding arithmetic never produces -- cvt.rn and the FPU
// emit the canonical 0x7FFF / 0x7FFFFFFF).  A consumer polls THE DATA ITSELF with relaxed gpu-scope loads until no
// element shows the pattern and multiplies straight out of the loaded registers: no arrival counter, no release /
// acquire fence pair, no staging copy -- the hand-off costs one L2 round trip after the producer's store has landed
// (a grid barrier + bulk copy was >= 3 100 cycles, tools/barrier_bench.cu).  Stores to these buffers are relaxed
// gpu-scope stores; every element is written once by exactly one thread and never changes afterwards, so a value
// that is not the pattern is final.  All CTAs are co-resident (cooperative launch) and every phase only waits for
// data of earlier phases, so the waits cannot form a cycle.  A poll that is still waiting ~2 s after the launch
// raises the abort flag (first word of `bar`), which makes every other poll give up too: the kernel then ends with
// garbage instead of hanging the device.
// =====================================================================================
constexpr uint32_t SENT = 0xFFFFFFFFu;
constexpr long long POLL_LIMIT_CYCLES = 4000000000ll;

__device__ __forceinline__ uint4 ldx16(const void* p) {
  uint4 r;
#ifdef CAPDEC_POLL_VOLATILE
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#else
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#endif
  return r;
}
__device__ __forceinline__ uint32_t ldx4(const void* p) {
  uint32_t r;
#ifdef CAPDEC_POLL_VOLATILE
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
#else
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
#endif
  return r;
}
__device__ __forceinline__ void stx16(void* p, const uint4& v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stx4(void* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stx4f(float* p, float v) { stx4(p, __float_as_uint(v)); }
__device__ __forceinline__ void stx2(bf16* p, bf16 v) {
  asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(p), "h"(__bfloat16_as_ushort(v)) : "memory");
}
// any bf16 lane / any float of the 16 bytes still the fill pattern?
__device__ __forceinline__ bool sent16(const uint4& v) {
  return (__vcmpeq2(v.x, SENT) | __vcmpeq2(v.y, SENT) | __vcmpeq2(v.z, SENT) | __vcmpeq2(v.w, SENT)) != 0u;
}
__device__ __forceinline__ bool sent32(const uint4& v) { return v.x == SENT || v.y == SENT || v.z == SENT || v.w == SENT; }

// called every 64 spins of a poll loop: true -> give up (somebody timed out, or this poll just did)
__device__ __noinline__ bool poll_stalled(unsigned* abortp, long long t_end) {
  if (ldx4(abortp) != 0u) return true;
  if (clock64() > t_end) {
    stx4(abortp, 1u);
    return true;
  }
  return false;
}
template <bool F32>
__device__ __forceinline__ uint4 poll16(const Grp& G, const void* p, uint4 v) {
  unsigned spins = 0;
  while (F32 ? sent32(v) : sent16(v)) {
    if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
    v = ldx16(p);
  }
  return v;
}
template <bool F32>
__device__ __forceinline__ uint4 poll16(const Grp& G, const void* p) { return poll16<F32>(G, p, ldx16(p)); }
__device__ __forceinline__ float poll4f(const Grp& G, const float* p) {
  uint32_t v = ldx4(p);
  unsigned spins = 0;
  while (v == SENT) {
    if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
    v = ldx4(p);
  }
  return __uint_as_float(v);
}

// N floats, `stride` elements apart, polled together in rounds
template <int N>
__device__ __forceinline__ void poll4f_n(const Grp& G, const float* p, int64_t stride, int n, float (&out)[N]) {
  uint32_t raw[N];
#pragma unroll
  for (int k = 0; k < N; ++k) raw[k] = k < n ? ldx4(p + (int64_t)k * stride) : 0u;
  for (unsigned spins = 0;;) {
    bool pend = false;
#pragma unroll
    for (int k = 0; k < N; ++k) pend = pend || raw[k] == SENT;
    if (!pend) break;
    if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (raw[k] == SENT) raw[k] = ldx4(p + (int64_t)k * stride);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) out[k] = __uint_as_float(raw[k]);
}

// One GEMM job of a row group with the activation operand taken STRAIGHT FROM THE EXCHANGE BUFFER in global memory
// (see above) into mma fragments: NH * 16 output features with the full K,
//   out[h] = sum_k A[row, k] * W[h*16 + j, k]       (row = G.tid / 16, j = G.tid % 16)
// A: NF "fills" of n rows x KF (= 64 * BPW * 4) bf16, fill kf at src + kf * fill_stride, rows KF elements apart
// (chunk-major [K/512][B][512] copies, or one dense row of 2F).  Warp ks multiplies the 16 rows with BPW 32-wide k
// blocks of every fill; each lane fetches 16 bytes (8 consecutive k) per row and block -- the same k permutation
// is used for the weight fragments (shared memory, resident), so no ldmatrix / transposition is needed.  One lane
// per warp first spins on the warp's first 16 bytes (row 0 is always live), so that a waiting group costs the L2
// one sector per warp and round trip; then all loads are issued at once and late elements are re-polled one by one.
template <int NH, int BPW, int NF>
__device__ __forceinline__ void gemm_job_df(Grp& G, const bf16* __restrict__ src, int64_t fill_stride, int n,
                                           const uint8_t* Ws, int wstride, float (&out)[NH]) {
  constexpr int KF = 64 * BPW * 4;
  const int lane = G.tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int ks = G.warp;
  const bool lo = g < n, hi = g + 8 < n;
  const uint8_t* a_lo = reinterpret_cast<const uint8_t*>(src) + ((size_t)g * KF + ks * (BPW * 32)) * 2 + 16 * c;
  const uint8_t* a_hi = a_lo + (size_t)8 * KF * 2;
  FSTAMP(100);
#ifndef CAPDEC_NO_PROBE
  if (lane == 0) (void)poll16<false>(G, a_lo);
  __syncwarp();
#endif
  FSTAMP(101);
  uint4 alo[NF][BPW], ahi[NF][BPW];
#pragma unroll
  for (int kf = 0; kf < NF; ++kf)
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
      alo[kf][j] = lo ? ldx16(a_lo + (size_t)kf * fill_stride * 2 + j * 64) : make_uint4(0, 0, 0, 0);
      ahi[kf][j] = hi ? ldx16(a_hi + (size_t)kf * fill_stride * 2 + j * 64) : make_uint4(0, 0, 0, 0);
    }
  float accb[2 * NH][BPW][4];
#pragma unroll
  for (int nt = 0; nt < 2 * NH; ++nt)
#pragma unroll
    for (int j = 0; j < BPW; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) accb[nt][j][i] = 0.f;
  const uint8_t* w_base = Ws + (size_t)g * wstride + ks * (BPW * 64) + 16 * c;
  // elements that had not arrived yet are re-requested IN ROUNDS, all pending ones of the lane together: one L2
  // round trip per round whatever their number (one after the other, 16 late elements cost 16 round trips)
  for (unsigned spins = 0;;) {
    bool pend = false;
#pragma unroll
    for (int kf = 0; kf < NF; ++kf)
#pragma unroll
      for (int j = 0; j < BPW; ++j) pend = pend || (lo && sent16(alo[kf][j])) || (hi && sent16(ahi[kf][j]));
    if (!pend) break;
    if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
#pragma unroll
    for (int kf = 0; kf < NF; ++kf)
#pragma unroll
      for (int j = 0; j < BPW; ++j) {
        if (lo && sent16(alo[kf][j])) alo[kf][j] = ldx16(a_lo + (size_t)kf * fill_stride * 2 + j * 64);
        if (hi && sent16(ahi[kf][j])) ahi[kf][j] = ldx16(a_hi + (size_t)kf * fill_stride * 2 + j * 64);
      }
  }
#pragma unroll
  for (int kf = 0; kf < NF; ++kf) {
    const uint8_t* wp = w_base + (size_t)kf * KF * 2;
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
      const uint4 x = alo[kf][j], y = ahi[kf][j];
#pragma unroll
      for (int nt = 0; nt < 2 * NH; ++nt) {
        const uint4 b = *reinterpret_cast<const uint4*>(wp + (size_t)nt * 8 * wstride + j * 64);
        mma_bf16(accb[nt][j], x.x, y.x, x.y, y.y, b.x, b.y);
        mma_bf16(accb[nt][j], x.z, y.z, x.w, y.w, b.z, b.w);
      }
    }
  }
  FSTAMP(103);
  float acc[2 * NH][4];
#pragma unroll
  for (int nt = 0; nt < 2 * NH; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float sum = accb[nt][0][i];
#pragma unroll
      for (int jb = 1; jb < BPW; ++jb) sum += accb[nt][jb][i];
      acc[nt][i] = sum;
    }
  const int row = G.tid >> 4, j = G.tid & 15;
  float* mine = G.red + (ks * GR + g) * REDLD + 2 * c;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    if (h > 0) gsync(G);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      *reinterpret_cast<float2*>(mine + q * 8) = make_float2(acc[2 * h + q][0], acc[2 * h + q][1]);
      *reinterpret_cast<float2*>(mine + 8 * REDLD + q * 8) = make_float2(acc[2 * h + q][2], acc[2 * h + q][3]);
    }
    gsync(G);
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < KSL; ++w) sum += G.red[(w * GR + row) * REDLD + j];
    out[h] = sum;
  }
  gsync(G);                                  // `red` is free again
  FSTAMP(104);
}

// Two GEMM jobs of a row group AT ONCE (backward phase H: the CTA's two (d-slice, K-chunk) jobs): different activation
// operands (n rows x 512 each), adjacent 16-row weight blocks in shared memory (job h = rows [16 h, 16 h + 16) of Ws).
// The operand loads of BOTH jobs are in flight together and share one poll phase -- run one after the other, the second
// job paid a second probe + load round trip (~600 cycles) on the CTAs that are last to finish the step.
__device__ __forceinline__ void gemm_job_df_pair(Grp& G, const bf16* __restrict__ src0, const bf16* __restrict__ src1, int n,
                                                 const uint8_t* Ws, int wstride, float (&out)[2]) {
  constexpr int BPW = 2, KF = 64 * BPW * 4;
  const int lane = G.tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int ks = G.warp;
  const bool lo = g < n, hi = g + 8 < n;
  const size_t off_lo = ((size_t)g * KF + ks * (BPW * 32)) * 2 + 16 * c, off_hi = off_lo + (size_t)8 * KF * 2;
  const uint8_t* a0 = reinterpret_cast<const uint8_t*>(src0);
  const uint8_t* a1 = reinterpret_cast<const uint8_t*>(src1);
#ifndef CAPDEC_NO_PROBE
  if (lane == 0) (void)poll16<false>(G, a0 + off_lo);
  __syncwarp();
#endif
  uint4 alo[2][BPW], ahi[2][BPW];
#pragma unroll
  for (int j = 0; j < BPW; ++j) {
    alo[0][j] = lo ? ldx16(a0 + off_lo + j * 64) : make_uint4(0, 0, 0, 0);
    ahi[0][j] = hi ? ldx16(a0 + off_hi + j * 64) : make_uint4(0, 0, 0, 0);
    alo[1][j] = lo ? ldx16(a1 + off_lo + j * 64) : make_uint4(0, 0, 0, 0);
    ahi[1][j] = hi ? ldx16(a1 + off_hi + j * 64) : make_uint4(0, 0, 0, 0);
  }
  float accb[4][BPW][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int j = 0; j < BPW; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) accb[nt][j][i] = 0.f;
  for (unsigned spins = 0;;) {                   // late elements: re-requested in rounds (see gemm_job_df)
    bool pend = false;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < BPW; ++j) pend = pend || (lo && sent16(alo[h][j])) || (hi && sent16(ahi[h][j]));
    if (!pend) break;
    if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint8_t* ap = h ? a1 : a0;
#pragma unroll
      for (int j = 0; j < BPW; ++j) {
        if (lo && sent16(alo[h][j])) alo[h][j] = ldx16(ap + off_lo + j * 64);
        if (hi && sent16(ahi[h][j])) ahi[h][j] = ldx16(ap + off_hi + j * 64);
      }
    }
  }
  const uint8_t* w_base = Ws + (size_t)g * wstride + ks * (BPW * 64) + 16 * c;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
      const uint4 x = alo[h][j], y = ahi[h][j];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int nt = 2 * h + q;
        const uint4 b = *reinterpret_cast<const uint4*>(w_base + (size_t)nt * 8 * wstride + j * 64);
        mma_bf16(accb[nt][j], x.x, y.x, x.y, y.y, b.x, b.y);
        mma_bf16(accb[nt][j], x.z, y.z, x.w, y.w, b.z, b.w);
      }
    }
  const int row = G.tid >> 4, jj = G.tid & 15;
  float* mine = G.red + (ks * GR + g) * REDLD + 2 * c;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h > 0) gsync(G);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int nt = 2 * h + q;
      *reinterpret_cast<float2*>(mine + q * 8) = make_float2(accb[nt][0][0] + accb[nt][1][0], accb[nt][0][1] + accb[nt][1][1]);
      *reinterpret_cast<float2*>(mine + 8 * REDLD + q * 8) = make_float2(accb[nt][0][2] + accb[nt][1][2], accb[nt][0][3] + accb[nt][1][3]);
    }
    gsync(G);
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < KSL; ++w) sum += G.red[(w * GR + row) * REDLD + jj];
    out[h] = sum;
  }
  gsync(G);                                  // `red` is free again
}

// copy `nrows` weight rows (K bf16 each, global pitch ldw elements) into shared memory rows of
// K*2 + WPAD bytes; rows at or beyond `valid` are zero-filled.  Whole CTA.
__device__ __noinline__ void load_weight_rows(uint8_t* Ws, const bf16* Wg, int64_t ldw, int K, int nrows,
                                              int valid) {
  const int vec_per_row = K / 8;
  const int wstride = K * 2 + WPAD;
  for (int i = threadIdx.x; i < nrows * vec_per_row; i += RT) {
    const int r = i / vec_per_row, cidx = i - r * vec_per_row;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < valid) val = __ldg(reinterpret_cast<const uint4*>(Wg + (int64_t)r * ldw) + cidx);
    *reinterpret_cast<uint4*>(Ws + (size_t)r * wstride + (size_t)cidx * 16) = val;
  }
}

// set up the row group of the calling thread; `stg`: 4 staging buffers, `red`: 2 x REDF floats, `al`: 2 x alw
// floats, `bars`: 4 mbarriers (initialised by thread 0 of the CTA before the first __syncthreads)
__device__ __forceinline__ void grp_init(Grp& G, uint8_t* stg, float* red, float* al, int alw, uint64_t* bars,
                                         unsigned* gbar) {
  G.g = threadIdx.x / GT;
  G.tid = threadIdx.x - G.g * GT;
  G.warp = G.tid >> 5;
  G.row0 = G.g * GR;
  G.vcta = G.g ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  G.pp.stg = stg + (size_t)G.g * 2 * STAGE;
  G.pp.stg_a = smem_u32(G.pp.stg);
  G.pp.full0 = smem_u32(&bars[2 * G.g]);
  G.pp.phase = 0;
  G.red = red + (size_t)G.g * REDF;
  G.al = al + (size_t)G.g * alw;
  G.abortp = gbar + 16;
  G.t_end = clock64() + POLL_LIMIT_CYCLES;
#ifdef CAPDEC_RECUR_FINE
  G.fidx = 0;
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
}

// The two row groups of a CTA run the same phase sequence; started together they also hit the L2-bound streaming
// phases (weighted sum, backward phase A) together and share the L2 slice throughput.  Starting group 1 a fraction of a
// step later keeps them out of phase for the whole launch: one group streams while the other multiplies.
__device__ __forceinline__ void group_skew(const Grp& G, int cycles) {
  if (G.g == 1 && cycles > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(200);
  }
}

struct FwdP {
  int B, T, P, E, A, M, D, F, NQ, NG1;
  int64_t R;                     // B*T
  const int32_t* len;            // [B] decode lengths, sorted descending
  // packed bf16 weights, K contiguous
  const bf16* Wcat1; int64_t ldD;   // [NG1][ldD]   [W_d ; W_beta ; W_ha^T]
  const bf16* Wxz; int64_t ldX;     // [NQ][ldX]    W_ia[M:, :]^T (already offset by M)
  const bf16* Wc; int64_t ld2F;     // [4][D][ld2F] [W_ic_g | W_hc_g]
  const float* b_cat1; const float* b_ih; const float* b_hh;
  const bf16* att1;                 // [B][P][A]
  const bf16* enc_cm;               // [B][E/256][P][256]   chunk-major copy of the features
  const float* w_f; const float* b_f;
  const float* v; const float* q;   // [B][NQ]
  const bf16* H0;                   // [B][D]
  bf16* Hall; bf16* Hd;             // (B, T, D)
  bf16* Ht;                         // [T][B][D]   time-major copy of h_t: the next step's G1 operand
  float* C;                         // [T+1][B][D]
  float* U;                         // [T][B][NQ]   in: Emb W_ia[:M] ; out: u
  float* g1;                        // [T][B][NG1]  att2 | beta_pre | p
  float* alphas;                    // (B, T, P)
  float* awe;                       // [T][B][E] or null
  bf16* z;                          // [T][B][E]
  bf16* zk;                         // [T][E/512][B][512]   chunk-major copy of z: the P3 operand
  bf16* m;                          // [4][R][2F]
  float* pre;                       // [T][B][4D]   (LSTM: z W_ih[:, M:]^T + Emb W_ih[:, :M]^T, [T][B][NQ])
  float* gates;                     // [T][B][4D]
  float* scores;                    // [T][B][pad4(P)]   per step: an exchange buffer (see "Dataflow exchange")
  unsigned* bar;                    // two counters, 128 bytes apart
  float dropout_p; const uint64_t* seed;
  long long* prof;                  // debug (CAPDEC_RECUR_PROF=1): [T][16] clock64 stamps of CTA 0, group 0
  int mask;                         // debug: bit i set -> phase i runs (G1, scores, wsum, P3, P4, cell)
  int skew;                         // cycles row group 1 starts after group 0 (see group_skew)
};

// LSTM = true: the pure_attention decoder (nn.LSTMCell on [emb ; z], pure_attention.py:143-146, gate order
// i,f,g,o): the G1 job yields [att2 | beta_pre | h W_hh^T], P3 adds z W_ih[:, M:]^T to the batched embedding
// part (result in `pre`), and the cell phase follows directly (no factor products, no P4).
//
// Phases of one step and what they wait for (dataflow, see "Dataflow exchange" above -- no grid barrier):
//   G1     h_{t-1} (Ht)                  -> g1 = [att2 | beta_pre | p],  m[:, F:] = p*q
//   scores att2 rows of g1               -> scores[t]          (att1 slice prefetched by TMA at the top of the step)
//   wsum   scores[t] row, beta_pre       -> alpha, awe, z, zk  (enc chunk streamed through the staging ring)
//   P3     zk[t]                         -> u (in place over U_emb), m[:, :F] = u*v      (LSTM: pre)
//   P4     m[t] (both halves)            -> pre[t]
//   cell   pre[t] (LSTM: pre + g1)       -> c_t (register + C), gates, h_t -> Hall / Hd / Ht
template <bool ATT, int NT1, bool LSTM = false>
__global__ void __launch_bounds__(RT, 1) recur_fwd_kernel(const __grid_constant__ FwdP p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT3 = 2, NT4 = 2;
  const int D = p.D, E = p.E, F = p.F, B = p.B, T = p.T, P = p.P, NQ = p.NQ, NG1 = p.NG1, A = p.A;
  const int lane = threadIdx.x & 31;
  // shared memory carve-up: 4 stages | W1 | W3 | W4 | red[2] | al[2] | lens | mbarriers[4]
  const int w1s = D * 2 + WPAD, w3s = E * 2 + WPAD, w4s = 2 * F * 2 + WPAD;
  const int alw = pad4i(P > 0 ? P : 4);
  uint8_t* W1s = smem + 4 * STAGE;
  uint8_t* W3s = W1s + (size_t)NT1 * 8 * w1s;
  uint8_t* W4s = W3s + (ATT ? (size_t)NT3 * 8 * w3s : 0);
  float* red = reinterpret_cast<float*>(W4s + (LSTM ? 0 : (size_t)NT4 * 8 * w4s));
  float* al = red + 2 * REDF;
  int* lens = reinterpret_cast<int*>(al + 2 * alw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(lens + ((B + 3) & ~3) + 2);
  bars = reinterpret_cast<uint64_t*>(((uintptr_t)bars + 7) & ~(uintptr_t)7);
  Grp G;
  grp_init(G, smem, red, al, alw, bars, p.bar);
  uint8_t* const stg = G.pp.stg;
  const int tid = G.tid, warp = G.warp, row0 = G.row0;

  // ---- which output features this CTA owns (both groups: same weights, different rows) ----
  const int f1 = blockIdx.x * NT1 * 8;                 // first G1 feature
  const int f3 = blockIdx.x * NT3 * 8;                 // first P3 feature (u)
  const int tpg = D / 8;
  const int t4 = blockIdx.x * NT4;                     // first P4 tile; tiles never straddle a gate
  const int gate4 = t4 / tpg, d4 = (t4 - gate4 * tpg) * 8;
  const bool has1 = f1 < NG1, has3 = ATT && f3 < NQ, has4 = !LSTM && t4 < 4 * tpg;
  // ---- one-time: weight slices -> shared memory ----
  if (has1) load_weight_rows(W1s, p.Wcat1 + (int64_t)f1 * p.ldD, p.ldD, D, NT1 * 8, NG1 - f1);
  if (has3) load_weight_rows(W3s, p.Wxz + (int64_t)f3 * p.ldX, p.ldX, E, NT3 * 8, NQ - f3);
  if (has4) load_weight_rows(W4s, p.Wc + ((int64_t)gate4 * D + d4) * p.ld2F, p.ld2F, 2 * F, NT4 * 8, D - d4);
  for (int i = threadIdx.x; i < B; i += RT) lens[i] = p.len[i];
  __syncthreads();
  if (row0 >= B) return;                               // this group has no rows at all (B <= 16)
  group_skew(G, p.skew);

  const int lrow = tid >> 4, ej = tid & 15;            // epilogue mapping of gemm_job (row inside the group)
  const int erow = row0 + lrow;                        // batch row
  const int col0 = ATT ? A + E : 0;
  const int Ppad = pad4i(P);
  const int chunks = ATT ? E / CHUNK : 1;
  const int grp = warp, col = lane;                    // weighted-sum mapping: warp = pixel group, lane = 16-byte column
  const float drop_p = p.dropout_p;
  const uint64_t seed = drop_p > 0.f ? __ldg(p.seed) : 0ull;
  const int a_lane = lane * 8;                         // scores: this lane's 8 attention channels (x2)
  const int nctas = gridDim.x;
  const int nfill3 = E / KC;
  // the cell element (row, d) of this thread is the same at every step (lengths only shrink the live prefix): the
  // cell state never leaves the thread's register
  const int ci = G.vcta * GT + tid;                    // i < n * D  <=>  this thread owns element (ci / D, ci % D)
  const int cbl = ci / D, cd = ci - cbl * D;
  float c_reg = (cbl < GR && row0 + cbl < B) ? __ldg(p.C + (int64_t)(row0 + cbl) * D + cd) : 0.f;

  int stamp = 0;
#define RECUR_STAMP()                                                                                  \
  do {                                                                                                 \
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[t * 16 + (stamp & 15)] = clock64();      \
    if (p.prof && t == T / 2 && G.tid == 0)                                                            \
      p.prof[(int64_t)T * 16 + ((int64_t)(G.g * gridDim.x + blockIdx.x)) * 16 + (stamp & 15)] = gtime_ns(); \
    FSTAMP(stamp);                                                                                     \
    ++stamp;                                                                                           \
  } while (0)
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    stamp = 0;
    RECUR_STAMP();
    // live rows of this group (lengths sorted descending): once it has none it never has any again
    const int n = __popc(__ballot_sync(0xffffffffu, lane < GR && row0 + lane < B && lens[min(row0 + lane, B - 1)] > t));
    if (n == 0) break;
    const int64_t tb = (int64_t)t * B;
    float* const sc_t = p.scores + tb * Ppad;
    // scores: this group's share of the (b, px) items of its rows = consecutive att1 rows = ONE bulk copy; att1 does
    // not depend on the step, so the copy is issued NOW and lands while G1 waits for h_{t-1} and multiplies
    const int items = ATT ? n * P : 0;
    const int per = (items + nctas - 1) / nctas;
    const int i0 = G.vcta * per, i1 = min(items, i0 + per);
    const int sci = min(SCI, P > 0 ? P : 1);           // items per fill: never more than one row's worth
    if (ATT && (p.mask & 2) && i0 < i1)
      stage_fill(G, 0, p.att1 + ((int64_t)row0 * P + i0) * A, (uint32_t)min(sci, i1 - i0) * A * 2);
    // ================= G1: [att2 | beta_pre | p] = h_{t-1} W_cat1^T + b, and p*q -> m =================
    if (has1 && (p.mask & 1)) {
      const bf16* hprev = (t == 0 ? p.H0 : p.Ht + (tb - B) * D) + (int64_t)row0 * D;
      float bias1[NT1 / 2], q1[NT1 / 2];
#pragma unroll
      for (int i = 0; i < NT1 / 2; ++i) {              // epilogue operands first: hidden behind the GEMM
        const int nf = f1 + i * 16 + ej;
        bias1[i] = nf < NG1 ? __ldg(p.b_cat1 + nf) : 0.f;
        q1[i] = (!LSTM && lrow < n && nf >= col0 && nf < NG1) ? __ldg(p.q + (int64_t)erow * NQ + (nf - col0)) : 0.f;
      }
      float out[NT1 / 2];
      gemm_job_df<NT1 / 2, 2, 1>(G, hprev, 0, n, W1s, w1s, out);
      if (lrow < n) {
        float* g1 = p.g1 + (tb + erow) * NG1;
#pragma unroll
        for (int i = 0; i < NT1 / 2; ++i) {
          const int nf = f1 + i * 16 + ej;
          if (nf < NG1) {
            const float val = out[i] + bias1[i];
            stx4f(g1 + nf, val);
            if (!LSTM && nf >= col0) {
              const int nn = nf - col0;
              const int gg = nn / F, f = nn - gg * F;
              stx2(p.m + ((int64_t)gg * p.R + tb + erow) * 2 * F + F + f, __float2bfloat16_rn(val * q1[i]));
            }
          }
        }
      }
    }
    if (!ATT) {
      // pure_scn: the input half u*v of the P4 operand (u = Emb W_ia is not recurrent)
      const int total = n * NQ;
#pragma unroll 1
      for (int i = G.vcta * GT + tid; i < total; i += nctas * GT) {
        const int bl = i / NQ, nf = i - bl * NQ;
        const int b = row0 + bl;
        const int gg = nf / F, f = nf - gg * F;
        const float u = __ldg(p.U + (tb + b) * NQ + nf);
        stx2(p.m + ((int64_t)gg * p.R + tb + b) * 2 * F + f, __float2bfloat16_rn(u * __ldg(p.v + (int64_t)b * NQ + nf)));
      }
    }
    RECUR_STAMP();
    if (ATT) {
      // ================= scores e[b, px] = w_f . relu(att1[b, px, :] + att2[b, :]) + b_f =================
      if ((p.mask & 2) && i0 < i1) {
        float wf[2][8];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int a = cc * 256 + a_lane;
          float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
          if (a < A) {
            w0 = __ldg(reinterpret_cast<const float4*>(p.w_f + a));
            w1 = __ldg(reinterpret_cast<const float4*>(p.w_f + a + 4));
          }
          wf[cc][0] = w0.x; wf[cc][1] = w0.y; wf[cc][2] = w0.z; wf[cc][3] = w0.w;
          wf[cc][4] = w1.x; wf[cc][5] = w1.y; wf[cc][6] = w1.z; wf[cc][7] = w1.w;
        }
        const float bfv = __ldg(p.b_f);
#pragma unroll 1
        for (int base = i0; base < i1; base += sci) {
          const int cnt = min(sci, i1 - base);
          if (base != i0) stage_fill(G, 0, p.att1 + ((int64_t)row0 * P + base) * A, (uint32_t)cnt * A * 2);
          // a fill's items lie in at most two consecutive rows (sci <= P): both att2 rows (written by the G1 phase of
          // other CTAs: polled) are requested at once
          const int bl0 = base / P;
          const int px_split = (bl0 + 1) * P - base;     // items at or beyond this index belong to row bl0 + 1
          float x2[2][2][8];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const bool need = r == 0 || px_split < cnt;
            const float* g1 = p.g1 + (tb + row0 + bl0 + r) * NG1;
            uint4 raw[2][2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int a = cc * 256 + a_lane;
              raw[cc][0] = raw[cc][1] = make_uint4(0, 0, 0, 0);
              if (need && a < A) {
                raw[cc][0] = ldx16(g1 + a);
                raw[cc][1] = ldx16(g1 + a + 4);
              }
            }
            for (unsigned spins = 0;;) {                 // late elements: re-requested in rounds (see gemm_job_df)
              bool pend = false;
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) pend = pend || sent32(raw[cc][0]) || sent32(raw[cc][1]);
              if (!pend) break;
              if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const int a = cc * 256 + a_lane;
                if (sent32(raw[cc][0])) raw[cc][0] = ldx16(g1 + a);
                if (sent32(raw[cc][1])) raw[cc][1] = ldx16(g1 + a + 4);
              }
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              x2[r][cc][0] = __uint_as_float(raw[cc][0].x); x2[r][cc][1] = __uint_as_float(raw[cc][0].y);
              x2[r][cc][2] = __uint_as_float(raw[cc][0].z); x2[r][cc][3] = __uint_as_float(raw[cc][0].w);
              x2[r][cc][4] = __uint_as_float(raw[cc][1].x); x2[r][cc][5] = __uint_as_float(raw[cc][1].y);
              x2[r][cc][6] = __uint_as_float(raw[cc][1].z); x2[r][cc][7] = __uint_as_float(raw[cc][1].w);
            }
          }
          pipe_wait(G.pp, 0);
          FSTAMP(120);
#pragma unroll 1
          for (int i = warp; i < cnt; i += GW) {
            const bool second = i >= px_split;
            const int bl = bl0 + (second ? 1 : 0);
            const int px = base + i - bl * P;
            float sacc = 0.f;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int a = cc * 256 + a_lane;
              if (a < A) {
                const uint4 raw = *reinterpret_cast<const uint4*>(stg + (size_t)(i * A + a) * 2);
                float f[8];
                unpack16(raw, f, bf16());
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  sacc = fmaf(wf[cc][k], fmaxf(f[k] + (second ? x2[1][cc][k] : x2[0][cc][k]), 0.f), sacc);
              }
            }
            sacc = warp_sum(sacc);
            if (lane == 0) stx4f(sc_t + (int64_t)(row0 + bl) * Ppad + px, sacc + bfv);
          }
          gsync(G);                                    // every warp is done with the staged att1 rows
        }
      }
      RECUR_STAMP();
      // ================= softmax + weighted sum over a 256-channel chunk + gate -> z =================
      // item = (row, chunk); its pixels stream through the two stages, 32 pixels (16 KB, one bulk copy from the
      // chunk-major feature copy) per fill; the first two fills are in flight while the row's scores are polled
      {
        const int items_w = (p.mask & 4) ? n * chunks : 0;
        const int nfill = (P + WPXS - 1) / WPXS;
#pragma unroll 1
        for (int item = G.vcta; item < items_w; item += nctas) {
          const int rl = item / chunks, chunk = item - rl * chunks;
          const int row = row0 + rl;
          const bf16* src = p.enc_cm + ((int64_t)row * chunks + chunk) * P * CHUNK;
          stage_fill(G, 0, src, (uint32_t)min(WPXS, P) * CHUNK * 2);
          if (nfill > 1) stage_fill(G, 1, src + (int64_t)WPXS * CHUNK, (uint32_t)min(WPXS, P - WPXS) * CHUNK * 2);
          // the gate pre-activation comes from this step's G1 phase: requested now, checked in the epilogue
          const float* g1b = p.g1 + (tb + row) * NG1 + A + chunk * CHUNK + col * 8;
          uint4 braw0 = make_uint4(0, 0, 0, 0), braw1 = braw0;
          if (grp == 0) {
            braw0 = ldx16(g1b);
            braw1 = ldx16(g1b + 4);
          }
          // softmax (every item of the row recomputes it: P <= GT values, one per thread); the scores of the row
          // come from up to ~9 other CTAs: polled
          float sv = -INFINITY;
          if (tid < P) sv = poll4f(G, sc_t + (int64_t)row * Ppad + tid);
          float mx = warp_max(sv);
          if (lane == 0) G.red[warp] = mx;
          gsync(G);
          mx = G.red[0];
#pragma unroll
          for (int w = 1; w < GW; ++w) mx = fmaxf(mx, G.red[w]);
          const float ex = tid < P ? __expf(sv - mx) : 0.f;
          float sum = warp_sum(ex);
          if (lane == 0) G.red[GW + warp] = sum;
          gsync(G);
          sum = 0.f;
#pragma unroll
          for (int w = 0; w < GW; ++w) sum += G.red[GW + w];
          const float alpha = __fdividef(ex, sum);
          if (tid < P) {
            G.al[tid] = alpha;
            if (chunk == 0) p.alphas[((int64_t)row * T + t) * P + tid] = alpha;
          }
          gsync(G);
          FSTAMP(110);
          float acc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 1
          for (int fi = 0; fi < nfill; ++fi) {
            const int s = fi & 1;
            const int px0 = fi * WPXS, cnt = min(WPXS, P - px0);
            pipe_wait(G.pp, s);
            FSTAMP(111);
            const uint8_t* base = stg + s * STAGE + col * 16;
#pragma unroll
            for (int u = 0; u < WPXS / GROUPS; ++u) {
              const int pl = grp + u * GROUPS;
              if (pl < cnt) {
                const uint4 raw = *reinterpret_cast<const uint4*>(base + (size_t)pl * CHUNK * 2);
                const float w = G.al[px0 + pl];
                float f[8];
                unpack16(raw, f, bf16());
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, f[k], acc[k]);
              }
            }
            if (fi + 2 < nfill) {
              gsync(G);
              const int pxn = (fi + 2) * WPXS;
              stage_fill(G, s, src + (int64_t)pxn * CHUNK, (uint32_t)min(WPXS, P - pxn) * CHUNK * 2);
            }
          }
          FSTAMP(112);
          // cross-group reduction through shared memory (red: >= 2 GW + (GROUPS-1) * NCOL * 8 floats)
          if (grp > 0) {
            float4* dst = reinterpret_cast<float4*>(G.red + 2 * GW + ((grp - 1) * NCOL + col) * 8);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
          }
          gsync(G);
          if (grp == 0) {
#pragma unroll 1
            for (int gq = 1; gq < GROUPS; ++gq) {
              const float4* sp = reinterpret_cast<const float4*>(G.red + 2 * GW + ((gq - 1) * NCOL + col) * 8);
              const float4 s0 = sp[0], s1 = sp[1];
              acc[0] += s0.x; acc[1] += s0.y; acc[2] += s0.z; acc[3] += s0.w;
              acc[4] += s1.x; acc[5] += s1.y; acc[6] += s1.z; acc[7] += s1.w;
            }
            const int e0 = chunk * CHUNK + col * 8;
            braw0 = poll16<true>(G, g1b, braw0);
            braw1 = poll16<true>(G, g1b + 4, braw1);
            const float bp[8] = {__uint_as_float(braw0.x), __uint_as_float(braw0.y), __uint_as_float(braw0.z),
                                 __uint_as_float(braw0.w), __uint_as_float(braw1.x), __uint_as_float(braw1.y),
                                 __uint_as_float(braw1.z), __uint_as_float(braw1.w)};
            float zv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) zv[k] = fsigmoid(bp[k]) * acc[k];
            const uint4 zp = pack16(zv, bf16());
            stx16(p.zk + (((int64_t)t * (E / KC) + e0 / KC) * B + row) * KC + (e0 % KC), zp);
            *reinterpret_cast<uint4*>(p.z + (tb + row) * E + e0) = zp;
            if (p.awe) {
              float* dst = p.awe + (tb + row) * E + e0;
              *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
          }
          gsync(G);
        }
      }
      RECUR_STAMP();
      // ================= P3: u = Emb W_ia[:M] + z W_ia[M:], and u*v -> m =================
      if (has3 && (p.mask & 8)) {
        // the epilogue's operands are requested before the GEMM: their L2 round trip hides behind it
        const int nf = f3 + ej;
        const bool ok3 = lrow < n && nf < NQ;
        float* U = p.U + (tb + erow) * NQ;
        const float u_emb = ok3 ? __ldg(U + nf) : 0.f;           // written by the batched GEMM before this kernel
        const float v3 = (!LSTM && ok3) ? __ldg(p.v + (int64_t)erow * NQ + nf) : 0.f;
        float out[1] = {0.f};
        const bf16* zsrc = p.zk + ((int64_t)t * nfill3 * B + row0) * KC;
        if (nfill3 == 4) {
          gemm_job_df<1, 2, 4>(G, zsrc, (int64_t)B * KC, n, W3s, w3s, out);
        } else {
#pragma unroll 1
          for (int kf = 0; kf < nfill3; ++kf) {
            float part[1];
            gemm_job_df<1, 2, 1>(G, zsrc + (int64_t)kf * B * KC, 0, n, W3s + (size_t)kf * KC * 2, w3s, part);
            out[0] += part[0];
          }
        }
        if (ok3) {
          const float val = out[0] + u_emb;
          if (LSTM) {
            stx4f(p.pre + (tb + erow) * NQ + nf, val);           // the cell phase of other CTAs reads it
          } else {
            U[nf] = val;                                         // kept for the backward (read by this thread only)
            const int gg = nf / F, f = nf - gg * F;
            stx2(p.m + ((int64_t)gg * p.R + tb + erow) * 2 * F + f, __float2bfloat16_rn(val * v3));
          }
        }
      }
      RECUR_STAMP();
    }
    if (!LSTM) {
      // ================= P4: pre_g = m_g [W_ic_g | W_hc_g]^T  (the n x 2F operand: both halves polled) =========
      if (has4 && (p.mask & 16)) {
        float out[1];
        gemm_job_df<1, 4, 1>(G, p.m + ((int64_t)gate4 * p.R + tb + row0) * 2 * F, 0, n, W4s, w4s, out);
        const int d = d4 + ej;
        if (lrow < n && d < D) stx4f(p.pre + (tb + erow) * 4 * D + (int64_t)gate4 * D + d, out[0]);
      }
      RECUR_STAMP();
    }
    // ================= LSTM pointwise (scn_cell.py:146-152), gate order i,f,o,c =================
    if ((p.mask & 32) && cbl < n) {
      const int b = row0 + cbl, d = cd;
      float x[4];
      if (LSTM) {
        // pre = (Emb W_ih[:M] + z W_ih[M:]) + h W_hh + b_ih + b_hh, torch gate order i,f,g,o
        const float* ua = (ATT ? p.pre : p.U) + (tb + b) * NQ + d;
        const float* hb = p.g1 + (tb + b) * NG1 + col0 + d;
        float uv[4], hv[4];
        poll4f_n<4>(G, ua, D, 4, uv);
        poll4f_n<4>(G, hb, D, 4, hv);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) x[gq] = uv[gq] + hv[gq] + __ldg(p.b_ih + gq * D + d) + __ldg(p.b_hh + gq * D + d);
        const float t2 = x[2]; x[2] = x[3]; x[3] = t2;        // -> i, f, o, g
      } else {
        float pv[4];
        poll4f_n<4>(G, p.pre + (tb + b) * 4 * D + d, D, 4, pv);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) x[gq] = pv[gq] + __ldg(p.b_ih + gq * D + d) + __ldg(p.b_hh + gq * D + d);
      }
      const float ig = fsigmoid(x[0]), fg = fsigmoid(x[1]), og = fsigmoid(x[2]);
      const float gg = ftanh(x[3]);
      const float c = fg * c_reg + ig * gg;
      const float h = og * ftanh(c);
      c_reg = c;
      const bf16 hb16 = __float2bfloat16_rn(h);
      stx2(p.Ht + (tb + b) * D + d, hb16);                    // the next step's G1 operand first
      p.C[(tb + B + b) * D + d] = c;
      float* gp = p.gates + (tb + b) * 4 * D + d;
      gp[0] = ig; gp[D] = fg; gp[2 * D] = og; gp[3 * D] = gg;
      const int64_t ho = ((int64_t)b * T + t) * D + d;
      p.Hall[ho] = hb16;
      if (drop_p > 0.f)
        p.Hd[ho] = __float2bfloat16_rn(h * dropout_scale(seed, ((uint64_t)b * T + t) * D + d, drop_p));
    }
    RECUR_STAMP();
  }
  FSTAMP_FLUSH();
#undef RECUR_STAMP
}

// =====================================================================================
// Reverse-time recurrence as one cooperative launch: the mirror image of
// recur_fwd_kernel, with the same two independent row groups per CTA.  Per step, six phases
// separated by the group's grid barrier:
//   C   LSTM pointwise backward (dh from the fc gradient + the recurrent gradient) -> dpre_t
//   W   [w_g | r_g] = dpre_g [W_ic_g | W_hc_g]  and the factor products du = w*v, dp = r*q,
//       dv_acc += w*u, dq_acc += r*p fused in the epilogue (full K per CTA: no atomics)
//   Z   dz = du W_ia[M:]^T
//   A   gate backward + partial dalpha_p = enc[p, chunk] . dawe   (streams enc, like the forward sum;
//       item = (row, 256-channel chunk): one warp owns a pixel's whole 512-byte run)
//   B   softmax backward, relu / score backward -> datt2, dw_f, de   (streams att1;
//       item = (row, 64 attention channels))
//   H   dh_{t-1} = [dp | dbeta_pre | datt2] [W_ha | W_beta^T | W_d^T]^T   (K = 4608 split in 512-wide
//       jobs over the CTAs; the partial sums meet in fp32 atomics on a zeroed slot)
// Operands that cross CTAs are written twice: in the layout the batched weight-gradient GEMMs read
// after the loop, and chunk-major ([K/512][B][512]) so that every staging fill is one bulk copy.
// =====================================================================================
struct BwdP {
  int B, T, P, E, A, M, D, F, NQ, NG1;
  int64_t R, ldPX;
  const int32_t* len;
  const bf16* WcT; int64_t ldD;       // [4][2F][ldD]   [W_ic_g^T ; W_hc_g^T]
  const bf16* Wxin; int64_t ldNQ;     // [E][ldNQ]      W_ia[M:, :] (already offset by M rows)
  const bf16* Whx; int64_t ldhx;      // [D][ldhx]      [W_ha | W_beta^T | W_d^T]   (LSTM: W_hh^T [D][4D])
  const bf16* Whx2; int64_t ldhx2;    // LSTM only: [W_beta^T | W_d^T]  [D][E+A]
  bf16* dbx; int64_t ldbx; int dbx_off;   // where [dbeta_pre | datt2] of a row go: dpx + 4F (SCN) or dba (LSTM)
  const float* dHfc;                  // (B, T, D) fp32: d loss / d h_t through the vocabulary projection
  const float* gates; const float* C;
  float* dc;                          // [B][D]
  float* dh_rec;                      // [B][D]: out, dh_0
  float* dhp;                         // [T][KH/512][B][D] K-chunk partials of the recurrent gradient (exchange)
  bf16* dpre;                         // [T][B][4D]
  bf16* dpre_gm;                      // [T][4][B][D]
  const float* U; const float* g1; const float* v; const float* q;
  bf16* du;                           // [T][B][NQ]
  bf16* duk;                          // [T][NQ/512][B][512]
  bf16* dpx;                          // [T][B][ldPX]   dp | dbeta_pre | datt2
  bf16* dpxk;                         // [T][KH/512][B][512]
  float* dv_acc; float* dq_acc;       // [B][NQ]
  float* dz;                          // [T][B][E]
  const float* awe; const float* alphas; const float* d_alphas;
  const bf16* enc_cm; const float* w_f;   // enc_cm [B][E/256][P][256]
  const bf16* att1_cm;                // [B][A/64][P][64]   eighth-major copy of att1
  float* part;                        // [T][B][E/256][pad4(P)] (exchange)
  float* de; float* dwf; float* dbf;  // [T][B][pad4(P)], [T][B][A], [T][B]
  unsigned* bar;
  float dropout_p; const uint64_t* seed;
  long long* prof;                    // debug (CAPDEC_RECUR_PROF=1): [T][16] clock64 stamps of CTA 0, group 0
  int skew;                           // cycles row group 1 starts after group 0 (see group_skew)
};

// LSTM = true (pure_attention): no factor products (phase W is skipped), dz = dpre W_ih[:, M:], the recurrent
// gradient contracts [dpre | dbeta_pre | datt2] with [W_hh | W_beta^T | W_d^T] held in two weight matrices, and
// the attention gradients go to the [dbeta_pre | datt2] buffer of that decoder.
template <bool ATT, bool LSTM = false>
__global__ void __launch_bounds__(RT, 1) recur_bwd_kernel(const __grid_constant__ BwdP p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int D = p.D, E = p.E, F = p.F, B = p.B, T = p.T, P = p.P, NQ = p.NQ, NG1 = p.NG1, A = p.A;
  const int lane = threadIdx.x & 31;
  const int wWs = D * 2 + WPAD, wZs = NQ * 2 + WPAD, wHs = KC * 2 + WPAD;
  const int alw = pad4i(P > 0 ? P : 4);
  // shared memory carve-up: 4 stages | WW | WZ | WH | red[2] | des[2] | lens | mbarriers[4]
  uint8_t* WWs = smem + 4 * STAGE;                        // 32 rows, K = D
  uint8_t* WZs = WWs + (size_t)32 * wWs;                  // 16 rows, K = NQ   (ATT only)
  uint8_t* WHs = WZs + (ATT ? (size_t)16 * wZs : 0);      // 2 x 16 rows, K = 512
  float* red = reinterpret_cast<float*>(WHs + (size_t)32 * wHs);
  float* desb = red + 2 * REDF;                           // [2][pad4(P)]
  int* lens = reinterpret_cast<int*>(desb + 2 * alw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(lens + ((B + 3) & ~3) + 2);
  bars = reinterpret_cast<uint64_t*>(((uintptr_t)bars + 7) & ~(uintptr_t)7);
  Grp G;
  grp_init(G, smem, red, desb, alw, bars, p.bar);
  uint8_t* const stg = G.pp.stg;
  float* const des = G.al;
  const int tid = G.tid, warp = G.warp, row0 = G.row0;
  // ---- roles (both groups: same weights, different rows) ----
  const int c = blockIdx.x;
  const int nctas = gridDim.x;
  const int spg = 2 * F / 32;                             // 32-feature slices per gate of the W job
  const bool hasW = !LSTM && c < 4 * spg;
  const int gateW = c / spg, fW0 = (c - gateW * spg) * 32;
  const bool hasZ = ATT && c * 16 < E;
  const int eZ0 = c * 16;
  const int col0 = ATT ? A + E : 0;
  const int KH = NQ + (ATT ? E + A : 0);
  const int nkc = KH / KC, nkq = NQ / KC;
  const int jobsH = (D / 16) * nkc;
  if (hasW) load_weight_rows(WWs, p.WcT + ((int64_t)gateW * 2 * F + fW0) * p.ldD, p.ldD, D, 32, 2 * F - fW0);
  if (hasZ) load_weight_rows(WZs, p.Wxin + (int64_t)eZ0 * p.ldNQ, p.ldNQ, NQ, 16, E - eZ0);
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int j = c + i * nctas;
    if (j < jobsH) {
      const int ds = j / nkc, kc = j - ds * nkc;
      if (LSTM && kc >= nkq)
        load_weight_rows(WHs + (size_t)i * 16 * wHs, p.Whx2 + (int64_t)ds * 16 * p.ldhx2 + (int64_t)(kc - nkq) * KC,
                         p.ldhx2, KC, 16, D - ds * 16);
      else
        load_weight_rows(WHs + (size_t)i * 16 * wHs, p.Whx + (int64_t)ds * 16 * p.ldhx + (int64_t)kc * KC, p.ldhx, KC, 16,
                         D - ds * 16);
    }
  }
  for (int i = threadIdx.x; i < B; i += RT) lens[i] = p.len[i];
  __syncthreads();
  if (row0 >= B) return;                                  // this group has no rows at all (B <= 16)
  group_skew(G, p.skew);

  const int lrow = tid >> 4, ej = tid & 15;
  const int erow = row0 + lrow;
  const int Ppad = pad4i(P);
  const int chunks = ATT ? E / CHUNK : 1;
  const int grp = warp, col = lane;                       // phase A mapping: warp = pixel group, lane = 16-byte column
  const float drop_p = p.dropout_p;
  const uint64_t seed = drop_p > 0.f ? __ldg(p.seed) : 0ull;
  const int nfillP = (P + WPXS - 1) / WPXS;
  bf16* const dwv = reinterpret_cast<bf16*>(G.red);        // phase A: dawe of the item, bf16 [CHUNK]
  float* const psum = G.red + CHUNK / 2;                   // phase A: [4 K quarters][pad4(P)] partial dalpha
  // cell element (row, d) of this thread: the same at every step, so dc never leaves its register
  const int ci = G.vcta * GT + tid;
  const int cbl = ci / D, cd = ci - cbl * D;
  const bool cell_mine = cbl < GR && row0 + cbl < B;
  float dc_reg = 0.f;
  int stamp = 0;
#define BSTAMP() do { if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[t * 16 + (stamp & 15)] = clock64(); \
    if (p.prof && t == T / 2 && G.tid == 0) p.prof[(int64_t)T * 16 + ((int64_t)(G.g * gridDim.x + blockIdx.x)) * 16 + (stamp & 15)] = gtime_ns(); \
    FSTAMP(stamp); ++stamp; } while (0)

#pragma unroll 1
  for (int t = T - 1; t >= 0; --t) {
    stamp = 0;
    BSTAMP();
    // live rows of this group; none yet (all its captions are shorter than t) -> nothing to do at this step
    const int n = __popc(__ballot_sync(0xffffffffu, lane < GR && row0 + lane < B && lens[min(row0 + lane, B - 1)] > t));
    if (n == 0) continue;
    const int64_t tb = (int64_t)t * B;
    // ================= C: LSTM pointwise backward =================
    // dh_t = fc gradient + the recurrent gradient of step t+1 = sum of the nkc K-chunk partials its H phase left in
    // dhp[t+1] (polled; a row that was not live at t+1 gets none)
    if (cell_mine && cbl < n) {
      const int b = row0 + cbl, d = cd;
      float dh = 0.f;
      if (t + 1 < T && lens[b] > t + 1) {
        const float* dp = p.dhp + (((int64_t)(t + 1) * nkc) * B + b) * D + d;
        float pv[16];
        poll4f_n<16>(G, dp, (int64_t)B * D, nkc, pv);
#pragma unroll
        for (int k = 0; k < 16; ++k) dh += pv[k];
      }
      {
        float g = __ldg(p.dHfc + ((int64_t)b * T + t) * D + d);
        if (drop_p > 0.f) g *= dropout_scale(seed, ((uint64_t)b * T + t) * D + d, drop_p);
        dh += g;
      }
      const int64_t i = (int64_t)cbl * D + d;
      const float* gp = p.gates + (tb + b) * 4 * D + d;
      const float ig = __ldg(gp), fg = __ldg(gp + D), og = __ldg(gp + 2 * D), gg = __ldg(gp + 3 * D);
      const float tc = ftanh(__ldg(p.C + (tb + B + row0) * D + i));
      const float dcn = dc_reg + dh * og * (1.f - tc * tc);
      const float dpo = dh * tc * og * (1.f - og), dpg = dcn * ig * (1.f - gg * gg);
      // pre-activation slots: i, f, o, c (SCN cell) or i, f, g, o (nn.LSTMCell)
      const float dpv[4] = {dcn * gg * ig * (1.f - ig), dcn * __ldg(p.C + (tb + row0) * D + i) * fg * (1.f - fg),
                            LSTM ? dpg : dpo, LSTM ? dpo : dpg};
      dc_reg = dcn * fg;
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        const bf16 x = __float2bfloat16_rn(dpv[gq]);
        stx2(p.dpre_gm + (((int64_t)t * 4 + gq) * B + b) * D + d, x);      // the W / Z / H operand of other CTAs
        p.dpre[(tb + b) * 4 * D + gq * D + d] = x;
      }
    }
    BSTAMP();
    // ================= W: [w | r] = dpre_g [W_ic_g | W_hc_g] and the factor products =================
    if (hasW) {
      // epilogue operands (factor, forward activation, running sum) are requested before the GEMM
      float fac[2], act[2], run[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int nf = fW0 + i * 16 + ej;
        fac[i] = act[i] = run[i] = 0.f;
        if (lrow < n && nf < 2 * F) {
          const bool is_w = nf < F;
          const int n4 = gateW * F + (is_w ? nf : nf - F);
          const int64_t k = (int64_t)erow * NQ + n4;
          fac[i] = __ldg((is_w ? p.v : p.q) + k);
          act[i] = is_w ? __ldg(p.U + (tb + erow) * NQ + n4) : __ldg(p.g1 + (tb + erow) * NG1 + col0 + n4);
          run[i] = (is_w ? p.dv_acc : p.dq_acc)[k];              // only this thread ever touches element k
        }
      }
      float out[2];
      gemm_job_df<2, 2, 1>(G, p.dpre_gm + (((int64_t)t * 4 + gateW) * B + row0) * D, 0, n, WWs, wWs, out);
      if (lrow < n) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int nf = fW0 + i * 16 + ej;
          if (nf < 2 * F) {
            const bool is_w = nf < F;
            const int n4 = gateW * F + (is_w ? nf : nf - F);
            const int64_t k = (int64_t)erow * NQ + n4;
            const float val = out[i];
            const bf16 x = __float2bfloat16_rn(val * fac[i]);
            if (is_w) {
              stx2(p.duk + (((int64_t)t * nkq + n4 / KC) * B + erow) * KC + (n4 % KC), x);   // the Z operand
              p.du[(tb + erow) * NQ + n4] = x;
              p.dv_acc[k] = run[i] + val * act[i];
            } else {
              stx2(p.dpxk + (((int64_t)t * nkc + n4 / KC) * B + erow) * KC + (n4 % KC), x);  // an H operand
              p.dpx[(tb + erow) * p.ldPX + n4] = x;
              p.dq_acc[k] = run[i] + val * act[i];
            }
          }
        }
      }
    }
    BSTAMP();
    if (ATT) {
      // ================= Z: dz = du W_ia[M:]^T   (LSTM: dpre W_ih[:, M:]) =================
      if (hasZ) {
        float out[1] = {0.f};
        const bf16* zsrc = (LSTM ? p.dpre_gm + (int64_t)t * 4 * B * D : p.duk + (int64_t)t * nkq * B * KC) + (int64_t)row0 * KC;
        if (nkq == 4) {
          gemm_job_df<1, 2, 4>(G, zsrc, (int64_t)B * KC, n, WZs, wZs, out);
        } else {
#pragma unroll 1
          for (int kf = 0; kf < nkq; ++kf) {
            float part1[1];
            gemm_job_df<1, 2, 1>(G, zsrc + (int64_t)kf * B * KC, 0, n, WZs + (size_t)kf * KC * 2, wZs, part1);
            out[0] += part1[0];
          }
        }
        const int e = eZ0 + ej;
        if (lrow < n && e < E) stx4f(p.dz + (tb + erow) * E + e, out[0]);     // phase A of other CTAs polls it
      }
      BSTAMP();
      // ================= A: gate backward + partial dalpha over one 256-channel chunk =================
      {
        const int items = n * chunks;
#pragma unroll 1
        for (int item = G.vcta; item < items; item += nctas) {
          const int rl = item / chunks, chunk = item - rl * chunks;
          const int row = row0 + rl;
          const bf16* src = p.enc_cm + ((int64_t)row * chunks + chunk) * P * CHUNK;
          stage_fill(G, 0, src, (uint32_t)min(WPXS, P) * CHUNK * 2);
          if (nfillP > 1) stage_fill(G, 1, src + (int64_t)WPXS * CHUNK, (uint32_t)min(WPXS, P - WPXS) * CHUNK * 2);
          // the saved forward activations (gate pre-activation, awe) do not depend on this step: requested before
          // dz is polled
          const int e0 = chunk * CHUNK + col * 8;
          float4 b0, b1, a0, a1;
          {
            const float* bpp = p.g1 + (tb + row) * NG1 + A + e0;
            const float* awp = p.awe + (tb + row) * E + e0;
            b0 = __ldg(reinterpret_cast<const float4*>(bpp)); b1 = __ldg(reinterpret_cast<const float4*>(bpp + 4));
            a0 = __ldg(reinterpret_cast<const float4*>(awp)); a1 = __ldg(reinterpret_cast<const float4*>(awp + 4));
          }
          float dawe[8];
          {
            const float* dzp = p.dz + (tb + row) * E + e0;
            uint4 z0 = ldx16(dzp), z1 = ldx16(dzp + 4);
            for (unsigned spins = 0; sent32(z0) || sent32(z1);) {          // late elements: re-requested together
              if ((++spins & 63u) == 0u && poll_stalled(G.abortp, G.t_end)) break;
              if (sent32(z0)) z0 = ldx16(dzp);
              if (sent32(z1)) z1 = ldx16(dzp + 4);
            }
            const float dzv[8] = {__uint_as_float(z0.x), __uint_as_float(z0.y), __uint_as_float(z0.z), __uint_as_float(z0.w),
                                  __uint_as_float(z1.x), __uint_as_float(z1.y), __uint_as_float(z1.z), __uint_as_float(z1.w)};
            const float bp[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const float aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float db[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float gate = fsigmoid(bp[k]);
              dawe[k] = dzv[k] * gate;
              db[k] = dzv[k] * aw[k] * gate * (1.0f - gate);
            }
            if (grp == 0) {
              const uint4 pk = pack16(db, bf16());
              stx16(p.dpxk + (((int64_t)t * nkc + nkq + e0 / KC) * B + row) * KC + (e0 % KC), pk);   // an H operand
              *reinterpret_cast<uint4*>(p.dbx + (tb + row) * p.ldbx + p.dbx_off + e0) = pk;
              // dawe as the (single useful) column of the mma B operand
              *reinterpret_cast<uint4*>(dwv + col * 8) = pack16(dawe, bf16());
            }
          }
          gsync(G);
          FSTAMP(130);
          // dalpha[px] = enc[px, chunk] . dawe on the tensor pipe: per fill, warp (pt, kq) multiplies the 16-pixel
          // tile pt with 64 channels (two 32-wide k blocks, same 16-byte-per-lane k permutation as gemm_job); only
          // column 0 of the n8 tile carries dawe.  ~10 instructions per warp and fill instead of ~120 scalar ones
          // (the scalar version was issue bound: 1 400 cycles per fill).
          const int pt = warp & 1, kq = warp >> 1;
          const int mg = lane >> 2, mc = lane & 3;
          uint4 bfrag[2];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            bfrag[j] = mg == 0 ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(dwv) + kq * 128 + j * 64 + 16 * mc)
                               : make_uint4(0, 0, 0, 0);
#pragma unroll 1
          for (int fi = 0; fi < nfillP; ++fi) {
            const int s = fi & 1;
            const int px0 = fi * WPXS, cnt = min(WPXS, P - px0);
            pipe_wait(G.pp, s);
            FSTAMP(131);
            if (pt * 16 < cnt) {                         // warp-uniform
              const uint8_t* ap = stg + s * STAGE + (size_t)(pt * 16 + mg) * (CHUNK * 2) + kq * 128 + 16 * mc;
              float acc[2][4];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
                const uint4 alo = *reinterpret_cast<const uint4*>(ap + j * 64);
                const uint4 ahi = *reinterpret_cast<const uint4*>(ap + 8 * (CHUNK * 2) + j * 64);
                mma_bf16(acc[j], alo.x, ahi.x, alo.y, ahi.y, bfrag[j].x, bfrag[j].y);
                mma_bf16(acc[j], alo.z, ahi.z, alo.w, ahi.w, bfrag[j].z, bfrag[j].w);
              }
              if (mc == 0) {                             // column 0 of the n8 tile
                const int pa = px0 + pt * 16 + mg;
                if (pa < P) psum[kq * Ppad + pa] = acc[0][0] + acc[1][0];
                if (pa + 8 < P) psum[kq * Ppad + pa + 8] = acc[0][2] + acc[1][2];
              }
            }
            if (fi + 2 < nfillP) {
              gsync(G);
              const int pxn = (fi + 2) * WPXS;
              stage_fill(G, s, src + (int64_t)pxn * CHUNK, (uint32_t)min(WPXS, P - pxn) * CHUNK * 2);
            }
          }
          gsync(G);
          if (tid < P)
            stx4f(p.part + (((int64_t)t * B + row) * chunks + chunk) * Ppad + tid,
                  (psum[tid] + psum[Ppad + tid]) + (psum[2 * Ppad + tid] + psum[3 * Ppad + tid]));   // phase B polls it
          gsync(G);
        }
      }
      BSTAMP();
      // ================= B: softmax backward + relu / score backward =================
      // item = (row, 64 attention channels): the eighth-major copy att1_cm makes the item's P x 64 slab
      // contiguous (two bulk copies), datt2 / dw_f of different items are disjoint, and every item redoes the
      // row's tiny softmax backward
      {
        const int eighths = A / QW;
        const int items = n * eighths;
        const bool live = G.vcta < items;
        const int rl = G.vcta / eighths, qa = G.vcta - rl * eighths;
        const int row = row0 + rl;
        const bf16* src = p.att1_cm + ((int64_t)row * eighths + qa) * P * QW;
        const int nfillB = (P + BPX - 1) / BPX;
        if (live) {
          stage_fill(G, 0, src, (uint32_t)min(BPX, P) * QW * 2);
          if (nfillB > 1) stage_fill(G, 1, src + (int64_t)BPX * QW, (uint32_t)min(BPX, P - BPX) * QW * 2);
        }
        // saved forward activations / loss gradients do not depend on this step's barrier: requested before it is crossed
        const int hl = lane & 15, hp = lane >> 4;      // half-warp per pixel, lane holds 4 attention features
        const int a0 = qa * QW + hl * 4;
        float d = 0.f, alp = 0.f;
        float4 x2 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = x2;
        if (live) {
          if (tid < P) {
            if (p.d_alphas) d = __ldg(p.d_alphas + ((int64_t)row * T + t) * P + tid);
            alp = __ldg(p.alphas + ((int64_t)row * T + t) * P + tid);
          }
          x2 = __ldg(reinterpret_cast<const float4*>(p.g1 + (tb + row) * NG1 + a0));
          w4 = __ldg(reinterpret_cast<const float4*>(p.w_f + a0));
        }
        if (live) {
          // dalpha = sum of the channel-chunk partials of phase A (other CTAs: polled) (+ external);
          // de = alpha (dalpha - alpha . dalpha)
          if (tid < P) {
            // independent loads: all in flight together (a running sum would serialise 8 L2 round trips)
            const float* pp0 = p.part + ((int64_t)t * B + row) * chunks * Ppad + tid;
            int cc = 0;
            for (; cc + 8 <= chunks; cc += 8) {
              float v[8];
              poll4f_n<8>(G, pp0 + (int64_t)cc * Ppad, Ppad, 8, v);
              d += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
            for (; cc < chunks; ++cc) d += poll4f(G, pp0 + (int64_t)cc * Ppad);
          }
          float dot = warp_sum(alp * d);
          if (lane == 0) G.red[warp] = dot;
          gsync(G);
          dot = 0.f;
#pragma unroll
          for (int w = 0; w < GW; ++w) dot += G.red[w];
          const float x = alp * (d - dot);
          if (tid < P) {
            des[tid] = x;
            if (qa == 0) p.de[(tb + row) * Ppad + tid] = x;
          }
          float sde = warp_sum(tid < P ? x : 0.f);
          if (lane == 0) G.red[GW + warp] = sde;
          gsync(G);
          if (tid == 0 && qa == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < GW; ++w) tot += G.red[GW + w];
            p.dbf[tb + row] = tot;
          }
          FSTAMP(140);
          float att2[4], wf[4], dacc[4], wacc[4];
          {
            att2[0] = x2.x; att2[1] = x2.y; att2[2] = x2.z; att2[3] = x2.w;
            wf[0] = w4.x; wf[1] = w4.y; wf[2] = w4.z; wf[3] = w4.w;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) { dacc[k] = 0.f; wacc[k] = 0.f; }
#pragma unroll 1
          for (int fi = 0; fi < nfillB; ++fi) {
            const int s = fi & 1;
            const int px0 = fi * BPX, cnt = min(BPX, P - px0);
            pipe_wait(G.pp, s);
            FSTAMP(141);
            const uint8_t* base = stg + s * STAGE + hl * 8;
#pragma unroll 4
            for (int pl = 2 * warp + hp; pl < cnt; pl += 2 * GW) {
              const float dep = des[px0 + pl];
              const uint2 raw = *reinterpret_cast<const uint2*>(base + (size_t)pl * QW * 2);
              const float f[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                                  __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u)};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float pre = f[k] + att2[k];
                const float on = pre > 0.f ? 1.f : 0.f;
                dacc[k] = fmaf(dep * wf[k], on, dacc[k]);
                wacc[k] = fmaf(dep, pre * on, wacc[k]);
              }
            }
            if (fi + 2 < nfillB) {
              gsync(G);
              const int pxn = (fi + 2) * BPX;
              stage_fill(G, s, src + (int64_t)pxn * QW, (uint32_t)min(BPX, P - pxn) * QW * 2);
            }
          }
          FSTAMP(142);
          gsync(G);                                              // the staging buffers become the reduction scratch
          float* slab = reinterpret_cast<float*>(stg);           // [2 GW half-warps][2][QW]
          const int hw = 2 * warp + hp;
          *reinterpret_cast<float4*>(slab + (hw * 2 + 0) * QW + hl * 4) = make_float4(dacc[0], dacc[1], dacc[2], dacc[3]);
          *reinterpret_cast<float4*>(slab + (hw * 2 + 1) * QW + hl * 4) = make_float4(wacc[0], wacc[1], wacc[2], wacc[3]);
          gsync(G);
          if (tid < 2 * QW) {
            const int which = tid / QW, aa = tid - which * QW;
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < 2 * GW; ++w) sum += slab[(w * 2 + which) * QW + aa];
            const int a = qa * QW + aa;
            if (which == 0) {
              const bf16 xb = __float2bfloat16_rn(sum);
              stx2(p.dpxk + (((int64_t)t * nkc + nkq + E / KC) * B + row) * KC + a, xb);      // an H operand
              p.dbx[(tb + row) * p.ldbx + p.dbx_off + E + a] = xb;
            } else {
              p.dwf[(tb + row) * A + a] = sum;
            }
          }
          gsync(G);
          if (tid == 0) fence_proxy_async();                     // the slab (generic writes) is overwritten by bulk copies next
        }
      }
    }
    BSTAMP();
    // ================= H: dh_{t-1} partial (chunk kc) = [dp | dbeta_pre | datt2] chunk . W_hx chunk^T =================
    // K = KH is cut in 512-wide jobs (at most two per CTA); job (d-slice, kc) leaves its partial sum in dhp[t][kc],
    // the C phase of step t-1 adds the nkc partials (no atomics, no zeroed accumulator)
    {
      const bf16* srcH[2] = {nullptr, nullptr};
      int dsH[2] = {0, 0}, kcH[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int j = c + i * nctas;
        if (j < jobsH) {
          dsH[i] = j / nkc;
          kcH[i] = j - dsH[i] * nkc;
          srcH[i] = ((LSTM && kcH[i] < nkq) ? p.dpre_gm + ((int64_t)t * 4 + kcH[i]) * B * D
                                            : p.dpxk + ((int64_t)t * nkc + kcH[i]) * B * KC) + (int64_t)row0 * KC;
        }
      }
      float outH[2] = {0.f, 0.f};
      if (srcH[0] && srcH[1]) {                  // both jobs: operand loads in flight together (CTA-uniform)
        gemm_job_df_pair(G, srcH[0], srcH[1], n, WHs, wHs, outH);
      } else if (srcH[0]) {
        float o1[1];
        gemm_job_df<1, 2, 1>(G, srcH[0], 0, n, WHs, wHs, o1);
        outH[0] = o1[0];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int d = dsH[i] * 16 + ej;
        if (srcH[i] && lrow < n && d < D) stx4f(p.dhp + (((int64_t)t * nkc + kcH[i]) * B + erow) * D + d, outH[i]);
      }
    }
    BSTAMP();
  }
  // dh_0 (gradient of init_h's output) = sum of the partials of step 0; dc_0 is what is left in the register
  if (cell_mine) {
    const int b = row0 + cbl, d = cd;
    float dh = 0.f;
    const float* dp = p.dhp + (int64_t)b * D + d;
    float pv[16];
    poll4f_n<16>(G, dp, (int64_t)B * D, nkc, pv);
#pragma unroll
    for (int k = 0; k < 16; ++k) dh += pv[k];
    p.dh_rec[(int64_t)b * D + d] = dh;
    p.dc[(int64_t)b * D + d] = dc_reg;
  }
  FSTAMP_FLUSH();
#undef BSTAMP
}

// dst[b][c][p][j] = src[b][p][c*cw + j]: every (pixel, cw-channel chunk) run becomes contiguous with its
// neighbours in p, so that a staging fill of the persistent kernels is ONE bulk copy
__global__ void chunk_major_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int P, int E, int cw) {
  const int vpr = E / 8, vpc = cw / 8, chunks = E / cw;
  const int64_t total = (int64_t)B * P * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bp = i / vpr;
    const int e8 = (int)(i - bp * vpr);
    const int b = (int)(bp / P), px = (int)(bp - (int64_t)b * P);
    const int c = e8 / vpc, j = e8 - c * vpc;
    dst[(((int64_t)b * chunks + c) * P + px) * vpc + j] = src[i];
  }
}

#ifdef CAPDEC_RECUR_FINE
void fine_reset() {
  unsigned z = 0;
  cudaMemcpyToSymbol(d_fine_idx, &z, sizeof z);
}
void fine_dump(const char* name) {
  unsigned n = 0;
  cudaMemcpyFromSymbol(&n, d_fine_idx, sizeof n);
  if (n > FINE_N) n = FINE_N;
  std::vector<unsigned long long> h(n);
  if (n) cudaMemcpyFromSymbol(h.data(), d_fine, n * sizeof(unsigned long long));
  // print the stamps of one step in the middle: a step starts at tag 0
  std::vector<unsigned> starts;
  for (unsigned i = 0; i < n; ++i)
    if ((h[i] >> 48) == 0) starts.push_back(i);
  if (starts.size() < 4) return;
  for (size_t which : {starts.size() / 2, starts.size() / 2 + 1}) {
    const unsigned b = starts[which], e = which + 1 < starts.size() ? starts[which + 1] : n;
    fprintf(stderr, "%s fine step #%zu:", name, which);
    for (unsigned i = b + 1; i <= e && i < n; ++i)
      fprintf(stderr, " %u:%lld", (unsigned)(h[i] >> 48), (long long)((h[i] & 0xffffffffffffull) - (h[i - 1] & 0xffffffffffffull)));
    fprintf(stderr, "\n");
  }
}
#else
void fine_reset() {}
void fine_dump(const char*) {}
#endif

// Rows (t, b) beyond a caption's length are never written inside the time loop, so in the exchange buffers they
// keep the fill pattern (a NaN).  Two of those buffers are read over ALL rows afterwards -- m by the batched
// weight-gradient GEMMs, g1 by the batched dAtt1 kernel -- and must hold zeros there (0 * NaN = NaN).
__global__ void dead_rows_zero_kernel(const int32_t* __restrict__ len, int B, bf16* m, int64_t R, int twoF, float* g1,
                                      int NG1) {
  const int t = blockIdx.x / B, b = blockIdx.x - t * B;
  if (t < __ldg(len + b)) return;
  const int64_t r = (int64_t)t * B + b;
  if (m) {
    const int v = twoF / 8;
    for (int i = threadIdx.x; i < 4 * v; i += blockDim.x) {
      const int g = i / v, j = i - g * v;
      reinterpret_cast<uint4*>(m + ((int64_t)g * R + r) * twoF)[j] = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < NG1 / 4; i += blockDim.x) reinterpret_cast<float4*>(g1 + r * NG1)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// debug: the phase boundaries of the middle step on EVERY CTA and both row groups (global timer, ns): per boundary the
// earliest / median / latest arrival relative to the earliest step start of that row group
void dump_all_cta_stamps(const char* name, const long long* h, int nctas) {
  for (int g = 0; g < 2; ++g) {
    long long t0 = 0;
    for (int c = 0; c < nctas; ++c) {
      const long long v = h[((int64_t)g * nctas + c) * 16];
      if (v && (!t0 || v < t0)) t0 = v;
    }
    if (!t0) continue;
    fprintf(stderr, "%s all-CTA stamps, row group %d (ns after the group's earliest step start; min / median / max):\n", name, g);
    for (int k = 0; k < 16; ++k) {
      std::vector<long long> v;
      for (int c = 0; c < nctas; ++c) {
        const long long x = h[((int64_t)g * nctas + c) * 16 + k];
        if (x) v.push_back(x - t0);
      }
      if (v.empty()) break;
      std::sort(v.begin(), v.end());
      fprintf(stderr, "  boundary %d: %lld / %lld / %lld  (%zu CTAs)\n", k, v.front(), v[v.size() / 2], v.back(), v.size());
    }
  }
}

struct DevInfo { int sms = 0; int smem_optin = 0; bool coop = false; };
DevInfo g_dev[64];
std::once_flag g_dev_once[64];

const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::call_once(g_dev_once[dev], [dev] {
    DevInfo& d = g_dev[dev];
    int v = 0;
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
    d.coop = v != 0;
  });
  return &g_dev[dev];
}

size_t fwd_smem_bytes(const RecurFwdArgs& a, int nt1, int nt3, int nt4) {
  size_t s = 4 * (size_t)STAGE;
  s += (size_t)nt1 * 8 * (a.D * 2 + WPAD);
  if (a.att) s += (size_t)nt3 * 8 * (a.E * 2 + WPAD);
  s += (size_t)nt4 * 8 * (2 * a.F * 2 + WPAD);
  s += (size_t)2 * REDF * 4;
  s += (size_t)2 * pad4i(a.P > 0 ? a.P : 4) * 4;
  s += (size_t)(((a.B + 3) & ~3) + 2) * 4;
  return s + 8 + 4 * 8 + 128;
}

int pick_nt(int tiles, int ctas) {
  if (tiles <= 2 * ctas) return 2;
  if (tiles <= 4 * ctas) return 4;
  return 0;
}

bool g_timing = false;
cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};     // [0,1] forward kernel, [2,3] backward kernel
bool g_timed[2] = {false, false};

bool persistent_enabled() {
  const char* s = getenv("CAPDEC_PERSISTENT");      // read per call: tests flip it
  return !(s && s[0] == '0');
}

// decide the tiling; returns false when the shape is outside what the persistent kernel covers
bool plan_fwd(const RecurFwdArgs& a, const DevInfo* di, int* nt1, int* nt3, int* nt4, size_t* smem) {
  if (!di || !di->coop || di->sms < 8) return false;
  if (a.B < 1 || a.B > 32 || a.T < 1) return false;
  if (a.D != KC || (!a.lstm && 2 * a.F != 2 * KC)) return false;   // G1 operand = one 512-wide fill, P4 operand = one 1024-wide fill
  if (a.lstm && !a.att) return false;
  if (a.att && (a.E % KC || a.E % CHUNK || a.A % 8 || a.A > 512 || a.P < 1 || a.P > GT)) return false;
  const int NQ = a.lstm ? 4 * a.D : 4 * a.F, NG1 = (a.att ? a.A + a.E : 0) + NQ;
  *nt1 = pick_nt(NG1 / 8, di->sms);
  *nt3 = a.att ? pick_nt(NQ / 8, di->sms) : 2;
  *nt4 = pick_nt(4 * (a.D / 8), di->sms);
  if (!*nt1 || *nt3 != 2 || *nt4 != 2) return false;
  if ((a.D / 8) % *nt4) return false;               // a CTA's P4 tiles stay inside one gate
  *smem = fwd_smem_bytes(a, *nt1, *nt3, a.lstm ? 0 : *nt4);
  return *smem <= (size_t)di->smem_optin;
}

}  // namespace

int chunk_major_copy(const void* src, void* dst, int B, int P, int E, int cw, cudaStream_t st) {
  CAPDEC_REQUIRE(E % cw == 0 && cw % 8 == 0, CAPDEC_ERR_BAD_SHAPE, "chunk_major_copy: E=%d cw=%d", E, cw);
  const DevInfo* di = dev_info();
  chunk_major_kernel<<<(di && di->sms > 0 ? di->sms : 148) * 8, 256, 0, st>>>((const uint4*)src, (uint4*)dst, B, P, E, cw);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

bool recur_fwd_supported(const RecurFwdArgs& a) {
  if (!persistent_enabled()) return false;
  int n1, n3, n4;
  size_t smem;
  return plan_fwd(a, dev_info(), &n1, &n3, &n4, &smem);
}

int recur_fwd(const RecurFwdArgs& a, cudaStream_t st) {
  const DevInfo* di = dev_info();
  int nt1, nt3, nt4;
  size_t smem;
  CAPDEC_REQUIRE(plan_fwd(a, di, &nt1, &nt3, &nt4, &smem), CAPDEC_ERR_BAD_SHAPE,
                 "recur_fwd: shape not covered by the persistent kernel");
  FwdP p;
  memset(&p, 0, sizeof p);
  p.B = a.B; p.T = a.T; p.P = a.P; p.E = a.E; p.A = a.A; p.M = a.M; p.D = a.D; p.F = a.F;
  p.NQ = a.lstm ? 4 * a.D : 4 * a.F; p.NG1 = (a.att ? a.A + a.E : 0) + p.NQ; p.R = (int64_t)a.B * a.T;
  p.len = a.len; p.Wcat1 = (const bf16*)a.Wcat1; p.ldD = a.ldD; p.Wxz = (const bf16*)a.Wxz; p.ldX = a.ldX;
  p.Wc = (const bf16*)a.Wc; p.ld2F = a.ld2F; p.b_cat1 = a.b_cat1; p.b_ih = a.b_ih; p.b_hh = a.b_hh;
  p.att1 = (const bf16*)a.att1; p.enc_cm = (const bf16*)a.enc_cm; p.w_f = a.w_f; p.b_f = a.b_f; p.v = a.v; p.q = a.q;
  p.H0 = (const bf16*)a.H0; p.Ht = (bf16*)a.Ht; p.zk = (bf16*)a.zk; p.Hall = (bf16*)a.Hall; p.Hd = (bf16*)a.Hd; p.C = a.C; p.U = a.U;
  p.g1 = a.g1; p.alphas = a.alphas; p.awe = a.awe; p.z = (bf16*)a.z; p.m = (bf16*)a.m; p.pre = a.pre;
  p.gates = a.gates; p.scores = a.scores; p.bar = a.bar; p.dropout_p = a.dropout_p; p.seed = a.seed;

  auto kernel = a.lstm ? (nt1 == 4 ? recur_fwd_kernel<true, 4, true> : recur_fwd_kernel<true, 2, true>)
                : a.att ? (nt1 == 4 ? recur_fwd_kernel<true, 4> : recur_fwd_kernel<true, 2>)
                        : (nt1 == 4 ? recur_fwd_kernel<false, 4> : recur_fwd_kernel<false, 2>);
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CAPDEC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RT, smem));
  CAPDEC_REQUIRE(per_sm >= 1, CAPDEC_ERR_CUDA, "recur_fwd: kernel does not fit one CTA per SM (smem %zu)", smem);
  CAPDEC_REQUIRE(a.ldH0 == a.D, CAPDEC_ERR_BAD_SHAPE, "recur_fwd: H0 must be dense");
  CAPDEC_REQUIRE((int64_t)GR * a.D <= (int64_t)di->sms * GT, CAPDEC_ERR_BAD_SHAPE, "recur_fwd: one cell element per thread");
  CAPDEC_REQUIRE(p.NG1 % 4 == 0 && (2 * a.F) % 8 == 0, CAPDEC_ERR_BAD_SHAPE, "recur_fwd: row widths");
  CAPDEC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, st));
  // exchange buffers of the time loop: the all-ones fill pattern their consumers poll on (see "Dataflow exchange")
  {
    const size_t R = (size_t)p.R;
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.Ht, 0xFF, R * a.D * 2, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.g1, 0xFF, R * p.NG1 * 4, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.pre, 0xFF, R * (size_t)4 * a.D * 4, st));     // [T][B][4D] (LSTM: NQ = 4D)
    if (!a.lstm) CAPDEC_CUDA_OK(cudaMemsetAsync(a.m, 0xFF, (size_t)4 * R * 2 * a.F * 2, st));
    if (a.att) {
      CAPDEC_CUDA_OK(cudaMemsetAsync(a.zk, 0xFF, R * a.E * 2, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(a.scores, 0xFF, (size_t)a.T * a.B * pad4i(a.P) * 4, st));
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(di->sms, 1, 1);
  cfg.blockDim = dim3(RT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.mask = 63;
  if (const char* mk = getenv("CAPDEC_RECUR_MASK")) p.mask = atoi(mk);
  p.skew = a.att ? RECUR_SKEW_FWD : 0;
  if (const char* sk = getenv("CAPDEC_RECUR_SKEW")) p.skew = atoi(sk);
  if (const char* sk = getenv("CAPDEC_RECUR_SKEW_FWD")) p.skew = atoi(sk);

  const char* prof_env = getenv("CAPDEC_RECUR_PROF");
  const bool prof = prof_env && prof_env[0] == '1';
  if (prof) {
    fine_reset();
    CAPDEC_CUDA_OK(cudaMalloc(&p.prof, (size_t)(a.T * 16 + 2 * di->sms * 16) * sizeof(long long)));
    CAPDEC_CUDA_OK(cudaMemsetAsync(p.prof, 0, (size_t)(a.T * 16 + 2 * di->sms * 16) * sizeof(long long), st));
  }
  if (g_timing) {
    for (int i = 0; i < 4; ++i)
      if (!g_ev[i]) CAPDEC_CUDA_OK(cudaEventCreate(&g_ev[i]));
    CAPDEC_CUDA_OK(cudaEventRecord(g_ev[0], st));
  }
  CAPDEC_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p));
  count_launch();
  if (g_timing) {
    CAPDEC_CUDA_OK(cudaEventRecord(g_ev[1], st));
    g_timed[0] = true;
  }
  if (a.ragged) {
    dead_rows_zero_kernel<<<a.B * a.T, 128, 0, st>>>(a.len, a.B, a.lstm ? nullptr : (bf16*)a.m, p.R, 2 * a.F, a.g1, p.NG1);
    CAPDEC_LAUNCH_OK();
  }
  if (prof) {
    // debug only: synchronises.  Prints per-phase cycles of CTA 0 (work, barrier wait) for a few steps.
    std::vector<long long> h((size_t)a.T * 16 + 2 * di->sms * 16);
    CAPDEC_CUDA_OK(cudaStreamSynchronize(st));
    CAPDEC_CUDA_OK(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.prof);
    for (int t = 0; t < a.T; t += (a.T > 8 ? a.T / 4 : 1)) {
      fprintf(stderr, "recur_fwd prof t=%d:", t);
      for (int k = 1; k < 16 && h[t * 16 + k]; ++k) fprintf(stderr, " %lld", h[t * 16 + k] - h[t * 16 + k - 1]);
      fprintf(stderr, "\n");
    }
    unsigned flag = 0;
    cudaMemcpy(&flag, a.bar + 16, sizeof flag, cudaMemcpyDeviceToHost);
    if (flag) fprintf(stderr, "recur_fwd: a dataflow poll TIMED OUT (abort flag set)\n");
    dump_all_cta_stamps("recur_fwd", h.data() + (size_t)a.T * 16, di->sms);
    fine_dump("recur_fwd");
  }
  return CAPDEC_OK;
}

namespace {
size_t bwd_smem_bytes(const RecurBwdArgs& a) {
  size_t s = 4 * (size_t)STAGE;
  s += (size_t)32 * (a.D * 2 + WPAD);
  if (a.att) s += (size_t)16 * ((a.lstm ? 4 * a.D : 4 * a.F) * 2 + WPAD);
  s += (size_t)32 * (KC * 2 + WPAD);
  s += (size_t)2 * REDF * 4;
  s += (size_t)2 * pad4i(a.P > 0 ? a.P : 4) * 4;
  s += (size_t)(((a.B + 3) & ~3) + 2) * 4;
  return s + 8 + 4 * 8 + 128;
}

bool plan_bwd(const RecurBwdArgs& a, const DevInfo* di, size_t* smem) {
  if (!di || !di->coop || di->sms < 8) return false;
  if (a.B < 1 || a.B > 32 || a.T < 1) return false;
  if (a.lstm && !a.att) return false;
  if (a.D != KC || (!a.lstm && (a.F % 16 || (4 * a.F) % KC))) return false;
  if (a.att && (a.E % KC || a.E % CHUNK || a.A != KC || a.A % QW || a.P < 1 || a.P > GT)) return false;
  const int NQ = a.lstm ? 4 * a.D : 4 * a.F, KH = NQ + (a.att ? a.E + a.A : 0);
  if (!a.lstm && 4 * (2 * a.F / 32) > di->sms) return false;    // W job: one 32-feature slice per CTA
  if (a.att && a.E / 16 > di->sms) return false;                // Z job: 16 features per CTA
  if ((a.D / 16) * (KH / KC) > 2 * di->sms) return false;       // H jobs: at most two per CTA
  if (a.att && GR * (a.A / QW) > di->sms) return false;         // B items: one (row, 64 channels) per CTA and row group
  *smem = bwd_smem_bytes(a);
  return *smem <= (size_t)di->smem_optin;
}
}  // namespace

bool recur_bwd_supported(const RecurBwdArgs& a) {
  if (!persistent_enabled()) return false;
  size_t smem;
  return plan_bwd(a, dev_info(), &smem);
}

int recur_bwd(const RecurBwdArgs& a, cudaStream_t st) {
  const DevInfo* di = dev_info();
  size_t smem;
  CAPDEC_REQUIRE(plan_bwd(a, di, &smem), CAPDEC_ERR_BAD_SHAPE, "recur_bwd: shape not covered by the persistent kernel");
  BwdP p;
  memset(&p, 0, sizeof p);
  p.B = a.B; p.T = a.T; p.P = a.P; p.E = a.E; p.A = a.A; p.M = a.M; p.D = a.D; p.F = a.F;
  p.NQ = a.lstm ? 4 * a.D : 4 * a.F; p.NG1 = (a.att ? a.A + a.E : 0) + p.NQ; p.R = (int64_t)a.B * a.T; p.ldPX = a.ldPX;
  p.Whx2 = (const bf16*)a.Whx2; p.ldhx2 = a.ldhx2; p.dbx = (bf16*)a.dbx; p.ldbx = a.ldbx; p.dbx_off = a.dbx_off;
  p.len = a.len; p.WcT = (const bf16*)a.WcT; p.ldD = a.ldD; p.Wxin = (const bf16*)a.Wxin; p.ldNQ = a.ldNQ;
  p.Whx = (const bf16*)a.Whx; p.ldhx = a.ldhx; p.dHfc = a.dHfc; p.gates = a.gates; p.C = a.C; p.dc = a.dc;
  p.dh_rec = a.dh_rec; p.dpre = (bf16*)a.dpre; p.dpre_gm = (bf16*)a.dpre_gm; p.U = a.U; p.g1 = a.g1; p.v = a.v;
  p.q = a.q; p.du = (bf16*)a.du; p.duk = (bf16*)a.duk; p.dpx = (bf16*)a.dpx; p.dpxk = (bf16*)a.dpxk;
  p.dv_acc = a.dv_acc; p.dq_acc = a.dq_acc; p.dz = a.dz; p.awe = a.awe; p.alphas = a.alphas; p.d_alphas = a.d_alphas;
  p.enc_cm = (const bf16*)a.enc_cm; p.att1_cm = (const bf16*)a.att1_cm; p.w_f = a.w_f; p.part = a.part; p.de = a.de;
  p.dwf = a.dwf; p.dbf = a.dbf; p.bar = a.bar; p.dropout_p = a.dropout_p; p.seed = a.seed; p.dhp = a.dhp;
  p.skew = a.att ? RECUR_SKEW_BWD : 0;
  if (const char* sk = getenv("CAPDEC_RECUR_SKEW")) p.skew = atoi(sk);
  if (const char* sk = getenv("CAPDEC_RECUR_SKEW_BWD")) p.skew = atoi(sk);
  auto kernel = a.lstm ? recur_bwd_kernel<true, true> : a.att ? recur_bwd_kernel<true> : recur_bwd_kernel<false>;
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CAPDEC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RT, smem));
  CAPDEC_REQUIRE(per_sm >= 1, CAPDEC_ERR_CUDA, "recur_bwd: kernel does not fit one CTA per SM (smem %zu)", smem);
  if (a.att && a.build_enc_cm) {      // the forward ran the per-step path: the chunk-major feature copy is missing
    chunk_major_kernel<<<di->sms * 8, 256, 0, st>>>((const uint4*)a.enc, (uint4*)a.enc_cm, a.B, a.P, a.E, CHUNK);
    CAPDEC_LAUNCH_OK();
  }
  if (a.att) {
    chunk_major_kernel<<<di->sms * 4, 256, 0, st>>>((const uint4*)a.att1, (uint4*)a.att1_cm, a.B, a.P, a.A, QW);
    CAPDEC_LAUNCH_OK();
  }
  CAPDEC_REQUIRE((int64_t)GR * a.D <= (int64_t)(di->sms - 8) * GT, CAPDEC_ERR_BAD_SHAPE, "recur_bwd: one cell element per thread");
  CAPDEC_CUDA_OK(cudaMemsetAsync(a.bar, 0, 256, st));
  // exchange buffers of the reverse loop: the all-ones fill pattern their consumers poll on (see "Dataflow exchange")
  {
    const size_t R = (size_t)p.R;
    const int KH = p.NQ + (a.att ? a.E + a.A : 0);
    CAPDEC_REQUIRE(KH / KC <= 16, CAPDEC_ERR_BAD_SHAPE, "recur_bwd: more than 16 K chunks of the recurrent gradient");
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.dpre_gm, 0xFF, R * 4 * a.D * 2, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.dpxk, 0xFF, R * (size_t)KH * 2, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(a.dhp, 0xFF, R * (size_t)(KH / KC) * a.D * 4, st));
    if (!a.lstm) CAPDEC_CUDA_OK(cudaMemsetAsync(a.duk, 0xFF, R * p.NQ * 2, st));
    if (a.att) {
      CAPDEC_CUDA_OK(cudaMemsetAsync(a.dz, 0xFF, R * a.E * 4, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(a.part, 0xFF, R * (size_t)(a.E / CHUNK) * pad4i(a.P) * 4, st));
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  // CAPDEC_RECUR_BWD_CTAS: launch on fewer SMs than the device has, so that a small NCCL kernel (the all-reduce of the
  // fc gradients, capdec/parallel.py) runs NEXT TO the reverse loop instead of waiting for it.  Every role mapping of
  // the kernel only needs >= 144 CTAs at the reference dims (plan_bwd's limits are re-checked for the reduced grid).
  int nctas = di->sms;
  if (const char* e = getenv("CAPDEC_RECUR_BWD_CTAS")) {
    const int want = atoi(e);
    DevInfo fewer = *di;
    fewer.sms = want;
    size_t smem2;
    if (want >= 8 && want < di->sms && plan_bwd(a, &fewer, &smem2)) nctas = want;
  }
  cfg.gridDim = dim3(nctas, 1, 1);
  cfg.blockDim = dim3(RT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const char* prof_env = getenv("CAPDEC_RECUR_PROF");
  const bool prof = prof_env && prof_env[0] == '1';
  if (prof) {
    fine_reset();
    CAPDEC_CUDA_OK(cudaMalloc(&p.prof, (size_t)(a.T * 16 + 2 * di->sms * 16) * sizeof(long long)));
    CAPDEC_CUDA_OK(cudaMemsetAsync(p.prof, 0, (size_t)(a.T * 16 + 2 * di->sms * 16) * sizeof(long long), st));
  }
  if (g_timing) {
    for (int i = 0; i < 4; ++i)
      if (!g_ev[i]) CAPDEC_CUDA_OK(cudaEventCreate(&g_ev[i]));
    CAPDEC_CUDA_OK(cudaEventRecord(g_ev[2], st));
  }
  CAPDEC_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p));
  count_launch();
  if (g_timing) {
    CAPDEC_CUDA_OK(cudaEventRecord(g_ev[3], st));
    g_timed[1] = true;
  }
  if (prof) {
    std::vector<long long> h((size_t)a.T * 16 + 2 * di->sms * 16);
    CAPDEC_CUDA_OK(cudaStreamSynchronize(st));
    CAPDEC_CUDA_OK(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.prof);
    for (int t = 0; t < a.T; t += (a.T > 8 ? a.T / 4 : 1)) {
      fprintf(stderr, "recur_bwd prof t=%d:", t);
      for (int k = 1; k < 16 && h[t * 16 + k]; ++k) fprintf(stderr, " %lld", h[t * 16 + k] - h[t * 16 + k - 1]);
      fprintf(stderr, "\n");
    }
    unsigned flag = 0;
    cudaMemcpy(&flag, a.bar + 16, sizeof flag, cudaMemcpyDeviceToHost);
    if (flag) fprintf(stderr, "recur_bwd: a dataflow poll TIMED OUT (abort flag set)\n");
    dump_all_cta_stamps("recur_bwd", h.data() + (size_t)a.T * 16, di->sms);
    fine_dump("recur_bwd");
  }
  return CAPDEC_OK;
}

void recur_timing(int enable) {
  g_timing = enable != 0;
  if (!enable) g_timed[0] = g_timed[1] = false;
}
float recur_last_ms(int which) {
  if (which < 0 || which > 1 || !g_timed[which]) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_ev[2 * which + 1]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, g_ev[2 * which], g_ev[2 * which + 1]) != cudaSuccess) return -1.f;
  return ms;
}

}  // namespace capdec
