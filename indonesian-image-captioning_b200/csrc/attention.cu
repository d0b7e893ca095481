// attention.cu -- Bahdanau soft-attention decode step (forward and backward).
//
// Reference math: models/attention.py:26-44 (att1 hoisted out of the time loop, SURVEY.md
// App. C-6) fused with the f_beta gate of models/decoders/attention_scn.py:147-148:
//
//   e_p   = w_f . relu(att1[p,:] + att2) + b_f          att2 = W_d h + b_d  (from the G1 GEMM)
//   alpha = softmax_p(e)
//   awe   = sum_p alpha_p enc[p,:]
//   z     = sigmoid(beta_pre) * awe                      beta_pre = W_beta h + b_beta (G1 GEMM)
//
// This is the bandwidth stage of the decoder: every row streams att1[b] (P*A) and enc[b] (P*E)
// once per step, and at B = 32 rows it is LATENCY that decides (32 MB per step is 3-5 us of L2
// bandwidth).  Each direction is therefore two short, wide kernels chained with programmatic
// dependent launch, every thread issuing all of its 16-byte feature loads before the first use
// (and before the PDL wait, since the features do not depend on the previous kernel):
//
//   forward   attn_scores_kernel   grid (pixel chunks, rows)    e_p for one chunk of pixels
//             attn_wsum_kernel     grid (channel chunks, rows)  softmax (redundant per CTA, 196
//                                                               values) + weighted sum + gate
//   backward  attn_bwd_a_kernel    grid (channel chunks, rows)  gate backward + partial
//                                                               dalpha_p = enc[p, chunk] . dawe
//             attn_bwd_b_kernel    grid (rows)                  softmax backward + relu/score
//                                                               backward -> datt2, dw_f, de
//   after the loop  attn_datt1_kernel   dAtt1[b,p,a] = w_f[a] sum_t de[t,b,p] 1[att1+att2_t > 0]
//
// The (P x A) fp32 gradient of att1 is NOT read-modified-written every step (that alone was
// 25.7 MB of traffic per step at B = 32): the steps only keep de (P floats per row) and the
// batched kernel rebuilds the relu masks from att1 and the saved att2_t.  The first version of
// this file used one thread-block cluster per row with a DSMEM score exchange; cluster launches
// and the two cluster barriers cost more than the second kernel does.
// All global feature loads are coalesced 16-byte L1-bypassing loads; reductions over the
// attention / channel dims use warp shuffles.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int NT = 256;           // threads of the wide kernels
constexpr int NW = NT / 32;
constexpr int NTB = 512;          // threads of attn_bwd_b_kernel (one CTA per row)
constexpr int NWB = NTB / 32;
constexpr int PXS = 4;            // pixels a warp keeps in flight (score-side kernels)
constexpr int UNR = 13;           // pixels a thread keeps in flight (channel-side kernels)

__device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }

// block-wide max / sum over NTHR threads through `red` (>= NTHR/32 floats)
template <int NTHR>
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = red[0];
#pragma unroll
  for (int w = 1; w < NTHR / 32; ++w) m = fmaxf(m, red[w]);
  return m;
}
template <int NTHR>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < NTHR / 32; ++w) s += red[w];
  return s;
}

// ---------------------------------------------------------------------------------------
// forward A: scores of one pixel chunk
// ---------------------------------------------------------------------------------------
struct ScoreArgs {
  const void* att1; const float* g1; int64_t ldg; const float* w_f; const float* b_f;
  float* scores; int rows_per_map, P, A, Pc;
};

// NCH = 16-byte chunks per lane along the attention dim: A <= 32 * VEC * NCH
// RPM = rows of one feature map handled by the SAME CTA (beam search: the k beams of an image share
//       att1, attention_scn.py:189 `expand`): the map is read once for all of them.  RPM == 1: one row
//       per CTA, map = row / rows_per_map.
template <typename FT, int NCH, int RPM>
__global__ void __launch_bounds__(NT)
attn_scores_kernel(ScoreArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  pdl_launch_dependents();
  const int row0 = RPM == 1 ? blockIdx.y : blockIdx.y * RPM;
  const int map = RPM == 1 ? row0 / a.rows_per_map : blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = a.P, A = a.A;
  const int p_begin = blockIdx.x * a.Pc;
  const int p_end = min(P, p_begin + a.Pc);
  const FT* att1 = (const FT*)a.att1 + (int64_t)map * P * A;
  // feature loads do not depend on the previous kernel: issue them before the PDL wait
  uint4 v[PXS][NCH];
  const int p0 = p_begin + warp;
#pragma unroll
  for (int i = 0; i < PXS; ++i) {
    const int pi = p0 + i * NW;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int a0 = (c * 32 + lane) * VEC;
      v[i][c] = make_uint4(0, 0, 0, 0);
      if (a0 < A && pi < p_end) v[i][c] = ld_stream16(att1 + (int64_t)pi * A + a0);
    }
  }
  pdl_wait();                               // g1 comes from the previous kernel of the stream
  float att2[RPM][NCH][VEC], wf[NCH][VEC];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      wf[c][k] = (a0 + k < A) ? a.w_f[a0 + k] : 0.f;
#pragma unroll
      for (int j = 0; j < RPM; ++j)
        att2[j][c][k] = (a0 + k < A) ? a.g1[(int64_t)(row0 + j) * a.ldg + a0 + k] : 0.f;
    }
  }
  const float bf = a.b_f[0];
  const int Ppad = pad4(P);
  for (int p = p0;; p += PXS * NW) {
#pragma unroll
    for (int i = 0; i < PXS; ++i) {
      const int pi = p + i * NW;
      float s[RPM];
#pragma unroll
      for (int j = 0; j < RPM; ++j) s[j] = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float f[VEC];
        unpack16(v[i][c], f, FT());
#pragma unroll
        for (int k = 0; k < VEC; ++k)
#pragma unroll
          for (int j = 0; j < RPM; ++j) s[j] = fmaf(wf[c][k], fmaxf(f[k] + att2[j][c][k], 0.f), s[j]);
      }
#pragma unroll
      for (int j = 0; j < RPM; ++j) {
        const float t = warp_sum(s[j]);
        if (lane == 0 && pi < p_end) a.scores[(int64_t)(row0 + j) * Ppad + pi] = t + bf;
      }
    }
    if (p + PXS * NW >= p_end) break;
    // more pixels than one pass covers (very large batches only): reload
#pragma unroll
    for (int i = 0; i < PXS; ++i) {
      const int pi = p + PXS * NW + i * NW;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int a0 = (c * 32 + lane) * VEC;
        v[i][c] = make_uint4(0, 0, 0, 0);
        if (a0 < A && pi < p_end) v[i][c] = ld_stream16(att1 + (int64_t)pi * A + a0);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward B: softmax + weighted sum over one channel chunk + gate
// ---------------------------------------------------------------------------------------
struct WsumArgs {
  const void* enc; const float* g1; int64_t ldg; int beta_col;
  const float* scores; float* alpha_out; int64_t alpha_stride;
  void* z_out; int64_t ldz; float* awe_out;
  int rows_per_map, P, E, ncol;       // ncol = 16-byte columns per CTA
};

template <typename FT, int RPM>
__global__ void __launch_bounds__(NT)
attn_wsum_kernel(WsumArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  extern __shared__ __align__(16) float smem_f[];
  pdl_launch_dependents();
  const int row0 = RPM == 1 ? blockIdx.y : blockIdx.y * RPM;
  const int map = RPM == 1 ? row0 / a.rows_per_map : blockIdx.y;
  const int tid = threadIdx.x;
  const int P = a.P, E = a.E, ncol = a.ncol;
  const int Ppad = pad4(P);
  float* al = smem_f;                 // [RPM][Ppad] alpha
  float* red = al + RPM * Ppad;       // [NT * VEC] cross-group reduction (+ block reductions)
  const int groups = NT / ncol;       // pixel groups; threads >= groups*ncol idle
  const int grp = tid / ncol, col = tid - grp * ncol;
  const bool active = grp < groups;
  const int e0 = (blockIdx.x * ncol + col) * VEC;
  const FT* src = (const FT*)a.enc + (int64_t)map * P * E + e0;
  // first wave of feature loads before the PDL wait (they do not depend on the scores)
  uint4 q[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int p = grp + u * groups;
    q[u] = make_uint4(0, 0, 0, 0);
    if (active && p < P) q[u] = ld_stream16(src + (int64_t)p * E);
  }
  pdl_wait();
  // ---- softmax over the P scores of every row of this CTA (196 values each) ----
#pragma unroll 1
  for (int j = 0; j < RPM; ++j) {
    const float* sc = a.scores + (int64_t)(row0 + j) * Ppad;
    float* alj = al + j * Ppad;
    float m = -INFINITY;
    for (int p = tid; p < P; p += NT) { const float s = sc[p]; alj[p] = s; m = fmaxf(m, s); }
    m = block_max<NT>(m, red);
    float sum = 0.f;
    for (int p = tid; p < P; p += NT) { const float e = expf(alj[p] - m); alj[p] = e; sum += e; }
    sum = block_sum<NT>(sum, red);
    const float inv = 1.0f / sum;
    for (int p = tid; p < P; p += NT) {
      const float v = alj[p] * inv;
      alj[p] = v;
      if (blockIdx.x == 0 && a.alpha_out) a.alpha_out[(int64_t)(row0 + j) * a.alpha_stride + p] = v;
    }
  }
  __syncthreads();
  // ---- weighted sums: every feature vector is used for all RPM rows ----
  float acc[RPM][VEC];
#pragma unroll
  for (int j = 0; j < RPM; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[j][k] = 0.f;
  for (int pb = 0; pb < P; pb += UNR * groups) {
    if (pb > 0) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int p = pb + grp + u * groups;
        q[u] = make_uint4(0, 0, 0, 0);
        if (active && p < P) q[u] = ld_stream16(src + (int64_t)p * E);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + grp + u * groups;
      float f[VEC];
      unpack16(q[u], f, FT());
#pragma unroll
      for (int j = 0; j < RPM; ++j) {
        const float w = (active && p < P) ? al[j * Ppad + p] : 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[j][k] = fmaf(w, f[k], acc[j][k]);
      }
    }
  }
#pragma unroll 1
  for (int j = 0; j < RPM; ++j) {
    float accj[VEC];
#pragma unroll
    for (int jj = 0; jj < RPM; ++jj)
      if (jj == j) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) accj[k] = acc[jj][k];
      }
    if (groups > 1) {
      __syncthreads();
      if (active && grp > 0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) red[((grp - 1) * ncol + col) * VEC + k] = accj[k];
      }
      __syncthreads();
      if (grp == 0) {
        for (int g = 1; g < groups; ++g)
#pragma unroll
          for (int k = 0; k < VEC; ++k) accj[k] += red[((g - 1) * ncol + col) * VEC + k];
      }
    }
    if (grp == 0) {
      const int row = row0 + j;
      const float* g1 = a.g1 + (int64_t)row * a.ldg;
      float zv[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float gate = 1.0f;
        if (a.beta_col >= 0) gate = sigmoidf_(g1[a.beta_col + e0 + k]);
        zv[k] = gate * accj[k];
      }
      if (a.awe_out) {
        float* dst = a.awe_out + (int64_t)row * E + e0;
#pragma unroll
        for (int k = 0; k < VEC; k += 4)
          *reinterpret_cast<float4*>(dst + k) = make_float4(accj[k], accj[k + 1], accj[k + 2], accj[k + 3]);
      }
      if (a.z_out) {
        FT* dst = (FT*)a.z_out + (int64_t)row * a.ldz + e0;
        *reinterpret_cast<uint4*>(dst) = pack16(zv, FT());
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward A, streaming variant for large batches (beam search): one CTA = (feature map, quarter of the
// pixels); its att1 rows are consecutive in memory, so ONE cp.async.bulk (<= 49 KB) stages them while
// the att2 vectors of the RPM rows (beams) are fetched; every staged pixel is used for all RPM rows.
// ---------------------------------------------------------------------------------------
constexpr int SS_QUARTERS = 4;

template <int RPM>
__global__ void __launch_bounds__(NT)
attn_scores_stream_kernel(ScoreArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, A = a.A;                                   // A == 512
  const int map = blockIdx.y, row0 = map * RPM;
  const int pq = (P + SS_QUARTERS - 1) / SS_QUARTERS;
  const int p_begin = blockIdx.x * pq, p_end = min(P, p_begin + pq);
  const int cnt = p_end - p_begin;
  if (cnt <= 0) return;
  const uint32_t full = smem_u32(&bar);
  if (tid == 0) {
    mbar_init(full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(full, (uint32_t)cnt * A * 2);
    bulk_g2s(smem_u32(smem_raw), (const bf16*)a.att1 + ((int64_t)map * P + p_begin) * A, (uint32_t)cnt * A * 2, full);
  }
  pdl_wait();                               // g1 comes from the previous kernel of the stream
  float att2[RPM][2][8], wf[2][8];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int a0 = (c * 32 + lane) * 8;
    const float4 w0 = *reinterpret_cast<const float4*>(a.w_f + a0), w1 = *reinterpret_cast<const float4*>(a.w_f + a0 + 4);
    wf[c][0] = w0.x; wf[c][1] = w0.y; wf[c][2] = w0.z; wf[c][3] = w0.w;
    wf[c][4] = w1.x; wf[c][5] = w1.y; wf[c][6] = w1.z; wf[c][7] = w1.w;
#pragma unroll
    for (int j = 0; j < RPM; ++j) {
      const float* g = a.g1 + (int64_t)(row0 + j) * a.ldg + a0;
      const float4 x0 = *reinterpret_cast<const float4*>(g), x1 = *reinterpret_cast<const float4*>(g + 4);
      att2[j][c][0] = x0.x; att2[j][c][1] = x0.y; att2[j][c][2] = x0.z; att2[j][c][3] = x0.w;
      att2[j][c][4] = x1.x; att2[j][c][5] = x1.y; att2[j][c][6] = x1.z; att2[j][c][7] = x1.w;
    }
  }
  const float bf = a.b_f[0];
  const int Ppad = pad4(P);
  __syncthreads();                          // barrier initialised before anyone waits on it
  mbar_wait(full, 0);
#pragma unroll 1
  for (int pl = warp; pl < cnt; pl += NW) {
    float s[RPM];
#pragma unroll
    for (int j = 0; j < RPM; ++j) s[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 raw = *reinterpret_cast<const uint4*>(smem_raw + ((size_t)pl * A + (c * 32 + lane) * 8) * 2);
      float f[8];
      unpack16(raw, f, bf16());
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < RPM; ++j) s[j] = fmaf(wf[c][k], fmaxf(f[k] + att2[j][c][k], 0.f), s[j]);
    }
#pragma unroll
    for (int j = 0; j < RPM; ++j) {
      const float tsum = warp_sum(s[j]);
      if (lane == 0) a.scores[(int64_t)(row0 + j) * Ppad + p_begin + pl] = tsum + bf;
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward B, streaming variant for large batches (beam search: 1 875 rows, 0.5 GB of features per step,
// HBM-bound): the chunk-major feature copy enc_cm [map][E/512][P][512] makes the P x 512 slab of a
// (map, chunk) item contiguous, so the pixels stream through a 3-stage shared-memory ring with ONE
// cp.async.bulk (32 pixels = 32 KB) per stage -- the bytes in flight no longer depend on registers,
// and two CTAs per SM overlap each other's softmax / epilogue.  RPM rows (beams) share the item.
// ---------------------------------------------------------------------------------------
constexpr int WS_STAGES = 3;
constexpr int WS_PX = 32;                       // pixels per stage
constexpr int WS_CH = 512;                      // channels per item
constexpr int WS_STAGE_BYTES = WS_PX * WS_CH * 2;

// barrier over the NT consumer threads (the producer warp is not part of it)
__device__ __forceinline__ void ws_csync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

template <int RPM>
__global__ void __launch_bounds__(NT + 32, 2)
attn_wsum_stream_kernel(WsumArgs a, int maps) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* stg = smem_raw;
  float* al = reinterpret_cast<float*>(stg + WS_STAGES * WS_STAGE_BYTES);     // [RPM][Ppad]
  const int P = a.P, E = a.E;
  const int Ppad = pad4(P);
  float* red = al + RPM * Ppad;                                               // [3 * 64 * 8] + block reductions
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 3 * 64 * 8 + 32);
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = E / WS_CH;
  const int nfill = (P + WS_PX - 1) / WS_PX;
  // PERSISTENT: this CTA walks over the items b, b + grid, ... (item = (map, chunk), chunks of a map adjacent) and
  // the ring keeps running ACROSS items: while the softmax prologue / the reduction epilogue of an item execute, the
  // next item's pixels are already in flight (a one-item CTA left its share of the bandwidth idle for ~25 % of its life)
  const int items = maps * chunks;
  const int my_items = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_fills = my_items * nfill;
  // full[s]: the stage's bytes have landed (TMA transaction count); empty[s]: all NW consumer warps are done with it.
  // A dedicated PRODUCER warp refills a stage as soon as it is empty: no CTA-wide barrier in the steady state (with a
  // __syncthreads per fill the warps spent 4 issue slots of 5 stalled on that barrier, ncu).
  const uint32_t bar0 = smem_u32(bars);          // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (WS_STAGES + s)
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (WS_STAGES + s), NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NW) {
    // ---- producer: the features do not depend on the previous kernel, so it starts before the PDL wait ----
    if (lane == 0) {
      for (int gf = 0; gf < total_fills; ++gf) {          // gf = fill index over all of this CTA's items
        const int k = gf / nfill, fi = gf - k * nfill;
        const int it = (int)blockIdx.x + k * (int)gridDim.x;
        const bf16* src = (const bf16*)a.enc + (int64_t)it * P * WS_CH;       // item it = (map, chunk) slab of enc_cm
        const int s = gf % WS_STAGES;
        mbar_wait(bar0 + 8 * (WS_STAGES + s), ((gf / WS_STAGES) & 1) ^ 1);
        const uint32_t bytes = (uint32_t)min(WS_PX, P - fi * WS_PX) * WS_CH * 2;
        mbar_expect_tx(bar0 + 8 * s, bytes);
        bulk_g2s(smem_u32(stg) + s * WS_STAGE_BYTES, src + (int64_t)fi * WS_PX * WS_CH, bytes, bar0 + 8 * s);
      }
    }
    return;
  }
  pdl_wait();
  const int grp = tid >> 6, col = tid & 63;      // weighted sums: thread = (pixel group of 4, 16-byte column of 64)
  int gf = 0;                                    // next fill to consume
#pragma unroll 1
  for (int k = 0; k < my_items; ++k) {
    const int it = (int)blockIdx.x + k * (int)gridDim.x;
    const int map = it / chunks, chunk = it - map * chunks;
    const int row0 = map * RPM;
    // ---- softmax over the P scores of every row of this item: one WARP per row (shuffle reductions only) ----
    for (int j = warp; j < RPM; j += NW) {
      const float* sc = a.scores + (int64_t)(row0 + j) * Ppad;
      float* alj = al + j * Ppad;
      float m = -INFINITY;
      for (int p = lane; p < P; p += 32) { const float sv = sc[p]; alj[p] = sv; m = fmaxf(m, sv); }
      m = warp_max(m);
      float sum = 0.f;
      for (int p = lane; p < P; p += 32) { const float e = expf(alj[p] - m); alj[p] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int p = lane; p < P; p += 32) {
        const float v = alj[p] * inv;
        alj[p] = v;
        if (chunk == 0 && a.alpha_out) a.alpha_out[(int64_t)(row0 + j) * a.alpha_stride + p] = v;
      }
    }
    ws_csync();
    float acc[RPM][8];
#pragma unroll
    for (int j = 0; j < RPM; ++j)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) acc[j][kk] = 0.f;
#pragma unroll 1
    for (int fi = 0; fi < nfill; ++fi, ++gf) {
      const int s = gf % WS_STAGES;
      const int px0 = fi * WS_PX, cnt = min(WS_PX, P - px0);
      mbar_wait(bar0 + 8 * s, (gf / WS_STAGES) & 1);
      const uint8_t* base = stg + s * WS_STAGE_BYTES + col * 16;
#pragma unroll
      for (int u = 0; u < WS_PX / 4; ++u) {
        const int pl = grp + u * 4;
        if (pl < cnt) {
          const uint4 raw = *reinterpret_cast<const uint4*>(base + (size_t)pl * WS_CH * 2);
          float f[8];
          unpack16(raw, f, bf16());
#pragma unroll
          for (int j = 0; j < RPM; ++j) {
            const float w = al[j * Ppad + px0 + pl];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) acc[j][kk] = fmaf(w, f[kk], acc[j][kk]);
          }
        }
      }
      __syncwarp();
      if (lane == 0)                                      // this warp is done with stage s
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (WS_STAGES + s)) : "memory");
    }
    // ---- cross-group reduction, gate, outputs ----
    const int e0 = chunk * WS_CH + col * 8;
#pragma unroll 1
    for (int j = 0; j < RPM; ++j) {
      float accj[8];
#pragma unroll
      for (int jj = 0; jj < RPM; ++jj)
        if (jj == j) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) accj[kk] = acc[jj][kk];
        }
      const int row = row0 + j;
      // gate pre-activations requested before the reduction barriers: their latency hides behind them
      float gpre[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        gpre[kk] = (grp == 0 && a.beta_col >= 0) ? __ldg(a.g1 + (int64_t)row * a.ldg + a.beta_col + e0 + kk) : 0.f;
      ws_csync();
      if (grp > 0) {
        float4* dst = reinterpret_cast<float4*>(red + ((grp - 1) * 64 + col) * 8);
        dst[0] = make_float4(accj[0], accj[1], accj[2], accj[3]);
        dst[1] = make_float4(accj[4], accj[5], accj[6], accj[7]);
      }
      ws_csync();
      if (grp == 0) {
#pragma unroll
        for (int g = 1; g < 4; ++g)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) accj[kk] += red[((g - 1) * 64 + col) * 8 + kk];
        float zv[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          float gate = 1.0f;
          if (a.beta_col >= 0) gate = sigmoidf_(gpre[kk]);
          zv[kk] = gate * accj[kk];
        }
        if (a.awe_out) {
          float* dst = a.awe_out + (int64_t)row * E + e0;
          *reinterpret_cast<float4*>(dst) = make_float4(accj[0], accj[1], accj[2], accj[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(accj[4], accj[5], accj[6], accj[7]);
        }
        if (a.z_out) *reinterpret_cast<uint4*>((bf16*)a.z_out + (int64_t)row * a.ldz + e0) = pack16(zv, bf16());
      }
    }
    ws_csync();                           // `al` and `red` are rewritten by the next item
  }
}

// ---------------------------------------------------------------------------------------
// backward (attention part)
//   dgate = dz*awe ; dawe = dz*gate ; dbeta_pre = dgate*gate*(1-gate)
//   dalpha_p = enc[p,:].dawe + dalpha_ext_p
//   de_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q)
//   drelu[p,a] = de_p w_f[a] 1[att1[p,a]+att2[a] > 0]
//   datt2[a] = sum_p drelu[p,a] ; dAtt1[p,a] += drelu[p,a] (after the loop, attn_datt1_kernel)
//   dw_f[a] += sum_p de_p relu(.)[p,a] ; db_f += sum_p de_p
// ---------------------------------------------------------------------------------------
struct BwdAArgs {
  const void* enc; const float* g1; int64_t ldg; int beta_col;
  const float* dz; int64_t lddz; const float* awe;
  void* dba; int64_t lddba;           // feature type: [dbeta_pre (E) | datt2 (A)]
  float* part;                        // [rows][EC][Ppad] partial dalpha
  int P, E, ncol, cpl;                // cpl = lanes per pixel group (power of two, >= ncol, <= 32)
};

template <typename FT>
__global__ void __launch_bounds__(NT)
attn_bwd_a_kernel(BwdAArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  pdl_launch_dependents();
  const int row = blockIdx.y, chunk = blockIdx.x, EC = gridDim.x;
  const int tid = threadIdx.x;
  const int P = a.P, E = a.E, ncol = a.ncol, cpl = a.cpl;
  const int Ppad = pad4(P);
  const int npg = NT / cpl;           // pixel groups per CTA
  const int pg = tid / cpl, col = tid - pg * cpl;
  const bool col_ok = col < ncol;
  const int e0 = (chunk * ncol + col) * VEC;
  const FT* src = (const FT*)a.enc + (int64_t)row * P * E + e0;
  uint4 q[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int p = pg + u * npg;
    q[u] = make_uint4(0, 0, 0, 0);
    if (col_ok && p < P) q[u] = ld_stream16(src + (int64_t)p * E);
  }
  pdl_wait();                         // dz comes from the previous kernel
  // ---- gate backward on this thread's 16-byte column ----
  float dawe[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) dawe[k] = 0.f;
  if (col_ok) {
    const float* g1 = a.g1 + (int64_t)row * a.ldg + a.beta_col + e0;
    const float* dz = a.dz + (int64_t)row * a.lddz + e0;
    const float* awe = a.awe + (int64_t)row * E + e0;
    float db[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float gate = sigmoidf_(g1[k]);
      const float d = dz[k];
      dawe[k] = d * gate;
      db[k] = d * awe[k] * gate * (1.0f - gate);
    }
    if (pg == 0) *reinterpret_cast<uint4*>((FT*)a.dba + (int64_t)row * a.lddba + e0) = pack16(db, FT());
  }
  // ---- partial dalpha_p over this channel chunk ----
  float* part = a.part + ((int64_t)row * EC + chunk) * Ppad;
  for (int pb = 0; pb < P; pb += UNR * npg) {
    if (pb > 0) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int p = pb + pg + u * npg;
        q[u] = make_uint4(0, 0, 0, 0);
        if (col_ok && p < P) q[u] = ld_stream16(src + (int64_t)p * E);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = pb + pg + u * npg;
      float f[VEC];
      unpack16(q[u], f, FT());
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VEC; ++k) s = fmaf(f[k], dawe[k], s);
      for (int o = cpl >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (col == 0 && p < P) part[p] = s;
    }
  }
}

struct BwdBArgs {
  const void* att1; const float* g1; int64_t ldg; const float* w_f;
  const float* alpha; int64_t alpha_stride; const float* dalpha_ext; int64_t dalpha_stride;
  const float* part; int EC;
  void* dba; int64_t lddba; int dba_col;       // datt2 -> dba[row][dba_col + a]
  float* de_out;                               // [rows][Ppad]
  float* dwf_part; float* dbf_part;            // [rows][A], [rows]
  int P, A;
};

template <typename FT, int NCH>
__global__ void __launch_bounds__(NTB, 1)
attn_bwd_b_kernel(BwdBArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  extern __shared__ __align__(16) float smem_f[];
  pdl_launch_dependents();
  const int row = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = a.P, A = a.A;
  const int Ppad = pad4(P), Apad = pad4(A);
  float* al = smem_f;                 // [Ppad] alpha
  float* de = al + Ppad;              // [Ppad]
  float* red = de + Ppad;             // [32]
  float* redA = red + 32;             // [NWB][2][Apad]
  const FT* att1 = (const FT*)a.att1 + (int64_t)row * P * A;
  // first wave of att1 loads before the PDL wait
  uint4 v[PXS][NCH];
#pragma unroll
  for (int i = 0; i < PXS; ++i) {
    const int pi = warp + i * NWB;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int a0 = (c * 32 + lane) * VEC;
      v[i][c] = make_uint4(0, 0, 0, 0);
      if (a0 < A && pi < P) v[i][c] = ld_stream16(att1 + (int64_t)pi * A + a0);
    }
  }
  pdl_wait();
  // ---- dalpha = sum of the channel-chunk partials (+ external), softmax backward ----
  float dot = 0.f;
  for (int p = tid; p < P; p += NTB) {
    float d = 0.f;
    for (int c = 0; c < a.EC; ++c) d += a.part[((int64_t)row * a.EC + c) * Ppad + p];
    if (a.dalpha_ext) d += a.dalpha_ext[(int64_t)row * a.dalpha_stride + p];
    const float al_p = a.alpha[(int64_t)row * a.alpha_stride + p];
    al[p] = al_p;
    de[p] = d;
    dot = fmaf(al_p, d, dot);
  }
  dot = block_sum<NTB>(dot, red);
  float sde = 0.f;
  for (int p = tid; p < P; p += NTB) {
    const float x = al[p] * (de[p] - dot);
    de[p] = x;
    sde += x;
    if (a.de_out) a.de_out[(int64_t)row * Ppad + p] = x;
  }
  sde = block_sum<NTB>(sde, red);
  if (tid == 0) a.dbf_part[row] = sde;
  __syncthreads();
  // ---- relu / score backward: warp per pixel, lane holds NCH*VEC attention features ----
  const float* g1 = a.g1 + (int64_t)row * a.ldg;
  float att2[NCH][VEC], wf[NCH][VEC], dacc[NCH][VEC], wacc[NCH][VEC];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      att2[c][k] = (a0 + k < A) ? g1[a0 + k] : 0.f;
      wf[c][k] = (a0 + k < A) ? a.w_f[a0 + k] : 0.f;
      dacc[c][k] = 0.f;
      wacc[c][k] = 0.f;
    }
  }
  for (int p = warp; p < P; p += PXS * NWB) {
    if (p != warp) {
#pragma unroll
      for (int i = 0; i < PXS; ++i) {
        const int pi = p + i * NWB;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int a0 = (c * 32 + lane) * VEC;
          v[i][c] = make_uint4(0, 0, 0, 0);
          if (a0 < A && pi < P) v[i][c] = ld_stream16(att1 + (int64_t)pi * A + a0);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PXS; ++i) {
      const int pi = p + i * NWB;
      const float dep = pi < P ? de[pi] : 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float f[VEC];
        unpack16(v[i][c], f, FT());
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float pre = f[k] + att2[c][k];
          const float on = pre > 0.f ? 1.f : 0.f;
          dacc[c][k] = fmaf(dep * wf[c][k], on, dacc[c][k]);
          wacc[c][k] = fmaf(dep, pre * on, wacc[c][k]);
        }
      }
    }
  }
  // cross-warp reduction of datt2 / dw_f
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
    if (a0 < A) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        redA[(warp * 2 + 0) * Apad + a0 + k] = dacc[c][k];
        redA[(warp * 2 + 1) * Apad + a0 + k] = wacc[c][k];
      }
    }
  }
  __syncthreads();
  FT* dba = (FT*)a.dba + (int64_t)row * a.lddba + a.dba_col;
  for (int i = tid; i < 2 * A; i += NTB) {
    const int which = i / A, aa = i - which * A;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NWB; ++w) s += redA[(w * 2 + which) * Apad + aa];
    if (which == 0) dba[aa] = from_f<FT>(s);
    else a.dwf_part[(int64_t)row * A + aa] = s;
  }
}

// ---------------------------------------------------------------------------------------
// after the loop: dAtt1[b,p,a] (+)= w_f[a] * sum_t de[t,b,p] * 1[att1[b,p,a] + att2[t,b,a] > 0]
// grid (pixel tiles, B); thread = attention feature(s); PT pixels per CTA in registers
// ---------------------------------------------------------------------------------------
constexpr int PT = 14;
struct Datt1Args {
  const void* att1; const float* g1; int64_t ldg; int64_t g1_step;      // att2[t][b] = g1 + t*g1_step + b*ldg
  const float* de; int64_t de_step;                                     // de[t][b][Ppad]
  const float* w_f; float* dAtt1; int accumulate; int T, P, A;
};

template <typename FT, int NA>
__global__ void __launch_bounds__(NT)
attn_datt1_kernel(Datt1Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const int b = blockIdx.y, p0 = blockIdx.x * PT;
  const int tid = threadIdx.x;
  const int T = a.T, P = a.P, A = a.A, Ppad = pad4(P);
  float* sde = smem_f;                // [T][PT]
  for (int i = tid; i < T * PT; i += NT) {
    const int t = i / PT, j = i - t * PT;
    sde[i] = (p0 + j < P) ? a.de[(int64_t)t * a.de_step + (int64_t)b * Ppad + p0 + j] : 0.f;
  }
  const FT* att1 = (const FT*)a.att1 + ((int64_t)b * P + p0) * A;
  float x[NA][PT], acc[NA][PT];
#pragma unroll
  for (int c = 0; c < NA; ++c) {
    const int aa = tid + c * NT;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      x[c][j] = (aa < A && p0 + j < P) ? to_f(att1[(int64_t)j * A + aa]) : 0.f;
      acc[c][j] = 0.f;
    }
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float* att2 = a.g1 + (int64_t)t * a.g1_step + (int64_t)b * a.ldg;
#pragma unroll
    for (int c = 0; c < NA; ++c) {
      const int aa = tid + c * NT;
      const float a2 = aa < A ? att2[aa] : 0.f;
#pragma unroll
      for (int j = 0; j < PT; ++j)
        acc[c][j] = fmaf(sde[t * PT + j], (x[c][j] + a2 > 0.f) ? 1.f : 0.f, acc[c][j]);
    }
  }
#pragma unroll
  for (int c = 0; c < NA; ++c) {
    const int aa = tid + c * NT;
    if (aa < A) {
      const float wf = a.w_f[aa];
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (p0 + j < P) {
          float* dst = a.dAtt1 + ((int64_t)b * P + p0 + j) * A + aa;
          const float val = wf * acc[c][j];
          *dst = a.accumulate ? *dst + val : val;
        }
      }
    }
  }
}

// ------------------------------- host side -------------------------------
// channel chunking shared by forward B and backward A: ncol 16-byte columns per CTA, <= 32
struct Chunking { int EC, ncol, cpl; };
Chunking pick_chunks(int E, int vec, int rows) {
  const int cols = E / vec;
  int EC = ceil_div(cols, 32);
  while (cols % EC != 0) ++EC;
  // small batches: more, narrower CTAs so that every thread's loads fit one wave
  while ((int64_t)rows * EC < 2 * 148 && (cols / EC) % 2 == 0 && cols / EC > 8) EC *= 2;
  Chunking c;
  c.EC = EC;
  c.ncol = cols / EC;
  c.cpl = 1;
  while (c.cpl < c.ncol) c.cpl *= 2;
  return c;
}

template <typename ArgsT>
int launch_attn(void (*kernel)(ArgsT), dim3 grid, int threads, size_t smem, cudaStream_t st,
                const ArgsT& args) {
  CAPDEC_CUDA_OK(launch_pdl(kernel, grid, dim3(threads, 1, 1), smem, st, 1, args));
  count_launch();
  return CAPDEC_OK;
}

}  // namespace

int attention_init() {
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_b_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_b_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_b_kernel<bf16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_b_kernel<bf16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  return CAPDEC_OK;
}

size_t attention_scratch_floats(int precision, int rows, int P, int E) {
  if (rows <= 0 || P <= 0 || E <= 0) return 0;
  const int vec = precision == CAPDEC_BF16 ? 8 : 4;
  if (E % vec != 0) return 0;
  const int Ppad = (P + 3) & ~3;
  // forward: scores [rows][Ppad]; backward: partial dalpha [rows][EC][Ppad] (EC is largest at rows = 1)
  const Chunking c = pick_chunks(E, vec, 1);
  return (size_t)rows * (size_t)c.EC * Ppad;
}

namespace {
template <int RPM>
int launch_wsum_stream(const WsumArgs& wa, int maps, cudaStream_t st) {
  auto kernel = attn_wsum_stream_kernel<RPM>;
  const size_t smem = (size_t)WS_STAGES * WS_STAGE_BYTES + (size_t)(RPM * ((wa.P + 3) & ~3) + 3 * 64 * 8 + 32) * 4 +
                      2 * WS_STAGES * 8 + 16;
  static std::once_flag once;
  static cudaError_t rc = cudaSuccess;
  std::call_once(once, [&] { rc = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); });
  CAPDEC_REQUIRE(rc == cudaSuccess && smem <= 112 * 1024, CAPDEC_ERR_CUDA, "attn_wsum_stream: smem %zu", smem);
  // persistent: two CTAs per SM (smem-limited), each walking over its share of the (map, chunk) items
  const int items = (wa.E / WS_CH) * maps;
  const int grid = items < 2 * 148 ? items : 2 * 148;
  CAPDEC_CUDA_OK(launch_pdl(kernel, dim3(grid, 1, 1), dim3(NT + 32, 1, 1), smem, st, 1, wa, maps));
  count_launch();
  return CAPDEC_OK;
}
}  // namespace

int attention_fwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* b_f, float* alpha_out,
                  int64_t alpha_stride, void* z_out, int64_t ldz, float* awe_out, int rows,
                  int rows_per_map, int P, int E, int A, float* scratch, cudaStream_t st, const void* enc_cm) {
  if (rows <= 0) return CAPDEC_OK;
  const int vec = precision == CAPDEC_BF16 ? 8 : 4;
  CAPDEC_REQUIRE(E % vec == 0 && A % vec == 0 && A <= 32 * vec * 4, CAPDEC_ERR_BAD_SHAPE,
                 "attention: need E,A multiples of %d and A <= %d (E=%d A=%d)", vec, 32 * vec * 4, E, A);
  CAPDEC_REQUIRE(scratch != nullptr, CAPDEC_ERR_BAD_ARG, "attention: scratch is NULL");
  // rows of one feature map (beam search) are handled by one CTA when the beam width has an instance
  const int rpm = (precision == CAPDEC_BF16 && A <= 32 * vec * 2 && rows_per_map >= 2 && rows_per_map <= 5 &&
                   rows % rows_per_map == 0) ? rows_per_map : 1;
  const int gy = rows / rpm;
  const char* force = getenv("CAPDEC_WSUM_STREAM");          // tests: 1 forces the streaming kernels at any size
  const bool want_stream = precision == CAPDEC_BF16 && rpm == rows_per_map && !(force && force[0] == '0') &&
                           (gy >= 148 || (force && force[0] == '1'));
  // ---- A: scores ----
  bool scores_done = false;
  if (want_stream && A == 512 && ((uintptr_t)g1 % 16) == 0 && (ldg % 4) == 0) {
    ScoreArgs sa{att1, g1, ldg, w_f, b_f, scratch, rows_per_map, P, A, 0};
    const size_t smem = (size_t)ceil_div(P, SS_QUARTERS) * A * 2;
    dim3 gs(SS_QUARTERS, gy, 1);
#define SS_LAUNCH(R_)                                                                                         \
    do {                                                                                                       \
      static std::once_flag once;                                                                              \
      static cudaError_t rc = cudaSuccess;                                                                     \
      std::call_once(once, [&] {                                                                               \
        rc = cudaFuncSetAttribute(attn_scores_stream_kernel<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
      });                                                                                                      \
      CAPDEC_REQUIRE(rc == cudaSuccess && smem <= 100 * 1024, CAPDEC_ERR_CUDA, "attn_scores_stream: smem %zu", smem); \
      CAPDEC_TRY(launch_attn(attn_scores_stream_kernel<R_>, gs, NT, smem, st, sa));                            \
    } while (0)
    switch (rpm) {
      case 1: SS_LAUNCH(1); break;
      case 2: SS_LAUNCH(2); break;
      case 3: SS_LAUNCH(3); break;
      case 4: SS_LAUNCH(4); break;
      default: SS_LAUNCH(5); break;
    }
#undef SS_LAUNCH
    scores_done = true;
  }
  int PC = ceil_div(P, PXS * NW);                                   // one pass of PXS*NW pixels per CTA ...
  while ((int64_t)gy * PC > 64 * 148 && PC > 1) PC = (PC + 1) / 2;      // ... unless the grid gets too large
  ScoreArgs sa{att1, g1, ldg, w_f, b_f, scratch, rows_per_map, P, A, ceil_div(P, PC)};
  PC = ceil_div(P, sa.Pc);
  const bool small = A <= 32 * vec * 2;
  dim3 gs(PC, gy, 1);
  if (scores_done) {
  } else if (precision == CAPDEC_BF16) {
    if (rpm == 2) CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 2, 2>, gs, NT, 0, st, sa));
    else if (rpm == 3) CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 2, 3>, gs, NT, 0, st, sa));
    else if (rpm == 4) CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 2, 4>, gs, NT, 0, st, sa));
    else if (rpm == 5) CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 2, 5>, gs, NT, 0, st, sa));
    else if (small) CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 2, 1>, gs, NT, 0, st, sa));
    else CAPDEC_TRY(launch_attn(attn_scores_kernel<bf16, 4, 1>, gs, NT, 0, st, sa));
  } else {
    if (small) CAPDEC_TRY(launch_attn(attn_scores_kernel<float, 2, 1>, gs, NT, 0, st, sa));
    else CAPDEC_TRY(launch_attn(attn_scores_kernel<float, 4, 1>, gs, NT, 0, st, sa));
  }
  // ---- B: softmax + weighted sum + gate ----
  if (want_stream && enc_cm != nullptr && E % WS_CH == 0) {
    // large batches: shared-memory ring fed by bulk copies from the chunk-major feature copy
    WsumArgs ws{enc_cm, g1, ldg, beta_col, scratch, alpha_out, alpha_stride, z_out, ldz, awe_out,
                rows_per_map, P, E, 0};
    switch (rpm) {
      case 1: return launch_wsum_stream<1>(ws, gy, st);
      case 2: return launch_wsum_stream<2>(ws, gy, st);
      case 3: return launch_wsum_stream<3>(ws, gy, st);
      case 4: return launch_wsum_stream<4>(ws, gy, st);
      case 5: return launch_wsum_stream<5>(ws, gy, st);
    }
  }
  const Chunking ch = pick_chunks(E, vec, gy);
  WsumArgs wa{enc, g1, ldg, beta_col, scratch, alpha_out, alpha_stride, z_out, ldz, awe_out,
              rows_per_map, P, E, ch.ncol};
  const size_t smem = (size_t)(rpm * ((P + 3) & ~3) + NT * 8) * sizeof(float);
  dim3 gw(ch.EC, gy, 1);
  if (precision == CAPDEC_BF16) {
    if (rpm == 2) return launch_attn(attn_wsum_kernel<bf16, 2>, gw, NT, smem, st, wa);
    if (rpm == 3) return launch_attn(attn_wsum_kernel<bf16, 3>, gw, NT, smem, st, wa);
    if (rpm == 4) return launch_attn(attn_wsum_kernel<bf16, 4>, gw, NT, smem, st, wa);
    if (rpm == 5) return launch_attn(attn_wsum_kernel<bf16, 5>, gw, NT, smem, st, wa);
    return launch_attn(attn_wsum_kernel<bf16, 1>, gw, NT, smem, st, wa);
  }
  return launch_attn(attn_wsum_kernel<float, 1>, gw, NT, smem, st, wa);
}

int attention_bwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* alpha, int64_t alpha_stride,
                  const float* dalpha_ext, int64_t dalpha_stride, const float* dz, int64_t lddz,
                  const float* awe, void* dba, int64_t lddba, float* de_out, float* dwf_part,
                  float* dbf_part, int rows, int P, int E, int A, float* scratch, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  const int vec = precision == CAPDEC_BF16 ? 8 : 4;
  CAPDEC_REQUIRE(E % vec == 0 && A % vec == 0 && A <= 32 * vec * 4 && beta_col >= 0,
                 CAPDEC_ERR_BAD_SHAPE, "attention bwd: unsupported dims E=%d A=%d", E, A);
  CAPDEC_REQUIRE(scratch != nullptr, CAPDEC_ERR_BAD_ARG, "attention bwd: scratch is NULL");
  const Chunking ch = pick_chunks(E, vec, rows);
  BwdAArgs aa{enc, g1, ldg, beta_col, dz, lddz, awe, dba, lddba, scratch, P, E, ch.ncol, ch.cpl};
  dim3 ga(ch.EC, rows, 1);
  if (precision == CAPDEC_BF16) CAPDEC_TRY(launch_attn(attn_bwd_a_kernel<bf16>, ga, NT, 0, st, aa));
  else CAPDEC_TRY(launch_attn(attn_bwd_a_kernel<float>, ga, NT, 0, st, aa));
  BwdBArgs ba{att1, g1, ldg, w_f, alpha, alpha_stride, dalpha_ext, dalpha_stride, scratch, ch.EC,
              dba, lddba, E, de_out, dwf_part, dbf_part, P, A};
  const int Ppad = (P + 3) & ~3, Apad = (A + 3) & ~3;
  const size_t smem = (size_t)(2 * Ppad + 32 + NWB * 2 * Apad) * sizeof(float);
  CAPDEC_REQUIRE(smem <= 160 * 1024, CAPDEC_ERR_BAD_SHAPE, "attention bwd: smem %zu too large", smem);
  const bool small = A <= 32 * vec * 2;
  dim3 gb(rows, 1, 1);
  if (precision == CAPDEC_BF16) {
    if (small) return launch_attn(attn_bwd_b_kernel<bf16, 2>, gb, NTB, smem, st, ba);
    return launch_attn(attn_bwd_b_kernel<bf16, 4>, gb, NTB, smem, st, ba);
  }
  if (small) return launch_attn(attn_bwd_b_kernel<float, 2>, gb, NTB, smem, st, ba);
  return launch_attn(attn_bwd_b_kernel<float, 4>, gb, NTB, smem, st, ba);
}

int attention_datt1(int precision, const void* att1, const float* g1, int64_t ldg, int64_t g1_step,
                    const float* de, int64_t de_step, const float* w_f, float* dAtt1, int accumulate,
                    int B, int T, int P, int A, cudaStream_t st) {
  if (B <= 0 || T <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(A <= 4 * NT, CAPDEC_ERR_BAD_SHAPE, "attention dAtt1: A=%d too large", A);
  Datt1Args a{att1, g1, ldg, g1_step, de, de_step, w_f, dAtt1, accumulate, T, P, A};
  const size_t smem = (size_t)T * PT * sizeof(float);
  dim3 grid(ceil_div(P, PT), B, 1);
  const int na = ceil_div(A, NT);
  if (precision == CAPDEC_BF16) {
    if (na <= 1) attn_datt1_kernel<bf16, 1><<<grid, NT, smem, st>>>(a);
    else if (na <= 2) attn_datt1_kernel<bf16, 2><<<grid, NT, smem, st>>>(a);
    else attn_datt1_kernel<bf16, 4><<<grid, NT, smem, st>>>(a);
  } else {
    if (na <= 1) attn_datt1_kernel<float, 1><<<grid, NT, smem, st>>>(a);
    else if (na <= 2) attn_datt1_kernel<float, 2><<<grid, NT, smem, st>>>(a);
    else attn_datt1_kernel<float, 4><<<grid, NT, smem, st>>>(a);
  }
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace capdec
