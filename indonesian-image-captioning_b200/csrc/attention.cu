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
// This is the HBM/L2-bandwidth kernel of the decoder: every row streams att1[b] (P*A) and
// enc[b] (P*E) once per step.  One thread-block CLUSTER handles one row: the CL CTAs of
// the cluster split the P pixels for the score phase and the E channels for the weighted
// sum; the 196 scores are exchanged through distributed shared memory (push model, one
// cluster barrier), so a row's features are read exactly once while B*CL CTAs (>= 2 per SM
// at B = 32) keep enough 16-byte loads in flight to cover the L2/HBM latency.  All global
// feature loads are coalesced and L1-bypassing; reductions over the attention dim and the
// softmax use warp shuffles.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace capdec {

namespace {

constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int CL_MAX = 8;
constexpr int PXB = 4;        // pixels a warp keeps in flight per iteration

struct FwdArgs {
  const void* att1; const void* enc; const float* g1; int64_t ldg; int beta_col;
  const float* w_f; const float* b_f; float* alpha_out; int64_t alpha_stride;
  void* z_out; int64_t ldz; float* awe_out; int rows, rows_per_map, P, E, A;
};

// NCH = 16-byte chunks per lane along the attention dim: A <= 32 * VEC * NCH
template <typename FT, int NCH>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(FwdArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  extern __shared__ __align__(16) float smem_f[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int row = blockIdx.x / CL;
  const int map = row / a.rows_per_map;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = a.P, E = a.E, A = a.A;
  const int Ppad = (P + 3) & ~3;
  float* sc = smem_f;              // [Ppad] scores (all pixels, after the exchange)
  float* al = sc + Ppad;           // [Ppad] alpha
  float* red = al + Ppad;          // [NTHREADS * VEC] cross-group reduction

  pdl_launch_dependents();
  if (CL > 1) cluster.barrier_arrive();     // "everyone has started" barrier, waited on before the push
  pdl_wait();                               // g1 / features come from the previous kernels of the stream

  const FT* att1 = (const FT*)a.att1 + (int64_t)map * P * A;
  const FT* enc = (const FT*)a.enc + (int64_t)map * P * E;
  const float* g1 = a.g1 + (int64_t)row * a.ldg;

  // ---------------- phase 1: scores for this CTA's pixel slice ----------------
  const int Pc = (P + CL - 1) / CL;
  const int p_begin = rank * Pc;
  const int p_end = min(P, p_begin + Pc);
  float att2[NCH][VEC], wf[NCH][VEC];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      att2[c][v] = (a0 + v < A) ? g1[a0 + v] : 0.f;
      wf[c][v] = (a0 + v < A) ? a.w_f[a0 + v] : 0.f;
    }
  }
  const float bf = a.b_f[0];
  for (int p = p_begin + warp; p < p_end; p += PXB * NWARPS) {
    uint4 v[PXB][NCH];
#pragma unroll
    for (int i = 0; i < PXB; ++i) {
      const int pi = p + i * NWARPS;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int a0 = (c * 32 + lane) * VEC;
        v[i][c] = make_uint4(0, 0, 0, 0);
        if (a0 < A && pi < p_end) v[i][c] = ld_stream16(att1 + (int64_t)pi * A + a0);
      }
    }
#pragma unroll
    for (int i = 0; i < PXB; ++i) {
      const int pi = p + i * NWARPS;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float f[VEC];
        unpack16(v[i][c], f, FT());
#pragma unroll
        for (int k = 0; k < VEC; ++k) s = fmaf(wf[c][k], fmaxf(f[k] + att2[c][k], 0.f), s);
      }
      s = warp_sum(s);
      if (lane == 0 && pi < p_end) sc[pi] = s + bf;
    }
  }
  __syncthreads();
  if (CL > 1) {
    cluster.barrier_wait();                 // all CTAs of the cluster are running: DSMEM is valid
    const int n_own = p_end - p_begin;
    for (int i = tid; i < n_own * (CL - 1); i += NTHREADS) {
      const int peer = (rank + 1 + i / n_own) % CL;
      const int p = p_begin + i % n_own;
      cluster.map_shared_rank(sc, peer)[p] = sc[p];
    }
    cluster.sync();                         // release/acquire: every CTA now holds all P scores
  }

  // ---------------- phase 2: softmax over the P pixels (warp 0) ----------------
  if (warp == 0) {
    float m = -INFINITY;
    for (int p = lane; p < P; p += 32) m = fmaxf(m, sc[p]);
    m = warp_max(m);
    float s = 0.f;
    for (int p = lane; p < P; p += 32) {
      const float e = expf(sc[p] - m);
      al[p] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int p = lane; p < P; p += 32) {
      const float v = al[p] * inv;
      al[p] = v;
      if (rank == 0 && a.alpha_out) a.alpha_out[(int64_t)row * a.alpha_stride + p] = v;
    }
  }
  __syncthreads();

  // ---------------- phase 3: awe over this CTA's channel slice ----------------
  const int Ec = E / CL;
  const int e_begin = rank * Ec;
  const int ncol = Ec / VEC;
  const int ncolPass = min(ncol, NTHREADS);
  const int groups = NTHREADS / ncolPass;
  const int grp = tid / ncolPass;
  const int cip = tid % ncolPass;
  constexpr int UNR = 8;
  for (int cb = 0; cb < ncol; cb += ncolPass) {
    const int col = cb + cip;
    const bool active = grp < groups && col < ncol;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    if (active) {
      const FT* src = enc + e_begin + col * VEC;
      int p = grp;
      for (; p + (UNR - 1) * groups < P; p += UNR * groups) {
        uint4 q[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) q[u] = ld_stream16(src + (int64_t)(p + u * groups) * E);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const float w = al[p + u * groups];
          float f[VEC];
          unpack16(q[u], f, FT());
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fmaf(w, f[v], acc[v]);
        }
      }
      for (; p < P; p += groups) {
        uint4 q0 = ld_stream16(src + (int64_t)p * E);
        const float w0 = al[p];
        float f[VEC];
        unpack16(q0, f, FT());
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(w0, f[v], acc[v]);
      }
    }
    if (groups > 1) {
      __syncthreads();
      if (active && grp > 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) red[((grp - 1) * ncolPass + cip) * VEC + v] = acc[v];
      }
      __syncthreads();
      if (active && grp == 0) {
        for (int g = 1; g < groups; ++g)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += red[((g - 1) * ncolPass + cip) * VEC + v];
      }
    }
    if (active && grp == 0) {
      const int e0 = e_begin + col * VEC;
      float zv[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float gate = 1.0f;
        if (a.beta_col >= 0) gate = sigmoidf_(g1[a.beta_col + e0 + v]);
        zv[v] = gate * acc[v];
      }
      if (a.awe_out) {
        float* dst = a.awe_out + (int64_t)row * E + e0;
#pragma unroll
        for (int v = 0; v < VEC; v += 4)
          *reinterpret_cast<float4*>(dst + v) = make_float4(acc[v], acc[v + 1], acc[v + 2], acc[v + 3]);
      }
      if (a.z_out) {
        FT* dst = (FT*)a.z_out + (int64_t)row * a.ldz + e0;
        *reinterpret_cast<uint4*>(dst) = pack16(zv, FT());
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward (attention part)
//   dgate = dz*awe ; dawe = dz*gate ; dbeta_pre = dgate*gate*(1-gate)
//   dalpha_p = enc[p,:].dawe + dalpha_ext_p
//   de_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q)
//   drelu[p,a] = de_p w_f[a] 1[att1[p,a]+att2[a] > 0]
//   datt2[a] = sum_p drelu[p,a] ; dAtt1[p,a] += drelu[p,a] (accumulated over time)
//   dw_f[a] += sum_p de_p relu(.)[p,a] ; db_f += sum_p de_p
// ---------------------------------------------------------------------------
struct BwdArgs {
  const void* att1; const void* enc; const float* g1; int64_t ldg; int beta_col;
  const float* w_f; const float* alpha; int64_t alpha_stride;
  const float* dalpha_ext; int64_t dalpha_stride;
  const float* dz; int64_t lddz; const float* awe;
  void* dba; int64_t lddba;           // feature type: [dbeta_pre (E) | datt2 (A)]
  float* dAtt1;                       // [rows][P][A] fp32, accumulated
  float* dwf_part; float* dbf_part;   // [rows][A], [rows]
  int rows, P, E, A;
};

// NCE = 16-byte chunks per lane along the CTA's channel slice: E/CL <= 32 * VEC * NCE
template <typename FT, int NCH, int NCE>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_kernel(BwdArgs a) {
  constexpr int VEC = FTraits<FT>::VEC;
  extern __shared__ __align__(16) float smem_f[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int row = blockIdx.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = a.P, E = a.E, A = a.A;
  const int Ppad = (P + 3) & ~3;
  const int Apad = (A + 3) & ~3;
  float* part = smem_f;                       // [CL_MAX][Ppad] partial dalpha from each CTA
  float* al = part + CL_MAX * Ppad;           // [Ppad] alpha
  float* de = al + Ppad;                      // [Ppad]
  float* redA = de + Ppad;                    // [NWARPS][2][Apad]
  float* partA = redA + NWARPS * 2 * Apad;    // [CL_MAX][2][Apad]  (used on rank 0)

  pdl_launch_dependents();
  if (CL > 1) cluster.barrier_arrive();
  pdl_wait();

  const FT* att1 = (const FT*)a.att1 + (int64_t)row * P * A;
  const FT* enc = (const FT*)a.enc + (int64_t)row * P * E;
  const float* g1 = a.g1 + (int64_t)row * a.ldg;
  const float* dz = a.dz + (int64_t)row * a.lddz;
  const float* awe = a.awe + (int64_t)row * E;
  FT* dba = (FT*)a.dba + (int64_t)row * a.lddba;

  // ---- phase A: gate backward on this CTA's channel slice; keep dawe in registers ----
  const int Ec = E / CL;
  const int e_begin = rank * Ec;
  const int ncol = Ec / VEC;                  // host guarantees ncol <= 32*NCE
  float dawe[NCE][VEC];
#pragma unroll
  for (int c = 0; c < NCE; ++c) {
    const int col = c * 32 + lane;
#pragma unroll
    for (int v = 0; v < VEC; ++v) dawe[c][v] = 0.f;
    if (col < ncol) {
      const int e0 = e_begin + col * VEC;
      float db[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float gate = sigmoidf_(g1[a.beta_col + e0 + v]);
        const float d = dz[e0 + v];
        dawe[c][v] = d * gate;
        db[v] = d * awe[e0 + v] * gate * (1.0f - gate);
      }
      if (warp == 0) *reinterpret_cast<uint4*>(dba + e0) = pack16(db, FT());
    }
  }
  for (int p = tid; p < P; p += NTHREADS) al[p] = a.alpha[(int64_t)row * a.alpha_stride + p];

  // ---- phase B: partial dalpha_p = enc[p, slice] . dawe[slice]  (warp per pixel) ----
  float* my_part = part + rank * Ppad;
  for (int p = warp; p < P; p += PXB * NWARPS) {
    uint4 v[PXB][NCE];
#pragma unroll
    for (int i = 0; i < PXB; ++i) {
      const int pi = p + i * NWARPS;
#pragma unroll
      for (int c = 0; c < NCE; ++c) {
        const int col = c * 32 + lane;
        v[i][c] = make_uint4(0, 0, 0, 0);
        if (col < ncol && pi < P) v[i][c] = ld_stream16(enc + (int64_t)pi * E + e_begin + col * VEC);
      }
    }
#pragma unroll
    for (int i = 0; i < PXB; ++i) {
      const int pi = p + i * NWARPS;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NCE; ++c) {
        float f[VEC];
        unpack16(v[i][c], f, FT());
#pragma unroll
        for (int k = 0; k < VEC; ++k) s = fmaf(f[k], dawe[c][k], s);
      }
      s = warp_sum(s);
      if (lane == 0 && pi < P) my_part[pi] = s;
    }
  }
  __syncthreads();
  if (CL > 1) {
    cluster.barrier_wait();
    for (int i = tid; i < P * (CL - 1); i += NTHREADS) {
      const int peer = (rank + 1 + i / P) % CL;
      const int p = i % P;
      cluster.map_shared_rank(part, peer)[rank * Ppad + p] = my_part[p];
    }
    cluster.sync();
  }

  // ---- phase C: softmax backward (warp 0), de_p for all pixels ----
  if (warp == 0) {
    float dot = 0.f;
    for (int p = lane; p < P; p += 32) {
      float d = 0.f;
      for (int r = 0; r < CL; ++r) d += part[r * Ppad + p];
      if (a.dalpha_ext) d += a.dalpha_ext[(int64_t)row * a.dalpha_stride + p];
      de[p] = d;
      dot = fmaf(al[p], d, dot);
    }
    dot = warp_sum(dot);
    float sde = 0.f;
    for (int p = lane; p < P; p += 32) {
      const float v = al[p] * (de[p] - dot);
      de[p] = v;
      sde += v;
    }
    sde = warp_sum(sde);
    if (lane == 0 && rank == 0) a.dbf_part[row] = sde;
  }
  __syncthreads();

  // ---- phase D: relu/score backward on this CTA's pixel slice ----
  const int Pc = (P + CL - 1) / CL;
  const int p_begin = rank * Pc;
  const int p_end = min(P, p_begin + Pc);
  float att2[NCH][VEC], wf[NCH][VEC], dacc[NCH][VEC], wacc[NCH][VEC];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      att2[c][v] = (a0 + v < A) ? g1[a0 + v] : 0.f;
      wf[c][v] = (a0 + v < A) ? a.w_f[a0 + v] : 0.f;
      dacc[c][v] = 0.f;
      wacc[c][v] = 0.f;
    }
  }
  float* dA = a.dAtt1 + (int64_t)row * P * A;
  for (int p = p_begin + warp; p < p_end; p += 2 * NWARPS) {
    const int p2 = p + NWARPS;
    const bool has2 = p2 < p_end;
    uint4 q1[NCH], q2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int a0 = (c * 32 + lane) * VEC;
      q1[c] = make_uint4(0, 0, 0, 0);
      q2[c] = make_uint4(0, 0, 0, 0);
      if (a0 < A) {
        q1[c] = ld_stream16(att1 + (int64_t)p * A + a0);
        if (has2) q2[c] = ld_stream16(att1 + (int64_t)p2 * A + a0);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i == 1 && !has2) break;
      const int pi = i == 0 ? p : p2;
      const float dep = de[pi];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int a0 = (c * 32 + lane) * VEC;
        if (a0 < A) {
          float f[VEC], dr[VEC];
          unpack16(i == 0 ? q1[c] : q2[c], f, FT());
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float pre = f[v] + att2[c][v];
            const float on = pre > 0.f ? 1.f : 0.f;
            dr[v] = dep * wf[c][v] * on;
            dacc[c][v] += dr[v];
            wacc[c][v] = fmaf(dep, pre * on, wacc[c][v]);
          }
          float* d = dA + (int64_t)pi * A + a0;
#pragma unroll
          for (int v = 0; v < VEC; v += 4) {
            float4 o = *reinterpret_cast<float4*>(d + v);
            o.x += dr[v]; o.y += dr[v + 1]; o.z += dr[v + 2]; o.w += dr[v + 3];
            *reinterpret_cast<float4*>(d + v) = o;
          }
        }
      }
    }
  }
  // cross-warp reduction of datt2 / dw_f partials
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int a0 = (c * 32 + lane) * VEC;
    if (a0 < A) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        redA[(warp * 2 + 0) * Apad + a0 + v] = dacc[c][v];
        redA[(warp * 2 + 1) * Apad + a0 + v] = wacc[c][v];
      }
    }
  }
  __syncthreads();
  float* dstA = (CL > 1) ? cluster.map_shared_rank(partA, 0) : partA;
  for (int i = tid; i < 2 * A; i += NTHREADS) {
    const int which = i / A, aa = i % A;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s += redA[(w * 2 + which) * Apad + aa];
    dstA[(rank * 2 + which) * Apad + aa] = s;
  }
  if (CL > 1) cluster.sync(); else __syncthreads();
  if (rank == 0) {
    for (int aa = tid; aa < A; aa += NTHREADS) {
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < CL; ++r) {
        s0 += partA[(r * 2 + 0) * Apad + aa];
        s1 += partA[(r * 2 + 1) * Apad + aa];
      }
      dba[E + aa] = from_f<FT>(s0);
      a.dwf_part[(int64_t)row * A + aa] = s1;
    }
  }
}

size_t fwd_smem(int P) { return (size_t)(2 * ((P + 3) & ~3) + NTHREADS * 8) * sizeof(float); }
size_t bwd_smem(int P, int A) {
  const int Ppad = (P + 3) & ~3, Apad = (A + 3) & ~3;
  return (size_t)(CL_MAX * Ppad + 2 * Ppad + NWARPS * 2 * Apad + CL_MAX * 2 * Apad) * sizeof(float);
}

template <typename ArgsT>
int launch_cluster(void (*kernel)(ArgsT), const ArgsT& args, int rows, int CL, size_t smem, cudaStream_t st) {
  CAPDEC_CUDA_OK(launch_pdl(kernel, dim3(rows * CL, 1, 1), dim3(NTHREADS, 1, 1), smem, st, CL, args));
  count_launch();
  return CAPDEC_OK;
}

int env_cluster() {
  static int v = [] {
    const char* s = getenv("CAPDEC_ATTN_CLUSTER");
    return s ? atoi(s) : 0;
  }();
  return v;
}

// cluster size: >= 2 CTAs per SM over the 148 SMs when the batch is small, bounded by the
// divisibility of the channel slice; CAPDEC_ATTN_CLUSTER overrides (tuning sweeps)
int pick_cluster(int rows, int E, int vec, int min_cl) {
  int cl = min_cl;
  const int want = env_cluster();
  if (want > 0) {
    while (cl < want && cl < CL_MAX && (E % (2 * cl * vec)) == 0) cl *= 2;
    return cl;
  }
  while (cl < CL_MAX && rows * cl < 2 * 148 - 40 && (E % (2 * cl * vec)) == 0) cl *= 2;
  return cl;
}

template <typename FT, int NCH>
int launch_bwd(const BwdArgs& a, int CL, int nce, size_t smem, cudaStream_t st) {
  if (nce <= 1) return launch_cluster(attn_bwd_kernel<FT, NCH, 1>, a, a.rows, CL, smem, st);
  if (nce <= 2) return launch_cluster(attn_bwd_kernel<FT, NCH, 2>, a, a.rows, CL, smem, st);
  return launch_cluster(attn_bwd_kernel<FT, NCH, 4>, a, a.rows, CL, smem, st);
}

template <typename FT, int NCH>
int set_bwd_attr() {
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<FT, NCH, 1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<FT, NCH, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CAPDEC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<FT, NCH, 4>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  return CAPDEC_OK;
}

}  // namespace

int attention_init() {
  CAPDEC_TRY((set_bwd_attr<float, 2>()));
  CAPDEC_TRY((set_bwd_attr<float, 4>()));
  CAPDEC_TRY((set_bwd_attr<bf16, 2>()));
  CAPDEC_TRY((set_bwd_attr<bf16, 4>()));
  return CAPDEC_OK;
}

int attention_fwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* b_f, float* alpha_out,
                  int64_t alpha_stride, void* z_out, int64_t ldz, float* awe_out, int rows,
                  int rows_per_map, int P, int E, int A, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  const int vec = precision == CAPDEC_BF16 ? 8 : 4;
  CAPDEC_REQUIRE(E % vec == 0 && A % vec == 0 && A <= 32 * vec * 4, CAPDEC_ERR_BAD_SHAPE,
                 "attention: need E,A multiples of %d and A <= %d (E=%d A=%d)", vec, 32 * vec * 4, E, A);
  FwdArgs a{att1, enc, g1, ldg, beta_col, w_f, b_f, alpha_out, alpha_stride, z_out, ldz, awe_out,
            rows, rows_per_map, P, E, A};
  const int CL = pick_cluster(rows, E, vec, 1);
  const bool small = A <= 32 * vec * 2;
  if (precision == CAPDEC_BF16) {
    if (small) return launch_cluster(attn_fwd_kernel<bf16, 2>, a, rows, CL, fwd_smem(P), st);
    return launch_cluster(attn_fwd_kernel<bf16, 4>, a, rows, CL, fwd_smem(P), st);
  }
  if (small) return launch_cluster(attn_fwd_kernel<float, 2>, a, rows, CL, fwd_smem(P), st);
  return launch_cluster(attn_fwd_kernel<float, 4>, a, rows, CL, fwd_smem(P), st);
}

int attention_bwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* alpha, int64_t alpha_stride,
                  const float* dalpha_ext, int64_t dalpha_stride, const float* dz, int64_t lddz,
                  const float* awe, void* dba, int64_t lddba, float* dAtt1, float* dwf_part,
                  float* dbf_part, int rows, int P, int E, int A, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  const int vec = precision == CAPDEC_BF16 ? 8 : 4;
  CAPDEC_REQUIRE(E % vec == 0 && A % vec == 0 && A <= 32 * vec * 4 && beta_col >= 0,
                 CAPDEC_ERR_BAD_SHAPE, "attention bwd: unsupported dims E=%d A=%d", E, A);
  // the E slice of one CTA must fit the per-lane register cache: E/CL/vec <= 32*4
  int min_cl = 1;
  while (min_cl < CL_MAX && (E / min_cl) > 32 * 4 * vec) min_cl *= 2;
  CAPDEC_REQUIRE((E / min_cl) <= 32 * 4 * vec && E % (min_cl * vec) == 0, CAPDEC_ERR_BAD_SHAPE,
                 "attention bwd: E=%d too large / not divisible", E);
  BwdArgs a{att1, enc, g1, ldg, beta_col, w_f, alpha, alpha_stride, dalpha_ext, dalpha_stride,
            dz, lddz, awe, dba, lddba, dAtt1, dwf_part, dbf_part, rows, P, E, A};
  const int CL = pick_cluster(rows, E, vec, min_cl);
  const int nce = ceil_div(E / CL / vec, 32);
  const size_t smem = bwd_smem(P, A);
  CAPDEC_REQUIRE(smem <= 160 * 1024, CAPDEC_ERR_BAD_SHAPE, "attention bwd: smem %zu too large", smem);
  const bool small = A <= 32 * vec * 2;
  if (precision == CAPDEC_BF16) {
    if (small) return launch_bwd<bf16, 2>(a, CL, nce, smem, st);
    return launch_bwd<bf16, 4>(a, CL, nce, smem, st);
  }
  if (small) return launch_bwd<float, 2>(a, CL, nce, smem, st);
  return launch_bwd<float, 4>(a, CL, nce, smem, st);
}

}  // namespace capdec
