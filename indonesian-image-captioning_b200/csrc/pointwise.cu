// pointwise.cu -- layout/packing, gather/scatter and LSTM/SCN pointwise kernels.
//
// Reference math (paths relative to the reference root):
//   cell_fwd / cell_bwd       models/scn_cell.py:146-152 (gate order i,f,o,c) and
//                             torch.nn.LSTMCell as used by pure_attention.py:143-146 (i,f,g,o)
//   scn_form_m                models/scn_cell.py:83-86, 134-143: (x W_ia) * (s W_ib), (h W_ha) * (s W_hb)
//   gather_features           attention_scn.py:113-120 (view, sort permutation) + :90 (pixel mean)
//   embedding_gather          attention_scn.py:124
#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return to_f(*p); }

// ------------------------------------------------------------------------------------
// transpose + cast through a 32x33 shared tile
// ------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void transpose_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int ni, int nj,
                                 int C, int64_t s_i, int64_t s_j, int64_t ldd, int64_t d_i,
                                 int64_t d_j) {
  __shared__ float tile[32][33];
  const int R = ni * nj;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      const int i = r / nj, j = r - i * nj;
      v = ldf(src + (int64_t)i * s_i + (int64_t)j * s_j + c);
    }
    tile[k][threadIdx.x] = v;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + threadIdx.x;
    if (r < R && c < C) {
      const int i = r / nj, j = r - i * nj;
      dst[(int64_t)c * ldd + (int64_t)i * d_i + (int64_t)j * d_j] = from_f<TD>(tile[threadIdx.x][k]);
    }
  }
}

template <typename TS, typename TD>
__global__ void copy_cast_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst,
                                 int64_t ldd, int R, int C) {
  const int64_t total = (int64_t)R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i - (int64_t)r * C);
    dst[(int64_t)r * ldd + c] = from_f<TD>(ldf(src + (int64_t)r * lds + c));
  }
}

// All weight packing of one call in ONE launch: a table of (fp32 source matrix -> feature-type copy or
// transpose) segments, 64 x 64 source tiles, one tile per CTA.
template <typename TD>
__global__ void pack_multi_kernel(const __grid_constant__ PackTable t) {
  constexpr int TS = 64;                      // source tile edge: 64 x 64 elements per CTA, 16 per thread
  __shared__ float tile[TS][TS + 1];
  int si = 0;
  while (si + 1 < t.n && (int)blockIdx.x >= t.seg[si + 1].tile0) ++si;
  const PackSeg& g = t.seg[si];
  const int local = blockIdx.x - g.tile0;
  const int tiles_c = (g.C + TS - 1) / TS;
  const int r0 = (local / tiles_c) * TS, c0 = (local % tiles_c) * TS;
  const float* src = g.src;
  TD* dst = (TD*)g.dst;
  float v[TS / 8][2];
#pragma unroll
  for (int i = 0; i < TS / 8; ++i) {
    const int r = r0 + threadIdx.y + 8 * i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + threadIdx.x + 32 * h;
      v[i][h] = (r < g.R && c < g.C) ? src[(int64_t)r * g.lds + c] : 0.f;
    }
  }
  if (!g.transpose) {
#pragma unroll
    for (int i = 0; i < TS / 8; ++i) {
      const int r = r0 + threadIdx.y + 8 * i;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + threadIdx.x + 32 * h;
        if (r < g.R && c < g.C) dst[(int64_t)r * g.ldd + c] = from_f<TD>(v[i][h]);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TS / 8; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) tile[threadIdx.y + 8 * i][threadIdx.x + 32 * h] = v[i][h];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TS / 8; ++i) {
    const int c = c0 + threadIdx.y + 8 * i;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + threadIdx.x + 32 * h;
      if (r < g.R && c < g.C) dst[(int64_t)c * g.ldd + r] = from_f<TD>(tile[threadIdx.x + 32 * h][threadIdx.y + 8 * i]);
    }
  }
}

// block (32, RY): thread (x, y) sums rows y, y + RY, ... of column n; four rows in flight per thread.
// The partial sums meet in a FIXED order (deterministic: the fp32 parity mode is bit-reproducible).
template <typename T, int RY>
__global__ void colsum_kernel(const T* __restrict__ X, int64_t ld, int R, int N, float* out,
                              int accumulate) {
  __shared__ float red[RY][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < N) {
    int r = threadIdx.y;
    for (; r + 3 * RY < R; r += 4 * RY) {
      s0 += ldf(X + (int64_t)r * ld + n);
      s1 += ldf(X + (int64_t)(r + RY) * ld + n);
      s2 += ldf(X + (int64_t)(r + 2 * RY) * ld + n);
      s3 += ldf(X + (int64_t)(r + 3 * RY) * ld + n);
    }
    for (; r < R; r += RY) s0 += ldf(X + (int64_t)r * ld + n);
  }
  red[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < RY; ++k) t += red[k][threadIdx.x];
    out[n] = accumulate ? out[n] + t : t;
  }
}

template <typename FT>
__global__ void gather_features_kernel(const float* __restrict__ enc, int64_t sb, int64_t sp,
                                       int64_t se, const int64_t* __restrict__ sort_ind,
                                       FT* __restrict__ enc_s, float* __restrict__ mean_f32,
                                       FT* __restrict__ mean_ft, int64_t ld_mean_ft, int B, int P,
                                       int E, FT* __restrict__ enc_cm, int cw) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t src_b = sort_ind ? sort_ind[b] : b;
  const float* src = enc + src_b * sb + (int64_t)e * se;
  FT* dst = enc_s + (int64_t)b * P * E + e;
  // optional second copy, chunk-major [B][E/cw][P][cw]: every (pixel range, cw-channel chunk) run is contiguous, so a
  // staging fill of the persistent / streaming attention kernels is ONE bulk copy (was a separate pass over enc_s)
  FT* dcm = enc_cm ? enc_cm + (((int64_t)b * (E / cw) + e / cw) * P) * cw + (e % cw) : nullptr;
  float s = 0.f;
  int p = 0;
  constexpr int U = 14;                         // pixels in flight per thread (196 = 14 * 14)
  for (; p + U <= P; p += U) {
    float v[U];
#pragma unroll
    for (int i = 0; i < U; ++i) v[i] = src[(int64_t)(p + i) * sp];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const FT x = from_f<FT>(v[i]);
      dst[(int64_t)(p + i) * E] = x;
      if (dcm) dcm[(int64_t)(p + i) * cw] = x;
      s += v[i];
    }
  }
  for (; p < P; ++p) {
    const float v0 = src[(int64_t)p * sp];
    const FT x = from_f<FT>(v0);
    dst[(int64_t)p * E] = x;
    if (dcm) dcm[(int64_t)p * cw] = x;
    s += v0;
  }
  const float m = s / (float)P;
  if (mean_f32) mean_f32[(int64_t)b * E + e] = m;
  if (mean_ft) mean_ft[(int64_t)b * ld_mean_ft + e] = from_f<FT>(m);
}

template <typename FT>
__global__ void embedding_gather_kernel(const float* __restrict__ emb,
                                        const int64_t* __restrict__ caps, int L, FT* __restrict__ Xe,
                                        int64_t ldx, int B, int T, int M, int V) {
  const int r = blockIdx.x;          // r = t*B + b
  const int t = r / B, b = r - t * B;
  int64_t w = caps[(int64_t)b * L + t];
  if (w < 0) w = 0;
  if (w >= V) w = V - 1;
  const float* src = emb + w * M;
  FT* dst = Xe + (int64_t)r * ldx;
  for (int i = threadIdx.x; i < M; i += blockDim.x) dst[i] = from_f<FT>(src[i]);
}

// Deterministic scatter: the block of row r = (t,b) is the LEADER of its token if no earlier valid row
// carries the same token; the leader sums every row of that token in row order and stores the result
// (no atomics -> bit-reproducible gradients, which the CUDA-graph replay test relies on).
__global__ void embedding_scatter_add_kernel(const float* __restrict__ dXe, int64_t ldx,
                                             const int64_t* __restrict__ caps, int L,
                                             const int32_t* __restrict__ len_d, float* dEmb, int B,
                                             int T, int M, int V) {
  extern __shared__ int s_rows[];              // rows that carry this block's token (ascending)
  __shared__ int s_n, s_dup;
  const int r = blockIdx.x;
  const int t = r / B, b = r - t * B;
  if (t >= len_d[b]) return;
  const int64_t w = caps[(int64_t)b * L + t];
  if (w < 0 || w >= V) return;
  if (threadIdx.x == 0) { s_n = 0; s_dup = 0; }
  __syncthreads();
  for (int r2 = threadIdx.x; r2 < r; r2 += blockDim.x) {
    const int t2 = r2 / B, b2 = r2 - t2 * B;
    if (t2 < len_d[b2] && caps[(int64_t)b2 * L + t2] == w) s_dup = 1;
  }
  __syncthreads();
  if (s_dup) return;
  // leader: collect the later rows with the same token, in ascending order (chunked ballot-free scan)
  const int R = B * T;
  for (int base = r; base < R; base += blockDim.x) {
    const int r2 = base + threadIdx.x;
    bool hit = false;
    if (r2 < R) {
      const int t2 = r2 / B, b2 = r2 - t2 * B;
      hit = t2 < len_d[b2] && caps[(int64_t)b2 * L + t2] == w;
    }
    // ordered compaction inside the chunk: thread i appends after all hits of threads < i
    __shared__ int s_flag[128];
    s_flag[threadIdx.x] = hit ? 1 : 0;
    __syncthreads();
    if (hit) {
      int pos = s_n;
      for (int i = 0; i < (int)threadIdx.x; ++i) pos += s_flag[i];
      s_rows[pos] = r2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int c = 0;
      for (int i = 0; i < (int)blockDim.x; ++i) c += s_flag[i];
      s_n += c;
    }
    __syncthreads();
  }
  const int n = s_n;
  float* dst = dEmb + w * M;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += dXe[(int64_t)s_rows[j] * ldx + i];
    dst[i] = acc;
  }
}

template <typename FT>
__global__ void expand_rows_kernel(const FT* __restrict__ src, int64_t lds, FT* __restrict__ dst,
                                   int64_t ldd, int k, int C) {
  const int r = blockIdx.x;
  const FT* s = src + (int64_t)(r / k) * lds;
  FT* d = dst + (int64_t)r * ldd;
  for (int i = threadIdx.x; i < C; i += blockDim.x) d[i] = s[i];
}

template <typename FT>
__global__ void scn_form_m_kernel(const float* __restrict__ u, int64_t ldu,
                                  const float* __restrict__ p, int64_t ldp,
                                  const float* __restrict__ v, const float* __restrict__ q,
                                  FT* __restrict__ m, int rows, int B, int F) {
  // B = rows per gate block of m: m[g][B][2F]
  pdl_prologue();
  const int64_t total = (int64_t)rows * 4 * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / (4 * F));
    const int n = (int)(i - (int64_t)b * 4 * F);     // n = g*F + f
    const int g = n / F, f = n - g * F;
    const float uv = u[(int64_t)b * ldu + n] * v[(int64_t)b * 4 * F + n];
    const float pq = p[(int64_t)b * ldp + n] * q[(int64_t)b * 4 * F + n];
    FT* dst = m + ((int64_t)g * B + b) * 2 * F;
    dst[f] = from_f<FT>(uv);
    dst[F + f] = from_f<FT>(pq);
  }
}

// gate slots in the pre-activation buffer for the two orders
__device__ __forceinline__ void gate_slots(int lstm_order, int& si, int& sf, int& so, int& sg) {
  si = 0; sf = 1;
  if (lstm_order) { sg = 2; so = 3; } else { so = 2; sg = 3; }
}

template <typename FT>
__global__ void cell_fwd_kernel(const float* __restrict__ preA, int64_t ldA,
                                const float* __restrict__ preB, int64_t ldB,
                                const float* __restrict__ b1, const float* __restrict__ b2,
                                int lstm_order, const float* __restrict__ c_prev,
                                float* __restrict__ c_new, float* __restrict__ gates,
                                FT* __restrict__ h_out, int64_t ldh, FT* __restrict__ hd_out,
                                float dropout_p, const uint64_t* __restrict__ seed_dev, int t, int T, int rows, int D) {
  pdl_prologue();
  int si, sf, so, sg;
  gate_slots(lstm_order, si, sf, so, sg);
  const int64_t total = (int64_t)rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / D), d = (int)(i - (int64_t)b * D);
    float pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x = preA[(int64_t)b * ldA + g * D + d];
      if (preB) x += preB[(int64_t)b * ldB + g * D + d];
      if (b1) x += b1[g * D + d];
      if (b2) x += b2[g * D + d];
      pre[g] = x;
    }
    const float ig = sigmoidf_(pre[si]), fg = sigmoidf_(pre[sf]), og = sigmoidf_(pre[so]);
    const float gg = tanhf(pre[sg]);
    const float c = fg * c_prev[i] + ig * gg;
    const float h = og * tanhf(c);
    c_new[i] = c;
    if (gates) {
      float* gp = gates + (int64_t)b * 4 * D + d;       // stored as [i | f | o | g~]
      gp[0] = ig; gp[D] = fg; gp[2 * D] = og; gp[3 * D] = gg;
    }
    h_out[(int64_t)b * ldh + d] = from_f<FT>(h);
    if (hd_out) {
      const float sc = dropout_scale(seed_dev[0], ((uint64_t)b * T + t) * D + d, dropout_p);
      hd_out[(int64_t)b * ldh + d] = from_f<FT>(h * sc);
    }
  }
}

template <typename FT>
__global__ void cell_bwd_kernel(const float* __restrict__ dh_fc, int64_t ld_dhfc,
                                const float* __restrict__ dh_rec, float* __restrict__ dc,
                                const float* __restrict__ gates, const float* __restrict__ c_prev,
                                const float* __restrict__ c_new, int lstm_order, float dropout_p,
                                const uint64_t* __restrict__ seed_dev, int t, int T, FT* __restrict__ dpre,
                                float* __restrict__ dpre_f32, int rows, int D) {
  pdl_prologue();
  int si, sf, so, sg;
  gate_slots(lstm_order, si, sf, so, sg);
  const int64_t total = (int64_t)rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / D), d = (int)(i - (int64_t)b * D);
    float dh = dh_rec[i];
    if (dh_fc) {
      float g = dh_fc[(int64_t)b * ld_dhfc + d];
      if (dropout_p > 0.f) g *= dropout_scale(seed_dev[0], ((uint64_t)b * T + t) * D + d, dropout_p);
      dh += g;
    }
    const float* gp = gates + (int64_t)b * 4 * D + d;
    const float ig = gp[0], fg = gp[D], og = gp[2 * D], gg = gp[3 * D];
    const float tc = tanhf(c_new[i]);
    const float dcn = dc[i] + dh * og * (1.f - tc * tc);
    const float dpo = dh * tc * og * (1.f - og);
    const float dpi = dcn * gg * ig * (1.f - ig);
    const float dpf = dcn * c_prev[i] * fg * (1.f - fg);
    const float dpg = dcn * ig * (1.f - gg * gg);
    dc[i] = dcn * fg;
    const int64_t o = (int64_t)b * 4 * D + d;
    dpre[o + (int64_t)si * D] = from_f<FT>(dpi);
    dpre[o + (int64_t)sf * D] = from_f<FT>(dpf);
    dpre[o + (int64_t)so * D] = from_f<FT>(dpo);
    dpre[o + (int64_t)sg * D] = from_f<FT>(dpg);
    if (dpre_f32) {
      dpre_f32[o + (int64_t)si * D] = dpi;
      dpre_f32[o + (int64_t)sf * D] = dpf;
      dpre_f32[o + (int64_t)so * D] = dpo;
      dpre_f32[o + (int64_t)sg * D] = dpg;
    }
  }
}

template <typename FT>
__global__ void scn_bwd_products_kernel(const float* __restrict__ wr, const float* __restrict__ u,
                                        int64_t ldu, const float* __restrict__ p, int64_t ldp,
                                        const float* __restrict__ v, const float* __restrict__ q,
                                        FT* __restrict__ du, FT* __restrict__ dp,
                                        float* __restrict__ dv_acc, float* __restrict__ dq_acc,
                                        int rows, int B, int F, int64_t lddp) {
  pdl_prologue();
  const int64_t total = (int64_t)rows * 4 * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / (4 * F));
    const int n = (int)(i - (int64_t)b * 4 * F);
    const int g = n / F, f = n - g * F;
    const float* src = wr + ((int64_t)g * B + b) * 2 * F;
    const float w = src[f], r = src[F + f];
    const int64_t k = (int64_t)b * 4 * F + n;
    du[k] = from_f<FT>(w * v[k]);
    dp[(int64_t)b * lddp + n] = from_f<FT>(r * q[k]);
    dv_acc[k] += w * u[(int64_t)b * ldu + n];
    dq_acc[k] += r * p[(int64_t)b * ldp + n];
  }
}

// ---- four-features-per-thread versions of the four per-step pointwise kernels (beam search: 1 875 rows per step, the
// per-step chains of the large shapes): 16-byte loads / stores and 32-bit index arithmetic.  The one-element-per-thread
// kernels above ran at a third of their own HBM roofline (scn_form_m 22.7 us for 50 MB, cell_fwd 14.6 us for 25 MB:
// profiles/r1g_decode_launches_summary.txt).  Launchers fall back to them when a row pitch or pointer is not 16-byte
// friendly. ----
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(bf16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<const uint32_t*>(&lo);
  v.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}

template <typename FT>
__global__ void scn_form_m_v4_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ p, int64_t ldp,
                                     const float* __restrict__ v, const float* __restrict__ q, FT* __restrict__ m,
                                     int rows, int B, int F) {
  pdl_prologue();
  const int total = rows * F;                       // groups of 4 features: (row, n / 4), n = g*F + f
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / F, n = (i - b * F) * 4;
    const int g = n / F, f = n - g * F;
    const float4 uu = ld4(u + (int64_t)b * ldu + n), vv = ld4(v + (int64_t)b * 4 * F + n);
    const float4 pp = ld4(p + (int64_t)b * ldp + n), qq = ld4(q + (int64_t)b * 4 * F + n);
    FT* dst = m + ((int64_t)g * B + b) * 2 * F + f;
    st4(dst, uu.x * vv.x, uu.y * vv.y, uu.z * vv.z, uu.w * vv.w);
    st4(dst + F, pp.x * qq.x, pp.y * qq.y, pp.z * qq.z, pp.w * qq.w);
  }
}

template <typename FT>
__global__ void cell_fwd_v4_kernel(const float* __restrict__ preA, int64_t ldA, const float* __restrict__ preB, int64_t ldB,
                                   const float* __restrict__ b1, const float* __restrict__ b2, int lstm_order,
                                   const float* __restrict__ c_prev, float* __restrict__ c_new, float* __restrict__ gates,
                                   FT* __restrict__ h_out, int64_t ldh, FT* __restrict__ hd_out, float dropout_p,
                                   const uint64_t* __restrict__ seed_dev, int t, int T, int rows, int D) {
  pdl_prologue();
  int si, sf, so, sg;
  gate_slots(lstm_order, si, sf, so, sg);
  const int D4 = D / 4, total = rows * D4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / D4, d = (i - b * D4) * 4;
    float pre[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float4 x = ld4(preA + (int64_t)b * ldA + g * D + d);
      if (preB) { const float4 y = ld4(preB + (int64_t)b * ldB + g * D + d); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
      if (b1) { const float4 y = ld4(b1 + g * D + d); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
      if (b2) { const float4 y = ld4(b2 + g * D + d); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
      pre[g][0] = x.x; pre[g][1] = x.y; pre[g][2] = x.z; pre[g][3] = x.w;
    }
    const float4 cp4 = ld4(c_prev + (int64_t)b * D + d);
    const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
    float ig[4], fg[4], og[4], gg[4], c[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ig[k] = sigmoidf_(pre[si][k]); fg[k] = sigmoidf_(pre[sf][k]); og[k] = sigmoidf_(pre[so][k]);
      gg[k] = tanhf(pre[sg][k]);
      c[k] = fg[k] * cp[k] + ig[k] * gg[k];
      h[k] = og[k] * tanhf(c[k]);
    }
    st4(c_new + (int64_t)b * D + d, c[0], c[1], c[2], c[3]);
    if (gates) {
      float* gp = gates + (int64_t)b * 4 * D + d;       // stored as [i | f | o | g~]
      st4(gp, ig[0], ig[1], ig[2], ig[3]);
      st4(gp + D, fg[0], fg[1], fg[2], fg[3]);
      st4(gp + 2 * D, og[0], og[1], og[2], og[3]);
      st4(gp + 3 * D, gg[0], gg[1], gg[2], gg[3]);
    }
    st4(h_out + (int64_t)b * ldh + d, h[0], h[1], h[2], h[3]);
    if (hd_out) {
      const uint64_t seed = seed_dev[0];
      float hs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) hs[k] = h[k] * dropout_scale(seed, ((uint64_t)b * T + t) * D + d + k, dropout_p);
      st4(hd_out + (int64_t)b * ldh + d, hs[0], hs[1], hs[2], hs[3]);
    }
  }
}

template <typename FT>
__global__ void cell_bwd_v4_kernel(const float* __restrict__ dh_fc, int64_t ld_dhfc, const float* __restrict__ dh_rec,
                                   float* __restrict__ dc, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                   const float* __restrict__ c_new, int lstm_order, float dropout_p,
                                   const uint64_t* __restrict__ seed_dev, int t, int T, FT* __restrict__ dpre,
                                   float* __restrict__ dpre_f32, int rows, int D) {
  pdl_prologue();
  int si, sf, so, sg;
  gate_slots(lstm_order, si, sf, so, sg);
  const int D4 = D / 4, total = rows * D4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / D4, d = (i - b * D4) * 4;
    const int64_t e = (int64_t)b * D + d;
    const float4 r4 = ld4(dh_rec + e);
    float dh[4] = {r4.x, r4.y, r4.z, r4.w};
    if (dh_fc) {
      const float4 g4 = ld4(dh_fc + (int64_t)b * ld_dhfc + d);
      float g[4] = {g4.x, g4.y, g4.z, g4.w};
      if (dropout_p > 0.f) {
        const uint64_t seed = seed_dev[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] *= dropout_scale(seed, ((uint64_t)b * T + t) * D + d + k, dropout_p);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) dh[k] += g[k];
    }
    const float* gp = gates + (int64_t)b * 4 * D + d;
    const float4 i4 = ld4(gp), f4 = ld4(gp + D), o4 = ld4(gp + 2 * D), g4 = ld4(gp + 3 * D);
    const float4 cn4 = ld4(c_new + e), cp4 = ld4(c_prev + e), dc4 = ld4(dc + e);
    const float ig[4] = {i4.x, i4.y, i4.z, i4.w}, fg[4] = {f4.x, f4.y, f4.z, f4.w}, og[4] = {o4.x, o4.y, o4.z, o4.w},
                gg[4] = {g4.x, g4.y, g4.z, g4.w}, cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w},
                dcv[4] = {dc4.x, dc4.y, dc4.z, dc4.w};
    float dpi[4], dpf[4], dpo[4], dpg[4], dco[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float tc = tanhf(cn[k]);
      const float dcn = dcv[k] + dh[k] * og[k] * (1.f - tc * tc);
      dpo[k] = dh[k] * tc * og[k] * (1.f - og[k]);
      dpi[k] = dcn * gg[k] * ig[k] * (1.f - ig[k]);
      dpf[k] = dcn * cp[k] * fg[k] * (1.f - fg[k]);
      dpg[k] = dcn * ig[k] * (1.f - gg[k] * gg[k]);
      dco[k] = dcn * fg[k];
    }
    st4(dc + e, dco[0], dco[1], dco[2], dco[3]);
    const int64_t o = (int64_t)b * 4 * D + d;
    st4(dpre + o + (int64_t)si * D, dpi[0], dpi[1], dpi[2], dpi[3]);
    st4(dpre + o + (int64_t)sf * D, dpf[0], dpf[1], dpf[2], dpf[3]);
    st4(dpre + o + (int64_t)so * D, dpo[0], dpo[1], dpo[2], dpo[3]);
    st4(dpre + o + (int64_t)sg * D, dpg[0], dpg[1], dpg[2], dpg[3]);
    if (dpre_f32) {
      st4(dpre_f32 + o + (int64_t)si * D, dpi[0], dpi[1], dpi[2], dpi[3]);
      st4(dpre_f32 + o + (int64_t)sf * D, dpf[0], dpf[1], dpf[2], dpf[3]);
      st4(dpre_f32 + o + (int64_t)so * D, dpo[0], dpo[1], dpo[2], dpo[3]);
      st4(dpre_f32 + o + (int64_t)sg * D, dpg[0], dpg[1], dpg[2], dpg[3]);
    }
  }
}

template <typename FT>
__global__ void scn_bwd_products_v4_kernel(const float* __restrict__ wr, const float* __restrict__ u, int64_t ldu,
                                           const float* __restrict__ p, int64_t ldp, const float* __restrict__ v,
                                           const float* __restrict__ q, FT* __restrict__ du, FT* __restrict__ dp,
                                           float* __restrict__ dv_acc, float* __restrict__ dq_acc, int rows, int B, int F,
                                           int64_t lddp) {
  pdl_prologue();
  const int total = rows * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / F, n = (i - b * F) * 4;
    const int g = n / F, f = n - g * F;
    const float* src = wr + ((int64_t)g * B + b) * 2 * F + f;
    const float4 w = ld4(src), r = ld4(src + F);
    const int64_t k = (int64_t)b * 4 * F + n;
    const float4 vv = ld4(v + k), qq = ld4(q + k);
    const float4 uu = ld4(u + (int64_t)b * ldu + n), pp = ld4(p + (int64_t)b * ldp + n);
    const float4 av = ld4(dv_acc + k), aq = ld4(dq_acc + k);
    st4(du + k, w.x * vv.x, w.y * vv.y, w.z * vv.z, w.w * vv.w);
    st4(dp + (int64_t)b * lddp + n, r.x * qq.x, r.y * qq.y, r.z * qq.z, r.w * qq.w);
    st4(dv_acc + k, av.x + w.x * uu.x, av.y + w.y * uu.y, av.z + w.z * uu.z, av.w + w.w * uu.w);
    st4(dq_acc + k, aq.x + r.x * pp.x, aq.y + r.y * pp.y, aq.z + r.z * pp.z, aq.w + r.w * pp.w);
  }
}

// all pointers 16-byte aligned and all pitches multiples of 4 elements?
template <typename... P>
inline bool aligned16(P... ptrs) {
  uintptr_t acc = 0;
  const uintptr_t v[] = {(uintptr_t)ptrs...};
  for (uintptr_t x : v) acc |= x;
  return (acc & 15u) == 0;
}
inline bool mult4(std::initializer_list<int64_t> v) {
  for (int64_t x : v)
    if (x % 4) return false;
  return true;
}

__global__ void concat_bias_kernel(float* dst, const float* a, int na, const float* b, int nb,
                                   int nzero) {
  const int n = na + nb + nzero;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < na) v = a[i];
    else if (i < na + nb) v = b[i - na];
    dst[i] = v;
  }
}

__global__ void zero_rows_kernel(float* x, const int32_t* __restrict__ len_d, int B, int T,
                                 int64_t row_elems) {
  const int r = blockIdx.x;       // r = b*T + t
  const int b = r / T, t = r - b * T;
  if (t < len_d[b]) return;
  float* p = x + (int64_t)r * row_elems;
  for (int64_t i = threadIdx.x; i < row_elems; i += blockDim.x) p[i] = 0.f;
}

inline int grid_for(int64_t total, int threads) {
  int64_t g = (total + threads - 1) / threads;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

#define DISPATCH_2FT(precision, src_ft, dst_ft, CALL)                                  \
  do {                                                                                 \
    const bool s_h = (src_ft) && (precision) == CAPDEC_BF16;                           \
    const bool d_h = (dst_ft) && (precision) == CAPDEC_BF16;                           \
    if (s_h && d_h) { CALL(bf16, bf16); }                                              \
    else if (s_h && !d_h) { CALL(bf16, float); }                                       \
    else if (!s_h && d_h) { CALL(float, bf16); }                                       \
    else { CALL(float, float); }                                                       \
  } while (0)

int transpose_cast(int precision, const void* src, int src_ft, void* dst, int dst_ft, int ni, int nj,
                   int C, int64_t s_i, int64_t s_j, int64_t ldd, int64_t d_i, int64_t d_j,
                   cudaStream_t st) {
  const int R = ni * nj;
  if (R <= 0 || C <= 0) return CAPDEC_OK;
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
#define CALL(TS, TD)                                                                          \
  transpose_kernel<TS, TD><<<grid, block, 0, st>>>((const TS*)src, (TD*)dst, ni, nj, C, s_i, s_j, \
                                                   ldd, d_i, d_j)
  DISPATCH_2FT(precision, src_ft, dst_ft, CALL);
#undef CALL
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int pack_multi(int precision, PackTable& t, cudaStream_t st) {
  if (t.n <= 0) return CAPDEC_OK;
  int tiles = 0;
  for (int i = 0; i < t.n; ++i) {
    t.seg[i].tile0 = tiles;
    tiles += ceil_div(t.seg[i].R, 64) * ceil_div(t.seg[i].C, 64);
  }
  if (tiles <= 0) return CAPDEC_OK;
  if (precision == CAPDEC_BF16) pack_multi_kernel<bf16><<<tiles, dim3(32, 8), 0, st>>>(t);
  else pack_multi_kernel<float><<<tiles, dim3(32, 8), 0, st>>>(t);
  CAPDEC_LAUNCH_OK();
  t.n = 0;
  return CAPDEC_OK;
}

int copy_cast(int precision, const void* src, int src_ft, int64_t lds, void* dst, int dst_ft,
              int64_t ldd, int R, int C, cudaStream_t st) {
  if (R <= 0 || C <= 0) return CAPDEC_OK;
  const int g = grid_for((int64_t)R * C, 256);
#define CALL(TS, TD) \
  copy_cast_kernel<TS, TD><<<g, 256, 0, st>>>((const TS*)src, lds, (TD*)dst, ldd, R, C)
  DISPATCH_2FT(precision, src_ft, dst_ft, CALL);
#undef CALL
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int colsum(int precision, const void* X, int x_ft, int64_t ld, int R, int N, float* out,
           int accumulate, cudaStream_t st) {
  if (N <= 0) return CAPDEC_OK;
  // few column blocks (N <= ~5000): 32 row lanes per block, otherwise the grid already fills the chip
  const bool tall = ceil_div(N, 32) < 296 && R >= 256;
  dim3 grid(ceil_div(N, 32)), block(32, tall ? 32 : 8);
  if (x_ft && precision == CAPDEC_BF16) {
    if (tall) colsum_kernel<bf16, 32><<<grid, block, 0, st>>>((const bf16*)X, ld, R, N, out, accumulate);
    else colsum_kernel<bf16, 8><<<grid, block, 0, st>>>((const bf16*)X, ld, R, N, out, accumulate);
  } else {
    if (tall) colsum_kernel<float, 32><<<grid, block, 0, st>>>((const float*)X, ld, R, N, out, accumulate);
    else colsum_kernel<float, 8><<<grid, block, 0, st>>>((const float*)X, ld, R, N, out, accumulate);
  }
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int gather_features(int precision, const float* enc, int64_t sb, int64_t sp, int64_t se,
                    const int64_t* sort_ind, void* enc_s, float* mean_f32, void* mean_ft,
                    int64_t ld_mean_ft, int B, int P, int E, cudaStream_t st, void* enc_cm, int cw) {
  CAPDEC_REQUIRE(!enc_cm || (cw > 0 && E % cw == 0), CAPDEC_ERR_BAD_SHAPE, "gather_features: E=%d cw=%d", E, cw);
  dim3 grid(ceil_div(E, 128), B);
  if (precision == CAPDEC_BF16)
    gather_features_kernel<bf16><<<grid, 128, 0, st>>>(enc, sb, sp, se, sort_ind, (bf16*)enc_s,
                                                       mean_f32, (bf16*)mean_ft, ld_mean_ft, B, P, E, (bf16*)enc_cm, cw);
  else
    gather_features_kernel<float><<<grid, 128, 0, st>>>(enc, sb, sp, se, sort_ind, (float*)enc_s,
                                                        mean_f32, (float*)mean_ft, ld_mean_ft, B, P, E, (float*)enc_cm, cw);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int embedding_gather(int precision, const float* emb, const int64_t* caps, int L, void* Xe,
                     int64_t ldx, int B, int T, int M, int V, cudaStream_t st) {
  if (B * T <= 0) return CAPDEC_OK;
  if (precision == CAPDEC_BF16)
    embedding_gather_kernel<bf16><<<B * T, 128, 0, st>>>(emb, caps, L, (bf16*)Xe, ldx, B, T, M, V);
  else
    embedding_gather_kernel<float><<<B * T, 128, 0, st>>>(emb, caps, L, (float*)Xe, ldx, B, T, M, V);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int embedding_scatter_add(const float* dXe, int64_t ldx, const int64_t* caps, int L,
                          const int32_t* len_d, float* dEmb, int B, int T, int M, int V,
                          cudaStream_t st) {
  if (B * T <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE((size_t)B * T * sizeof(int) <= 160 * 1024, CAPDEC_ERR_BAD_SHAPE,
                 "embedding scatter: B*T=%d rows exceed the shared row list", B * T);
  static bool attr_set = false;
  if (!attr_set) {
    CAPDEC_CUDA_OK(cudaFuncSetAttribute(embedding_scatter_add_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  embedding_scatter_add_kernel<<<B * T, 128, (size_t)B * T * sizeof(int), st>>>(dXe, ldx, caps, L, len_d,
                                                                             dEmb, B, T, M, V);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int expand_rows(int precision, const void* src, int64_t lds, void* dst, int64_t ldd, int G, int k, int C,
                cudaStream_t st) {
  if (G * k <= 0) return CAPDEC_OK;
  if (precision == CAPDEC_BF16)
    expand_rows_kernel<bf16><<<G * k, 128, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, k, C);
  else
    expand_rows_kernel<float><<<G * k, 128, 0, st>>>((const float*)src, lds, (float*)dst, ldd, k, C);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int scn_form_m(int precision, const float* u, int64_t ldu, const float* p, int64_t ldp,
               const float* v, const float* q, void* m, int rows, int B, int F, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  if (aligned16(u, p, v, q, m) && mult4({ldu, ldp, F}) && (int64_t)rows * F < (1ll << 30)) {
    const int g4 = grid_for((int64_t)rows * F, 256);
    if (precision == CAPDEC_BF16)
      CAPDEC_CUDA_OK(launch_pdl(scn_form_m_v4_kernel<bf16>, dim3(g4), dim3(256), 0, st, 1, u, ldu, p, ldp, v, q, (bf16*)m, rows, B, F));
    else
      CAPDEC_CUDA_OK(launch_pdl(scn_form_m_v4_kernel<float>, dim3(g4), dim3(256), 0, st, 1, u, ldu, p, ldp, v, q, (float*)m, rows, B, F));
    CAPDEC_LAUNCH_OK();
    return CAPDEC_OK;
  }
  const int g = grid_for((int64_t)rows * 4 * F, 256);
  if (precision == CAPDEC_BF16)
    CAPDEC_CUDA_OK(launch_pdl(scn_form_m_kernel<bf16>, dim3(g), dim3(256), 0, st, 1, u, ldu, p, ldp, v, q, (bf16*)m, rows, B, F));
  else
    CAPDEC_CUDA_OK(launch_pdl(scn_form_m_kernel<float>, dim3(g), dim3(256), 0, st, 1, u, ldu, p, ldp, v, q, (float*)m, rows, B, F));
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int cell_fwd(int precision, const float* preA, int64_t ldA, const float* preB, int64_t ldB,
             const float* b1, const float* b2, int lstm_order, const float* c_prev, float* c_new,
             float* gates, void* h_out, int64_t ldh, void* hd_out, float dropout_p, const uint64_t* seed,
             int t, int T, int rows, int D, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  if (aligned16(preA, preB, b1, b2, c_prev, c_new, gates, h_out, hd_out) && mult4({ldA, ldB, ldh, D}) &&
      (int64_t)rows * D < (1ll << 31)) {
    const int g4 = grid_for((int64_t)rows * D / 4, 128);
    if (precision == CAPDEC_BF16)
      CAPDEC_CUDA_OK(launch_pdl(cell_fwd_v4_kernel<bf16>, dim3(g4), dim3(128), 0, st, 1, preA, ldA, preB, ldB, b1, b2, lstm_order, c_prev,
                                c_new, gates, (bf16*)h_out, ldh, (bf16*)hd_out, dropout_p, seed, t, T, rows, D));
    else
      CAPDEC_CUDA_OK(launch_pdl(cell_fwd_v4_kernel<float>, dim3(g4), dim3(128), 0, st, 1, preA, ldA, preB, ldB, b1, b2, lstm_order, c_prev,
                                c_new, gates, (float*)h_out, ldh, (float*)hd_out, dropout_p, seed, t, T, rows, D));
    CAPDEC_LAUNCH_OK();
    return CAPDEC_OK;
  }
  const int g = grid_for((int64_t)rows * D, 128);
  if (precision == CAPDEC_BF16)
    CAPDEC_CUDA_OK(launch_pdl(cell_fwd_kernel<bf16>, dim3(g), dim3(128), 0, st, 1, preA, ldA, preB, ldB, b1, b2, lstm_order, c_prev, c_new,
                                             gates, (bf16*)h_out, ldh, (bf16*)hd_out, dropout_p, seed,
                                             t, T, rows, D));
  else
    CAPDEC_CUDA_OK(launch_pdl(cell_fwd_kernel<float>, dim3(g), dim3(128), 0, st, 1, preA, ldA, preB, ldB, b1, b2, lstm_order, c_prev, c_new,
                                              gates, (float*)h_out, ldh, (float*)hd_out, dropout_p,
                                              seed, t, T, rows, D));
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int cell_bwd(int precision, const float* dh_fc, int64_t ld_dhfc, const float* dh_rec, float* dc,
             const float* gates, const float* c_prev, const float* c_new, int lstm_order,
             float dropout_p, const uint64_t* seed, int t, int T, void* dpre, float* dpre_f32, int rows,
             int D, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  if (aligned16(dh_fc, dh_rec, dc, gates, c_prev, c_new, dpre, dpre_f32) && mult4({ld_dhfc, D}) &&
      (int64_t)rows * D < (1ll << 31)) {
    const int g4 = grid_for((int64_t)rows * D / 4, 128);
    if (precision == CAPDEC_BF16)
      CAPDEC_CUDA_OK(launch_pdl(cell_bwd_v4_kernel<bf16>, dim3(g4), dim3(128), 0, st, 1, dh_fc, ld_dhfc, dh_rec, dc, gates, c_prev, c_new,
                                lstm_order, dropout_p, seed, t, T, (bf16*)dpre, dpre_f32, rows, D));
    else
      CAPDEC_CUDA_OK(launch_pdl(cell_bwd_v4_kernel<float>, dim3(g4), dim3(128), 0, st, 1, dh_fc, ld_dhfc, dh_rec, dc, gates, c_prev, c_new,
                                lstm_order, dropout_p, seed, t, T, (float*)dpre, dpre_f32, rows, D));
    CAPDEC_LAUNCH_OK();
    return CAPDEC_OK;
  }
  const int g = grid_for((int64_t)rows * D, 128);
  if (precision == CAPDEC_BF16)
    CAPDEC_CUDA_OK(launch_pdl(cell_bwd_kernel<bf16>, dim3(g), dim3(128), 0, st, 1, dh_fc, ld_dhfc, dh_rec, dc, gates, c_prev, c_new,
                                             lstm_order, dropout_p, seed, t, T, (bf16*)dpre, dpre_f32,
                                             rows, D));
  else
    CAPDEC_CUDA_OK(launch_pdl(cell_bwd_kernel<float>, dim3(g), dim3(128), 0, st, 1, dh_fc, ld_dhfc, dh_rec, dc, gates, c_prev, c_new,
                                              lstm_order, dropout_p, seed, t, T, (float*)dpre,
                                              dpre_f32, rows, D));
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int scn_bwd_products(int precision, const float* wr, const float* u, int64_t ldu, const float* p,
                     int64_t ldp, const float* v, const float* q, void* du, void* dp,
                     float* dv_acc, float* dq_acc, int rows, int B, int F, int64_t lddp, cudaStream_t st) {
  if (rows <= 0) return CAPDEC_OK;
  if (aligned16(wr, u, p, v, q, du, dp, dv_acc, dq_acc) && mult4({ldu, ldp, lddp, F}) && (int64_t)rows * F < (1ll << 30)) {
    const int g4 = grid_for((int64_t)rows * F, 256);
    if (precision == CAPDEC_BF16)
      CAPDEC_CUDA_OK(launch_pdl(scn_bwd_products_v4_kernel<bf16>, dim3(g4), dim3(256), 0, st, 1, wr, u, ldu, p, ldp, v, q, (bf16*)du,
                                (bf16*)dp, dv_acc, dq_acc, rows, B, F, lddp));
    else
      CAPDEC_CUDA_OK(launch_pdl(scn_bwd_products_v4_kernel<float>, dim3(g4), dim3(256), 0, st, 1, wr, u, ldu, p, ldp, v, q, (float*)du,
                                (float*)dp, dv_acc, dq_acc, rows, B, F, lddp));
    CAPDEC_LAUNCH_OK();
    return CAPDEC_OK;
  }
  const int g = grid_for((int64_t)rows * 4 * F, 256);
  if (precision == CAPDEC_BF16)
    CAPDEC_CUDA_OK(launch_pdl(scn_bwd_products_kernel<bf16>, dim3(g), dim3(256), 0, st, 1, wr, u, ldu, p, ldp, v, q, (bf16*)du, (bf16*)dp,
                                                     dv_acc, dq_acc, rows, B, F, lddp));
  else
    CAPDEC_CUDA_OK(launch_pdl(scn_bwd_products_kernel<float>, dim3(g), dim3(256), 0, st, 1, wr, u, ldu, p, ldp, v, q, (float*)du,
                                                      (float*)dp, dv_acc, dq_acc, rows, B, F, lddp));
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int concat_bias(float* dst, const float* a, int na, const float* b, int nb, int nzero,
                cudaStream_t st) {
  const int n = na + nb + nzero;
  if (n <= 0) return CAPDEC_OK;
  concat_bias_kernel<<<grid_for(n, 256), 256, 0, st>>>(dst, a, na, b, nb, nzero);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int zero_rows_beyond_len(float* x, const int32_t* len_d, int B, int T, int64_t row_elems,
                         cudaStream_t st) {
  if (B * T <= 0) return CAPDEC_OK;
  zero_rows_kernel<<<B * T, 256, 0, st>>>(x, len_d, B, T, row_elems);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace capdec
