// gemm_simt.cu -- fp32 FFMA GEMM engine (the "fp32 mode" of the decoder).
//
//   out[r, n] = sum_k X[r, k] * W[n, k]  (+ bias[n]) (+ addm[r, n])
//
// Both operands are K-contiguous ("NT" form): X is an activation matrix, W a packed
// weight matrix with one output feature per row (the nn.Linear layout of the
// reference, models/attention.py:18-22; the SCN factor matrices are packed into this
// form by pack.cu).  fp32 mode exists for parity: BASELINE.json asks for logits within
// 1e-4 of the reference's fp32 arithmetic, which TF32/bf16 tensor-core inputs cannot
// give, so this engine multiplies and accumulates in plain fp32.
#include "common.cuh"

namespace capdec {

namespace {

constexpr int BK = 16;

template <int BR, int BN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw,
                 float* out, int64_t ldo, const float* __restrict__ bias,
                 const float* addm, int64_t ldadd, int rows, int N, int K,
                 int64_t sX, int64_t sW, int64_t sO, int64_t sBias, int64_t sAdd, int vec_ok) {
  constexpr int TM = BR / 16, TN = BN / 16;
  constexpr int PAD = 4;
  __shared__ __align__(16) float Xs[BK][BR + PAD];
  __shared__ __align__(16) float Ws[BK][BN + PAD];

  const int z = blockIdx.z;
  X += (int64_t)z * sX;
  W += (int64_t)z * sW;
  out += (int64_t)z * sO;
  if (bias) bias += (int64_t)z * sBias;
  if (addm) addm += (int64_t)z * sAdd;

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int r0 = blockIdx.y * BR, n0 = blockIdx.x * BN;

  // loader mapping: one float4 (4 consecutive k) per thread per tile
  const int lrow = tid >> 2, lkq = (tid & 3) * 4;
  const bool x_loader = lrow < BR;
  const bool w_loader = lrow < BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 xr = make_float4(0, 0, 0, 0), wr = make_float4(0, 0, 0, 0);
  auto fetch = [&](int k0) {
    xr = make_float4(0, 0, 0, 0);
    wr = make_float4(0, 0, 0, 0);
    const int k = k0 + lkq;
    if (x_loader && r0 + lrow < rows) {
      const float* p = X + (int64_t)(r0 + lrow) * ldx + k;
      if (vec_ok && k + 3 < K) {
        xr = *reinterpret_cast<const float4*>(p);
      } else {
        if (k + 0 < K) xr.x = p[0];
        if (k + 1 < K) xr.y = p[1];
        if (k + 2 < K) xr.z = p[2];
        if (k + 3 < K) xr.w = p[3];
      }
    }
    if (w_loader && n0 + lrow < N) {
      const float* p = W + (int64_t)(n0 + lrow) * ldw + k;
      if (vec_ok && k + 3 < K) {
        wr = *reinterpret_cast<const float4*>(p);
      } else {
        if (k + 0 < K) wr.x = p[0];
        if (k + 1 < K) wr.y = p[1];
        if (k + 2 < K) wr.z = p[2];
        if (k + 3 < K) wr.w = p[3];
      }
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    __syncthreads();
    if (x_loader) {
      Xs[lkq + 0][lrow] = xr.x; Xs[lkq + 1][lrow] = xr.y;
      Xs[lkq + 2][lrow] = xr.z; Xs[lkq + 3][lrow] = xr.w;
    }
    if (w_loader) {
      Ws[lkq + 0][lrow] = wr.x; Ws[lkq + 1][lrow] = wr.y;
      Ws[lkq + 2][lrow] = wr.z; Ws[lkq + 3][lrow] = wr.w;
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
    // two-level summation: each 16-wide k-tile is summed on its own and then added to the running
    // total, which keeps the fp32 rounding error near that of a blocked BLAS instead of growing with K
    float part[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = Xs[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] += part[i][j];
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = r0 + ty * TM + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (addm) v += addm[(int64_t)r * ldadd + n];
      out[(int64_t)r * ldo + n] = v;
    }
  }
}

}  // namespace

int gemm_simt(const GemmArgs& a, cudaStream_t st) {
  if (a.rows <= 0 || a.N <= 0) return CAPDEC_OK;
  CAPDEC_REQUIRE(a.X && a.W && a.out && a.K > 0, CAPDEC_ERR_BAD_ARG, "gemm_simt: null operand");
  const float* X = (const float*)a.X;
  const float* W = (const float*)a.W;
  int vec_ok = (a.ldx % 4 == 0) && (a.ldw % 4 == 0) && (((uintptr_t)X) % 16 == 0) &&
               (((uintptr_t)W) % 16 == 0) && (a.sX % 4 == 0) && (a.sW % 4 == 0);
  if (a.rows <= 32) {
    dim3 grid(ceil_div(a.N, 32), ceil_div(a.rows, 32), a.batch);
    gemm_simt_kernel<32, 32><<<grid, 256, 0, st>>>(X, a.ldx, W, a.ldw, (float*)a.out, a.ldo, a.bias,
                                                   a.addm, a.ldadd, a.rows, a.N, a.K, a.sX, a.sW,
                                                   a.sO, a.sBias, a.sAdd, vec_ok);
  } else {
    dim3 grid(ceil_div(a.N, 64), ceil_div(a.rows, 64), a.batch);
    gemm_simt_kernel<64, 64><<<grid, 256, 0, st>>>(X, a.ldx, W, a.ldw, (float*)a.out, a.ldo, a.bias,
                                                   a.addm, a.ldadd, a.rows, a.N, a.K, a.sX, a.sW,
                                                   a.sO, a.sBias, a.sAdd, vec_ok);
  }
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace capdec
