// loss.cu -- the loss glue of the training loop as two fused kernels + a finaliser.
//
// Reference: trains/attention_scn.py:219-235
//   targets = caps_sorted[:, 1:] ; pack_padded_sequence(scores/targets, decode_lengths)
//   loss = CrossEntropyLoss(mean over N = sum(decode_lengths))
//        + alpha_c * mean_{b,p} (1 - sum_t alphas[b,t,p])^2
// The packing is only a row selection: row (b,t) takes part iff t < decode_len[b].
#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int NT = 256;
constexpr int ARS = 4;            // time slices of the alpha-regulariser kernels (ARS * NT threads per CTA)

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < NT / 32; ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < NT / 32; ++i) r += sh[i];
  __syncthreads();
  return r;
}

// one CTA per (b,t) row: log-sum-exp and the target's negative log-likelihood
__global__ void __launch_bounds__(NT)
ce_fwd_kernel(const float* __restrict__ pred, const int64_t* __restrict__ caps,
              const int32_t* __restrict__ len_d, int T, int V, int L, float* __restrict__ lse_out,
              float* __restrict__ nll_out) {
  __shared__ float sh[NT / 32];
  const int r = blockIdx.x;
  const int b = r / T, t = r - b * T;
  if (t >= len_d[b]) {
    if (threadIdx.x == 0) { lse_out[r] = 0.f; nll_out[r] = 0.f; }
    return;
  }
  const float* x = pred + (int64_t)r * V;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += NT) m = fmaxf(m, x[i]);
  m = block_max(m, sh);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += NT) s += expf(x[i] - m);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    const float lse = logf(s) + m;
    int64_t tgt = caps[(int64_t)b * L + t + 1];
    if (tgt < 0) tgt = 0;
    if (tgt >= V) tgt = V - 1;
    lse_out[r] = lse;
    nll_out[r] = lse - x[tgt];
  }
}

// Top-k accuracy count (utils/metric.py:25-39 `accuracy`: scores.topk(k) contains the target).  One CTA per
// row: the target is a hit iff fewer than k logits rank before it (larger value; equal value and smaller
// index -- the order of a descending stable sort).  Two row sources: the (B,T,V) predictions with the
// captions / lengths of the training step (caps != NULL; rows beyond a caption's length do not count), or
// packed (N,V) scores with (N) targets, as the reference's loss glue builds them.
__global__ void __launch_bounds__(NT)
topk_hits_kernel(const float* __restrict__ scores, int64_t ld, const int64_t* __restrict__ targets,
                 const int64_t* __restrict__ caps, const int32_t* __restrict__ len_d, int T, int L, int V, int k,
                 int* __restrict__ hits) {
  __shared__ float sh[NT / 32];
  const int r = blockIdx.x;
  int64_t tgt;
  if (caps) {
    const int b = r / T, t = r - b * T;
    if (t >= len_d[b]) return;
    tgt = caps[(int64_t)b * L + t + 1];
  } else {
    tgt = targets[r];
  }
  if (tgt < 0 || tgt >= V) return;              // out-of-range label: never in the top k
  const float* x = scores + (int64_t)r * ld;
  const float xt = x[tgt];
  int before = 0;
  for (int i = threadIdx.x; i < V; i += NT) {
    const float v = x[i];
    before += (v > xt || (v == xt && i < (int)tgt)) ? 1 : 0;
  }
  const float tot = block_sum((float)before, sh);     // exact: counts stay far below 2^24
  if (threadIdx.x == 0 && tot < (float)k) atomicAdd(hits, 1);
}

// one CTA per caption b: sum_p (1 - sum_t alpha[b,t,p])^2
__global__ void __launch_bounds__(ARS * NT)
alpha_reg_fwd_kernel(const float* __restrict__ alphas, int T, int P, float* __restrict__ regpart) {
  // one CTA per caption, 4 x NT threads: thread (pixel, time slice) sums every 4th step, so the dependent chain is
  // T/4 loads long (a single 50-long chain per pixel made this 0.3 MB kernel 13 us)
  __shared__ float sh[ARS * NT / 32];
  __shared__ float part[ARS][NT];
  const int b = blockIdx.x;
  const float* a = alphas + (int64_t)b * T * P;
  const int pl = threadIdx.x % NT, ts = threadIdx.x / NT;
  float acc = 0.f;
  for (int p0 = 0; p0 < P; p0 += NT) {
    const int p = p0 + pl;
    float s = 0.f;
    if (p < P)
      for (int t = ts; t < T; t += ARS) s += a[(int64_t)t * P + p];
    part[ts][pl] = s;
    __syncthreads();
    if (ts == 0 && p < P) {
      float tot = part[0][pl];
#pragma unroll
      for (int k = 1; k < ARS; ++k) tot += part[k][pl];
      const float d = 1.f - tot;
      acc += d * d;
    }
    __syncthreads();
  }
  acc = warp_sum(acc);                               // zero in the threads of time slices > 0
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < ARS * NT / 32; ++i) tot += sh[i];
    regpart[b] = tot;
  }
}

__global__ void __launch_bounds__(NT)
loss_finalize_kernel(const float* __restrict__ nll, int n_rows, const float* __restrict__ regpart,
                     int B, int P, int n_tokens, float alpha_c, float* __restrict__ loss_out) {
  __shared__ float sh[NT / 32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_rows; i += NT) s += nll[i];
  s = block_sum(s, sh);
  float r = 0.f;
  if (regpart)
    for (int i = threadIdx.x; i < B; i += NT) r += regpart[i];
  r = block_sum(r, sh);
  if (threadIdx.x == 0) {
    const float ce = s / (float)n_tokens;
    const float reg = regpart ? alpha_c * r / ((float)B * (float)P) : 0.f;
    loss_out[0] = ce + reg;
    loss_out[1] = ce;
    loss_out[2] = reg;
  }
}

// d logits = g/N (softmax - onehot) for active rows, 0 otherwise
template <typename FT>
__global__ void __launch_bounds__(NT)
ce_bwd_kernel(const float* __restrict__ pred, const int64_t* __restrict__ caps,
              const int32_t* __restrict__ len_d, const float* __restrict__ lse, int T, int V, int L,
              float scale, const float* __restrict__ gscale_dev, float* __restrict__ d_pred,
              FT* __restrict__ d_ft, int64_t ldq) {
  if (gscale_dev) scale *= gscale_dev[0];
  const int r = blockIdx.x;
  const int b = r / T, t = r - b * T;
  const bool active = t < len_d[b];
  const float* x = pred + (int64_t)r * V;
  float* dp = d_pred ? d_pred + (int64_t)r * V : nullptr;
  FT* dq = d_ft ? d_ft + (int64_t)r * ldq : nullptr;
  if (!active) {
    for (int i = threadIdx.x; i < V; i += NT) {
      if (dp) dp[i] = 0.f;
      if (dq) dq[i] = from_f<FT>(0.f);
    }
    return;
  }
  const float l = lse[r];
  int64_t tgt = caps[(int64_t)b * L + t + 1];
  if (tgt < 0) tgt = 0;
  if (tgt >= V) tgt = V - 1;
  for (int i = threadIdx.x; i < V; i += NT) {
    float g = expf(x[i] - l);
    if (i == (int)tgt) g -= 1.f;
    g *= scale;
    if (dp) dp[i] = g;
    if (dq) dq[i] = from_f<FT>(g);
  }
}

__global__ void __launch_bounds__(ARS * NT)
alpha_reg_bwd_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ len_d, int T,
                     int P, float scale, const float* __restrict__ gscale_dev,
                     float* __restrict__ d_alphas) {
  if (gscale_dev) scale *= gscale_dev[0];
  __shared__ float part[ARS][NT];
  const int b = blockIdx.x;
  const float* a = alphas + (int64_t)b * T * P;
  float* d = d_alphas + (int64_t)b * T * P;
  const int len = len_d[b];
  const int pl = threadIdx.x % NT, ts = threadIdx.x / NT;      // (pixel, time slice), as in the forward
  for (int p0 = 0; p0 < P; p0 += NT) {
    const int p = p0 + pl;
    float s = 0.f;
    if (p < P)
      for (int t = ts; t < T; t += ARS) s += a[(int64_t)t * P + p];
    part[ts][pl] = s;
    __syncthreads();
    if (p < P) {
      float tot = part[0][pl];
#pragma unroll
      for (int k = 1; k < ARS; ++k) tot += part[k][pl];
      const float g = -2.f * (1.f - tot) * scale;
      for (int t = ts; t < T; t += ARS) d[(int64_t)t * P + p] = t < len ? g : 0.f;
    }
    __syncthreads();
  }
}

}  // namespace

int loss_fwd(const CapdecDims& d, const float* pred, const float* alphas, const int64_t* caps,
             const int32_t* len_d, int n_tokens, float alpha_c, float* loss_out, float* lse_out,
             cudaStream_t st) {
  const int R = d.B * d.T;
  float* nll = lse_out + R;
  float* regpart = alphas ? lse_out + 2 * R : nullptr;
  ce_fwd_kernel<<<R, NT, 0, st>>>(pred, caps, len_d, d.T, d.V, d.L, lse_out, nll);
  CAPDEC_LAUNCH_OK();
  if (alphas) {
    alpha_reg_fwd_kernel<<<d.B, ARS * NT, 0, st>>>(alphas, d.T, d.P, regpart);
    CAPDEC_LAUNCH_OK();
  }
  loss_finalize_kernel<<<1, NT, 0, st>>>(nll, R, regpart, d.B, d.P, n_tokens, alpha_c, loss_out);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int loss_bwd(const CapdecDims& d, const float* pred, const float* alphas, const int64_t* caps,
             const int32_t* len_d, int n_tokens, float alpha_c, float gscale, const float* gscale_dev,
             const float* lse, float* d_pred, void* d_logits_ft, int64_t ldq, float* d_alphas,
             cudaStream_t st) {
  const int R = d.B * d.T;
  const float scale = gscale / (float)n_tokens;
  if (d_pred || d_logits_ft) {
    if (d.precision == CAPDEC_BF16)
      ce_bwd_kernel<bf16><<<R, NT, 0, st>>>(pred, caps, len_d, lse, d.T, d.V, d.L, scale, gscale_dev,
                                            d_pred, (bf16*)d_logits_ft, ldq);
    else
      ce_bwd_kernel<float><<<R, NT, 0, st>>>(pred, caps, len_d, lse, d.T, d.V, d.L, scale, gscale_dev,
                                             d_pred, (float*)d_logits_ft, ldq);
    CAPDEC_LAUNCH_OK();
  }
  if (alphas && d_alphas) {
    alpha_reg_bwd_kernel<<<d.B, ARS * NT, 0, st>>>(alphas, len_d, d.T, d.P,
                                             gscale * alpha_c / ((float)d.B * (float)d.P), gscale_dev,
                                             d_alphas);
    CAPDEC_LAUNCH_OK();
  }
  return CAPDEC_OK;
}

int topk_hits(const float* scores, int64_t ld, const int64_t* targets, const int64_t* caps, const int32_t* len_d,
              int rows, int T, int L, int V, int k, int* hits, cudaStream_t st) {
  CAPDEC_REQUIRE(scores && hits && (targets || (caps && len_d)) && rows >= 0 && V > 0 && k >= 1, CAPDEC_ERR_BAD_ARG,
                 "topk_hits: bad argument");
  CAPDEC_CUDA_OK(cudaMemsetAsync(hits, 0, sizeof(int), st));
  if (rows == 0) return CAPDEC_OK;
  topk_hits_kernel<<<rows, NT, 0, st>>>(scores, ld, targets, caps, len_d, T > 0 ? T : 1, L, V, k, hits);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}


// the keep factors (0 or 1/(1-p)) the decoder applies to h_t before fc in training mode: the same counter-based
// hash of (seed, (b*T + t)*D + d) the forward and backward kernels evaluate in place (common.cuh dropout_scale)
__global__ void dropout_mask_kernel(uint64_t seed, float p, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = dropout_scale(seed, (uint64_t)i, p);
}

int dropout_mask(uint64_t seed, float p, int64_t n, float* out, cudaStream_t st) {
  CAPDEC_REQUIRE(out && n >= 0 && p >= 0.f && p < 1.f, CAPDEC_ERR_BAD_ARG, "dropout_mask: bad argument");
  if (n == 0) return CAPDEC_OK;
  const int64_t blocks = (n + 255) / 256;
  dropout_mask_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(seed, p, n, out);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace capdec
