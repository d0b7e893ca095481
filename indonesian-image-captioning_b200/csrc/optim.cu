// optim.cu -- fused gradient clip + Adam over all decoder parameters in ONE launch.
//
// Reference: trains/attention_scn.py:244-252 -- `clip_gradient(decoder_optimizer, grad_clip)`
// (utils/optimizer.py:1-11: every .grad clamped element-wise to [-clip, clip], in place) followed by
// `decoder_optimizer.step()` with torch.optim.Adam(lr) defaults (betas 0.9/0.999, eps 1e-8, no weight
// decay, no amsgrad).  The stock path is one clamp kernel per tensor plus the multi-kernel foreach Adam
// over 23 tensors; here a segment table drives one grid over every element: p, g, m, v are read once and
// p, m, v (and the clamped g, as the reference leaves it) written once -- 0.87 GB at 27.2 M parameters,
// an HBM-bound pass.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int AT = 256;
constexpr int AV = 4;                       // elements per thread per iteration (float4)
constexpr int ACHUNK = AT * AV * 4;         // elements per CTA chunk

struct AdamTable {
  CapdecAdamSeg seg[CAPDEC_ADAM_MAX_SEGS];
  int chunk0[CAPDEC_ADAM_MAX_SEGS + 1];
  int n;
};

struct AdamHyper {      // derived on the host in double precision, like the python floats of torch.optim.Adam
  float beta2, omb1, omb2, eps, weight_decay, grad_clip, step_size, bc2_sqrt;
  int write_clipped;
};

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, const AdamHyper& h) {
  if (h.grad_clip > 0.f) g = fminf(fmaxf(g, -h.grad_clip), h.grad_clip);
  float gg = g;
  if (h.weight_decay != 0.f) gg = fmaf(h.weight_decay, p, gg);
  m = fmaf(h.omb1, gg - m, m);                        // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(h.omb2 * gg, gg, v * h.beta2);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p - h.step_size * (m / denom);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(AT) clip_adam_kernel(const __grid_constant__ AdamTable t,
                                                       const __grid_constant__ AdamHyper h) {
  int si = 0;
  while (si + 1 < t.n && (int)blockIdx.x >= t.chunk0[si + 1]) ++si;
  const CapdecAdamSeg& s = t.seg[si];
  const int64_t base = (int64_t)(blockIdx.x - t.chunk0[si]) * ACHUNK;
  const int64_t n = s.n;
  const bool vec = (((uintptr_t)s.p | (uintptr_t)s.g | (uintptr_t)s.m | (uintptr_t)s.v) & 15) == 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t i = base + ((int64_t)it * AT + threadIdx.x) * AV;
    if (i >= n) break;
    if (vec && i + AV <= n) {
      float4 p = *reinterpret_cast<float4*>(s.p + i), g = *reinterpret_cast<const float4*>(s.g + i);
      float4 m = *reinterpret_cast<float4*>(s.m + i), v = *reinterpret_cast<float4*>(s.v + i);
      adam_elem(p.x, g.x, m.x, v.x, h); adam_elem(p.y, g.y, m.y, v.y, h);
      adam_elem(p.z, g.z, m.z, v.z, h); adam_elem(p.w, g.w, m.w, v.w, h);
      *reinterpret_cast<float4*>(s.p + i) = p;
      *reinterpret_cast<float4*>(s.m + i) = m;
      *reinterpret_cast<float4*>(s.v + i) = v;
      if (h.write_clipped) *reinterpret_cast<float4*>(s.g + i) = g;
    } else {
      for (int k = 0; k < AV && i + k < n; ++k) {
        float p = s.p[i + k], g = s.g[i + k], m = s.m[i + k], v = s.v[i + k];
        adam_elem(p, g, m, v, h);
        s.p[i + k] = p; s.m[i + k] = m; s.v[i + k] = v;
        if (h.write_clipped) s.g[i + k] = g;
      }
    }
  }
}

}  // namespace

int clip_adam_step(const CapdecAdamSeg* segs, int n_segs, double lr, double beta1, double beta2, double eps,
                   double weight_decay, double grad_clip, int step, int write_clipped, cudaStream_t st) {
  CAPDEC_REQUIRE(segs && n_segs >= 0 && step >= 1, CAPDEC_ERR_BAD_ARG, "clip_adam_step: bad argument");
  int done = 0;
  while (done < n_segs) {                      // more tensors than one table holds: several launches
    AdamTable t;
    memset(&t, 0, sizeof t);
    int chunks = 0;
    while (done < n_segs && t.n < CAPDEC_ADAM_MAX_SEGS) {
      const CapdecAdamSeg& s = segs[done++];
      if (s.n <= 0) continue;
      CAPDEC_REQUIRE(s.p && s.g && s.m && s.v, CAPDEC_ERR_BAD_ARG, "clip_adam_step: null tensor");
      t.seg[t.n] = s;
      t.chunk0[t.n] = chunks;
      chunks += (int)((s.n + ACHUNK - 1) / ACHUNK);
      ++t.n;
    }
    t.chunk0[t.n] = chunks;
    if (chunks == 0) continue;
    AdamHyper h;
    h.beta2 = (float)beta2; h.omb1 = (float)(1.0 - beta1); h.omb2 = (float)(1.0 - beta2); h.eps = (float)eps;
    h.weight_decay = (float)weight_decay; h.grad_clip = (float)grad_clip;
    h.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
    h.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    h.write_clipped = write_clipped;
    clip_adam_kernel<<<chunks, AT, 0, st>>>(t, h);
    CAPDEC_LAUNCH_OK();
  }
  return CAPDEC_OK;
}

}  // namespace capdec
