// beam.cu -- device-side kernels of the batched beam search (reference `sample`:
// models/decoders/attention_scn.py:160-296, pure_scn.py:142-249, pure_attention.py:153-281).
//
// The reference searches ONE image at a time and goes through the host every step
// (`enumerate(next_word_inds)`, :262).  Here G independent searches advance together, all
// bookkeeping stays on the device and the host is touched once at the end:
//   * beam rows of image g are rows g*k .. g*k+s-1 (s = live beams, shrinking as beams finish)
//   * no history is copied when beams are re-ordered: every step stores back-pointers
//     (parent slot, word) and the attention maps of all rows; the winning caption is
//     reconstructed by back-tracking in beam_finalize_kernel
//   * selection = log-softmax over the vocabulary + running score + top-k over the s*V
//     candidates of an image (:235-253), in descending order like torch.topk(sorted=True)
#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int NT = 256;
constexpr int KMAX = 8;

template <typename FT>
__global__ void beam_embed_kernel(const float* __restrict__ emb, const int32_t* __restrict__ prev_word,
                                  FT* __restrict__ Xe, int64_t ldx, int M, int V) {
  const int r = blockIdx.x;
  int w = prev_word[r];
  if (w < 0) w = 0;
  if (w >= V) w = V - 1;
  const float* src = emb + (int64_t)w * M;
  FT* dst = Xe + (int64_t)r * ldx;
  for (int i = threadIdx.x; i < M; i += blockDim.x) dst[i] = from_f<FT>(src[i]);
}

__global__ void beam_init_kernel(int32_t* prev_word, float* score, int32_t* live, int32_t* krem,
                                 int32_t* has_done, float* best_score, int32_t* best_t,
                                 int32_t* best_parent, int G, int k, int32_t start_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G * k) {
    prev_word[i] = start_id;
    score[i] = 0.f;
  }
  if (i < G) {
    live[i] = k;
    krem[i] = k;
    has_done[i] = 0;
    best_score[i] = -INFINITY;
    best_t[i] = -1;
    best_parent[i] = 0;
  }
}

// h_state[g*k + j] = h_new[g*k + src_row[g*k + j]] for the surviving beams (attention_scn.py:278-285)
template <typename FT>
__global__ void beam_gather_state_kernel(const FT* __restrict__ h_new, const float* __restrict__ c_new,
                                         FT* __restrict__ h_state, float* __restrict__ c_state,
                                         const int32_t* __restrict__ src_row,
                                         const int32_t* __restrict__ live, int k, int D, int64_t ldh) {
  const int r = blockIdx.x;
  const int g = r / k, j = r - g * k;
  if (j >= live[g]) return;
  const int src = g * k + src_row[r];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    h_state[(int64_t)r * ldh + i] = h_new[(int64_t)src * ldh + i];
    c_state[(int64_t)r * D + i] = c_new[(int64_t)src * D + i];
  }
}

struct ArgMax {
  float v;
  int idx;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  // larger value wins; ties -> smaller flat index
  if (b.v > a.v || (b.v == a.v && b.idx < a.idx)) return b;
  return a;
}

// sorted insertion of `cand` into a descending list of KL entries (value, then smaller index)
template <int KL>
__device__ __forceinline__ void list_insert(ArgMax (&l)[KL], ArgMax cand) {
  const ArgMax last = l[KL - 1];
  if (cand.v > last.v || (cand.v == last.v && cand.idx < last.idx)) {
    l[KL - 1] = cand;
#pragma unroll
    for (int q = KL - 1; q > 0; --q) {
      const ArgMax lo = l[q], hi = l[q - 1];
      if (lo.v > hi.v || (lo.v == hi.v && lo.idx < hi.idx)) { l[q] = hi; l[q - 1] = lo; }
    }
  }
}

// one CTA per image; KL = length of the per-thread candidate lists (>= beam size).
// FAST (bf16 mode): ONE pass per row -- online log-sum-exp and the row's KL largest raw logits per thread
// (within a row, ordering by logit = ordering by log-probability); the exact mode keeps the three-pass
// arithmetic that the fp32 token-parity tests pin.
// FAST = 2: `logits` is the partial buffer written by the fused vocabulary kernel (gemm_tc_vocab_topk):
// [row][slots][2 + 2 KL] = {max, sum exp, KL largest logits, their indices} per run of vocabulary tiles one CTA of
// that kernel walked; the row's log-sum-exp and candidates are merged from its valid slots (the schedule is
// recomputed from the plan), the (rows x V) logits are never materialised.
template <int KL, int FAST>
__global__ void __launch_bounds__(NT)
beam_select_kernel(const float* __restrict__ logits, int V, VocabTopkPlan vp, int k, int t, int32_t end_id,
                   const float* __restrict__ score_in, float* __restrict__ score_out,
                   int32_t* __restrict__ prev_word, int32_t* __restrict__ src_row,
                   int32_t* __restrict__ live, int32_t* __restrict__ krem,
                   int32_t* __restrict__ has_done, float* __restrict__ best_score,
                   int32_t* __restrict__ best_t, int32_t* __restrict__ best_parent,
                   int32_t* __restrict__ bp_parent, int32_t* __restrict__ bp_word,
                   int32_t* __restrict__ tr_parent, int32_t* __restrict__ tr_word,
                   float* __restrict__ tr_score, int n_steps, int G) {
  __shared__ float red_v[NT / 32];
  __shared__ int red_i[NT / 32];
  __shared__ float s_max[KMAX], s_logsum[KMAX], s_score[KMAX];
  __shared__ int s_sel[KMAX];
  __shared__ float s_selv[KMAX];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kr = krem[g];
  const int s = live[g];
  if (tr_parent) {
    for (int i = tid; i < k; i += NT) {
      const int64_t o = ((int64_t)g * n_steps + t) * k + i;
      tr_parent[o] = -1; tr_word[o] = -1; tr_score[o] = 0.f;
    }
  }
  if (kr <= 0 || s <= 0) return;               // this image is finished
  const int ns = (t == 0) ? 1 : s;             // first step: all k rows are identical, use row 0 (:242-244)
  const float* base = logits + (int64_t)g * k * V;

  ArgMax mine[KL];
#pragma unroll
  for (int q = 0; q < KL; ++q) mine[q] = ArgMax{-INFINITY, 0x7fffffff};
  if (FAST) {
    for (int j = 0; j < ns; ++j) {
      const float* x = base + (int64_t)j * V;
      ArgMax rowl[KL];
#pragma unroll
      for (int q = 0; q < KL; ++q) rowl[q] = ArgMax{-INFINITY, 0x7fffffff};
      float m = -INFINITY, sum = 0.f;
      if (FAST == 2) {
        constexpr int PW = 2 + 2 * KL;
        const int row = g * k + j, rtile = row / 128;
        const int run_lo = (int)(((int64_t)rtile * vp.tiles_r * vp.grid) / vp.num_tiles);
        const int run_hi = (int)((((int64_t)(rtile + 1) * vp.tiles_r - 1) * vp.grid) / vp.num_tiles);
        const int nslot = 2 * (run_hi - run_lo + 1);          // two column halves per run
        const float* pr = logits + ((int64_t)row * vp.slots) * PW;
        for (int tt = tid; tt < nslot; tt += NT) {
          const float* q = pr + (int64_t)tt * PW;
          const float mt = q[0], st = q[1];
          const float mn = fmaxf(m, mt);
          if (mn > -INFINITY) {
            sum = sum * __expf(m - mn) + st * __expf(mt - mn);
            m = mn;
          }
#pragma unroll
          for (int e = 0; e < KL; ++e) {
            const int idx = __float_as_int(q[2 + KL + e]);
            if (idx != 0x7fffffff) list_insert<KL>(rowl, ArgMax{q[2 + e], idx});
          }
        }
      }
      const int V4 = (FAST == 2) ? 0 : ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? (V >> 2) : 0;   // 16-byte loads when aligned
      for (int i4 = tid; i4 < V4; i4 += NT) {
        const float4 q4 = *reinterpret_cast<const float4*>(x + 4 * i4);
        const float xs[4] = {q4.x, q4.y, q4.z, q4.w};
        const float mn = fmaxf(fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3])), m);
        sum = sum * __expf(m - mn) + (__expf(xs[0] - mn) + __expf(xs[1] - mn)) + (__expf(xs[2] - mn) + __expf(xs[3] - mn));
        m = mn;
#pragma unroll
        for (int e = 0; e < 4; ++e) list_insert<KL>(rowl, ArgMax{xs[e], 4 * i4 + e});
      }
      for (int i = 4 * V4 + tid; i < (FAST == 2 ? 0 : V); i += NT) {
        const float xv = x[i];
        const float mn = fmaxf(m, xv);
        sum = sum * __expf(m - mn) + __expf(xv - mn);
        m = mn;
        list_insert<KL>(rowl, ArgMax{xv, i});
      }
      float mb = warp_max(m);
      if (lane == 0) red_v[warp] = mb;
      __syncthreads();
      mb = red_v[0];
      for (int w = 1; w < NT / 32; ++w) mb = fmaxf(mb, red_v[w]);
      __syncthreads();
      float sb = warp_sum(m > -INFINITY ? sum * __expf(m - mb) : 0.f);   // threads without elements contribute 0
      if (lane == 0) red_v[warp] = sb;
      __syncthreads();
      sb = 0.f;
      for (int w = 0; w < NT / 32; ++w) sb += red_v[w];
      __syncthreads();
      const float lj = logf(sb), sj = score_in[g * k + j];
#pragma unroll
      for (int q = 0; q < KL; ++q)
        if (rowl[q].idx != 0x7fffffff) list_insert<KL>(mine, ArgMax{sj + ((rowl[q].v - mb) - lj), j * V + rowl[q].idx});
    }
  } else {
  // log-sum-exp of each considered row
  for (int j = 0; j < ns; ++j) {
    const float* x = base + (int64_t)j * V;
    float m = -INFINITY;
    for (int i = tid; i < V; i += NT) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if (lane == 0) red_v[warp] = m;
    __syncthreads();
    m = red_v[0];
    for (int w = 1; w < NT / 32; ++w) m = fmaxf(m, red_v[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < V; i += NT) sum += expf(x[i] - m);
    sum = warp_sum(sum);
    if (lane == 0) red_v[warp] = sum;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < NT / 32; ++w) tot += red_v[w];
      s_max[j] = m;
      s_logsum[j] = logf(tot);
      s_score[j] = score_in[g * k + j];
    }
    __syncthreads();
  }

  // top-kr of score_j + log_softmax(row j)[v] over the ns*V candidates, descending (value, then
  // smaller flat index: a strict total order, so the result does not depend on the schedule).
  // ONE pass over the logits: every thread keeps the kr best of its own candidates in registers
  // (sorted), then kr block-wide arg-max rounds pop the winners off the thread-local lists.
  // (the lists hold KL >= kr entries: a superset of what is needed, with static register indexing)
  for (int j = 0; j < ns; ++j) {
    const float* x = base + (int64_t)j * V;
    const float sj = s_score[j], mj = s_max[j], lj = s_logsum[j];
    for (int v = tid; v < V; v += NT) list_insert<KL>(mine, ArgMax{sj + ((x[v] - mj) - lj), j * V + v});
  }
  }
  for (int r = 0; r < kr; ++r) {
    ArgMax best = mine[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
      best = better(best, other);
    }
    if (lane == 0) { red_v[warp] = best.v; red_i[warp] = best.idx; }
    __syncthreads();
    ArgMax b{red_v[0], red_i[0]};
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) b = better(b, ArgMax{red_v[w], red_i[w]});
    if (tid == 0) {
      s_sel[r] = b.idx;
      s_selv[r] = b.v;
    }
    if (mine[0].idx == b.idx) {      // the owner pops its head (flat indices are unique)
#pragma unroll
      for (int q = 0; q + 1 < KL; ++q) mine[q] = mine[q + 1];
      mine[KL - 1] = ArgMax{-INFINITY, 0x7fffffff};
    }
    __syncthreads();
  }

  if (tid == 0) {
    int nlive = 0, ndone = 0;
    for (int r = 0; r < kr; ++r) {
      const int idx = s_sel[r];
      const int parent = idx / V;               // `//` restatement of :252
      const int word = idx - parent * V;
      const float sc = s_selv[r];
      if (tr_parent) {
        const int64_t o = ((int64_t)g * n_steps + t) * k + r;
        tr_parent[o] = parent; tr_word[o] = word; tr_score[o] = sc;
      }
      if (word == end_id) {
        ++ndone;
        // complete_seqs_scores.index(max(...)) (:292): the FIRST maximum in completion order
        if (!has_done[g] || sc > best_score[g]) {
          has_done[g] = 1;
          best_score[g] = sc;
          best_t[g] = t;
          best_parent[g] = parent;
        }
      } else {
        const int o = g * k + nlive;
        score_out[o] = sc;
        prev_word[o] = word;
        src_row[o] = parent;
        bp_parent[(int64_t)t * G * k + o] = parent;
        bp_word[(int64_t)t * G * k + o] = word;
        ++nlive;
      }
    }
    live[g] = nlive;
    krem[g] = kr - ndone;
  }
}

// back-track the winning caption of each image
__global__ void beam_finalize_kernel(int k, int n_steps, int P, int32_t start_id, int32_t end_id,
                                     const float* __restrict__ score, const int32_t* __restrict__ live,
                                     const int32_t* __restrict__ has_done,
                                     const float* __restrict__ best_score,
                                     const int32_t* __restrict__ best_t,
                                     const int32_t* __restrict__ best_parent,
                                     const int32_t* __restrict__ bp_parent,
                                     const int32_t* __restrict__ bp_word,
                                     const float* __restrict__ alpha_hist, int G,
                                     int32_t* __restrict__ out_seq, int32_t* __restrict__ out_len,
                                     float* __restrict__ out_score, int32_t* __restrict__ out_completed,
                                     float* __restrict__ out_alpha) {
  __shared__ int s_row[64];      // alpha source row per sequence position (n_steps + 2 <= 64)
  __shared__ int s_len;
  const int g = blockIdx.x;
  const int L = n_steps + 1;     // out_seq pitch
  if (threadIdx.x == 0) {
    int len, p, t_last;
    if (has_done[g]) {
      t_last = best_t[g];
      len = t_last + 2;
      out_seq[(int64_t)g * L + len - 1] = end_id;
      p = best_parent[g];
      s_row[len - 1] = t_last * G * k + g * k + p;
      out_score[g] = best_score[g];
      out_completed[g] = 1;
    } else {
      // upstream raises ValueError here (App. C-4); defined fallback: best live beam, first maximum
      const int s = live[g];
      int bj = 0;
      float bv = -INFINITY;
      for (int j = 0; j < s; ++j)
        if (score[g * k + j] > bv) { bv = score[g * k + j]; bj = j; }
      t_last = n_steps;          // positions 1..n_steps all come from the back-pointers
      len = n_steps + 1;
      p = bj;
      out_score[g] = bv;
      out_completed[g] = 0;
    }
    // tokens generated at steps tau = t_last-1 .. 0 live in the back-pointer tables
    for (int tau = t_last - 1; tau >= 0; --tau) {
      const int o = tau * G * k + g * k + p;
      const int pp = bp_parent[o];
      out_seq[(int64_t)g * L + tau + 1] = bp_word[o];
      s_row[tau + 1] = tau * G * k + g * k + pp;
      p = pp;
    }
    out_seq[(int64_t)g * L] = start_id;
    for (int i = len; i < L; ++i) out_seq[(int64_t)g * L + i] = 0;
    out_len[g] = len;
    s_len = len;
  }
  __syncthreads();
  if (out_alpha) {
    const int len = s_len;
    for (int pos = 0; pos < L; ++pos) {
      float* dst = out_alpha + ((int64_t)g * L + pos) * P;
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        float v = 0.f;
        if (pos == 0) v = 1.f;                            // seqs_alpha starts as ones (:204)
        else if (pos < len && alpha_hist) v = alpha_hist[(int64_t)s_row[pos] * P + i];
        dst[i] = v;
      }
    }
  }
}

}  // namespace

int beam_init(int32_t* prev_word, float* score, int32_t* live, int32_t* krem, int32_t* has_done,
              float* best_score, int32_t* best_t, int32_t* best_parent, int G, int k, int32_t start_id,
              cudaStream_t st) {
  beam_init_kernel<<<ceil_div(G * k, 256), 256, 0, st>>>(prev_word, score, live, krem, has_done,
                                                         best_score, best_t, best_parent, G, k, start_id);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int beam_embed(int precision, const float* emb, const int32_t* prev_word, void* Xe, int64_t ldx, int rows,
               int M, int V, cudaStream_t st) {
  if (precision == CAPDEC_BF16)
    beam_embed_kernel<bf16><<<rows, 128, 0, st>>>(emb, prev_word, (bf16*)Xe, ldx, M, V);
  else
    beam_embed_kernel<float><<<rows, 128, 0, st>>>(emb, prev_word, (float*)Xe, ldx, M, V);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int beam_gather_state(int precision, const void* h_new, const float* c_new, void* h_state, float* c_state,
                      const int32_t* src_row, const int32_t* live, int rows, int k, int D, int64_t ldh,
                      cudaStream_t st) {
  if (precision == CAPDEC_BF16)
    beam_gather_state_kernel<bf16><<<rows, 128, 0, st>>>((const bf16*)h_new, c_new, (bf16*)h_state, c_state,
                                                         src_row, live, k, D, ldh);
  else
    beam_gather_state_kernel<float><<<rows, 128, 0, st>>>((const float*)h_new, c_new, (float*)h_state,
                                                          c_state, src_row, live, k, D, ldh);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int beam_select(const float* logits, int V, int G, int k, int t, int32_t end_id, const float* score_in,
                float* score_out, int32_t* prev_word, int32_t* src_row, int32_t* live, int32_t* krem,
                int32_t* has_done, float* best_score, int32_t* best_t, int32_t* best_parent,
                int32_t* bp_parent, int32_t* bp_word, int32_t* tr_parent, int32_t* tr_word,
                float* tr_score, int n_steps, cudaStream_t st, int fast, const VocabTopkPlan* part) {
  CAPDEC_REQUIRE(k >= 1 && k <= KMAX, CAPDEC_ERR_BAD_SHAPE, "beam size must be 1..%d (got %d)", KMAX, k);
  VocabTopkPlan vp{0, 0, 0, 0};
  if (part) vp = *part;
#define BS_LAUNCH(KL_, FAST_)                                                                              \
  beam_select_kernel<KL_, FAST_><<<G, NT, 0, st>>>(logits, V, vp, k, t, end_id, score_in, score_out, prev_word,   \
                                                   src_row, live, krem, has_done, best_score, best_t,         \
                                                   best_parent, bp_parent, bp_word, tr_parent, tr_word,       \
                                                   tr_score, n_steps, G)
  if (part) { if (k <= 4) BS_LAUNCH(4, 2); else BS_LAUNCH(KMAX, 2); }
  else if (k <= 4) { if (fast) BS_LAUNCH(4, 1); else BS_LAUNCH(4, 0); }
  else { if (fast) BS_LAUNCH(KMAX, 1); else BS_LAUNCH(KMAX, 0); }
#undef BS_LAUNCH
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

int beam_finalize(int G, int k, int n_steps, int P, int32_t start_id, int32_t end_id, const float* score,
                  const int32_t* live, const int32_t* has_done, const float* best_score,
                  const int32_t* best_t, const int32_t* best_parent, const int32_t* bp_parent,
                  const int32_t* bp_word, const float* alpha_hist, int32_t* out_seq, int32_t* out_len,
                  float* out_score, int32_t* out_completed, float* out_alpha, cudaStream_t st) {
  CAPDEC_REQUIRE(n_steps + 2 <= 64, CAPDEC_ERR_BAD_SHAPE, "max_steps too large (%d)", n_steps);
  beam_finalize_kernel<<<G, 64, 0, st>>>(k, n_steps, P, start_id, end_id, score, live, has_done,
                                         best_score, best_t, best_parent, bp_parent, bp_word, alpha_hist, G,
                                         out_seq, out_len, out_score, out_completed, out_alpha);
  CAPDEC_LAUNCH_OK();
  return CAPDEC_OK;
}

}  // namespace capdec
