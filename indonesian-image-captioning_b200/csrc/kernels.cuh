// kernels.cuh -- host-callable launchers shared between the translation units of libcapdec.
#pragma once
#include "common.cuh"

namespace capdec {

// ---- attention.cu ----
int attention_init();
// scratch floats both directions need for `rows` rows (forward: scores; backward: partial dalpha)
size_t attention_scratch_floats(int precision, int rows, int P, int E);
int attention_fwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* b_f, float* alpha_out,
                  int64_t alpha_stride, void* z_out, int64_t ldz, float* awe_out, int rows,
                  int rows_per_map, int P, int E, int A, float* scratch, cudaStream_t st,
                  const void* enc_cm = nullptr);      // optional chunk-major feature copy [map][E/512][P][512]
// dst[b][c][p][j] = src[b][p][c*cw + j]   (bf16; chunk-major copies for one-bulk-copy staging fills)
int chunk_major_copy(const void* src, void* dst, int B, int P, int E, int cw, cudaStream_t st);
// de_out (rows, pad4(P)) receives the softmax-input gradient of every pixel (consumed by attention_datt1)
int attention_bwd(int precision, const void* att1, const void* enc, const float* g1, int64_t ldg,
                  int beta_col, const float* w_f, const float* alpha, int64_t alpha_stride,
                  const float* dalpha_ext, int64_t dalpha_stride, const float* dz, int64_t lddz,
                  const float* awe, void* dba, int64_t lddba, float* de_out, float* dwf_part,
                  float* dbf_part, int rows, int P, int E, int A, float* scratch, cudaStream_t st);
// dAtt1[b,p,a] (+)= w_f[a] sum_t de[t,b,p] 1[att1[b,p,a] + att2[t,b,a] > 0] ; att2[t][b] = g1 + t*g1_step + b*ldg,
// de[t][b] = de + t*de_step + b*pad4(P)
int attention_datt1(int precision, const void* att1, const float* g1, int64_t ldg, int64_t g1_step,
                    const float* de, int64_t de_step, const float* w_f, float* dAtt1, int accumulate,
                    int B, int T, int P, int A, cudaStream_t st);

// ---- pointwise.cu ----
// dst[c*ldd + i*d_i + j*d_j] = src[i*s_i + j*s_j + c]   for i<ni, j<nj, c<C   (transpose + cast)
// src_ft/dst_ft: 0 = fp32, 1 = the feature type of `precision`
int transpose_cast(int precision, const void* src, int src_ft, void* dst, int dst_ft, int ni, int nj,
                   int C, int64_t s_i, int64_t s_j, int64_t ldd, int64_t d_i, int64_t d_j,
                   cudaStream_t st);
// dst[r*ldd + c] = src[r*lds + c]  (cast), r<R, c<C
int copy_cast(int precision, const void* src, int src_ft, int64_t lds, void* dst, int dst_ft,
              int64_t ldd, int R, int C, cudaStream_t st);
// One launch for many fp32 -> feature-type matrix copies / transposes (the per-call weight packing):
//   transpose == 0: dst[r*ldd + c] = src[r*lds + c]     transpose == 1: dst[c*ldd + r] = src[r*lds + c]
struct PackSeg { const float* src; void* dst; int R, C; int64_t lds, ldd; int transpose; int tile0; };
struct PackTable {
  PackSeg seg[40];
  int n = 0;
  bool add(const float* src, int64_t lds, void* dst, int64_t ldd, int R, int C, int transpose) {
    if (n >= 40 || R <= 0 || C <= 0) return R <= 0 || C <= 0;
    seg[n++] = PackSeg{src, dst, R, C, lds, ldd, transpose, 0};
    return true;
  }
};
int pack_multi(int precision, PackTable& t, cudaStream_t st);
// out[n] (=|+=) sum_r X[r*ld + n]
int colsum(int precision, const void* X, int x_ft, int64_t ld, int R, int N, float* out,
           int accumulate, cudaStream_t st);
// gather rows by sort_ind, convert to the feature type, and mean over pixels
int gather_features(int precision, const float* enc, int64_t sb, int64_t sp, int64_t se,
                    const int64_t* sort_ind, void* enc_s, float* mean_f32, void* mean_ft,
                    int64_t ld_mean_ft, int B, int P, int E, cudaStream_t st,
                    void* enc_cm = nullptr, int cw = 0);   // optional chunk-major copy [B][E/cw][P][cw] in the same pass
// Xe[(t*B + b)*ldx + :] = emb[caps[b*L + t]]
int embedding_gather(int precision, const float* emb, const int64_t* caps, int L, void* Xe,
                     int64_t ldx, int B, int T, int M, int V, cudaStream_t st);
// dEmb[caps[b][t]] += dXe[(t*B+b)] for t < len[b]
int embedding_scatter_add(const float* dXe, int64_t ldx, const int64_t* caps, int L,
                          const int32_t* len_d, float* dEmb, int B, int T, int M, int V,
                          cudaStream_t st);
// SCN: m[g][b][0:F] = u*v, m[g][b][F:2F] = p*q
int scn_form_m(int precision, const float* u, int64_t ldu, const float* p, int64_t ldp,
               const float* v, const float* q, void* m, int rows, int B, int F, cudaStream_t st);
// LSTM pointwise.  pre = preA + preB + b1 + b2 ; lstm_order: 0 = (i,f,o,c) SCN, 1 = (i,f,g,o) torch
int cell_fwd(int precision, const float* preA, int64_t ldA, const float* preB, int64_t ldB,
             const float* b1, const float* b2, int lstm_order, const float* c_prev, float* c_new,
             float* gates, void* h_out, int64_t ldh, void* hd_out, float dropout_p, const uint64_t* seed,
             int t, int T, int rows, int D, cudaStream_t st);
int cell_bwd(int precision, const float* dh_fc, int64_t ld_dhfc, const float* dh_rec, float* dc,
             const float* gates, const float* c_prev, const float* c_new, int lstm_order,
             float dropout_p, const uint64_t* seed, int t, int T, void* dpre, float* dpre_f32, int rows,
             int D, cudaStream_t st);
// SCN backward pointwise: du = w*v, dp = r*q (row pitch lddp), dv_acc += w*u, dq_acc += r*p ;
// wr = [g][b][w(F)|r(F)]
int scn_bwd_products(int precision, const float* wr, const float* u, int64_t ldu, const float* p,
                     int64_t ldp, const float* v, const float* q, void* du, void* dp,
                     float* dv_acc, float* dq_acc, int rows, int B, int F, int64_t lddp, cudaStream_t st);
// out[i] = a[i] (+ b[i]) ; small helpers
int concat_bias(float* dst, const float* a, int na, const float* b, int nb, int nzero,
                cudaStream_t st);
int zero_rows_beyond_len(float* x, const int32_t* len_d, int B, int T, int64_t row_elems,
                         cudaStream_t st);

// dst[(g*k + j)*ldd + c] = src[g*lds + c]   (feature type both sides; beam rows share their image's row)
int expand_rows(int precision, const void* src, int64_t lds, void* dst, int64_t ldd, int G, int k, int C,
                cudaStream_t st);

// ---- recur.cu: persistent (one cooperative launch for all T steps) SCN decoder recurrence, bf16 ----
struct RecurFwdArgs {
  int att = 0;                       // 1: attention_scn / pure_attention, 0: pure_scn
  int lstm = 0;                      // 1: pure_attention (nn.LSTMCell, F unused)
  int B = 0, T = 0, P = 0, E = 0, A = 0, M = 0, D = 0, F = 0;
  const int32_t* len = nullptr;      // device [B]
  const void* Wcat1 = nullptr; int64_t ldD = 0;
  const void* Wxz = nullptr; int64_t ldX = 0;      // W_ia[M:, :]^T, i.e. Wp_xq + M
  const void* Wc = nullptr; int64_t ld2F = 0;
  const float* b_cat1 = nullptr; const float* b_ih = nullptr; const float* b_hh = nullptr;
  const void* att1 = nullptr; const void* enc = nullptr; const float* w_f = nullptr; const float* b_f = nullptr;
  const float* v = nullptr; const float* q = nullptr;
  const void* H0 = nullptr; int64_t ldH0 = 0;
  void* Hall = nullptr; void* Hd = nullptr;
  void* Ht = nullptr;                // [T][B][D] time-major copy of h (G1 operand of the next step)
  void* zk = nullptr;                // [T][E/512][B][512] chunk-major copy of z (P3 operand)
  void* enc_cm = nullptr;            // [B][E/256][P][256] chunk-major copy of enc (built by recur_fwd)
  float* C = nullptr; float* U = nullptr; float* g1 = nullptr; float* alphas = nullptr; float* awe = nullptr;
  void* z = nullptr; void* m = nullptr; float* pre = nullptr; float* gates = nullptr;
  float* scores = nullptr;           // [T][B][pad4(P)] per-step attention scores (exchange buffer)
  int ragged = 0;                    // some caption is shorter than T (or the lengths are unknown to the host)
  unsigned* bar = nullptr;           // 256 bytes: abort flag of the dataflow polls (zeroed by recur_fwd)
  float dropout_p = 0.f; const uint64_t* seed = nullptr;
};
bool recur_fwd_supported(const RecurFwdArgs& a);     // shape / device / CAPDEC_PERSISTENT check
int recur_fwd(const RecurFwdArgs& a, cudaStream_t st);
struct RecurBwdArgs {
  int att = 0, lstm = 0;
  int B = 0, T = 0, P = 0, E = 0, A = 0, M = 0, D = 0, F = 0;
  int64_t ldPX = 0;
  const int32_t* len = nullptr;
  const void* WcT = nullptr; int64_t ldD = 0;       // Wp_cT
  const void* Wxin = nullptr; int64_t ldNQ = 0;     // Wp_xin + M rows
  const void* Whx = nullptr; int64_t ldhx = 0;      // Wp_hx (LSTM: Wp_hq)
  const void* Whx2 = nullptr; int64_t ldhx2 = 0;    // LSTM: Wp_b6
  void* dbx = nullptr; int64_t ldbx = 0; int dbx_off = 0;   // destination of [dbeta_pre | datt2]
  const float* dHfc = nullptr; const float* gates = nullptr; const float* C = nullptr;
  float* dc = nullptr; float* dh_rec = nullptr;     // [B][D] each: out, gradients of c_0 / h_0
  float* dhp = nullptr;              // [T][KH/512][B][D] K-chunk partials of the recurrent gradient (exchange buffer)
  void* dpre = nullptr; void* dpre_gm = nullptr;
  const float* U = nullptr; const float* g1 = nullptr; const float* v = nullptr; const float* q = nullptr;
  void* du = nullptr; void* duk = nullptr; void* dpx = nullptr; void* dpxk = nullptr;
  float* dv_acc = nullptr; float* dq_acc = nullptr; float* dz = nullptr;
  const float* awe = nullptr; const float* alphas = nullptr; const float* d_alphas = nullptr;
  const void* enc_cm = nullptr; const void* att1 = nullptr; const float* w_f = nullptr;
  void* att1_cm = nullptr;           // [B][A/64][P][64] eighth-major copy of att1 (built by recur_bwd)
  const void* enc = nullptr; int build_enc_cm = 0;   // build enc_cm from enc first (forward ran per-step kernels)
  float* part = nullptr;             // [T][B][E/256][pad4(P)] per-step partial dalpha (exchange buffer)
  float* de = nullptr; float* dwf = nullptr; float* dbf = nullptr;
  unsigned* bar = nullptr; float dropout_p = 0.f; const uint64_t* seed = nullptr;
};
bool recur_bwd_supported(const RecurBwdArgs& a);
int recur_bwd(const RecurBwdArgs& a, cudaStream_t st);
void recur_timing(int enable);
float recur_last_ms(int which);

// ---- beam.cu ----
int beam_init(int32_t* prev_word, float* score, int32_t* live, int32_t* krem, int32_t* has_done,
              float* best_score, int32_t* best_t, int32_t* best_parent, int G, int k, int32_t start_id,
              cudaStream_t st);
int beam_embed(int precision, const float* emb, const int32_t* prev_word, void* Xe, int64_t ldx, int rows,
               int M, int V, cudaStream_t st);
int beam_gather_state(int precision, const void* h_new, const float* c_new, void* h_state, float* c_state,
                      const int32_t* src_row, const int32_t* live, int rows, int k, int D, int64_t ldh,
                      cudaStream_t st);
int beam_select(const float* logits, int V, int G, int k, int t, int32_t end_id, const float* score_in,
                float* score_out, int32_t* prev_word, int32_t* src_row, int32_t* live, int32_t* krem,
                int32_t* has_done, float* best_score, int32_t* best_t, int32_t* best_parent,
                int32_t* bp_parent, int32_t* bp_word, int32_t* tr_parent, int32_t* tr_word,
                float* tr_score, int n_steps, cudaStream_t st, int fast = 0,     // fast: single-pass (bf16 mode)
                const VocabTopkPlan* part = nullptr);   // non-null: `logits` is the partial buffer of gemm_tc_vocab_topk
// fused vocabulary projection + log-softmax statistics + top-kl candidates per 128-entry vocabulary tile (gemm_tc.cu)
size_t vocab_topk_part_floats(int rows, int V, int kl);
int gemm_tc_vocab_topk(const void* H, int64_t ldh, int rows, const void* Wfc, int64_t ldw, int V, int K,
                       const float* bias, float* part, int kl, cudaStream_t st);
int beam_finalize(int G, int k, int n_steps, int P, int32_t start_id, int32_t end_id, const float* score,
                  const int32_t* live, const int32_t* has_done, const float* best_score,
                  const int32_t* best_t, const int32_t* best_parent, const int32_t* bp_parent,
                  const int32_t* bp_word, const float* alpha_hist, int32_t* out_seq, int32_t* out_len,
                  float* out_score, int32_t* out_completed, float* out_alpha, cudaStream_t st);

int topk_hits(const float* scores, int64_t ld, const int64_t* targets, const int64_t* caps, const int32_t* len_d,
              int rows, int T, int L, int V, int k, int* hits, cudaStream_t st);

int dropout_mask(uint64_t seed, float p, int64_t n, float* out, cudaStream_t st);

// ---- optim.cu ----
int clip_adam_step(const CapdecAdamSeg* segs, int n_segs, double lr, double beta1, double beta2, double eps,
                   double weight_decay, double grad_clip, int step, int write_clipped, cudaStream_t st);

// ---- loss.cu ----
int loss_fwd(const CapdecDims& d, const float* pred, const float* alphas, const int64_t* caps,
             const int32_t* len_d, int n_tokens, float alpha_c, float* loss_out, float* lse_out,
             cudaStream_t st);
int loss_bwd(const CapdecDims& d, const float* pred, const float* alphas, const int64_t* caps,
             const int32_t* len_d, int n_tokens, float alpha_c, float gscale, const float* gscale_dev,
             const float* lse, float* d_pred, void* d_logits_ft, int64_t ldq, float* d_alphas,
             cudaStream_t st);

}  // namespace capdec

// ---- decoder.cu ----
namespace capdec {
size_t workspace_bytes(const CapdecDims& d, int with_bwd);
int forward_train(const CapdecDims& d, const CapdecParams& w, const float* enc, int64_t sb, int64_t sp,
                  int64_t se, const int64_t* sort_ind, const float* tags, const int64_t* caps,
                  const int32_t* len_h, float dropout_p, uint64_t seed, int save_bwd, int phases,
                  float* predictions, float* alphas, void* workspace, size_t ws_bytes,
                  cudaStream_t st);
int backward(const CapdecDims& d, const CapdecParams& w,
             const int32_t* len_h, float dropout_p, const float* d_pred,
             const void* d_logits_ft, const float* d_alphas, const float* alphas,
             const CapdecParams& g, void* workspace, size_t ws_bytes, int phases, cudaStream_t st);
}  // namespace capdec
