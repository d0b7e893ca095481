// capi.cu -- the extern "C" surface of libcapdec.so (see include/capdec.h).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* s = getenv("CAPDEC_PDL");
    return !(s && s[0] == '0');
  }();
  return on;
}

static std::once_flag g_init_once;
static int g_init_rc = CAPDEC_OK;

static int do_init() {
  int dev = 0;
  CAPDEC_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CAPDEC_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  CAPDEC_REQUIRE(prop.major == 10, CAPDEC_ERR_UNSUPPORTED,
                 "libcapdec is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  CAPDEC_TRY(attention_init());
  CAPDEC_TRY(gemm_tc_init());
  return CAPDEC_OK;
}

int beam_search(const CapdecDims& d, const CapdecParams& w, const float* enc, int64_t enc_sb, int64_t enc_sp,
                int64_t enc_se, const float* tags, int G,
                int k, int max_steps, int32_t start_id, int32_t end_id, int32_t* out_seq,
                int32_t* out_len, float* out_score, int32_t* out_completed, float* out_alpha,
                int32_t* trace_parent, int32_t* trace_word, float* trace_score, void* workspace,
                size_t ws_bytes, cudaStream_t st);
size_t beam_workspace_bytes(const CapdecDims& d, int G, int k, int max_steps);

}  // namespace capdec

using namespace capdec;

extern "C" {

int capdec_version(void) { return CAPDEC_VERSION; }
const char* capdec_last_error(void) { return get_error(); }

unsigned long long capdec_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int capdec_topk_hits(const float* scores, int64_t ld, const int64_t* targets, const int64_t* caps_sorted,
                     const int32_t* decode_len_d, int rows, int T, int L, int V, int k, int32_t* hits_out,
                     void* stream) {
  CAPDEC_TRY(capdec_init());
  return topk_hits(scores, ld, targets, caps_sorted, decode_len_d, rows, T, L, V, k, hits_out, (cudaStream_t)stream);
}

int capdec_clip_adam_step(const CapdecAdamSeg* segs, int n_segs, double lr, double beta1, double beta2, double eps,
                          double weight_decay, double grad_clip, int step, int write_clipped, void* stream) {
  CAPDEC_TRY(capdec_init());
  return clip_adam_step(segs, n_segs, lr, beta1, beta2, eps, weight_decay, grad_clip, step, write_clipped,
                        (cudaStream_t)stream);
}

int capdec_dropout_mask(uint64_t dropout_seed, float dropout_p, int64_t n, float* mask_out, void* stream) {
  CAPDEC_TRY(capdec_init());
  return dropout_mask(dropout_seed, dropout_p, n, mask_out, (cudaStream_t)stream);
}

void capdec_recur_timing(int enable) { recur_timing(enable); }
float capdec_recur_last_ms(int which) { return recur_last_ms(which); }

int capdec_init(void) {
  std::call_once(g_init_once, [] { g_init_rc = do_init(); });
  return g_init_rc;
}

size_t capdec_workspace_bytes(const CapdecDims* dims, int with_backward) {
  if (!dims) return 0;
  return workspace_bytes(*dims, with_backward);
}

int capdec_forward_train(const CapdecDims* dims, const CapdecParams* params, const float* enc,
                         int64_t enc_sb, int64_t enc_sp, int64_t enc_se, const int64_t* sort_ind,
                         const float* tags, const int64_t* caps_sorted, const int32_t* decode_len_h,
                         float dropout_p, uint64_t dropout_seed, int save_for_backward, int phases,
                         float* predictions, float* alphas, void* workspace, size_t workspace_bytes,
                         void* stream) {
  CAPDEC_REQUIRE(dims && params && predictions && workspace, CAPDEC_ERR_BAD_ARG, "capdec_forward_train: null argument");
  CAPDEC_REQUIRE((phases != 0 && !(phases & 1)) || (enc && caps_sorted && decode_len_h), CAPDEC_ERR_BAD_ARG,
                 "capdec_forward_train: the input phase needs enc, caps_sorted and decode_len_h");
  CAPDEC_REQUIRE(dims->kind == CAPDEC_PURE_SCN || alphas, CAPDEC_ERR_BAD_ARG, "alphas output is NULL");
  CAPDEC_REQUIRE(dims->kind == CAPDEC_PURE_ATTENTION || tags || (phases != 0 && !(phases & 1)), CAPDEC_ERR_BAD_ARG,
                 "tags is NULL");
  CAPDEC_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, CAPDEC_ERR_BAD_ARG, "dropout_p out of range");
  CAPDEC_TRY(capdec_init());
  return forward_train(*dims, *params, enc, enc_sb, enc_sp, enc_se, sort_ind, tags, caps_sorted,
                       decode_len_h, dropout_p, dropout_seed, save_for_backward, phases ? phases : 3, predictions,
                       dims->kind == CAPDEC_PURE_SCN ? nullptr : alphas, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

int capdec_backward(const CapdecDims* dims, const CapdecParams* params,
                    const int32_t* decode_len_h, float dropout_p,
                    const float* d_predictions, const void* d_logits_ft,
                    const float* d_alphas, const float* alphas, const CapdecParams* grads,
                    void* workspace, size_t workspace_bytes, int phases, void* stream) {
  CAPDEC_REQUIRE(dims && params && grads && workspace, CAPDEC_ERR_BAD_ARG, "capdec_backward: null argument");
  CAPDEC_REQUIRE(phases >= 0 && phases <= 63, CAPDEC_ERR_BAD_ARG, "capdec_backward: bad phases");
  CAPDEC_REQUIRE(dims->kind == CAPDEC_PURE_SCN || alphas, CAPDEC_ERR_BAD_ARG, "alphas is NULL");
  CAPDEC_TRY(capdec_init());
  return backward(*dims, *params, decode_len_h, dropout_p,
                  d_predictions, d_logits_ft, d_alphas, alphas, *grads, workspace, workspace_bytes, phases,
                  (cudaStream_t)stream);
}

int capdec_loss_fwd(const CapdecDims* dims, const float* predictions, const float* alphas,
                    const int64_t* caps_sorted, const int32_t* decode_len_d, int32_t n_tokens,
                    float alpha_c, float* loss_out, float* lse_out, void* stream) {
  CAPDEC_REQUIRE(dims && predictions && caps_sorted && decode_len_d && loss_out && lse_out &&
                     n_tokens > 0,
                 CAPDEC_ERR_BAD_ARG, "capdec_loss_fwd: bad argument");
  return loss_fwd(*dims, predictions, alphas, caps_sorted, decode_len_d, n_tokens, alpha_c, loss_out,
                  lse_out, (cudaStream_t)stream);
}

int capdec_loss_bwd(const CapdecDims* dims, const float* predictions, const float* alphas,
                    const int64_t* caps_sorted, const int32_t* decode_len_d, int32_t n_tokens,
                    float alpha_c, float gscale, const float* gscale_dev, const float* lse,
                    float* d_predictions, void* d_logits_ft, float* d_alphas, void* stream) {
  CAPDEC_REQUIRE(dims && predictions && caps_sorted && decode_len_d && lse && n_tokens > 0,
                 CAPDEC_ERR_BAD_ARG, "capdec_loss_bwd: bad argument");
  return loss_bwd(*dims, predictions, alphas, caps_sorted, decode_len_d, n_tokens, alpha_c, gscale, gscale_dev,
                  lse, d_predictions, d_logits_ft, round_up(dims->V, 8), d_alphas, (cudaStream_t)stream);
}

size_t capdec_beam_workspace_bytes(const CapdecDims* dims, int G, int k, int max_steps) {
  if (!dims) return 0;
  return beam_workspace_bytes(*dims, G, k, max_steps);
}

int capdec_beam_search_strided(const CapdecDims* dims, const CapdecParams* params, const float* enc,
                               int64_t enc_sb, int64_t enc_sp, int64_t enc_se,
                               const float* tags, int G, int k, int max_steps, int32_t start_id,
                               int32_t end_id, int32_t* out_seq, int32_t* out_len, float* out_score,
                               int32_t* out_completed, float* out_alpha, int32_t* trace_parent,
                               int32_t* trace_word, float* trace_score, void* workspace,
                               size_t workspace_bytes, void* stream) {
  CAPDEC_REQUIRE(dims && params && enc && out_seq && out_len && out_score && out_completed && workspace,
                 CAPDEC_ERR_BAD_ARG, "capdec_beam_search: null argument");
  CAPDEC_TRY(capdec_init());
  return beam_search(*dims, *params, enc, enc_sb, enc_sp, enc_se, tags, G, k, max_steps, start_id, end_id, out_seq,
                     out_len, out_score, out_completed, out_alpha, trace_parent, trace_word, trace_score,
                     workspace, workspace_bytes, (cudaStream_t)stream);
}

int capdec_beam_search(const CapdecDims* dims, const CapdecParams* params, const float* enc,
                       const float* tags, int G, int k, int max_steps, int32_t start_id,
                       int32_t end_id, int32_t* out_seq, int32_t* out_len, float* out_score,
                       int32_t* out_completed, float* out_alpha, int32_t* trace_parent,
                       int32_t* trace_word, float* trace_score, void* workspace,
                       size_t workspace_bytes, void* stream) {
  CAPDEC_REQUIRE(dims, CAPDEC_ERR_BAD_ARG, "capdec_beam_search: null argument");
  return capdec_beam_search_strided(dims, params, enc, (int64_t)dims->P * dims->E, dims->E, 1, tags, G, k, max_steps,
                                    start_id, end_id, out_seq, out_len, out_score, out_completed, out_alpha,
                                    trace_parent, trace_word, trace_score, workspace, workspace_bytes, stream);
}

int capdec_gemm(int precision, const void* X, int64_t ldx, const void* W, int64_t ldw, void* out,
                int64_t ldo, int out_ft, const float* bias, const float* addm, int64_t ldadd, int rows,
                int N, int K, int batch, int64_t sX, int64_t sW, int64_t sO, int splitk, void* stream) {
  CAPDEC_TRY(capdec_init());
  GemmArgs a;
  a.splitk = splitk;
  a.X = X; a.ldx = ldx; a.W = W; a.ldw = ldw; a.out = out; a.ldo = ldo; a.out_ft = out_ft;
  a.bias = bias; a.addm = addm; a.ldadd = ldadd; a.rows = rows; a.N = N; a.K = K;
  a.batch = batch < 1 ? 1 : batch; a.sX = sX; a.sW = sW; a.sO = sO;
  return gemm(precision, a, (cudaStream_t)stream);
}

int capdec_gemm_tn(const void* XT, int64_t ldx, const void* WT, int64_t ldw, float* out, int64_t ldo,
                   int rows, int N, int K, int batch, int64_t sX, int64_t sW, int64_t sO, void* stream) {
  CAPDEC_REQUIRE(XT && WT && out && rows > 0 && N > 0 && K > 0, CAPDEC_ERR_BAD_ARG, "capdec_gemm_tn: bad argument");
  CAPDEC_TRY(capdec_init());
  GemmArgs a;
  a.tn = 3;
  a.X = XT; a.ldx = ldx; a.W = WT; a.ldw = ldw; a.out = out; a.ldo = ldo; a.rows = rows; a.N = N; a.K = K;
  a.batch = batch < 1 ? 1 : batch; a.sX = sX; a.sW = sW; a.sO = sO;
  return gemm(CAPDEC_BF16, a, (cudaStream_t)stream);
}

size_t capdec_attention_scratch_floats(int precision, int rows, int P, int E) {
  return attention_scratch_floats(precision, rows, P, E);
}

int capdec_attention_step(int precision, const void* att1, const void* enc, const float* g1,
                          int64_t ldg, int beta_col, const float* w_f, const float* b_f,
                          float* alpha_out, int64_t alpha_stride, void* z_out, float* awe_out, int rows,
                          int rows_per_map, int P, int E, int A, float* scratch, void* stream) {
  CAPDEC_REQUIRE(att1 && enc && g1 && w_f && b_f && scratch && rows_per_map >= 1, CAPDEC_ERR_BAD_ARG,
                 "capdec_attention_step: bad argument");
  CAPDEC_TRY(capdec_init());
  return attention_fwd(precision, att1, enc, g1, ldg, beta_col, w_f, b_f, alpha_out, alpha_stride, z_out,
                       E, awe_out, rows, rows_per_map, P, E, A, scratch, (cudaStream_t)stream);
}

int capdec_attention_bwd_step(int precision, const void* att1, const void* enc, const float* g1,
                              int64_t ldg, int beta_col, const float* w_f, const float* alpha,
                              int64_t alpha_stride, const float* dalpha_ext, int64_t dalpha_stride,
                              const float* dz, const float* awe, void* dba, int64_t lddba, float* dAtt1,
                              float* dwf_part, float* dbf_part, int rows, int P, int E, int A,
                              float* scratch, void* stream) {
  CAPDEC_REQUIRE(att1 && enc && g1 && w_f && alpha && dz && awe && dba && dAtt1 && dwf_part && dbf_part &&
                     scratch,
                 CAPDEC_ERR_BAD_ARG, "capdec_attention_bwd_step: null argument");
  CAPDEC_TRY(capdec_init());
  // scratch = [partial dalpha | de (rows x pad4(P))]
  float* de = scratch + attention_scratch_floats(precision, rows, P, E);
  CAPDEC_TRY(attention_bwd(precision, att1, enc, g1, ldg, beta_col, w_f, alpha, alpha_stride, dalpha_ext,
                           dalpha_stride, dz, E, awe, dba, lddba, de, dwf_part, dbf_part, rows, P, E, A,
                           scratch, (cudaStream_t)stream));
  return attention_datt1(precision, att1, g1, ldg, 0, de, 0, w_f, dAtt1, 1, rows, 1, P, A,
                         (cudaStream_t)stream);
}

// ---- SCNCell.forward on fp32 master weights (unit entry; models/scn_cell.py:52-154) ----
namespace {
struct CellPlan {
  int fsz;
  int64_t ldX, ldS, ldD, ld2F;
  size_t WiaT, WibT, WhaT, WhbT, Wc, xF, sF, hF, u, v, p, q, m, pre, hO, total;
};
CellPlan cell_plan(int precision, int n, int X, int D, int F, int S) {
  CellPlan c;
  c.fsz = precision == CAPDEC_BF16 ? 2 : 4;
  c.ldX = round_up(X, 8); c.ldS = round_up(S, 8); c.ldD = round_up(D, 8); c.ld2F = round_up(2 * F, 8);
  size_t cur = 0;
  auto take = [&](size_t b) { size_t at = cur; cur += (size_t)round_up((int64_t)b, 256); return at; };
  const size_t f = c.fsz;
  c.WiaT = take((size_t)4 * F * c.ldX * f); c.WibT = take((size_t)4 * F * c.ldS * f);
  c.WhaT = take((size_t)4 * F * c.ldD * f); c.WhbT = take((size_t)4 * F * c.ldS * f);
  c.Wc = take((size_t)4 * D * c.ld2F * f);
  c.xF = take((size_t)n * c.ldX * f); c.sF = take((size_t)n * c.ldS * f); c.hF = take((size_t)n * c.ldD * f);
  c.u = take((size_t)n * 4 * F * 4); c.v = take((size_t)n * 4 * F * 4);
  c.p = take((size_t)n * 4 * F * 4); c.q = take((size_t)n * 4 * F * 4);
  c.m = take((size_t)4 * n * 2 * F * f); c.pre = take((size_t)n * 4 * D * 4);
  c.hO = take((size_t)n * c.ldD * f);
  c.total = cur;
  return c;
}
}  // namespace

size_t capdec_scn_cell_workspace_bytes(int precision, int rows, int X, int D, int F, int S) {
  if (rows <= 0 || X <= 0 || D <= 0 || F <= 0 || S <= 0) return 0;
  return cell_plan(precision, rows, X, D, F, S).total;
}

int capdec_scn_cell_step(int precision, int rows, int X, int D, int F, int S, const float* w_ia,
                         const float* w_ib, const float* w_ic, const float* w_ha, const float* w_hb,
                         const float* w_hc, const float* b_ih, const float* b_hh, const float* x,
                         const float* s, const float* h, const float* c, float* h_out, float* c_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  CAPDEC_REQUIRE(w_ia && w_ib && w_ic && w_ha && w_hb && w_hc && x && s && h && c && h_out && c_out &&
                     workspace,
                 CAPDEC_ERR_BAD_ARG, "capdec_scn_cell_step: null argument");
  CAPDEC_REQUIRE(D % 8 == 0 && F % 8 == 0, CAPDEC_ERR_BAD_SHAPE,
                 "hidden_size and factor_size must be multiples of 8 (D=%d F=%d)", D, F);
  CAPDEC_TRY(capdec_init());
  const CellPlan cp = cell_plan(precision, rows, X, D, F, S);
  CAPDEC_REQUIRE(workspace_bytes >= cp.total, CAPDEC_ERR_WORKSPACE, "workspace %zu < %zu",
                 workspace_bytes, cp.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int pr = precision, NQ = 4 * F, n = rows;
  CAPDEC_TRY(transpose_cast(pr, w_ia, 0, ws + cp.WiaT, 1, 1, X, NQ, 0, NQ, cp.ldX, 0, 1, st));
  CAPDEC_TRY(transpose_cast(pr, w_ib, 0, ws + cp.WibT, 1, 1, S, NQ, 0, NQ, cp.ldS, 0, 1, st));
  CAPDEC_TRY(transpose_cast(pr, w_ha, 0, ws + cp.WhaT, 1, 1, D, NQ, 0, NQ, cp.ldD, 0, 1, st));
  CAPDEC_TRY(transpose_cast(pr, w_hb, 0, ws + cp.WhbT, 1, 1, S, NQ, 0, NQ, cp.ldS, 0, 1, st));
  for (int g = 0; g < 4; ++g) {
    CAPDEC_TRY(copy_cast(pr, w_ic + g * F, 0, NQ, ws + cp.Wc + (size_t)g * D * cp.ld2F * cp.fsz, 1,
                         cp.ld2F, D, F, st));
    CAPDEC_TRY(copy_cast(pr, w_hc + g * F, 0, NQ,
                         ws + cp.Wc + ((size_t)g * D * cp.ld2F + F) * cp.fsz, 1, cp.ld2F, D, F, st));
  }
  CAPDEC_TRY(copy_cast(pr, x, 0, X, ws + cp.xF, 1, cp.ldX, n, X, st));
  CAPDEC_TRY(copy_cast(pr, s, 0, S, ws + cp.sF, 1, cp.ldS, n, S, st));
  CAPDEC_TRY(copy_cast(pr, h, 0, D, ws + cp.hF, 1, cp.ldD, n, D, st));
  auto gm = [&](const void* Xp, int64_t ldx, const void* Wp, int64_t ldw, void* out, int64_t ldo, int N,
                int K, int batch = 1, int64_t sX = 0, int64_t sW = 0, int64_t sO = 0) {
    GemmArgs a;
    a.X = Xp; a.ldx = ldx; a.W = Wp; a.ldw = ldw; a.out = out; a.ldo = ldo; a.rows = n; a.N = N;
    a.K = K; a.batch = batch; a.sX = sX; a.sW = sW; a.sO = sO;
    return gemm(pr, a, st);
  };
  CAPDEC_TRY(gm(ws + cp.xF, cp.ldX, ws + cp.WiaT, cp.ldX, ws + cp.u, NQ, NQ, X));
  CAPDEC_TRY(gm(ws + cp.sF, cp.ldS, ws + cp.WibT, cp.ldS, ws + cp.v, NQ, NQ, S));
  CAPDEC_TRY(gm(ws + cp.hF, cp.ldD, ws + cp.WhaT, cp.ldD, ws + cp.p, NQ, NQ, D));
  CAPDEC_TRY(gm(ws + cp.sF, cp.ldS, ws + cp.WhbT, cp.ldS, ws + cp.q, NQ, NQ, S));
  CAPDEC_TRY(scn_form_m(pr, (float*)(ws + cp.u), NQ, (float*)(ws + cp.p), NQ, (float*)(ws + cp.v),
                        (float*)(ws + cp.q), ws + cp.m, n, n, F, st));
  CAPDEC_TRY(gm(ws + cp.m, 2 * F, ws + cp.Wc, cp.ld2F, ws + cp.pre, 4 * D, D, 2 * F, 4, (int64_t)n * 2 * F,
                (int64_t)D * cp.ld2F, D));
  CAPDEC_TRY(cell_fwd(pr, (float*)(ws + cp.pre), 4 * D, nullptr, 0, b_ih, b_hh, 0, c, c_out, nullptr,
                      ws + cp.hO, cp.ldD, nullptr, 0.f, nullptr, 0, 1, n, D, st));
  CAPDEC_TRY(copy_cast(pr, ws + cp.hO, 1, cp.ldD, h_out, 0, D, n, D, st));
  return CAPDEC_OK;
}

}  // extern "C"
