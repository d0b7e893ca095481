/*
 * capdec.h -- C ABI of libcapdec.so, the B200 (sm_100a) caption-decoder hot path.
 *
 * This is the drop-in boundary for ONE path of rayandrew/indonesian-image-captioning:
 * the caption decoder (SCN-LSTM cell + Bahdanau soft attention + vocab projection +
 * masked cross-entropy; teacher-forced forward/backward and beam search).  The
 * reference is pure Python/PyTorch and has no FFI; the "interface each entry point
 * replaces" is therefore the reference Python method whose tensor math it takes over
 * (paths relative to the reference root):
 *
 *   capdec_forward_train   models/decoders/attention_scn.py:95-158   AttentionSCN.forward
 *                          models/decoders/pure_scn.py:87-140        PureSCN.forward
 *                          models/decoders/pure_attention.py:90-151  PureAttention.forward
 *                          (+ models/scn_cell.py:52-154, models/attention.py:26-44)
 *   capdec_backward        what torch autograd derives from the three forwards above
 *                          (SURVEY.md App. A.2), incl. grad wrt the returned alphas
 *   capdec_loss_fwd/_bwd   trains/attention_scn.py:219-235 (packed CE + alpha regulariser)
 *   capdec_beam_search     attention_scn.py:160-296, pure_scn.py:142-249,
 *                          pure_attention.py:153-281  (`sample`, batched over images)
 *   capdec_scn_cell_step   models/scn_cell.py:52-154   SCNCell.forward   (unit entry)
 *   capdec_attention_step  models/attention.py:26-44   Attention.forward (unit entry)
 *   capdec_gemm            test entry for the two GEMM engines (SIMT fp32 / tcgen05 bf16)
 *
 * Conventions
 *   - plain C structs of device pointers and ints; no torch / C++ types cross the ABI
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates device memory and never synchronises the device
 *   - every launch goes to the cudaStream_t passed by the caller (as void*)
 *   - return value: 0 = ok, negative = capdec_err; capdec_last_error() gives a
 *     thread-local message.  No exceptions or aborts cross the ABI.
 *   - all matrices are row-major and dense unless a stride is given
 */
#ifndef CAPDEC_H
#define CAPDEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define CAPDEC_VERSION 100

enum capdec_kind { CAPDEC_ATTENTION_SCN = 0, CAPDEC_PURE_SCN = 1, CAPDEC_PURE_ATTENTION = 2 };
enum capdec_precision { CAPDEC_FP32 = 0, CAPDEC_BF16 = 1 };

enum capdec_err {
  CAPDEC_OK = 0,
  CAPDEC_ERR_BAD_SHAPE = -1,     /* inconsistent / unsupported dimension            */
  CAPDEC_ERR_BAD_ARG = -2,       /* null pointer, bad enum                          */
  CAPDEC_ERR_WORKSPACE = -3,     /* workspace too small                             */
  CAPDEC_ERR_CUDA = -4,          /* a CUDA runtime / driver call failed             */
  CAPDEC_ERR_UNSUPPORTED = -5    /* device is not sm_100 / feature missing          */
};

/* Problem dimensions.  Names follow the reference: P pixels (14*14), E encoder_dim,
 * A attention_dim, M embed_dim, D decoder_dim, F factored_dim, S semantic_dim,
 * V vocab_size, L caption pitch (max_caption_length, 52).  B rows are captions
 * already sorted by decreasing length; T = max(decode_lengths). */
typedef struct CapdecDims {
  int32_t kind;        /* capdec_kind */
  int32_t precision;   /* capdec_precision: arithmetic of the GEMM operands / features */
  int32_t B, T, P, E, A, M, D, F, S, V, L;
} CapdecDims;

/* Device pointers to the fp32 master parameters, laid out exactly like the reference
 * state_dict.  Unused members (e.g. attention for pure_scn) are NULL.
 * For kind == PURE_ATTENTION (nn.LSTMCell): w_ia = weight_ih (4D, M+E), w_ha = weight_hh
 * (4D, D), gate order i,f,g,o; w_ib/w_ic/w_hb/w_hc are NULL.
 * The same struct with non-const use carries the gradients in capdec_backward. */
typedef struct CapdecParams {
  float *enc_att_w, *enc_att_b;      /* attention.encoder_att  (A,E) (A)  */
  float *dec_att_w, *dec_att_b;      /* attention.decoder_att  (A,D) (A)  */
  float *full_att_w, *full_att_b;    /* attention.full_att     (1,A) (1)  */
  float *emb;                        /* embedding.weight       (V,M)      */
  float *w_ia, *w_ib, *w_ic;         /* decode_step.weight_i{a,b,c} (X,4F) (S,4F) (D,4F) */
  float *w_ha, *w_hb, *w_hc;         /* decode_step.weight_h{a,b,c} (D,4F) (S,4F) (D,4F) */
  float *b_ih, *b_hh;                /* decode_step.bias_{ih,hh}    (4D)               */
  float *init_h_w, *init_h_b;        /* init_h (D,E) (D) */
  float *init_c_w, *init_c_b;        /* init_c (D,E) (D) */
  float *f_beta_w, *f_beta_b;        /* f_beta (E,D) (E) */
  float *fc_w, *fc_b;                /* fc     (V,D) (V) */
} CapdecParams;

int capdec_version(void);
const char* capdec_last_error(void);
/* Number of kernels this library has launched in this process (bench.py `gpu_launches`). */
unsigned long long capdec_launch_count(void);

/* One-time per-process/device setup (function attributes, driver entry points).
 * Idempotent and thread-safe.  Returns CAPDEC_ERR_UNSUPPORTED off sm_100. */
int capdec_init(void);

/* Bytes of caller-provided workspace needed by forward_train (+backward if
 * `with_backward`) for these dims.  0 on bad dims. */
size_t capdec_workspace_bytes(const CapdecDims* dims, int with_backward);

/* Teacher-forced forward.
 *   enc            (B,P,E) fp32, UNSORTED, arbitrary strides in elements (enc_sb, enc_sp, enc_se)
 *   sort_ind       (B) int64: sorted row i reads image sort_ind[i]
 *   tags           (B,S) fp32, used in the given (unsorted) order -- reference quirk App. C-1
 *   caps_sorted    (B,L) int64
 *   decode_len_h   HOST array (B) int32, non-increasing, decode_len_h[0] == T
 *   dropout_p / dropout_seed: p == 0 -> eval.  Mask = counter-based hash of (seed,b,t,d).
 *   predictions    (B,T,V) fp32 out (rows beyond each length are zero)
 *   alphas         (B,T,P) fp32 out (NULL for pure_scn)
 *   workspace      capdec_workspace_bytes() bytes, 256-B aligned; holds the packed weights
 *                  and the per-step activations that capdec_backward consumes.
 *   phases         0 or 3 = everything.  1 = INPUT PHASE only: the launches that read the
 *                  caller-owned input pointers (enc, sort_ind, tags, caps_sorted, lengths, seed)
 *                  and stage them in the workspace.  2 = COMPUTE PHASE only: touches nothing but
 *                  params, workspace, predictions and alphas, so with fixed pointers it can be
 *                  captured ONCE into a CUDA graph (cudaStreamBeginCapture on `stream`) and
 *                  replayed every iteration after a fresh phase-1 call.  The compute phase can be cut
 *                  in two: 4 = its PROLOGUE alone (weight packing, time-invariant products, embedding
 *                  projection: everything before the recurrence), 2|8 = the rest (recurrence +
 *                  vocabulary projection) after a prologue that has already been launched -- the host
 *                  can queue 1|4 BEFORE it learns the decode lengths of a repeated shape and keep
 *                  the GPU busy while it waits for them (capdec/functional.py `speculate`).
 *                  16 (with 2 and/or 4) = LENGTH-INDEPENDENT compute phases: the launches take the decode lengths
 *                  only from the device copy the input phase staged (decode_len_h may be NULL), so ONE captured
 *                  graph serves every batch with the same (B, T) whatever its lengths are.  Needs the persistent
 *                  recurrence kernels (CAPDEC_BF16 and a shape csrc/recur.cu covers), else CAPDEC_ERR_UNSUPPORTED.
 *                  In a compute-only call (no bit 1) enc / sort_ind / tags / caps_sorted are not read.
 *                  32 = re-stage decode_len_h on the device and nothing else (after an input phase that ran
 *                  before the lengths were known). */
int capdec_forward_train(const CapdecDims* dims, const CapdecParams* params,
                         const float* enc, int64_t enc_sb, int64_t enc_sp, int64_t enc_se,
                         const int64_t* sort_ind, const float* tags,
                         const int64_t* caps_sorted, const int32_t* decode_len_h,
                         float dropout_p, uint64_t dropout_seed, int save_for_backward, int phases,
                         float* predictions, float* alphas,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Test / parity hook for the dropout between h_t and fc (reference: `self.fc(self.dropout(h))`,
 * attention_scn.py:154; nn.Dropout scales the kept elements by 1/(1-p)).  torch's mask RNG cannot be matched,
 * so the decoder draws its mask from a counter-based hash; this entry writes the keep factors the forward
 * and backward kernels of a call with the same (dropout_seed, dropout_p) apply: mask_out[i] is 0 or
 * 1/(1-p) for the flat index i = (b*T + t)*D + d of the sorted batch, n = B*T*D.  Feeding it to an
 * independent implementation as the dropout mask reproduces the training-mode arithmetic exactly. */
int capdec_dropout_mask(uint64_t dropout_seed, float dropout_p, int64_t n, float* mask_out, void* stream);

/* Reverse-time backward of capdec_forward_train (same dims / workspace, later on the same
 * stream).  Reads only params, workspace (tags / captions / dropout seed were staged there by the
 * forward), alphas and the d_* inputs -> graph-capturable like the compute phase.
 *   d_predictions  (B,T,V) fp32 gradient of the returned scores (may be NULL if d_logits_ft given)
 *   d_logits_ft    optional (B*T, ldq) gradient already in the GEMM feature type (bf16 when
 *                  precision==BF16, fp32 otherwise), ldq = V rounded up to 8; written by
 *                  capdec_loss_bwd in the fused-loss path
 *   d_alphas       (B,T,P) fp32 gradient of the returned alphas, or NULL
 *   alphas         the (B,T,P) tensor capdec_forward_train wrote
 *   grads          every non-NULL member is OVERWRITTEN with the dense gradient of that
 *                  parameter (reference shapes; bias_ih.grad == bias_hh.grad).
 *   decode_len_h   HOST lengths as given to the forward, or NULL = length-independent launch (the device copy
 *                  staged by the forward is used; persistent kernels required, see `phases` 16 above)
 *   phases         0 = everything, else a bit set of the backward's stages IN PRODUCTION ORDER of the gradients,
 *                  so that a data-parallel caller can all-reduce one bucket of gradients while the next is being
 *                  computed (capdec/parallel.py):
 *                    1  fc.weight, fc.bias (and dH_fc)           2  the reverse-time recurrence
 *                    4  weight_ia / weight_ih, embedding.weight  8  the other cell weights, bias_ih, bias_hh,
 *                                                                   init_h.*, init_c.*
 *                    16 f_beta.*, attention.decoder_att.*,       32 attention.encoder_att.* (the long dAtt1
 *                       attention.full_att.*                        chain with the fewest bytes: last)
 *                  Stages must run in this order on one stream; each is graph-capturable on its own. */
int capdec_backward(const CapdecDims* dims, const CapdecParams* params,
                    const int32_t* decode_len_h, float dropout_p,
                    const float* d_predictions, const void* d_logits_ft, const float* d_alphas,
                    const float* alphas, const CapdecParams* grads,
                    void* workspace, size_t workspace_bytes, int phases, void* stream);

/* Loss glue: packed cross entropy (mean over N = sum(decode_len)) + alpha_c * mean_{b,p}
 * (1 - sum_t alpha)^2.  loss_out[0] = total, [1] = CE part, [2] = regulariser part.
 * lse_out: caller scratch of (2*B*T + B) floats; its first B*T entries (row log-sum-exp)
 * are what capdec_loss_bwd needs. */
int capdec_loss_fwd(const CapdecDims* dims, const float* predictions, const float* alphas,
                    const int64_t* caps_sorted, const int32_t* decode_len_d, int32_t n_tokens,
                    float alpha_c, float* loss_out, float* lse_out, void* stream);
/* Gradients of the loss: d_predictions (fp32 (B,T,V), may be NULL) and/or d_logits_ft (GEMM
 * feature type, pitch V rounded up to 8, may be NULL) and d_alphas (may be NULL).  The upstream
 * gradient is gscale (host) times *gscale_dev (device scalar, may be NULL -> 1). */
int capdec_loss_bwd(const CapdecDims* dims, const float* predictions, const float* alphas,
                    const int64_t* caps_sorted, const int32_t* decode_len_d, int32_t n_tokens,
                    float alpha_c, float gscale, const float* gscale_dev, const float* lse,
                    float* d_predictions, void* d_logits_ft, float* d_alphas, void* stream);

/* Batched beam search (reference `sample`, one independent search per image; all G searches
 * advance together and the host is not consulted inside the loop).
 *   enc (G,P,E) fp32 contiguous, tags (G,S) fp32 (NULL for pure_attention)
 *   k beams (<= 8); n_steps decode steps (<= 62): the reference breaks on `step > 50` AFTER
 *   processing that step (attention_scn.py:288), i.e. n_steps = 51
 *   out_seq   (G, n_steps+1) int32 incl. <start>, zero padded; out_len (G) int32; out_score (G) fp32
 *   out_completed (G) int32: 1 if some beam emitted <end>; 0 -> the reference raises ValueError
 *              and the result is the best LIVE beam (first maximum), DESIGN.md
 *   out_alpha (G, n_steps+1, P) fp32 or NULL; entry 0 is all ones (attention_scn.py:204)
 *   trace_parent/word (G, n_steps, k) int32 and trace_score fp32, or NULL: the top-k picks of every
 *             step in torch.topk order (-1 / 0 once an image has finished)
 *   dims->B, T, L are ignored.
 * CAPDEC_BF16: `scores = F.log_softmax(self.fc(h), dim=1)` + `scores.view(-1).topk(k)` (attention_scn.py:235-253)
 * run as ONE fused vocabulary kernel (projection + log-softmax statistics + top-k candidates; the (rows x V) logits
 * are never written) followed by a per-image merge; CAPDEC_FP32 keeps GEMM + the exact three-pass selection that the
 * token-parity tests pin (environment switches: INTEGRATION.md). */
size_t capdec_beam_workspace_bytes(const CapdecDims* dims, int G, int k, int n_steps);
int capdec_beam_search(const CapdecDims* dims, const CapdecParams* params,
                       const float* enc, const float* tags, int G, int k, int n_steps,
                       int32_t start_id, int32_t end_id,
                       int32_t* out_seq, int32_t* out_len, float* out_score, int32_t* out_completed,
                       float* out_alpha, int32_t* trace_parent, int32_t* trace_word,
                       float* trace_score,
                       void* workspace, size_t workspace_bytes, void* stream);
/* The same with STRIDED encoder features: element (g, p, e) at enc[g * enc_sb + p * enc_sp + e * enc_se].  The
 * reference's EncoderCaption returns a permuted view that is physically NCHW (models/encoders/caption.py:43; SURVEY.md
 * App. C-22): the prologue's gather casts / reorders it in its one pass instead of a dense fp32 copy being made first
 * (the `encoder_out.view(1, -1, encoder_dim)` + `expand` of attention_scn.py:176-189 never copies either). */
int capdec_beam_search_strided(const CapdecDims* dims, const CapdecParams* params,
                               const float* enc, int64_t enc_sb, int64_t enc_sp, int64_t enc_se,
                               const float* tags, int G, int k, int n_steps,
                               int32_t start_id, int32_t end_id,
                               int32_t* out_seq, int32_t* out_len, float* out_score, int32_t* out_completed,
                               float* out_alpha, int32_t* trace_parent, int32_t* trace_word,
                               float* trace_score,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Top-k accuracy count (reference utils/metric.py:25-39 `accuracy(scores, targets, k)` as
 * called on the packed scores in trains/attention_scn.py:255, 338).  hits_out[0] = number of rows whose target
 * is among the k largest logits (ties: the smaller index ranks first).  Two forms:
 *   packed:   scores (rows, V) with pitch ld, targets (rows) int64; caps_sorted = decode_len_d = NULL
 *   unpacked: scores = the (B,T,V) predictions of capdec_forward_train (ld = V, rows = B*T), targets = NULL,
 *             caps_sorted (B,L) and decode_len_d (B): row (b,t) counts iff t < decode_len[b], target
 *             caps_sorted[b, t+1] -- no pack_padded_sequence copy of the logits is needed.
 * The count stays on the device: the caller decides when to read it (the reference's .item() every iteration
 * is a host sync). */
int capdec_topk_hits(const float* scores, int64_t ld, const int64_t* targets, const int64_t* caps_sorted,
                     const int32_t* decode_len_d, int rows, int T, int L, int V, int k, int32_t* hits_out,
                     void* stream);

/* Fused gradient clip + Adam step over up to any number of fp32 tensors (reference:
 * utils/optimizer.py:1-11 `clip_gradient` = in-place clamp of every .grad to [-grad_clip, grad_clip], then
 * torch.optim.Adam.step(), trains/attention_scn.py:244-252).  One launch per CAPDEC_ADAM_MAX_SEGS tensors.
 *   segs          HOST array of n_segs entries: device pointers p (parameter), g (gradient), m (exp_avg),
 *                 v (exp_avg_sq), n elements each
 *   step          1-based step count (bias corrections 1 - beta^step, as torch)
 *   grad_clip     <= 0: no clamp
 *   write_clipped != 0: the clamped gradient is written back to g, as the reference leaves it */
#define CAPDEC_ADAM_MAX_SEGS 48
typedef struct CapdecAdamSeg {
  float* p; float* g; float* m; float* v;
  int64_t n;
} CapdecAdamSeg;
int capdec_clip_adam_step(const CapdecAdamSeg* segs, int n_segs, double lr, double beta1, double beta2, double eps,
                          double weight_decay, double grad_clip, int step, int write_clipped, void* stream);

/* Measurement aid for bench.py: with timing enabled, the persistent recurrence kernels of the next
 * capdec_forward_train / capdec_backward calls (csrc/recur.cu: the whole `for t` loop of
 * attention_scn.py:139-156, resp. its reverse-time gradient, in one cooperative launch each) are
 * bracketed by CUDA events on the launching stream.  capdec_recur_last_ms(which) (0 = forward
 * kernel, 1 = backward kernel) waits for the second event and returns the kernel's device time in
 * milliseconds (< 0: nothing was timed).  Do not enable it while the stream is being captured. */
void capdec_recur_timing(int enable);
float capdec_recur_last_ms(int which);

/* ---- unit entry points (single kernels, used by the parity tests) ---- */

/* out[rows,N] (ldo) = X[rows,K] (ldx) . W[N,K]^T (ldw) (+ bias[N]) (+ addm[rows,N] (ldadd)).
 * precision FP32: X,W fp32, SIMT FFMA engine.  BF16: X,W bf16, tcgen05+TMA engine.
 * out_ft != 0: out is written in the feature type, else fp32.  batch>1: element strides.
 * splitk (tcgen05 engine only): 0 = off, -1 = auto, n = n K-slices reduced with fp32 atomics
 * into `out`, which the caller must then have PRE-INITIALISED (zeros, or addm aliased to out). */
int capdec_gemm(int precision, const void* X, int64_t ldx, const void* W, int64_t ldw,
                void* out, int64_t ldo, int out_ft, const float* bias,
                const float* addm, int64_t ldadd,
                int rows, int N, int K, int batch, int64_t sX, int64_t sW, int64_t sO,
                int splitk, void* stream);

/* Weight-gradient product on TRANSPOSED bf16 operands (tcgen05 engine, MN-major UMMA descriptors):
 *   out[r, n] (fp32, ldo) = sum_k XT[k, r] * WT[k, n]      XT [K][rows] pitch ldx, WT [K][N] pitch ldw.
 * This is dW = dY^T X of an nn.Linear whose inputs X and output gradients dY are stored row = sample
 * (what torch autograd computes for the reference's `@` / nn.Linear weights) without a transposition
 * pass.  batch > 1: element strides sX / sW / sO between the batch members. */
int capdec_gemm_tn(const void* XT, int64_t ldx, const void* WT, int64_t ldw, float* out, int64_t ldo,
                   int rows, int N, int K, int batch, int64_t sX, int64_t sW, int64_t sO, void* stream);

/* Scratch floats capdec_attention_step needs for `rows` rows; capdec_attention_bwd_step needs this
 * plus rows * ((P + 3) & ~3). */
size_t capdec_attention_scratch_floats(int precision, int rows, int P, int E);

/* Attention.forward on prepared features: att1 (G,P,A) / enc (G,P,E) in the feature type,
 * g1 (rows, ldg) fp32 with att2 at column 0 and the f_beta pre-activation at column
 * beta_col (beta_col < 0: no gate, z = awe).  Row r uses feature map r / rows_per_map. */
int capdec_attention_step(int precision, const void* att1, const void* enc,
                          const float* g1, int64_t ldg, int beta_col,
                          const float* w_f, const float* b_f,
                          float* alpha_out, int64_t alpha_stride,
                          void* z_out, float* awe_out,
                          int rows, int rows_per_map, int P, int E, int A, float* scratch, void* stream);

/* Backward of capdec_attention_step for `rows` rows with one feature map each (training):
 * inputs dz (rows,E) = d loss / d z, the saved awe (rows,E) and alpha (rows, alpha_stride),
 * optional dalpha_ext (gradient arriving directly at alpha).  Outputs: dba (rows, lddba) in the
 * feature type = [d beta_pre (E) | d att2 (A)], dAtt1 (rows,P,A) fp32 ACCUMULATED INTO,
 * dwf_part (rows,A) and dbf_part (rows) per-row partials of the full_att gradients. */
int capdec_attention_bwd_step(int precision, const void* att1, const void* enc,
                              const float* g1, int64_t ldg, int beta_col, const float* w_f,
                              const float* alpha, int64_t alpha_stride,
                              const float* dalpha_ext, int64_t dalpha_stride,
                              const float* dz, const float* awe,
                              void* dba, int64_t lddba, float* dAtt1, float* dwf_part,
                              float* dbf_part, int rows, int P, int E, int A, float* scratch, void* stream);

/* SCNCell.forward for `rows` rows on fp32 master weights (packs them internally into
 * `workspace`): h_out/c_out (rows,D).  x (rows,X). */
size_t capdec_scn_cell_workspace_bytes(int precision, int rows, int X, int D, int F, int S);
int capdec_scn_cell_step(int precision, int rows, int X, int D, int F, int S,
                         const float* w_ia, const float* w_ib, const float* w_ic,
                         const float* w_ha, const float* w_hb, const float* w_hc,
                         const float* b_ih, const float* b_hh,
                         const float* x, const float* s, const float* h, const float* c,
                         float* h_out, float* c_out,
                         void* workspace, size_t workspace_bytes, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CAPDEC_H */
