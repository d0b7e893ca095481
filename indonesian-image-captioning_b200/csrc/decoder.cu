// decoder.cu -- host orchestration of the teacher-forced decoder forward and the
// reverse-time backward (one call = one mini-batch; every launch goes to the caller's
// stream; no allocation, no synchronisation).
//
// Reference: models/decoders/attention_scn.py:95-158 (AttentionSCN.forward),
// pure_scn.py:87-140, pure_attention.py:90-151, models/scn_cell.py:52-154,
// models/attention.py:26-44; backward = SURVEY.md App. A.2.
//
// Structure of one decode step (attention_scn), each line one kernel:
//   G1   [att2 | beta_pre | p] = h_{t-1} . [W_d ; W_beta ; W_ha^T]^T + [b_d ; b_beta ; 0]
//   ATT  alpha, awe, z = sigmoid(beta_pre) * awe                       (attention.cu)
//   P3   u = U_emb[t] + z . W_ia[M:]            (U_emb = Emb[caps] . W_ia[:M], batched over t)
//   FM   m_g = [u_g * v_g | p_g * q_g]          (v = s W_ib, q = s W_hb: hoisted, App. C-7)
//   P4   pre_g = m_g . [W_ic_g | W_hc_g]^T      (4 gates = 4 GEMM batches)
//   CELL i,f,o,g~ -> c_t, h_t (+ dropout copy for fc)
// The time-invariant products (att1, v, q, U_emb, h0/c0) and the vocabulary projection
// over all (b,t) rows are batched GEMMs outside the loop.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

inline int64_t pad8(int64_t x) { return round_up(x, 8); }

bool fused_epilogue_enabled() {
  const char* s = getenv("CAPDEC_FUSED_EPILOGUE");      // read per call: tests flip it
  return s && s[0] == '1';
}

struct Plan {
  CapdecDims d;
  bool att, scn, bwd;
  int X, NQ, NG1, fsz;             // fsz = sizeof(feature type)
  int64_t ldE, ldD, ldX, ldS, ld2F, ldV, ldNQ, ldEA, ldM, ldR, ldB, ldBP, ldA, ldPX;
  bool fused;                      // bf16 SCN path with the fused GEMM epilogues (gemm_tc.cu)
  // byte offsets into the workspace
  struct Off {
    // packed weights (feature type) + fp32 bias vectors
    size_t Wp_e, Wp_cat1, b_cat1, Wp_xq, Wp_ibT, Wp_hbT, Wp_c, Wp_init, Wp_fc;
    size_t Wp_fcT, Wp_cT, Wp_hq, Wp_xin, Wp_b6, Wp_hx;
    // forward activations
    size_t enc_s, att1, mean, meanF, tagsF, v, q, Xe, U, g1, awe, z, m, pre, gates, C, H0, Hall,
        Hd, lenD, seedD, capsD, counters, att_scr, bar, Ht, zk, enc_cm, scores_t;
    // backward buffers
    size_t dlogF, dHfc, dh_rec, dc, dpre, wr, du, dpx, de, dv_acc, dq_acc, dz, dba, dAtt1, dwf, dbf, dXe;
    size_t dpre_gm, duk, dpxk, att1_cm;     // chunk-major operand copies of the persistent backward (recur.cu)
    size_t dhp, part_t;                     // ... and its per-step exchange buffers
    size_t tA, tB, tC;             // transposed-operand scratch
    size_t total;
  } o;
};

int make_plan(const CapdecDims& d, bool with_bwd, Plan* p) {
  CAPDEC_REQUIRE(d.kind >= 0 && d.kind <= 2 && (d.precision == 0 || d.precision == 1),
                 CAPDEC_ERR_BAD_ARG, "bad kind/precision");
  CAPDEC_REQUIRE(d.B > 0 && d.T > 0 && d.V > 1 && d.L > d.T && d.M > 0 && d.D > 0 && d.E > 0,
                 CAPDEC_ERR_BAD_SHAPE, "bad dims B=%d T=%d V=%d L=%d", d.B, d.T, d.V, d.L);
  p->d = d;
  p->att = d.kind != CAPDEC_PURE_SCN;
  p->scn = d.kind != CAPDEC_PURE_ATTENTION;
  p->bwd = with_bwd;
  CAPDEC_REQUIRE(d.D % 8 == 0 && d.M % 8 == 0 && d.E % 8 == 0, CAPDEC_ERR_BAD_SHAPE,
                 "decoder_dim, embed_dim, encoder_dim must be multiples of 8 (D=%d M=%d E=%d)", d.D,
                 d.M, d.E);
  if (p->scn)
    CAPDEC_REQUIRE(d.F > 0 && d.F % 8 == 0 && d.S > 0, CAPDEC_ERR_BAD_SHAPE,
                   "factored_dim must be a multiple of 8 (F=%d S=%d)", d.F, d.S);
  if (p->att)
    CAPDEC_REQUIRE(d.A > 0 && d.A % 8 == 0 && d.P > 0, CAPDEC_ERR_BAD_SHAPE,
                   "attention_dim must be a multiple of 8 (A=%d P=%d)", d.A, d.P);
  p->X = p->att ? d.M + d.E : d.M;
  p->NQ = p->scn ? 4 * d.F : 4 * d.D;
  p->NG1 = (p->att ? d.A + d.E : 0) + p->NQ;
  p->fsz = d.precision == CAPDEC_BF16 ? 2 : 4;
  p->ldE = pad8(d.E); p->ldD = pad8(d.D); p->ldX = pad8(p->X); p->ldS = pad8(d.S);
  p->ld2F = pad8(2 * d.F); p->ldV = pad8(d.V); p->ldNQ = pad8(p->NQ); p->ldEA = pad8(d.E + d.A);
  p->ldM = pad8(d.M); p->ldR = pad8((int64_t)d.B * d.T); p->ldB = pad8(d.B);
  p->ldBP = pad8((int64_t)d.B * d.P); p->ldA = pad8(d.A);
  p->ldPX = pad8(p->NQ + (p->att ? d.E + d.A : 0));
  // opt-in (CAPDEC_FUSED_EPILOGUE=1): on B200 the ticketed last-CTA epilogues cost as much as the
  // pointwise kernels they replace, so the default keeps the plain split-K GEMMs
  p->fused = fused_epilogue_enabled() && d.precision == CAPDEC_BF16 && p->scn && gemm_tc_cell_fusable(d.B);

  const size_t f = p->fsz;
  const int64_t B = d.B, T = d.T, P = d.P, E = d.E, A = d.A, M = d.M, D = d.D, F = d.F, S = d.S,
                V = d.V, X = p->X, NQ = p->NQ, NG1 = p->NG1, R = B * T;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t at = cur; cur += (size_t)round_up((int64_t)bytes, 256); return at; };
  Plan::Off& o = p->o;
  memset(&o, 0, sizeof o);
  // ---- packed weights ----
  if (p->att) o.Wp_e = take(A * p->ldE * f);
  o.Wp_cat1 = take(NG1 * p->ldD * f);
  o.b_cat1 = take(NG1 * 4);
  o.Wp_xq = take(NQ * p->ldX * f);                  // SCN: W_ia^T [4F][X] ; LSTM: W_ih [4D][X]
  if (p->scn) {
    o.Wp_ibT = take(NQ * p->ldS * f);
    o.Wp_hbT = take(NQ * p->ldS * f);
    o.Wp_c = take(4 * D * p->ld2F * f);
  }
  o.Wp_init = take(2 * D * p->ldE * f);
  o.Wp_fc = take(V * p->ldD * f);
  if (with_bwd) {
    o.Wp_fcT = take(D * p->ldV * f);
    if (p->scn) o.Wp_cT = take(4 * 2 * F * p->ldD * f);
    if (!p->scn) o.Wp_hq = take(D * p->ldNQ * f);   // LSTM: W_hh^T [D][4D]
    o.Wp_xin = take(X * p->ldNQ * f);               // SCN: W_ia [X][4F] ; LSTM: W_ih^T [X][4D]
    if (p->att && !p->scn) o.Wp_b6 = take(D * p->ldEA * f);
    if (p->scn) o.Wp_hx = take(D * p->ldPX * f);     // [W_ha | W_beta^T | W_d^T]
  }
  // ---- forward activations ----
  o.enc_s = take(B * P * E * f);
  if (p->att) o.att1 = take(B * P * A * f);
  o.mean = take(B * E * 4);
  o.meanF = take(B * p->ldE * f);
  if (p->scn) {
    o.tagsF = take(B * p->ldS * f);
    o.v = take(B * NQ * 4);
    o.q = take(B * NQ * 4);
  }
  o.Xe = take(R * p->ldM * f);
  o.U = take(R * NQ * 4);
  o.g1 = take(R * NG1 * 4);
  if (p->att) {
    o.awe = take(R * E * 4);
    o.z = take(R * E * f);
  }
  if (p->scn) o.m = take(4 * R * 2 * F * f);          // [gate][t*B + b][u*v | p*q]
  // SCN: gate pre-activations (per step: split-K GEMMs accumulate into zeroed slots); LSTM + persistent kernel: the
  // input-side product of the step
  o.pre = take(R * 4 * D * 4);
  o.gates = take(R * 4 * D * 4);
  o.C = take((T + 1) * B * D * 4);
  o.H0 = take(B * p->ldD * f);
  o.Hall = take(R * D * f);
  o.Hd = take(R * D * f);
  o.lenD = take(B * 4);
  o.seedD = take(8);
  o.capsD = take((size_t)B * d.L * 8);
  if (p->att) o.att_scr = take(attention_scratch_floats(d.precision, (int)B, (int)P, (int)E) * 4);
  o.counters = take((size_t)GEMM_TC_MAX_TILE_COUNTERS * 4);   // split-K tickets of the fused GEMM epilogues
  o.bar = take(256);                                          // grid-barrier counter of the persistent kernels
  if (d.precision == CAPDEC_BF16 && (p->scn || p->att)) {     // operand copies of the persistent kernel (recur.cu)
    o.Ht = take(R * D * f);
    if (p->att) {
      o.zk = take(R * E * f);
      o.enc_cm = take(B * P * E * f);
      o.scores_t = take(R * ((P + 3) / 4 * 4) * 4);    // per-step attention scores
    }
  }
  if (with_bwd) {
    o.dlogF = take(R * p->ldV * f);
    o.dHfc = take(R * D * 4);
    o.dh_rec = take((T + 1) * B * D * 4);            // dhs[t] = d loss / d h_{t-1} from step t ; dhs[T] = 0
    o.dc = take(B * D * 4);
    o.dpre = take(R * 4 * D * f);
    if (p->scn) {
      o.wr = take(R * 4 * 2 * F * 4);
      o.du = take(R * NQ * f);
      o.dpx = take(R * p->ldPX * f);                 // per row [dp | dbeta_pre | datt2]
      o.dv_acc = take(B * NQ * 4);
      o.dq_acc = take(B * NQ * 4);
    }
    if (p->att) {
      o.dz = take(R * E * 4);
      o.dba = take(R * p->ldEA * f);
      o.dAtt1 = take(B * P * A * 4);
      o.de = take(R * ((P + 3) / 4 * 4) * 4);          // softmax-input gradients of every step
      o.dwf = take(R * A * 4);
      o.dbf = take(R * 4);
    }
    o.dXe = take(R * M * 4);
    if (d.precision == CAPDEC_BF16 && (p->scn || p->att)) {
      o.dpre_gm = take(R * 4 * D * f);
      if (p->scn) o.duk = take(R * NQ * f);
      o.dpxk = take(R * p->ldPX * f);
      o.dhp = take(R * ((p->NQ + (p->att ? E + A : 0)) / 512 + 1) * D * 4);
      if (p->att) {
        o.att1_cm = take(B * P * A * f);
        o.part_t = take(R * ((E + 255) / 256) * ((P + 3) / 4 * 4) * 4);
      }
    }
    // transposed operands for the weight-gradient GEMMs (K = rows).  tA/tB are sized for
    // the largest pair used together, tC for the shared H_prev^T.
    int64_t widest = V;
    if (NQ > widest) widest = NQ;
    if (4 * D > widest) widest = 4 * D;
    if (E + A > widest) widest = E + A;
    if (X > widest) widest = X;
    int64_t kmax = p->ldR;
    size_t tA = (size_t)widest * kmax * f;
    size_t tB = (size_t)(NQ > 4 * D ? NQ : 4 * D) * kmax * f;
    auto grow = [](size_t& x, size_t v) { if (v > x) x = v; };
    grow(tB, (size_t)M * kmax * f);
    grow(tB, (size_t)E * kmax * f);
    grow(tB, (size_t)D * kmax * f);
    if (p->scn) grow(tB, (size_t)(8 * F) * kmax * f);
    grow(tA, (size_t)S * p->ldB * f);
    grow(tA, (size_t)D * p->ldB * f);
    grow(tB, (size_t)NQ * p->ldB * f);
    grow(tB, (size_t)E * p->ldB * f);
    if (p->att) {
      grow(tA, (size_t)A * p->ldBP * f);
      grow(tA, (size_t)B * P * p->ldA * f);          // bf16 copy of dAtt1 [B*P][A] (transposed-operand GEMM)
      grow(tB, (size_t)E * p->ldBP * f);
    }
    o.tA = take(tA);
    o.tB = take(tB);
    o.tC = take((size_t)D * kmax * f);
  }
  o.total = cur;
  return CAPDEC_OK;
}

struct Ctx {
  Plan p;
  uint8_t* ws;
  cudaStream_t st;
  int prec;
  template <typename T = void> T* at(size_t off) const { return (T*)(ws + off); }
  // element offset inside a feature-type buffer
  void* ft(size_t off, int64_t elem) const { return ws + off + (size_t)elem * p.fsz; }
};

int G_(const Ctx& c, const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo,
      int out_ft, const float* bias, const float* addm, int64_t ldadd, int rows, int N, int K,
      int rows_alloc = 0, int batch = 1, int64_t sX = 0, int64_t sW = 0, int64_t sO = 0, int splitk = 0) {
  GemmArgs a;
  a.X = X; a.ldx = ldx; a.W = W; a.ldw = ldw; a.out = out; a.ldo = ldo; a.out_ft = out_ft;
  a.bias = bias; a.addm = addm; a.ldadd = ldadd; a.rows = rows; a.N = N; a.K = K;
  a.rows_alloc = rows_alloc; a.batch = batch; a.sX = sX; a.sW = sW; a.sO = sO; a.splitk = splitk;
  return gemm(c.prec, a, c.st);
}

// tcgen05 engine with transposed operand(s) (bf16 only): tn bit 0 = W is W^T [K][N], bit 1 = X is X^T [K][rows]
int GT_(const Ctx& c, int tn, const void* X, int64_t ldx, const void* W, int64_t ldw, float* out, int64_t ldo,
        int rows, int N, int K, int batch = 1, int64_t sX = 0, int64_t sW = 0, int64_t sO = 0) {
  GemmArgs a;
  a.tn = tn;
  a.X = X; a.ldx = ldx; a.W = W; a.ldw = ldw; a.out = out; a.ldo = ldo; a.rows = rows; a.N = N; a.K = K;
  a.batch = batch; a.sX = sX; a.sW = sW; a.sO = sO;
  return gemm(CAPDEC_BF16, a, c.st);
}

// fp32 master weights -> packed feature-type operands ([N_out][K] , K contiguous)
int pack_weights(const Ctx& c, const CapdecParams& w) {
  const Plan& p = c.p;
  const CapdecDims& d = p.d;
  const int pr = c.prec;
  cudaStream_t st = c.st;
  const int D = d.D, E = d.E, A = d.A, F = d.F, S = d.S, V = d.V, X = p.X, NQ = p.NQ;
  int row = 0;     // row cursor inside Wp_cat1
  PackTable tbl;
  bool ok = true;
  // dst[r][c] = src[r][c] (CC) / dst[c][r] = src[r][c] (TT), fp32 master -> feature type, all in one launch
  auto CC = [&](const float* src, int64_t lds, void* dst, int64_t ldd, int R_, int C_) { ok = ok && tbl.add(src, lds, dst, ldd, R_, C_, 0); };
  auto TT = [&](const float* src, int64_t lds, void* dst, int64_t ldd, int R_, int C_) { ok = ok && tbl.add(src, lds, dst, ldd, R_, C_, 1); };
  if (p.att) {
    CC(w.enc_att_w, E, c.at(p.o.Wp_e), p.ldE, A, E);
    CC(w.dec_att_w, D, c.ft(p.o.Wp_cat1, 0), p.ldD, A, D);
    CC(w.f_beta_w, D, c.ft(p.o.Wp_cat1, (int64_t)A * p.ldD), p.ldD, E, D);
    row = A + E;
    CAPDEC_TRY(concat_bias(c.at<float>(p.o.b_cat1), w.dec_att_b, A, w.f_beta_b, E, NQ, st));
  } else {
    CAPDEC_TRY(concat_bias(c.at<float>(p.o.b_cat1), nullptr, 0, nullptr, 0, NQ, st));
  }
  if (p.scn) {
    // W_ha (D,4F) -> rows [row, row+4F) of cat1 as W_ha^T
    TT(w.w_ha, NQ, c.ft(p.o.Wp_cat1, (int64_t)row * p.ldD), p.ldD, D, NQ);
    TT(w.w_ia, NQ, c.at(p.o.Wp_xq), p.ldX, X, NQ);
    TT(w.w_ib, NQ, c.at(p.o.Wp_ibT), p.ldS, S, NQ);
    TT(w.w_hb, NQ, c.at(p.o.Wp_hbT), p.ldS, S, NQ);
    for (int g = 0; g < 4; ++g) {
      // Wp_c[g] = [ W_ic[:, gF:(g+1)F] | W_hc[:, gF:(g+1)F] ]   (D x 2F)
      CC(w.w_ic + g * F, NQ, c.ft(p.o.Wp_c, (int64_t)g * D * p.ld2F), p.ld2F, D, F);
      CC(w.w_hc + g * F, NQ, c.ft(p.o.Wp_c, (int64_t)g * D * p.ld2F + F), p.ld2F, D, F);
    }
  } else {
    // LSTM: weight_hh (4D,D) and weight_ih (4D,X) are already [N_out][K]
    CC(w.w_ha, D, c.ft(p.o.Wp_cat1, (int64_t)row * p.ldD), p.ldD, NQ, D);
    CC(w.w_ia, X, c.at(p.o.Wp_xq), p.ldX, NQ, X);
  }
  CC(w.init_h_w, E, c.ft(p.o.Wp_init, 0), p.ldE, D, E);
  CC(w.init_c_w, E, c.ft(p.o.Wp_init, (int64_t)D * p.ldE), p.ldE, D, E);
  CC(w.fc_w, D, c.at(p.o.Wp_fc), p.ldD, V, D);
  if (p.bwd) {
    TT(w.fc_w, D, c.at(p.o.Wp_fcT), p.ldV, V, D);
    if (p.scn) {
      for (int g = 0; g < 4; ++g) {
        // Wp_cT[g] = [ W_ic_g^T ; W_hc_g^T ]  (2F x D)
        TT(w.w_ic + g * F, NQ, c.ft(p.o.Wp_cT, (int64_t)g * 2 * F * p.ldD), p.ldD, D, F);
        TT(w.w_hc + g * F, NQ, c.ft(p.o.Wp_cT, ((int64_t)g * 2 * F + F) * p.ldD), p.ldD, D, F);
      }
      CC(w.w_ia, NQ, c.at(p.o.Wp_xin), p.ldNQ, X, NQ);
    } else {
      TT(w.w_ha, D, c.at(p.o.Wp_hq), p.ldNQ, NQ, D);
      TT(w.w_ia, X, c.at(p.o.Wp_xin), p.ldNQ, NQ, X);
    }
    if (p.scn) {
      // Wp_hx = [ W_ha | W_beta^T | W_d^T ]   (D x (4F + E + A)): ONE GEMM gives the recurrent dh
      CC(w.w_ha, NQ, c.at(p.o.Wp_hx), p.ldPX, D, NQ);
      if (p.att) {
        TT(w.f_beta_w, D, c.ft(p.o.Wp_hx, NQ), p.ldPX, E, D);
        TT(w.dec_att_w, D, c.ft(p.o.Wp_hx, NQ + E), p.ldPX, A, D);
      }
    }
    if (p.att && !p.scn) {
      // Wp_b6 = [ W_beta^T | W_d^T ]   (D x (E+A))
      TT(w.f_beta_w, D, c.ft(p.o.Wp_b6, 0), p.ldEA, E, D);
      TT(w.dec_att_w, D, c.ft(p.o.Wp_b6, E), p.ldEA, A, D);
    }
  }
  CAPDEC_REQUIRE(ok, CAPDEC_ERR_BAD_ARG, "pack_weights: segment table overflow");
  CAPDEC_TRY(pack_multi(pr, tbl, st));
  return CAPDEC_OK;
}

}  // namespace

size_t workspace_bytes(const CapdecDims& d, int with_bwd) {
  Plan p;
  if (make_plan(d, with_bwd != 0, &p) != CAPDEC_OK) return 0;
  return p.o.total;
}

int forward_train(const CapdecDims& d, const CapdecParams& w, const float* enc, int64_t sb, int64_t sp,
                  int64_t se, const int64_t* sort_ind, const float* tags, const int64_t* caps,
                  const int32_t* len_h, float dropout_p, uint64_t seed, int save_bwd, int phases,
                  float* predictions, float* alphas, void* workspace, size_t ws_bytes,
                  cudaStream_t st) {
  // phases: bit 0 = input phase (reads enc / tags / captions / lengths / seed into the workspace;
  // the only part that touches caller-owned INPUT pointers), bit 1 = everything else (parameters,
  // workspace and outputs only -> capturable once into a CUDA graph and replayed)
  Ctx c;
  CAPDEC_TRY(make_plan(d, save_bwd != 0, &c.p));
  CAPDEC_REQUIRE(ws_bytes >= c.p.o.total, CAPDEC_ERR_WORKSPACE, "workspace %zu < %zu", ws_bytes,
                 c.p.o.total);
  CAPDEC_REQUIRE(((uintptr_t)workspace % 256) == 0, CAPDEC_ERR_BAD_ARG, "workspace must be 256-B aligned");
  // phases bit 4 (16): LENGTH-INDEPENDENT compute phases -- the launches read the decode lengths only from the
  // device copy the input phase staged, so one captured graph serves every batch with the same (B, T)
  const bool len_free = (phases & 16) != 0;
  CAPDEC_REQUIRE(len_h || (len_free && !(phases & 1)), CAPDEC_ERR_BAD_ARG, "decode_len_h is NULL");
  if (len_h) {
    CAPDEC_REQUIRE(len_h[0] == d.T, CAPDEC_ERR_BAD_SHAPE, "decode_len[0]=%d != T=%d", len_h[0], d.T);
    for (int b = 1; b < d.B; ++b)
      CAPDEC_REQUIRE(len_h[b] <= len_h[b - 1] && len_h[b] >= 1, CAPDEC_ERR_BAD_SHAPE,
                     "decode lengths must be sorted descending and >= 1");
  }
  c.ws = (uint8_t*)workspace;
  c.st = st;
  c.prec = d.precision;
  const Plan& p = c.p;
  const Plan::Off& o = p.o;
  const int pr = c.prec;
  const int B = d.B, T = d.T, P = d.P, E = d.E, A = d.A, M = d.M, D = d.D, F = d.F, S = d.S, V = d.V,
            NQ = p.NQ, NG1 = p.NG1;
  const int64_t R = (int64_t)B * T;
  const bool ragged = len_free || len_h[B - 1] != T;

  std::vector<int> bt(T, B);
  for (int t = 0; t < T && !len_free; ++t) {
    int n = 0;
    while (n < B && len_h[n] > t) ++n;
    bt[t] = n;
  }
  if ((phases & 32) && !(phases & 1)) {     // re-stage the decode lengths alone (after a speculative input phase)
    CAPDEC_REQUIRE(len_h, CAPDEC_ERR_BAD_ARG, "decode_len_h is NULL");
    CAPDEC_CUDA_OK(cudaMemcpyAsync(c.at(o.lenD), len_h, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  }
  if (phases & 1) {
    CAPDEC_CUDA_OK(cudaMemcpyAsync(c.at(o.lenD), len_h, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CAPDEC_CUDA_OK(cudaMemcpyAsync(c.at(o.seedD), &seed, 8, cudaMemcpyHostToDevice, st));
    CAPDEC_CUDA_OK(cudaMemcpyAsync(c.at(o.capsD), caps, (size_t)B * d.L * 8, cudaMemcpyDeviceToDevice, st));
    // sorted, converted features + pixel mean (attention_scn.py:113-120, :90)
    // ... and, for the persistent kernels, the chunk-major copy of the features in the same pass
    const bool want_cm = o.enc_cm != 0 && E % 256 == 0;
    CAPDEC_TRY(gather_features(pr, enc, sb, sp, se, sort_ind, c.at(o.enc_s), c.at<float>(o.mean),
                               c.at(o.meanF), p.ldE, B, P, E, st, want_cm ? c.at(o.enc_cm) : nullptr, 256));
    if (p.scn)   // tags are NOT permuted (App. C-1): row i of the sorted batch uses tags[i]
      CAPDEC_TRY(copy_cast(pr, tags, 0, S, c.at(o.tagsF), 1, p.ldS, B, S, st));
  }
  // compute phases: PROLOGUE = everything before the recurrence (weight packing, time-invariant products,
  // embedding projection, buffer zeroing), REST = recurrence + vocabulary projection.  bit 1 (2) asks for
  // both unless bit 3 (8) says the prologue has already been launched; bit 2 (4) asks for the prologue alone
  const bool do_pro = (phases & 4) || ((phases & 2) && !(phases & 8));
  const bool do_rest = (phases & 2) != 0;
  if (!do_pro && !do_rest) return CAPDEC_OK;
  const int64_t* capsD = c.at<int64_t>(o.capsD);
  const bool fused = p.fused;
  const bool drop = dropout_p > 0.f;
  // the recurrence as ONE persistent cooperative kernel (recur.cu) when the shape is covered
  bool persistent = false;
  RecurFwdArgs ra;
  if (pr == CAPDEC_BF16 && (p.scn || p.att) && !fused && (!p.att || alphas)) {
    ra.att = p.att ? 1 : 0;
    ra.lstm = p.scn ? 0 : 1;
    ra.B = B; ra.T = T; ra.P = P; ra.E = E; ra.A = A; ra.M = M; ra.D = D; ra.F = F;
    ra.len = c.at<int32_t>(o.lenD);
    ra.Wcat1 = c.at(o.Wp_cat1); ra.ldD = p.ldD;
    ra.Wxz = c.ft(o.Wp_xq, M); ra.ldX = p.ldX;
    ra.Wc = p.scn ? c.at(o.Wp_c) : nullptr; ra.ld2F = p.ld2F;
    ra.b_cat1 = c.at<float>(o.b_cat1); ra.b_ih = w.b_ih; ra.b_hh = w.b_hh;
    if (p.att) {
      ra.att1 = c.at(o.att1); ra.enc = c.at(o.enc_s); ra.w_f = w.full_att_w; ra.b_f = w.full_att_b;
      ra.alphas = alphas; ra.awe = save_bwd ? c.at<float>(o.awe) : nullptr; ra.z = c.at(o.z);
      ra.scores = c.at<float>(o.scores_t);
    }
    ra.ragged = ragged ? 1 : 0;
    ra.v = p.scn ? c.at<float>(o.v) : nullptr; ra.q = p.scn ? c.at<float>(o.q) : nullptr;
    ra.Ht = c.at(o.Ht); ra.zk = p.att ? c.at(o.zk) : nullptr; ra.enc_cm = p.att ? c.at(o.enc_cm) : nullptr;
    ra.H0 = c.at(o.H0); ra.ldH0 = p.ldD; ra.Hall = c.at(o.Hall); ra.Hd = drop ? c.at(o.Hd) : nullptr;
    ra.C = c.at<float>(o.C); ra.U = c.at<float>(o.U); ra.g1 = c.at<float>(o.g1); ra.m = p.scn ? c.at(o.m) : nullptr;
    ra.pre = c.at<float>(o.pre); ra.gates = c.at<float>(o.gates); ra.bar = c.at<unsigned>(o.bar);
    ra.dropout_p = dropout_p; ra.seed = c.at<uint64_t>(o.seedD);
    persistent = recur_fwd_supported(ra);
  }
  CAPDEC_REQUIRE(persistent || !len_free, CAPDEC_ERR_UNSUPPORTED,
                 "length-independent launch needs the persistent recurrence kernels (bf16, shape covered by recur.cu)");
  const int SK = pr == CAPDEC_BF16 ? -1 : 0;
  int* counters = c.at<int>(o.counters);

  if (do_pro) {
  CAPDEC_TRY(pack_weights(c, w));

  // ---------------- prologue: time-invariant products ----------------
  if (p.att)   // att1 = enc . W_e^T + b_e      (attention.py:35, hoisted)
    CAPDEC_TRY(G_(c, c.at(o.enc_s), E, c.at(o.Wp_e), p.ldE, c.at(o.att1), A, 1, w.enc_att_b, nullptr, 0,
                 B * P, A, E));
  // h0 -> H0 (feature type), c0 -> C[0] (fp32)   (attention_scn.py:90-92)
  CAPDEC_TRY(G_(c, c.at(o.meanF), p.ldE, c.ft(o.Wp_init, 0), p.ldE, c.at(o.H0), p.ldD, 1, w.init_h_b,
               nullptr, 0, B, D, E));
  CAPDEC_TRY(G_(c, c.at(o.meanF), p.ldE, c.ft(o.Wp_init, (int64_t)D * p.ldE), p.ldE, c.at(o.C), D, 0,
               w.init_c_b, nullptr, 0, B, D, E));
  if (p.scn) {   // v = s W_ib, q = s W_hb   (scn_cell.py:78-81, 134-143)
    CAPDEC_TRY(G_(c, c.at(o.tagsF), p.ldS, c.at(o.Wp_ibT), p.ldS, c.at(o.v), NQ, 0, nullptr, nullptr, 0, B,
                 NQ, S));
    CAPDEC_TRY(G_(c, c.at(o.tagsF), p.ldS, c.at(o.Wp_hbT), p.ldS, c.at(o.q), NQ, 0, nullptr, nullptr, 0, B,
                 NQ, S));
  }
  // embeddings of the teacher tokens and their input-side projection, all (t,b) rows at once
  CAPDEC_TRY(embedding_gather(pr, w.emb, capsD, d.L, c.at(o.Xe), p.ldM, B, T, M, V, st));
  if (ragged) {
    // rows beyond a caption's length are never written by the step kernels; the batched GEMMs over
    // all (t,b) rows must see finite (zero) operands there
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.Hall), 0, (size_t)R * D * p.fsz, st));
    if (dropout_p > 0.f) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.Hd), 0, (size_t)R * D * p.fsz, st));
    if (p.att) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.z), 0, (size_t)R * E * p.fsz, st));
    // (the persistent kernel fills m with its exchange pattern and zeroes the dead rows itself)
    if (p.scn && !persistent) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.m), 0, (size_t)R * 4 * 2 * F * p.fsz, st));
  }
  if (fused && !p.att) {
    // pure_scn: the whole input side is non-recurrent -- u = Emb W_ia AND the left half u*v of the
    // P4 operand for every (t,b) row come out of this one GEMM
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.U), 0, (size_t)R * NQ * 4, st));
    GemmArgs a;
    a.X = c.at(o.Xe); a.ldx = p.ldM; a.W = c.at(o.Wp_xq); a.ldw = p.ldX; a.out = c.at(o.U); a.ldo = NQ;
    a.rows = (int)R; a.N = NQ; a.K = M; a.epi = EPI_P3;
    a.e.fa = c.at<float>(o.v); a.e.m = c.at(o.m); a.e.mB = (int)R; a.e.F = F; a.e.vB = B;
    CAPDEC_TRY(gemm(pr, a, st));
  } else {
    CAPDEC_TRY(G_(c, c.at(o.Xe), p.ldM, c.at(o.Wp_xq), p.ldX, c.at(o.U), NQ, 0, nullptr, nullptr, 0, (int)R,
                  NQ, M));
  }
  if (alphas) CAPDEC_CUDA_OK(cudaMemsetAsync(alphas, 0, (size_t)R * P * 4, st));
  if (!persistent && SK) {
    // split-K GEMMs of the tcgen05 engine accumulate with atomics into pre-zeroed buffers
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.g1), 0, (size_t)R * NG1 * 4, st));
    if (p.scn) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.pre), 0, (size_t)R * 4 * D * 4, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)GEMM_TC_MAX_TILE_COUNTERS * 4, st));
  }
  }   // do_pro
  if (!do_rest) return CAPDEC_OK;
  if (persistent) CAPDEC_TRY(recur_fwd(ra, st));

  // ---------------- the recurrence, one kernel chain per step ----------------
  for (int t = 0; t < (persistent ? 0 : T); ++t) {
    const int n = bt[t];
    const void* hprev = t == 0 ? c.at(o.H0) : c.ft(o.Hall, (int64_t)(t - 1) * D);
    const int64_t ldh = t == 0 ? p.ldD : (int64_t)T * D;
    float* g1 = c.at<float>(o.g1) + (int64_t)t * B * NG1;
    float* U = c.at<float>(o.U) + (int64_t)t * B * NQ;
    void* m = p.scn ? c.ft(o.m, (int64_t)t * B * 2 * F) : nullptr;     // gate stride R*2F
    const float* pcol = g1 + (p.att ? A + E : 0);
    const float* c_prev = c.at<float>(o.C) + (int64_t)t * B * D;
    float* c_new = c.at<float>(o.C) + (int64_t)(t + 1) * B * D;
    float* gates = c.at<float>(o.gates) + (int64_t)t * B * 4 * D;
    void* hout = c.ft(o.Hall, (int64_t)t * D);
    void* hdout = drop ? c.ft(o.Hd, (int64_t)t * D) : nullptr;
    {
      // G1: [att2 | beta_pre | p] = h_{t-1} [W_d ; W_beta ; W_ha^T]^T (+ fused p*q -> right half of m)
      GemmArgs a;
      a.X = hprev; a.ldx = ldh; a.W = c.at(o.Wp_cat1); a.ldw = p.ldD; a.out = g1; a.ldo = NG1;
      a.bias = c.at<float>(o.b_cat1); a.rows = n; a.N = NG1; a.K = D; a.rows_alloc = B; a.splitk = SK;
      if (fused) {
        a.epi = EPI_G1;
        a.e.fa = c.at<float>(o.q); a.e.m = m; a.e.mB = (int)R; a.e.F = F; a.e.col0 = p.att ? A + E : 0;
        a.counters = counters;
      }
      CAPDEC_TRY(gemm(pr, a, st));
    }
    if (p.att) {
      void* z = c.ft(o.z, (int64_t)t * B * E);
      float* awe = save_bwd ? c.at<float>(o.awe) + (int64_t)t * B * E : nullptr;
      CAPDEC_TRY(attention_fwd(pr, c.at(o.att1), c.at(o.enc_s), g1, NG1, A, w.full_att_w, w.full_att_b,
                               alphas + (int64_t)t * P, (int64_t)T * P, z, E, awe, n, 1, P, E, A,
                               c.at<float>(o.att_scr), st));
      // u (in place over U_emb[t]) += z . W_x[:, M:]^T   (+ fused u*v -> left half of m)
      GemmArgs a;
      a.X = z; a.ldx = E; a.W = c.ft(o.Wp_xq, M); a.ldw = p.ldX; a.out = U; a.ldo = NQ; a.addm = U;
      a.ldadd = NQ; a.rows = n; a.N = NQ; a.K = E; a.rows_alloc = B; a.splitk = SK;
      if (fused) {
        a.epi = EPI_P3;
        a.e.fa = c.at<float>(o.v); a.e.m = m; a.e.mB = (int)R; a.e.F = F; a.e.vB = 0;
        a.counters = counters;
      }
      CAPDEC_TRY(gemm(pr, a, st));
    }
    if (fused) {
      // P4 + LSTM pointwise: pre_g = m_g [W_ic_g | W_hc_g]^T for the 4 gates of one d tile in one CTA
      GemmArgs a;
      a.X = m; a.ldx = 2 * F; a.sX = R * 2 * F; a.W = c.at(o.Wp_c); a.ldw = p.ld2F;
      a.sW = (int64_t)D * p.ld2F; a.batch = 4; a.rows = n; a.N = D; a.K = 2 * F; a.rows_alloc = B;
      a.splitk = SK; a.epi = EPI_CELL; a.counters = counters;
      a.abuf = c.at<float>(o.pre) + (int64_t)t * B * 4 * D; a.a_ld = 4 * D; a.a_sa = D;
      a.e.b1 = w.b_ih; a.e.b2 = w.b_hh; a.e.c_prev = c_prev; a.e.c_new = c_new; a.e.gates = gates;
      a.e.h_out = hout; a.e.ldh = (int64_t)T * D; a.e.hd_out = hdout; a.e.dropout_p = dropout_p;
      a.e.seed = c.at<uint64_t>(o.seedD); a.e.t = t; a.e.T = T; a.e.D = D; a.e.lstm_order = 0;
      CAPDEC_TRY(gemm(pr, a, st));
    } else if (p.scn) {
      CAPDEC_TRY(scn_form_m(pr, U, NQ, pcol, NG1, c.at<float>(o.v), c.at<float>(o.q), m, n, (int)R, F, st));
      float* pre = c.at<float>(o.pre) + (int64_t)t * B * 4 * D;
      CAPDEC_TRY(G_(c, m, 2 * F, c.at(o.Wp_c), p.ld2F, pre, 4 * D, 0, nullptr, nullptr, 0, n, D, 2 * F, B, 4,
                    R * 2 * F, (int64_t)D * p.ld2F, D, SK));
      CAPDEC_TRY(cell_fwd(pr, pre, 4 * D, nullptr, 0, w.b_ih, w.b_hh, 0, c_prev, c_new,
                          gates, hout, (int64_t)T * D, hdout, dropout_p, c.at<uint64_t>(o.seedD), t, T, n, D, st));
    } else {
      CAPDEC_TRY(cell_fwd(pr, U, NQ, pcol, NG1, w.b_ih, w.b_hh, 1, c_prev, c_new, gates, hout,
                          (int64_t)T * D, hdout, dropout_p, c.at<uint64_t>(o.seedD), t, T, n, D, st));
    }
  }
  // ---------------- vocabulary projection over all (b,t) rows ----------------
  CAPDEC_TRY(G_(c, drop ? c.at(o.Hd) : c.at(o.Hall), D, c.at(o.Wp_fc), p.ldD, predictions, V, 0, w.fc_b,
               nullptr, 0, (int)R, V, D));
  if (ragged)
    CAPDEC_TRY(zero_rows_beyond_len(predictions, c.at<int32_t>(o.lenD), B, T, V, st));
  return CAPDEC_OK;
}

int backward(const CapdecDims& d, const CapdecParams& w,
             const int32_t* len_h, float dropout_p, const float* d_pred,
             const void* d_logits_ft, const float* d_alphas, const float* alphas,
             const CapdecParams& g, void* workspace, size_t ws_bytes, int phases, cudaStream_t st) {
  // phases (0 = all): the backward in production order of the gradients, so that a data-parallel caller can start
  // the all-reduce of one bucket of the flat gradient buffer while the next one is still being computed:
  //   1  fc.weight / fc.bias and dH_fc                 (before the reverse loop)
  //   2  the reverse-time recurrence
  //   4  weight_ia (weight_ih) and embedding.weight    (the two largest gradients after fc)
  //   8  the other cell weights, both cell biases, init_h / init_c
  //   16 f_beta, decoder_att, full_att            32 encoder_att (the long dAtt1 chain, the smallest bucket: last)
  // Every phase reads only params, workspace, alphas and the d_* inputs -> each is graph-capturable on its own.
  if (phases == 0) phases = 63;
  const bool len_free = len_h == nullptr;   // lengths only on the device: persistent recurrence kernels required
  Ctx c;
  CAPDEC_TRY(make_plan(d, true, &c.p));
  CAPDEC_REQUIRE(ws_bytes >= c.p.o.total, CAPDEC_ERR_WORKSPACE, "workspace %zu < %zu", ws_bytes,
                 c.p.o.total);
  CAPDEC_REQUIRE(d_pred || d_logits_ft, CAPDEC_ERR_BAD_ARG, "backward: no logits gradient");
  c.ws = (uint8_t*)workspace;
  c.st = st;
  c.prec = d.precision;
  const Plan& p = c.p;
  const Plan::Off& o = p.o;
  const int pr = c.prec;
  const int B = d.B, T = d.T, P = d.P, E = d.E, A = d.A, M = d.M, D = d.D, F = d.F, S = d.S, V = d.V,
            X = p.X, NQ = p.NQ, NG1 = p.NG1;
  const int64_t R = (int64_t)B * T;
  const bool ragged = len_free || len_h[B - 1] != T;
  const bool drop = dropout_p > 0.f;
  std::vector<int> bt(T, B);
  for (int t = 0; t < T && !len_free; ++t) {
    int n = 0;
    while (n < B && len_h[n] > t) ++n;
    bt[t] = n;
  }

  // ---- gradient wrt the logits in the GEMM operand type ----
  const void* dlog;
  int64_t lddl;
  if (d_logits_ft) { dlog = d_logits_ft; lddl = p.ldV; }
  else if (pr == CAPDEC_FP32) { dlog = d_pred; lddl = V; }
  else {
    if (phases & 1) CAPDEC_TRY(copy_cast(pr, d_pred, 0, V, c.at(o.dlogF), 1, p.ldV, (int)R, V, st));
    dlog = c.at(o.dlogF); lddl = p.ldV;
  }
  const bool tn = pr == CAPDEC_BF16;
  if (phases & 1) {
  // dH_fc[(b,t), :] = dlogits . W_fc
  // K = V is long and the output small (R x D): split K four ways in bf16 mode (fp32 atomics into a zeroed buffer)
  const int sk_fc = pr == CAPDEC_BF16 ? 4 : 0;
  if (sk_fc) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dHfc), 0, (size_t)R * D * 4, st));
  CAPDEC_TRY(G_(c, dlog, lddl, c.at(o.Wp_fcT), p.ldV, c.at(o.dHfc), D, 0, nullptr, nullptr, 0, (int)R, D, V, 0, 1, 0, 0,
                0, sk_fc));
  // fc.weight.grad = dlogits^T . dropout(H) ; fc.bias.grad = colsum(dlogits)     (rows in (b,t) order)
  // bf16: the operands stay as they are ([sample][feature]) -- transposed-operand GEMM, no transposition pass
  if (tn) {
    CAPDEC_TRY(GT_(c, 3, dlog, lddl, drop ? c.at(o.Hd) : c.at(o.Hall), D, g.fc_w, D, V, D, (int)R));
  } else {
    CAPDEC_TRY(transpose_cast(pr, dlog, 1, c.at(o.tA), 1, 1, (int)R, V, 0, lddl, p.ldR, 0, 1, st));
    CAPDEC_TRY(transpose_cast(pr, drop ? c.at(o.Hd) : c.at(o.Hall), 1, c.at(o.tB), 1, 1, (int)R, D, 0, D,
                              p.ldR, 0, 1, st));
    CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.fc_w, D, 0, nullptr, nullptr, 0, V, D, (int)R));
  }
  CAPDEC_TRY(colsum(pr, dlog, 1, lddl, (int)R, V, g.fc_b, 0, st));
  }   // phase 1

  const int SK = pr == CAPDEC_BF16 ? -1 : 0;
  const bool fused = p.fused;
  // where the loop leaves dp (gradient wrt p = h W_ha) and [dbeta_pre | datt2]
  // SCN: one row of dpx = [dp | dbeta_pre | datt2], so ONE GEMM with [W_ha | W_beta^T | W_d^T] gives dh
  const void* dp_all = p.scn ? c.at(o.dpx) : nullptr;
  const int64_t lddp = p.ldPX;
  const void* dba_all = p.scn ? c.ft(o.dpx, NQ) : c.at(o.dba);
  const int64_t lddba = p.scn ? p.ldPX : p.ldEA;
  int* counters = c.at<int>(o.counters);
  const int Ppad = (P + 3) / 4 * 4;
  const int Ri = (int)R;
  if (phases & 2) {
  // will the reverse loop run as the persistent kernel?  It then needs none of the zeroed accumulators of the split-K
  // kernel chains (it writes dh_0 / dc_0 itself and fills its exchange buffers in its launcher)
  bool will_persist = false;
  if (pr == CAPDEC_BF16 && (p.scn || p.att) && !fused && (!p.att || alphas)) {
    RecurBwdArgs probe;
    probe.att = p.att ? 1 : 0; probe.lstm = p.scn ? 0 : 1;
    probe.B = B; probe.T = T; probe.P = P; probe.E = E; probe.A = A; probe.M = M; probe.D = D; probe.F = F;
    will_persist = recur_bwd_supported(probe);
  }
  if (!will_persist) {
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dh_rec), 0, (size_t)(T + 1) * B * D * 4, st));
    if (SK) {
      if (p.scn) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.wr), 0, (size_t)R * 4 * 2 * F * 4, st));
      if (p.att) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dz), 0, (size_t)R * E * 4, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)GEMM_TC_MAX_TILE_COUNTERS * 4, st));
    }
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dc), 0, (size_t)B * D * 4, st));
  }
  if (p.scn) {
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dv_acc), 0, (size_t)B * NQ * 4, st));
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dq_acc), 0, (size_t)B * NQ * 4, st));
  }
  // de feeds the batched dAtt1 kernel over ALL (t,b) rows: rows beyond a caption's length must be zero
  if (p.att && (ragged || !will_persist)) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.de), 0, (size_t)R * Ppad * 4, st));
  if (ragged) {
    CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dpre), 0, (size_t)R * 4 * D * p.fsz, st));
    if (p.scn) {
      CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.du), 0, (size_t)R * NQ * p.fsz, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dpx), 0, (size_t)R * p.ldPX * p.fsz, st));
    }
    if (p.att) {
      if (!p.scn) CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dba), 0, (size_t)R * p.ldEA * p.fsz, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dwf), 0, (size_t)R * A * 4, st));
      CAPDEC_CUDA_OK(cudaMemsetAsync(c.at(o.dbf), 0, (size_t)R * 4, st));
    }
  }

  // ---------------- reverse-time recurrence as ONE persistent cooperative kernel (recur.cu) ----------------
  bool persistent = false;
  if (pr == CAPDEC_BF16 && (p.scn || p.att) && !fused) {
    RecurBwdArgs rb;
    rb.att = p.att ? 1 : 0;
    rb.lstm = p.scn ? 0 : 1;
    rb.B = B; rb.T = T; rb.P = P; rb.E = E; rb.A = A; rb.M = M; rb.D = D; rb.F = F; rb.ldPX = p.ldPX;
    rb.len = c.at<int32_t>(o.lenD);
    rb.WcT = p.scn ? c.at(o.Wp_cT) : nullptr; rb.ldD = p.ldD;
    rb.Wxin = c.ft(o.Wp_xin, (int64_t)M * p.ldNQ); rb.ldNQ = p.ldNQ;
    if (p.scn) {
      rb.Whx = c.at(o.Wp_hx); rb.ldhx = p.ldPX;
      rb.dbx = c.at(o.dpx); rb.ldbx = p.ldPX; rb.dbx_off = NQ;
    } else {
      rb.Whx = c.at(o.Wp_hq); rb.ldhx = p.ldNQ;          // W_hh^T [D][4D]
      rb.Whx2 = c.at(o.Wp_b6); rb.ldhx2 = p.ldEA;        // [W_beta^T | W_d^T] [D][E+A]
      rb.dbx = c.at(o.dba); rb.ldbx = p.ldEA; rb.dbx_off = 0;
    }
    rb.dHfc = c.at<float>(o.dHfc); rb.gates = c.at<float>(o.gates); rb.C = c.at<float>(o.C);
    rb.dc = c.at<float>(o.dc); rb.dh_rec = c.at<float>(o.dh_rec); rb.dhp = c.at<float>(o.dhp);
    rb.dpre = c.at(o.dpre); rb.dpre_gm = c.at(o.dpre_gm);
    rb.U = c.at<float>(o.U); rb.g1 = c.at<float>(o.g1);
    if (p.scn) {
      rb.v = c.at<float>(o.v); rb.q = c.at<float>(o.q);
      rb.du = c.at(o.du); rb.duk = c.at(o.duk); rb.dpx = c.at(o.dpx);
      rb.dv_acc = c.at<float>(o.dv_acc); rb.dq_acc = c.at<float>(o.dq_acc);
    }
    rb.dpxk = c.at(o.dpxk);
    if (p.att) {
      rb.dz = c.at<float>(o.dz); rb.awe = c.at<float>(o.awe); rb.alphas = alphas; rb.d_alphas = d_alphas;
      rb.enc_cm = c.at(o.enc_cm); rb.att1 = c.at(o.att1); rb.att1_cm = c.at(o.att1_cm); rb.w_f = w.full_att_w;
      rb.enc = c.at(o.enc_s); rb.build_enc_cm = E % 256 == 0 ? 0 : 1;       // the forward's input phase wrote it
      rb.part = c.at<float>(o.part_t); rb.de = c.at<float>(o.de); rb.dwf = c.at<float>(o.dwf);
      rb.dbf = c.at<float>(o.dbf);
    }
    rb.bar = c.at<unsigned>(o.bar); rb.dropout_p = dropout_p; rb.seed = c.at<uint64_t>(o.seedD);
    if ((!p.att || alphas) && recur_bwd_supported(rb)) {
      persistent = true;
      CAPDEC_TRY(recur_bwd(rb, st));
    }
  }
  CAPDEC_REQUIRE(persistent || !len_free, CAPDEC_ERR_UNSUPPORTED,
                 "length-independent launch needs the persistent recurrence kernels (bf16, shape covered by recur.cu)");
  // ---------------- reverse-time recurrence, one kernel chain per step ----------------
  auto cell_bwd_step = [&](int t, const float* dh_in) {
    return cell_bwd(pr, c.at<float>(o.dHfc) + (int64_t)t * D, (int64_t)T * D, dh_in, c.at<float>(o.dc),
                    c.at<float>(o.gates) + (int64_t)t * B * 4 * D, c.at<float>(o.C) + (int64_t)t * B * D,
                    c.at<float>(o.C) + (int64_t)(t + 1) * B * D, p.scn ? 0 : 1, dropout_p,
                    c.at<uint64_t>(o.seedD), t, T, c.ft(o.dpre, (int64_t)t * B * 4 * D), nullptr, bt[t], D, st);
  };
  if (fused)    // the last step's pointwise backward; every earlier one rides on the dh GEMM below
    CAPDEC_TRY(cell_bwd_step(T - 1, c.at<float>(o.dh_rec) + (int64_t)T * B * D));
  for (int t = persistent ? -1 : T - 1; t >= 0; --t) {
    const int n = bt[t];
    float* g1 = c.at<float>(o.g1) + (int64_t)t * B * NG1;
    const float* pcol = g1 + (p.att ? A + E : 0);
    const float* U = c.at<float>(o.U) + (int64_t)t * B * NQ;
    void* dpre = c.ft(o.dpre, (int64_t)t * B * 4 * D);
    float* dz = p.att ? c.at<float>(o.dz) + (int64_t)t * B * E : nullptr;
    float* de_t = p.att ? c.at<float>(o.de) + (int64_t)t * B * Ppad : nullptr;
    if (fused) {
      void* du_t = c.ft(o.du, (int64_t)t * B * NQ);
      void* dpx_t = c.ft(o.dpx, (int64_t)t * B * p.ldPX);
      {
        // [w_g | r_g] = dpre_g [W_ic_g | W_hc_g] with the factor products fused in the epilogue:
        // du = w*v, dp = r*q, dv_acc += w*u, dq_acc += r*p
        GemmArgs a;
        a.X = dpre; a.ldx = 4 * D; a.sX = D; a.W = c.at(o.Wp_cT); a.ldw = p.ldD;
        a.sW = (int64_t)2 * F * p.ldD; a.batch = 4; a.rows = n; a.N = 2 * F; a.K = D; a.rows_alloc = B;
        a.splitk = SK; a.epi = EPI_WR; a.counters = counters;
        a.abuf = c.at<float>(o.wr) + (int64_t)t * 4 * B * 2 * F; a.a_ld = 2 * F; a.a_sz = (int64_t)B * 2 * F;
        a.e.fa = c.at<float>(o.v); a.e.fb = c.at<float>(o.q); a.e.fc = U; a.e.ldc = NQ; a.e.fd = pcol;
        a.e.ldd = NG1; a.e.du = du_t; a.e.lddu = NQ; a.e.dp = dpx_t; a.e.lddp = p.ldPX;
        a.e.dv_acc = c.at<float>(o.dv_acc); a.e.dq_acc = c.at<float>(o.dq_acc); a.e.F = F;
        CAPDEC_TRY(gemm(pr, a, st));
      }
      if (p.att) {
        // dz = du . W_x[M:, :]^T ; attention backward writes [dbeta_pre | datt2] next to dp
        CAPDEC_TRY(G_(c, du_t, NQ, c.ft(o.Wp_xin, (int64_t)M * p.ldNQ), p.ldNQ, dz, E, 0, nullptr,
                      nullptr, 0, n, E, NQ, B, 1, 0, 0, 0, SK));
        CAPDEC_TRY(attention_bwd(pr, c.at(o.att1), c.at(o.enc_s), g1, NG1, A, w.full_att_w,
                                 alphas + (int64_t)t * P, (int64_t)T * P,
                                 d_alphas ? d_alphas + (int64_t)t * P : nullptr, (int64_t)T * P,
                                 dz, E, c.at<float>(o.awe) + (int64_t)t * B * E,
                                 c.ft(o.dpx, (int64_t)t * B * p.ldPX + NQ), p.ldPX,
                                 de_t, c.at<float>(o.dwf) + (int64_t)t * B * A,
                                 c.at<float>(o.dbf) + (int64_t)t * B, n, P, E, A, c.at<float>(o.att_scr), st));
      }
      {
        // dh_{t-1} = [dp | dbeta_pre | datt2] [W_ha | W_beta^T | W_d^T]^T, and (t > 0) the LSTM
        // pointwise backward of step t-1 in the epilogue
        GemmArgs a;
        a.X = dpx_t; a.ldx = p.ldPX; a.W = c.at(o.Wp_hx); a.ldw = p.ldPX; a.rows = n; a.N = D;
        a.K = NQ + (p.att ? E + A : 0); a.rows_alloc = B; a.splitk = SK;
        if (t > 0) {
          const int tp = t - 1;
          a.epi = EPI_DHCELL; a.counters = counters;
          a.abuf = c.at<float>(o.dh_rec) + (int64_t)t * B * D; a.a_ld = D;     // zeroed slot of step t
          a.e.rows_epi = bt[tp];
          a.e.dh_fc = c.at<float>(o.dHfc) + (int64_t)tp * D; a.e.ld_dhfc = (int64_t)T * D;
          a.e.gates = c.at<float>(o.gates) + (int64_t)tp * B * 4 * D;
          a.e.c_prev = c.at<float>(o.C) + (int64_t)tp * B * D;
          a.e.c_new_r = c.at<float>(o.C) + (int64_t)(tp + 1) * B * D;
          a.e.dc = c.at<float>(o.dc); a.e.dpre = c.ft(o.dpre, (int64_t)tp * B * 4 * D);
          a.e.dropout_p = dropout_p; a.e.seed = c.at<uint64_t>(o.seedD); a.e.t = tp; a.e.T = T; a.e.D = D;
          a.e.lstm_order = 0;
        } else {
          a.out = c.at<float>(o.dh_rec); a.ldo = D;       // dh0 -> init_h gradients
        }
        CAPDEC_TRY(gemm(pr, a, st));
      }
      continue;
    }
    // ---- default path: plain (split-K) GEMMs + pointwise kernels, chained with PDL ----
    // dhs[t+1]: recurrent gradient flowing into h_t ; dhs[t]: what this step sends to h_{t-1}
    const float* dh_in = c.at<float>(o.dh_rec) + (int64_t)(t + 1) * B * D;
    float* dh_rec = c.at<float>(o.dh_rec) + (int64_t)t * B * D;
    CAPDEC_TRY(cell_bwd_step(t, dh_in));
    if (p.scn) {
      float* wr = c.at<float>(o.wr) + (int64_t)t * 4 * B * 2 * F;
      void* du_t = c.ft(o.du, (int64_t)t * B * NQ);
      void* dpx_t = c.ft(o.dpx, (int64_t)t * B * p.ldPX);
      // [w_g | r_g] = dpre_g . [W_ic_g | W_hc_g] ; du = w*v, dp = r*q, dv_acc += w*u, dq_acc += r*p
      CAPDEC_TRY(G_(c, dpre, 4 * D, c.at(o.Wp_cT), p.ldD, wr, 2 * F, 0, nullptr, nullptr, 0, n, 2 * F,
                   D, B, 4, D, (int64_t)2 * F * p.ldD, (int64_t)B * 2 * F, SK));
      CAPDEC_TRY(scn_bwd_products(pr, wr, U, NQ, pcol, NG1, c.at<float>(o.v),
                                  c.at<float>(o.q), du_t, dpx_t, c.at<float>(o.dv_acc),
                                  c.at<float>(o.dq_acc), n, B, F, p.ldPX, st));
      if (p.att) {
        // dz = du . W_x[M:, :]^T ; attention backward writes [dbeta_pre | datt2] next to dp
        CAPDEC_TRY(G_(c, du_t, NQ, c.ft(o.Wp_xin, (int64_t)M * p.ldNQ), p.ldNQ, dz, E, 0, nullptr,
                      nullptr, 0, n, E, NQ, B, 1, 0, 0, 0, SK));
        CAPDEC_TRY(attention_bwd(pr, c.at(o.att1), c.at(o.enc_s), g1, NG1, A, w.full_att_w,
                                 alphas + (int64_t)t * P, (int64_t)T * P,
                                 d_alphas ? d_alphas + (int64_t)t * P : nullptr, (int64_t)T * P,
                                 dz, E, c.at<float>(o.awe) + (int64_t)t * B * E,
                                 c.ft(o.dpx, (int64_t)t * B * p.ldPX + NQ), p.ldPX,
                                 de_t, c.at<float>(o.dwf) + (int64_t)t * B * A,
                                 c.at<float>(o.dbf) + (int64_t)t * B, n, P, E, A, c.at<float>(o.att_scr), st));
      }
      // dh_{t-1} = [dp | dbeta_pre | datt2] . [W_ha | W_beta^T | W_d^T]^T
      CAPDEC_TRY(G_(c, dpx_t, p.ldPX, c.at(o.Wp_hx), p.ldPX, dh_rec, D, 0, nullptr, nullptr, 0, n, D,
                    NQ + (p.att ? E + A : 0), B, 1, 0, 0, 0, SK));
    } else {
      // LSTM: dh_{t-1} (recurrent part) = dpre . W_hh
      CAPDEC_TRY(G_(c, dpre, NQ, c.at(o.Wp_hq), p.ldNQ, dh_rec, D, 0, nullptr, nullptr, 0, n, D, NQ, B, 1, 0, 0, 0, SK));
      // dz = dpre . W_ih[:, M:]
      CAPDEC_TRY(G_(c, dpre, NQ, c.ft(o.Wp_xin, (int64_t)M * p.ldNQ), p.ldNQ, dz, E, 0, nullptr,
                   nullptr, 0, n, E, NQ, B, 1, 0, 0, 0, SK));
      void* dba = c.ft(o.dba, (int64_t)t * B * p.ldEA);
      CAPDEC_TRY(attention_bwd(pr, c.at(o.att1), c.at(o.enc_s), g1, NG1, A, w.full_att_w,
                               alphas + (int64_t)t * P, (int64_t)T * P,
                               d_alphas ? d_alphas + (int64_t)t * P : nullptr, (int64_t)T * P,
                               dz, E, c.at<float>(o.awe) + (int64_t)t * B * E, dba, p.ldEA,
                               de_t, c.at<float>(o.dwf) + (int64_t)t * B * A,
                               c.at<float>(o.dbf) + (int64_t)t * B, n, P, E, A, c.at<float>(o.att_scr), st));
      // dh_{t-1} += [dbeta_pre | datt2] . [W_beta^T | W_d^T]^T
      CAPDEC_TRY(G_(c, dba, p.ldEA, c.at(o.Wp_b6), p.ldEA, dh_rec, D, 0, nullptr, dh_rec, D, n, D, E + A, B, 1,
                   0, 0, 0, SK));
    }
  }

  // H_prev^T [D][R]: column (t,b) = h_{t-1}[b]  (t=0 -> H0, else Hall[b][t-1]); used by phases 8 and 16
  CAPDEC_TRY(transpose_cast(pr, c.at(o.H0), 1, c.at(o.tC), 1, 1, B, D, 0, p.ldD, p.ldR, 0, 1, st));
  if (T > 1)
    CAPDEC_TRY(transpose_cast(pr, c.at(o.Hall), 1, c.ft(o.tC, B), 1, T - 1, B, D, D, (int64_t)T * D, p.ldR,
                              B, 1, st));
  }   // phase 2

  // ---------------- weight gradients: batched GEMMs over all (t,b) rows ----------------
  // X-side operand = (d out-feature)^T [N_out][R], W-side = (input)^T [K_in][R]; K = R rows.
  // bf16: the operands stay as they are ([sample][feature]) -- transposed-operand GEMM, no transposition pass
  if (phases & 4) {
    // ---- weight_ia.grad [X][4F] = [Xe | z]^T . du  (LSTM: weight_ih.grad [4D][X] = dpre^T . [Xe | z]) and
    // embedding.weight.grad: dXe = du . W_ia[:M]^T, scattered to the consumed token rows ----
    if (tn) {
      if (p.scn) {
        CAPDEC_TRY(GT_(c, 3, c.at(o.Xe), p.ldM, c.at(o.du), NQ, g.w_ia, NQ, M, NQ, Ri));
        if (p.att)
          CAPDEC_TRY(GT_(c, 3, c.at(o.z), E, c.at(o.du), NQ, g.w_ia + (int64_t)M * NQ, NQ, E, NQ, Ri));
        CAPDEC_TRY(G_(c, c.at(o.du), NQ, c.at(o.Wp_xin), p.ldNQ, c.at(o.dXe), M, 0, nullptr, nullptr, 0, Ri, M, NQ));
      } else {
        CAPDEC_TRY(GT_(c, 3, c.at(o.dpre), NQ, c.at(o.Xe), p.ldM, g.w_ia, X, NQ, M, Ri));
        if (p.att) CAPDEC_TRY(GT_(c, 3, c.at(o.dpre), NQ, c.at(o.z), E, g.w_ia + M, X, NQ, E, Ri));
        CAPDEC_TRY(G_(c, c.at(o.dpre), NQ, c.at(o.Wp_xin), p.ldNQ, c.at(o.dXe), M, 0, nullptr, nullptr, 0, Ri, M, NQ));
      }
    } else if (p.scn) {
      CAPDEC_TRY(transpose_cast(pr, c.at(o.du), 1, c.at(o.tB), 1, 1, Ri, NQ, 0, NQ, p.ldR, 0, 1, st));
      CAPDEC_TRY(transpose_cast(pr, c.at(o.Xe), 1, c.at(o.tA), 1, 1, Ri, M, 0, p.ldM, p.ldR, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.w_ia, NQ, 0, nullptr, nullptr, 0, M, NQ, Ri));
      if (p.att) {
        CAPDEC_TRY(transpose_cast(pr, c.at(o.z), 1, c.at(o.tA), 1, 1, Ri, E, 0, E, p.ldR, 0, 1, st));
        CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.w_ia + (int64_t)M * NQ, NQ, 0, nullptr,
                     nullptr, 0, E, NQ, Ri));
      }
      CAPDEC_TRY(G_(c, c.at(o.du), NQ, c.at(o.Wp_xin), p.ldNQ, c.at(o.dXe), M, 0, nullptr, nullptr, 0, Ri, M, NQ));
    } else {
      CAPDEC_TRY(transpose_cast(pr, c.at(o.dpre), 1, c.at(o.tA), 1, 1, Ri, 4 * D, 0, 4 * D, p.ldR, 0, 1, st));
      CAPDEC_TRY(transpose_cast(pr, c.at(o.Xe), 1, c.at(o.tB), 1, 1, Ri, M, 0, p.ldM, p.ldR, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.w_ia, X, 0, nullptr, nullptr, 0, NQ, M, Ri));
      if (p.att) {
        CAPDEC_TRY(transpose_cast(pr, c.at(o.z), 1, c.at(o.tB), 1, 1, Ri, E, 0, E, p.ldR, 0, 1, st));
        CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.w_ia + M, X, 0, nullptr, nullptr, 0, NQ, E, Ri));
      }
      CAPDEC_TRY(G_(c, c.at(o.dpre), NQ, c.at(o.Wp_xin), p.ldNQ, c.at(o.dXe), M, 0, nullptr, nullptr, 0, Ri, M, NQ));
    }
    if (g.emb) {
      CAPDEC_CUDA_OK(cudaMemsetAsync(g.emb, 0, (size_t)V * M * 4, st));
      CAPDEC_TRY(embedding_scatter_add(c.at<float>(o.dXe), M, c.at<int64_t>(o.capsD), d.L, c.at<int32_t>(o.lenD), g.emb, B, T,
                                       M, V, st));
    }
  }   // phase 4

  if (phases & 8) {
    // bias_ih.grad == bias_hh.grad = colsum(dpre)
    CAPDEC_TRY(colsum(pr, c.at(o.dpre), 1, 4 * D, Ri, 4 * D, g.b_ih, 0, st));
    CAPDEC_CUDA_OK(cudaMemcpyAsync(g.b_hh, g.b_ih, (size_t)4 * D * 4, cudaMemcpyDeviceToDevice, st));
    if (tn) {
      if (p.scn) {
        // weight_ic.grad[:, gF:(g+1)F] = dpre_g^T . (u_g*v_g) ; weight_hc.grad likewise with (p_g*q_g)
        CAPDEC_TRY(GT_(c, 3, c.at(o.dpre), 4 * D, c.at(o.m), 2 * F, g.w_ic, NQ, D, F, Ri, 4, D, R * 2 * F, F));
        CAPDEC_TRY(GT_(c, 3, c.at(o.dpre), 4 * D, c.ft(o.m, F), 2 * F, g.w_hc, NQ, D, F, Ri, 4, D, R * 2 * F, F));
        // weight_ha.grad [D][4F] = H_prev^T . dp
        CAPDEC_TRY(GT_(c, 1, c.at(o.tC), p.ldR, dp_all, lddp, g.w_ha, NQ, D, NQ, Ri));
      } else {
        // LSTM: weight_hh.grad [4D][D] = dpre^T . H_prev
        CAPDEC_TRY(GT_(c, 2, c.at(o.dpre), NQ, c.at(o.tC), p.ldR, g.w_ha, D, NQ, D, Ri));
      }
    } else {
      // dpre^T [4D][R]
      CAPDEC_TRY(transpose_cast(pr, c.at(o.dpre), 1, c.at(o.tA), 1, 1, Ri, 4 * D, 0, 4 * D, p.ldR, 0, 1, st));
      if (p.scn) {
        // m^T: per gate [2F][R]
        for (int gg = 0; gg < 4; ++gg)
          CAPDEC_TRY(transpose_cast(pr, c.ft(o.m, (int64_t)gg * R * 2 * F), 1,
                                    c.ft(o.tB, (int64_t)gg * 2 * F * p.ldR), 1, 1, Ri, 2 * F, 0, 2 * F,
                                    p.ldR, 0, 1, st));
        CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tB), p.ldR, g.w_ic, NQ, 0, nullptr, nullptr, 0, D, F, Ri, 0, 4,
                     (int64_t)D * p.ldR, (int64_t)2 * F * p.ldR, F));
        CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.ft(o.tB, (int64_t)F * p.ldR), p.ldR, g.w_hc, NQ, 0, nullptr, nullptr,
                     0, D, F, Ri, 0, 4, (int64_t)D * p.ldR, (int64_t)2 * F * p.ldR, F));
        // weight_ha.grad [D][4F] = H_prev^T . dp
        CAPDEC_TRY(transpose_cast(pr, dp_all, 1, c.at(o.tB), 1, 1, Ri, NQ, 0, lddp, p.ldR, 0, 1, st));
        CAPDEC_TRY(G_(c, c.at(o.tC), p.ldR, c.at(o.tB), p.ldR, g.w_ha, NQ, 0, nullptr, nullptr, 0, D, NQ, Ri));
      } else {
        CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tC), p.ldR, g.w_ha, D, 0, nullptr, nullptr, 0, NQ, D, Ri));
      }
    }
    if (p.scn) {
      // weight_ib.grad [S][4F] = s^T . sum_t dv ; weight_hb.grad = s^T . sum_t dq   (B rows: tiny; the tag matrix
      // as the forward saw it: the feature-type copy kept in the workspace)
      CAPDEC_TRY(transpose_cast(pr, c.at(o.tagsF), 1, c.at(o.tA), 1, 1, B, S, 0, p.ldS, p.ldB, 0, 1, st));
      CAPDEC_TRY(transpose_cast(pr, c.at(o.dv_acc), 0, c.at(o.tB), 1, 1, B, NQ, 0, NQ, p.ldB, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldB, c.at(o.tB), p.ldB, g.w_ib, NQ, 0, nullptr, nullptr, 0, S, NQ, B));
      CAPDEC_TRY(transpose_cast(pr, c.at(o.dq_acc), 0, c.at(o.tB), 1, 1, B, NQ, 0, NQ, p.ldB, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldB, c.at(o.tB), p.ldB, g.w_hb, NQ, 0, nullptr, nullptr, 0, S, NQ, B));
    }
    // init_h / init_c: dh0 = dh_rec, dc0 = dc after the last reverse step
    CAPDEC_TRY(transpose_cast(pr, c.at(o.mean), 0, c.at(o.tB), 1, 1, B, E, 0, E, p.ldB, 0, 1, st));
    CAPDEC_TRY(transpose_cast(pr, c.at(o.dh_rec), 0, c.at(o.tA), 1, 1, B, D, 0, D, p.ldB, 0, 1, st));
    CAPDEC_TRY(G_(c, c.at(o.tA), p.ldB, c.at(o.tB), p.ldB, g.init_h_w, E, 0, nullptr, nullptr, 0, D, E, B));
    CAPDEC_TRY(colsum(pr, c.at(o.dh_rec), 0, D, B, D, g.init_h_b, 0, st));
    CAPDEC_TRY(transpose_cast(pr, c.at(o.dc), 0, c.at(o.tA), 1, 1, B, D, 0, D, p.ldB, 0, 1, st));
    CAPDEC_TRY(G_(c, c.at(o.tA), p.ldB, c.at(o.tB), p.ldB, g.init_c_w, E, 0, nullptr, nullptr, 0, D, E, B));
    CAPDEC_TRY(colsum(pr, c.at(o.dc), 0, D, B, D, g.init_c_b, 0, st));
  }   // phase 8

  if (phases & 16) {
  if (p.att) {
    // f_beta / decoder_att: [dbeta_pre | datt2]^T . H_prev
    if (tn) {
      CAPDEC_TRY(GT_(c, 2, dba_all, lddba, c.at(o.tC), p.ldR, g.f_beta_w, D, E, D, Ri));
      CAPDEC_TRY(GT_(c, 2, (const uint8_t*)dba_all + (size_t)E * p.fsz, lddba, c.at(o.tC), p.ldR, g.dec_att_w, D, A, D, Ri));
    } else {
      CAPDEC_TRY(transpose_cast(pr, dba_all, 1, c.at(o.tA), 1, 1, Ri, E + A, 0, lddba, p.ldR, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldR, c.at(o.tC), p.ldR, g.f_beta_w, D, 0, nullptr, nullptr, 0, E, D, Ri));
      CAPDEC_TRY(G_(c, c.ft(o.tA, (int64_t)E * p.ldR), p.ldR, c.at(o.tC), p.ldR, g.dec_att_w, D, 0, nullptr,
                   nullptr, 0, A, D, Ri));
    }
    CAPDEC_TRY(colsum(pr, dba_all, 1, lddba, Ri, E, g.f_beta_b, 0, st));
    CAPDEC_TRY(colsum(pr, (const uint8_t*)dba_all + (size_t)E * p.fsz, 1, lddba, Ri, A, g.dec_att_b, 0, st));
    // full_att: per-row partials reduced over all (t,b)
    CAPDEC_TRY(colsum(pr, c.at(o.dwf), 0, A, Ri, A, g.full_att_w, 0, st));
    CAPDEC_TRY(colsum(pr, c.at(o.dbf), 0, 1, Ri, 1, g.full_att_b, 0, st));
  }
  }   // phase 16

  if (phases & 32) {
  if (p.att) {
    // dAtt1[b,p,:] = w_f * sum_t de[t,b,p] 1[att1[b,p,:] + att2_t[b,:] > 0]  (masks rebuilt, not stored)
    CAPDEC_TRY(attention_datt1(pr, c.at(o.att1), c.at<float>(o.g1), NG1, (int64_t)B * NG1, c.at<float>(o.de),
                               (int64_t)B * Ppad, w.full_att_w, c.at<float>(o.dAtt1), 0, B, T, P, A, st));
    // encoder_att: dAtt1^T . enc  (K = B*P pixel rows)
    const int BP = B * P;
    CAPDEC_TRY(colsum(pr, c.at(o.dAtt1), 0, A, BP, A, g.enc_att_b, 0, st));
    if (tn) {
      CAPDEC_TRY(copy_cast(pr, c.at(o.dAtt1), 0, A, c.at(o.tA), 1, p.ldA, BP, A, st));      // fp32 -> bf16, same layout
      CAPDEC_TRY(GT_(c, 3, c.at(o.tA), p.ldA, c.at(o.enc_s), E, g.enc_att_w, E, A, E, BP));
    } else {
      CAPDEC_TRY(transpose_cast(pr, c.at(o.dAtt1), 0, c.at(o.tA), 1, 1, BP, A, 0, A, p.ldBP, 0, 1, st));
      CAPDEC_TRY(transpose_cast(pr, c.at(o.enc_s), 1, c.at(o.tB), 1, 1, BP, E, 0, E, p.ldBP, 0, 1, st));
      CAPDEC_TRY(G_(c, c.at(o.tA), p.ldBP, c.at(o.tB), p.ldBP, g.enc_att_w, E, 0, nullptr, nullptr, 0, A, E, BP));
    }
  }
  }   // phase 32
  return CAPDEC_OK;
}


// =====================================================================================
// Batched beam search (reference `sample`: attention_scn.py:160-296, pure_scn.py:142-249,
// pure_attention.py:153-281).  G independent searches of k beams advance together as
// R = G*k rows; image g owns rows g*k .. g*k+k-1 and all of them read the SAME feature map
// (the reference merely `expand`s it, :189), so the attention kernel runs with
// rows_per_map = k.  Every step is the training step's kernel sequence in eval mode followed
// by the vocabulary GEMM, the fused log-softmax + top-k selection (beam.cu) and the re-ordering
// of the recurrent state; the host is never consulted inside the loop.
// =====================================================================================
namespace {

struct BeamPlan {
  Plan w;                 // weight offsets only (built for B = 1, T = 1)
  int R;
  struct Off {
    size_t enc_f, enc_cm, att1, mean, meanF, meanX, tagsG, tagsX, v, q, H, C, Hn, Cn, Xe, xz, U, g1, z, m, pre,
        logits, vpart, alpha_hist, att_scr, prev_word, scoreA, scoreB, src_row, live, krem, has_done, best_score,
        best_t, best_parent, bp_parent, bp_word, total;
  } o;
};

int make_beam_plan(const CapdecDims& d_in, int G, int k, int n_steps, bool want_alpha, BeamPlan* bp) {
  CAPDEC_REQUIRE(G > 0 && k >= 1 && k <= 8 && n_steps >= 1 && n_steps <= 62, CAPDEC_ERR_BAD_SHAPE,
                 "beam search: bad G=%d k=%d steps=%d", G, k, n_steps);
  CapdecDims d = d_in;
  d.B = 1; d.T = 1; d.L = 2;
  CAPDEC_TRY(make_plan(d, false, &bp->w));
  const Plan& p = bp->w;
  const int64_t R = (int64_t)G * k;
  bp->R = (int)R;
  const size_t f = p.fsz;
  const int64_t P = d.P, E = d.E, A = d.A, D = d.D, F = d.F, V = d.V, NQ = p.NQ, NG1 = p.NG1;
  size_t cur = p.o.enc_s;            // first byte after the packed weights
  auto take = [&](size_t bytes) { size_t at = cur; cur += (size_t)round_up((int64_t)bytes, 256); return at; };
  BeamPlan::Off& o = bp->o;
  memset(&o, 0, sizeof o);
  o.enc_f = take((size_t)G * P * E * f);
  if (p.att && d.precision == CAPDEC_BF16 && E % 512 == 0) o.enc_cm = take((size_t)G * P * E * f);   // streaming weighted sum
  if (p.att) o.att1 = take((size_t)G * P * A * f);
  o.mean = take((size_t)G * E * 4);
  o.meanF = take((size_t)G * p.ldE * f);
  o.meanX = take((size_t)R * p.ldE * f);
  if (p.scn) {
    o.tagsG = take((size_t)G * p.ldS * f);
    o.tagsX = take((size_t)R * p.ldS * f);
    o.v = take((size_t)R * NQ * 4);
    o.q = take((size_t)R * NQ * 4);
  }
  o.H = take((size_t)R * p.ldD * f);
  o.C = take((size_t)R * D * 4);
  o.Hn = take((size_t)R * p.ldD * f);
  o.Cn = take((size_t)R * D * 4);
  o.Xe = take((size_t)R * p.ldM * f);
  o.U = take((size_t)R * NQ * 4);
  o.g1 = take((size_t)R * NG1 * 4);
  if (p.att) o.z = take((size_t)R * E * f);
  if (p.att) o.xz = take((size_t)R * (d.M + E) * f);     // [emb_t | z] rows: ONE K = M+E input-side GEMM per step
  if (p.scn) {
    o.m = take((size_t)4 * R * 2 * F * f);
    o.pre = take((size_t)R * 4 * D * 4);
  }
  o.logits = take((size_t)R * V * 4);
  if (d.precision == CAPDEC_BF16) o.vpart = take(vocab_topk_part_floats((int)R, (int)V, k <= 4 ? 4 : 8) * 4);
  if (p.att) o.att_scr = take(attention_scratch_floats(d.precision, (int)R, (int)P, (int)E) * 4);
  if (p.att) o.alpha_hist = take((size_t)(want_alpha ? n_steps : 1) * R * P * 4);
  o.prev_word = take((size_t)R * 4);
  o.scoreA = take((size_t)R * 4);
  o.scoreB = take((size_t)R * 4);
  o.src_row = take((size_t)R * 4);
  o.live = take((size_t)G * 4);
  o.krem = take((size_t)G * 4);
  o.has_done = take((size_t)G * 4);
  o.best_score = take((size_t)G * 4);
  o.best_t = take((size_t)G * 4);
  o.best_parent = take((size_t)G * 4);
  o.bp_parent = take((size_t)n_steps * R * 4);
  o.bp_word = take((size_t)n_steps * R * 4);
  o.total = cur;
  return CAPDEC_OK;
}

}  // namespace

size_t beam_workspace_bytes(const CapdecDims& d, int G, int k, int n_steps) {
  BeamPlan bp;
  // sized for the alpha history; a search without it needs less but never more
  if (make_beam_plan(d, G, k, n_steps, true, &bp) != CAPDEC_OK) return 0;
  return bp.o.total;
}

int beam_search(const CapdecDims& d, const CapdecParams& w, const float* enc, int64_t enc_sb, int64_t enc_sp,
                int64_t enc_se, const float* tags, int G, int k, int n_steps, int32_t start_id, int32_t end_id, int32_t* out_seq,
                int32_t* out_len, float* out_score, int32_t* out_completed, float* out_alpha,
                int32_t* trace_parent, int32_t* trace_word, float* trace_score, void* workspace,
                size_t ws_bytes, cudaStream_t st) {
  BeamPlan bp;
  const bool want_alpha = out_alpha != nullptr && d.kind != CAPDEC_PURE_SCN;
  CAPDEC_TRY(make_beam_plan(d, G, k, n_steps, want_alpha, &bp));
  CAPDEC_REQUIRE(ws_bytes >= bp.o.total, CAPDEC_ERR_WORKSPACE, "beam workspace %zu < %zu", ws_bytes,
                 bp.o.total);
  CAPDEC_REQUIRE(((uintptr_t)workspace % 256) == 0, CAPDEC_ERR_BAD_ARG, "workspace must be 256-B aligned");
  CAPDEC_REQUIRE(start_id >= 0 && start_id < d.V && end_id >= 0 && end_id < d.V, CAPDEC_ERR_BAD_ARG,
                 "start/end token outside the vocabulary");
  Ctx c;
  c.p = bp.w;
  c.ws = (uint8_t*)workspace;
  c.st = st;
  c.prec = d.precision;
  const Plan& p = c.p;
  const BeamPlan::Off& o = bp.o;
  const int pr = c.prec;
  const int R = bp.R, P = d.P, E = d.E, A = d.A, M = d.M, D = d.D, F = d.F, S = d.S, V = d.V, NQ = p.NQ,
            NG1 = p.NG1;
  CAPDEC_REQUIRE(!p.scn || tags, CAPDEC_ERR_BAD_ARG, "beam search: tags is NULL");

  CAPDEC_TRY(pack_weights(c, w));
  // ---- per-image prologue (attention_scn.py:176-214) ----
  const bool have_cm = p.att && pr == CAPDEC_BF16 && E % 512 == 0;
  // strided views accepted (the real encoder output is physically NCHW): the gather casts /
  // reorders in its one pass, no dense fp32 copy is made first
  CAPDEC_TRY(gather_features(pr, enc, enc_sb, enc_sp, enc_se, nullptr, c.at(o.enc_f), c.at<float>(o.mean),
                             c.at(o.meanF), p.ldE, G, P, E, st, have_cm ? c.at(o.enc_cm) : nullptr, 512));
  if (p.att)
    CAPDEC_TRY(G_(c, c.at(o.enc_f), E, c.at(p.o.Wp_e), p.ldE, c.at(o.att1), A, 1, w.enc_att_b, nullptr, 0,
                  G * P, A, E));

  CAPDEC_TRY(expand_rows(pr, c.at(o.meanF), p.ldE, c.at(o.meanX), p.ldE, G, k, E, st));
  CAPDEC_TRY(G_(c, c.at(o.meanX), p.ldE, c.ft(p.o.Wp_init, 0), p.ldE, c.at(o.H), p.ldD, 1, w.init_h_b,
                nullptr, 0, R, D, E));
  CAPDEC_TRY(G_(c, c.at(o.meanX), p.ldE, c.ft(p.o.Wp_init, (int64_t)D * p.ldE), p.ldE, c.at(o.C), D, 0,
                w.init_c_b, nullptr, 0, R, D, E));
  if (p.scn) {
    CAPDEC_TRY(copy_cast(pr, tags, 0, S, c.at(o.tagsG), 1, p.ldS, G, S, st));
    CAPDEC_TRY(expand_rows(pr, c.at(o.tagsG), p.ldS, c.at(o.tagsX), p.ldS, G, k, S, st));
    CAPDEC_TRY(G_(c, c.at(o.tagsX), p.ldS, c.at(p.o.Wp_ibT), p.ldS, c.at(o.v), NQ, 0, nullptr, nullptr, 0,
                  R, NQ, S));
    CAPDEC_TRY(G_(c, c.at(o.tagsX), p.ldS, c.at(p.o.Wp_hbT), p.ldS, c.at(o.q), NQ, 0, nullptr, nullptr, 0,
                  R, NQ, S));
  }
  int32_t* prev_word = c.at<int32_t>(o.prev_word);
  int32_t* src_row = c.at<int32_t>(o.src_row);
  int32_t* live = c.at<int32_t>(o.live);
  int32_t* krem = c.at<int32_t>(o.krem);
  int32_t* has_done = c.at<int32_t>(o.has_done);
  float* best_score = c.at<float>(o.best_score);
  int32_t* best_t = c.at<int32_t>(o.best_t);
  int32_t* best_parent = c.at<int32_t>(o.best_parent);
  int32_t* bpp = c.at<int32_t>(o.bp_parent);
  int32_t* bpw = c.at<int32_t>(o.bp_word);
  float* score_in = c.at<float>(o.scoreA);
  float* score_out = c.at<float>(o.scoreB);
  CAPDEC_TRY(beam_init(prev_word, score_in, live, krem, has_done, best_score, best_t, best_parent, G, k,
                       start_id, st));
  CAPDEC_CUDA_OK(cudaMemsetAsync(score_out, 0, (size_t)R * 4, st));
  CAPDEC_CUDA_OK(cudaMemsetAsync(bpp, 0, (size_t)n_steps * R * 4, st));
  CAPDEC_CUDA_OK(cudaMemsetAsync(bpw, 0, (size_t)n_steps * R * 4, st));

  // bf16 mode: single-pass selection (CAPDEC_BEAM_EXACT=1 keeps the three-pass arithmetic of the parity mode)
  const char* exact_env = getenv("CAPDEC_BEAM_EXACT");
  const bool fast_select = pr == CAPDEC_BF16 && !(exact_env && exact_env[0] == '1');
  // ... and the vocabulary projection, the log-softmax statistics and the per-tile top-k candidates are ONE kernel
  // (gemm_tc_vocab_topk): the (rows x V) logits are never written (CAPDEC_BEAM_FUSED=0: separate GEMM + selection)
  const char* fused_env = getenv("CAPDEC_BEAM_FUSED");
  const bool fused_vocab = fast_select && !(fused_env && fused_env[0] == '0');
  const int kl = k <= 4 ? 4 : 8;
  const VocabTopkPlan vplan = vocab_topk_plan(R, V);
  // ---- the search loop (attention_scn.py:216-290) ----
  // attention decoders: the embedding and the gated context are written side by side ([emb_t | z], the cell's
  // input as the reference concatenates it, attention_scn.py:232), so the input-side product is ONE GEMM over
  // K = M + E instead of a K = M GEMM plus a read-modify-write K = E GEMM over the fp32 U
  const bool cat_xz = p.att && (M * (int)p.fsz) % 16 == 0;
  const int64_t ldXZ = (int64_t)M + E;
  for (int t = 0; t < n_steps; ++t) {
    float* g1 = c.at<float>(o.g1);
    float* U = c.at<float>(o.U);
    if (cat_xz) {
      CAPDEC_TRY(beam_embed(pr, w.emb, prev_word, c.at(o.xz), ldXZ, R, M, V, st));
    } else {
      CAPDEC_TRY(beam_embed(pr, w.emb, prev_word, c.at(o.Xe), p.ldM, R, M, V, st));
      CAPDEC_TRY(G_(c, c.at(o.Xe), p.ldM, c.at(p.o.Wp_xq), p.ldX, U, NQ, 0, nullptr, nullptr, 0, R, NQ, M));
    }
    CAPDEC_TRY(G_(c, c.at(o.H), p.ldD, c.at(p.o.Wp_cat1), p.ldD, g1, NG1, 0, c.at<float>(p.o.b_cat1),
                  nullptr, 0, R, NG1, D));
    const float* pcol = g1 + (p.att ? A + E : 0);
    if (p.att) {
      float* alpha_t = c.at<float>(o.alpha_hist) + (want_alpha ? (int64_t)t * R * P : 0);
      CAPDEC_TRY(attention_fwd(pr, c.at(o.att1), c.at(o.enc_f), g1, NG1, A, w.full_att_w, w.full_att_b,
                               alpha_t, P, cat_xz ? c.ft(o.xz, M) : c.at(o.z), cat_xz ? ldXZ : (int64_t)E, nullptr, R,
                               k, P, E, A, c.at<float>(o.att_scr), st, have_cm ? c.at(o.enc_cm) : nullptr));
      if (cat_xz)
        CAPDEC_TRY(G_(c, c.at(o.xz), ldXZ, c.at(p.o.Wp_xq), p.ldX, U, NQ, 0, nullptr, nullptr, 0, R, NQ, M + E));
      else
        CAPDEC_TRY(G_(c, c.at(o.z), E, c.ft(p.o.Wp_xq, M), p.ldX, U, NQ, 0, nullptr, U, NQ, R, NQ, E));
    }
    if (p.scn) {
      CAPDEC_TRY(scn_form_m(pr, U, NQ, pcol, NG1, c.at<float>(o.v), c.at<float>(o.q), c.at(o.m), R, R, F, st));
      CAPDEC_TRY(G_(c, c.at(o.m), 2 * F, c.at(p.o.Wp_c), p.ld2F, c.at(o.pre), 4 * D, 0, nullptr, nullptr, 0,
                    R, D, 2 * F, R, 4, (int64_t)R * 2 * F, (int64_t)D * p.ld2F, D));
      CAPDEC_TRY(cell_fwd(pr, c.at<float>(o.pre), 4 * D, nullptr, 0, w.b_ih, w.b_hh, 0, c.at<float>(o.C),
                          c.at<float>(o.Cn), nullptr, c.at(o.Hn), p.ldD, nullptr, 0.f, nullptr, 0, 1, R, D, st));
    } else {
      CAPDEC_TRY(cell_fwd(pr, U, NQ, pcol, NG1, w.b_ih, w.b_hh, 1, c.at<float>(o.C), c.at<float>(o.Cn),
                          nullptr, c.at(o.Hn), p.ldD, nullptr, 0.f, nullptr, 0, 1, R, D, st));
    }
    // scores = log_softmax(fc(h)) (:235-236; eval mode: dropout is the identity)
    if (fused_vocab) {
      CAPDEC_TRY(gemm_tc_vocab_topk(c.at(o.Hn), p.ldD, R, c.at(p.o.Wp_fc), p.ldD, V, D, w.fc_b, c.at<float>(o.vpart),
                                    kl, st));
      CAPDEC_TRY(beam_select(c.at<float>(o.vpart), V, G, k, t, end_id, score_in, score_out, prev_word,
                             src_row, live, krem, has_done, best_score, best_t, best_parent, bpp, bpw,
                             trace_parent, trace_word, trace_score, n_steps, st, 1, &vplan));
    } else {
      CAPDEC_TRY(G_(c, c.at(o.Hn), p.ldD, c.at(p.o.Wp_fc), p.ldD, c.at(o.logits), V, 0, w.fc_b, nullptr, 0,
                    R, V, D));
      CAPDEC_TRY(beam_select(c.at<float>(o.logits), V, G, k, t, end_id, score_in, score_out, prev_word,
                             src_row, live, krem, has_done, best_score, best_t, best_parent, bpp, bpw,
                             trace_parent, trace_word, trace_score, n_steps, st, fast_select ? 1 : 0));
    }
    CAPDEC_TRY(beam_gather_state(pr, c.at(o.Hn), c.at<float>(o.Cn), c.at(o.H), c.at<float>(o.C), src_row,
                                 live, R, k, D, p.ldD, st));
    float* tmp = score_in; score_in = score_out; score_out = tmp;
  }
  CAPDEC_TRY(beam_finalize(G, k, n_steps, P, start_id, end_id, score_in, live, has_done, best_score, best_t,
                           best_parent, bpp, bpw, want_alpha ? c.at<float>(o.alpha_hist) : nullptr, out_seq,
                           out_len, out_score, out_completed, out_alpha, st));
  return CAPDEC_OK;
}

}  // namespace capdec
