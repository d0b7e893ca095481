"""CPU oracle for the caption-decoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (PyTorch CPU tensor ops, any float dtype)
of the reference decoder algorithm of rayandrew/indonesian-image-captioning.
It is the checker for the CUDA path, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it.  The product (`indonesian-image-captioning_b200/`) never
imports anything from `oracle/` and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c: "parity unpinned" by the reference itself).  The oracle is therefore
pinned against outputs of the *live reference modules* run in the build
container: `tests/golden/make_golden.py` imports `/root/reference`, runs
forward / loss / backward / beam search on seeded inputs and commits the
input+output vectors under `tests/golden/`; `tests/test_oracle_golden.py`
replays them through this file and demands bit-level (fp32 `==`) agreement for
the forward and <=1e-6 for gradients.

Each function cites the reference file:line it follows (paths relative to the
reference root).  Parameters are plain dicts keyed exactly like the reference
`state_dict()`, so a reference checkpoint can be fed in
unchanged.

The op *structure* deliberately follows the reference as written (att1 and the
tag projections recomputed at every step, twelve small matmuls per cell call,
autograd for the backward) so that timing this file is an honest port of the
reference's CPU cost (bench.py `cpu_baseline.kind == "port"`).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence

Params = Dict[str, torch.Tensor]

ATTENTION_SCN = "attention_scn"
PURE_SCN = "pure_scn"
PURE_ATTENTION = "pure_attention"
KINDS = (ATTENTION_SCN, PURE_SCN, PURE_ATTENTION)


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def gate_blocks(w: torch.Tensor, n: int) -> List[torch.Tensor]:
    """Four column blocks (2-D) / four chunks (1-D) in the cell's gate order
    i, f, o, c.  Follows utils/tensor.py:1-15 (1-D) and :18-42 (2-D)."""
    if w.dim() == 1:
        return [w[g * n:(g + 1) * n] for g in range(4)]
    return [w[:, g * n:(g + 1) * n] for g in range(4)]


def scn_cell(p: Params, prefix: str, x: torch.Tensor, s: torch.Tensor,
             h: torch.Tensor, c: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """One SCN-LSTM step.  Follows models/scn_cell.py:52-110 (input side) and
    :112-154 (recurrent side + LSTM pointwise).  Gate order i,f,o,c; two
    biases per gate; tag projections recomputed on every call as upstream."""
    F_ = p[prefix + "weight_ia"].shape[1] // 4
    D_ = p[prefix + "weight_ic"].shape[0]
    ia = gate_blocks(p[prefix + "weight_ia"], F_)
    ib = gate_blocks(p[prefix + "weight_ib"], F_)
    ic = gate_blocks(p[prefix + "weight_ic"], F_)
    ha = gate_blocks(p[prefix + "weight_ha"], F_)
    hb = gate_blocks(p[prefix + "weight_hb"], F_)
    hc = gate_blocks(p[prefix + "weight_hc"], F_)
    bi = gate_blocks(p[prefix + "bias_ih"], D_)
    bh = gate_blocks(p[prefix + "bias_hh"], D_)
    pre = []
    for g in range(4):
        # scn_cell.py:73-86  x-side factor product
        below = ((x @ ia[g]) * (s @ ib[g])) @ ic[g].t() + bi[g]
        # scn_cell.py:134-144  h-side factor product
        rec = (h @ ha[g]) * (s @ hb[g])
        pre.append(rec @ hc[g].t() + below + bh[g])
    # scn_cell.py:146-152
    i_g = torch.sigmoid(pre[0])
    f_g = torch.sigmoid(pre[1])
    o_g = torch.sigmoid(pre[2])
    g_g = torch.tanh(pre[3])
    c_new = f_g * c + i_g * g_g
    h_new = o_g * torch.tanh(c_new)
    return h_new, c_new


def lstm_cell(p: Params, prefix: str, x: torch.Tensor,
              h: torch.Tensor, c: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stock LSTM cell used by PureAttention (models/decoders/pure_attention.py:40-41,
    143-146 call torch.nn.LSTMCell).  Gate order i,f,g,o (torch convention)."""
    gates = x @ p[prefix + "weight_ih"].t() + p[prefix + "bias_ih"] \
        + h @ p[prefix + "weight_hh"].t() + p[prefix + "bias_hh"]
    D_ = h.shape[1]
    i_g = torch.sigmoid(gates[:, 0 * D_:1 * D_])
    f_g = torch.sigmoid(gates[:, 1 * D_:2 * D_])
    g_g = torch.tanh(gates[:, 2 * D_:3 * D_])
    o_g = torch.sigmoid(gates[:, 3 * D_:4 * D_])
    c_new = f_g * c + i_g * g_g
    h_new = o_g * torch.tanh(c_new)
    return h_new, c_new


def soft_attention(p: Params, enc: torch.Tensor, h: torch.Tensor,
                   att1: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Bahdanau soft attention.  Follows models/attention.py:26-44; att1 is
    recomputed here on every call exactly as upstream (attention.py:35).
    `att1` given (decoder_forward(hoist=True)): the time-invariant projection
    is passed in and the weighted sum runs as a batched matrix product instead
    of materialising enc * alpha -- the same sums, used by the full-size parity
    tests so that the checker finishes in seconds."""
    hoisted = att1 is not None
    if not hoisted:
        att1 = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    att2 = F.linear(h, p["attention.decoder_att.weight"], p["attention.decoder_att.bias"])
    e = F.linear(torch.relu(att1 + att2.unsqueeze(1)),
                 p["attention.full_att.weight"], p["attention.full_att.bias"]).squeeze(2)
    alpha = torch.softmax(e, dim=1)
    if hoisted:
        awe = torch.bmm(alpha.unsqueeze(1), enc).squeeze(1)
    else:
        awe = (enc * alpha.unsqueeze(2)).sum(dim=1)
    return awe, alpha


def init_hidden_state(p: Params, enc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/decoders/attention_scn.py:82-93 (identical in the other two)."""
    m = enc.mean(dim=1)
    return (F.linear(m, p["init_h.weight"], p["init_h.bias"]),
            F.linear(m, p["init_c.weight"], p["init_c.bias"]))


def _decode_step(kind: str, p: Params, enc: Optional[torch.Tensor], s: Optional[torch.Tensor],
                 emb_t: torch.Tensor, h: torch.Tensor, c: torch.Tensor,
                 att1: Optional[torch.Tensor] = None):
    """One decoder step shared by forward() and sample().
    attention_scn.py:144-153 / pure_scn.py:134-136 / pure_attention.py:136-146."""
    alpha = None
    if kind == PURE_SCN:
        x = emb_t
    else:
        awe, alpha = soft_attention(p, enc, h, att1)
        gate = torch.sigmoid(F.linear(h, p["f_beta.weight"], p["f_beta.bias"]))
        x = torch.cat([emb_t, gate * awe], dim=1)
    if kind == PURE_ATTENTION:
        h, c = lstm_cell(p, "decode_step.", x, h, c)
    else:
        h, c = scn_cell(p, "decode_step.", x, s, h, c)
    return h, c, alpha


# --------------------------------------------------------------------------
# teacher-forced forward
# --------------------------------------------------------------------------
def decoder_forward(kind: str, p: Params, encoder_out: torch.Tensor,
                    semantic_input: Optional[torch.Tensor],
                    encoded_captions: torch.Tensor, caption_lengths: torch.Tensor,
                    dropout_masks: Optional[torch.Tensor] = None,
                    sort_ind: Optional[torch.Tensor] = None, hoist: bool = False):
    """Teacher-forced unroll.  Follows attention_scn.py:95-158,
    pure_scn.py:87-140, pure_attention.py:90-151.

    * rows are sorted by caption length (descending); encoder_out and captions
      are permuted by sort_ind, the tag matrix is NOT (attention_scn.py:119-120
      vs :152) -- reproduced on purpose.
    * `dropout_masks` (B, T, D) of already-scaled keep factors stands in for
      nn.Dropout between h and fc (:154); None == eval mode.
    * `sort_ind` may be forced (the sort is unstable on ties, App. C-2).
    * `hoist`: compute the time-invariant att1 once instead of at every step
      (attention.py:35 inside the loop of attention_scn.py:144; App. C-6) -- same
      arithmetic per element, ~50x less CPU work at T=50; tests/test_oracle_golden.py
      checks it against the as-written structure.
    Returns the reference tuple; PureSCN's has no alphas (pure_scn.py:140).
    """
    assert kind in KINDS
    B = encoder_out.size(0)
    E = encoder_out.size(-1)
    enc = encoder_out.reshape(B, -1, E)
    P = enc.size(1)
    lens = caption_lengths.squeeze(1)
    if sort_ind is None:
        lens, sort_ind = lens.sort(dim=0, descending=True)
    else:
        lens = lens[sort_ind]
    enc = enc[sort_ind]
    caps = encoded_captions[sort_ind]
    emb = p["embedding.weight"][caps]                                   # :124
    h, c = init_hidden_state(p, enc)                                    # :127
    decode_lengths = (lens - 1).tolist()                                # :131
    T = max(decode_lengths)
    V = p["fc.weight"].shape[0]
    predictions = torch.zeros(B, T, V, dtype=enc.dtype)                 # :134-137
    alphas = torch.zeros(B, T, P, dtype=enc.dtype)
    att1_all = None
    if hoist and kind != PURE_SCN:
        att1_all = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    for t in range(T):                                                  # :142
        bt = sum(l > t for l in decode_lengths)
        s_t = None if semantic_input is None else semantic_input[:bt]
        h, c, alpha = _decode_step(kind, p, enc[:bt] if kind != PURE_SCN else None,
                                   s_t, emb[:bt, t, :], h[:bt], c[:bt],
                                   None if att1_all is None else att1_all[:bt])
        h_out = h if dropout_masks is None else h * dropout_masks[:bt, t, :]
        predictions[:bt, t, :] = F.linear(h_out, p["fc.weight"], p["fc.bias"])   # :154-155
        if alpha is not None:
            alphas[:bt, t, :] = alpha                                   # :156
    if kind == PURE_SCN:
        return predictions, caps, decode_lengths, sort_ind
    return predictions, caps, decode_lengths, alphas, sort_ind


def caption_loss(scores: torch.Tensor, caps_sorted: torch.Tensor,
                 decode_lengths: Sequence[int], alphas: Optional[torch.Tensor],
                 alpha_c: float = 1.0) -> torch.Tensor:
    """Loss glue of the training loop: packed (time-major) cross entropy, mean
    over N = sum(decode_lengths), plus the doubly-stochastic attention
    regulariser.  Follows trains/attention_scn.py:219-235 (pure_scn: :216-229,
    no regulariser)."""
    targets = caps_sorted[:, 1:]
    ps = pack_padded_sequence(scores, list(decode_lengths), batch_first=True).data
    pt = pack_padded_sequence(targets, list(decode_lengths), batch_first=True).data
    loss = F.cross_entropy(ps, pt)
    if alphas is not None:
        loss = loss + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    return loss


# --------------------------------------------------------------------------
# beam search
# --------------------------------------------------------------------------
def beam_search(kind: str, p: Params, encoder_out: torch.Tensor,
                tag_out: Optional[torch.Tensor], beam_size: int,
                start_id: int, end_id: int, max_steps: int = 50):
    """Beam search for ONE image.  Follows attention_scn.py:160-296,
    pure_scn.py:142-249, pure_attention.py:153-281 with the one-token
    restatement `//` for the parent index (:252; the upstream `/` floored on
    the torch version the code was written for).

    Returns a dict:
      seq, alphas      -- what upstream returns (alphas only for attention kinds)
      completed        -- True iff some beam emitted <end>; upstream raises
                          ValueError at :292 otherwise.  In that case seq/alphas
                          are the best LIVE beam (the defined fallback shared
                          with the CUDA path).
      trace            -- per step (parents, words, scores) as python lists
      score            -- cumulative log-prob of the returned sequence
    """
    k = beam_size
    V = p["fc.weight"].shape[0]
    E = encoder_out.size(-1)
    side = encoder_out.size(1)
    enc = encoder_out.reshape(1, -1, E)
    P = enc.size(1)
    enc = enc.expand(k, P, E)                                           # :189
    tags = None if tag_out is None else tag_out.expand(k, tag_out.size(1))
    prev_words = torch.full((k, 1), start_id, dtype=torch.long)         # :194-195
    seqs = prev_words
    top_scores = torch.zeros(k, 1, dtype=enc.dtype)
    with_alpha = kind != PURE_SCN
    seqs_alpha = torch.ones(k, 1, side, side, dtype=enc.dtype) if with_alpha else None  # :204
    done_seqs: List[List[int]] = []
    done_alpha: List = []
    done_scores: List[float] = []
    trace = []
    step = 1
    h, c = init_hidden_state(p, enc)                                    # :214
    while True:
        emb = p["embedding.weight"][prev_words].squeeze(1)              # :219
        h, c, alpha = _decode_step(kind, p, enc if with_alpha else None, tags, emb, h, c)
        scores = F.log_softmax(F.linear(h, p["fc.weight"], p["fc.bias"]), dim=1)  # :235-236
        scores = top_scores.expand_as(scores) + scores                  # :239
        if step == 1:
            top_scores, top_words = scores[0].topk(k, 0, True, True)    # :242-244
        else:
            top_scores, top_words = scores.view(-1).topk(k, 0, True, True)  # :248-249
        parents = top_words // V                                        # :252 (restated)
        words = top_words % V                                           # :253
        trace.append((parents.tolist(), words.tolist(), top_scores.tolist()))
        seqs = torch.cat([seqs[parents], words.unsqueeze(1)], dim=1)    # :256-257
        if with_alpha:
            a = alpha.view(-1, side, side)
            seqs_alpha = torch.cat([seqs_alpha[parents], a[parents].unsqueeze(1)], dim=1)
        live = [i for i, w in enumerate(words.tolist()) if w != end_id]  # :262-265
        dead = sorted(set(range(len(words))) - set(live))
        if dead:                                                        # :268-271
            done_seqs.extend(seqs[dead].tolist())
            if with_alpha:
                done_alpha.extend(seqs_alpha[dead].tolist())
            done_scores.extend(top_scores[dead].tolist())
        k -= len(dead)                                                  # :272
        if k == 0:
            break
        seqs = seqs[live]                                               # :278-285
        if with_alpha:
            seqs_alpha = seqs_alpha[live]
        h = h[parents[live]]
        c = c[parents[live]]
        enc = enc[parents[live]]
        if tags is not None:
            tags = tags[parents[live]]
        top_scores = top_scores[live].unsqueeze(1)
        prev_words = words[live].unsqueeze(1)
        if step > max_steps:                                            # :288-290
            break
        step += 1
    out = {"trace": trace, "completed": len(done_scores) > 0}
    if done_scores:
        i = done_scores.index(max(done_scores))                         # :292
        out["seq"] = done_seqs[i]
        out["alphas"] = done_alpha[i] if with_alpha else None
        out["score"] = done_scores[i]
    else:
        # upstream: ValueError("max() arg is an empty sequence").  Defined
        # fallback: best live beam (first max), documented in DESIGN.md.
        flat = top_scores.view(-1).tolist()
        i = flat.index(max(flat))
        out["seq"] = seqs[i].tolist()
        out["alphas"] = seqs_alpha[i].tolist() if with_alpha else None
        out["score"] = flat[i]
    return out


# --------------------------------------------------------------------------
# synthetic inputs / parameter construction
# --------------------------------------------------------------------------
def param_shapes(kind: str, *, attention_dim=512, embed_dim=512, decoder_dim=512,
                 factored_dim=512, semantic_dim=1000, vocab_size=10000,
                 encoder_dim=2048) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of the three decoders (reference
    constructors attention_scn.py:28-56, pure_scn.py:26-51, pure_attention.py:25-52)."""
    A, M, D, Fd, S, V, E = (attention_dim, embed_dim, decoder_dim, factored_dim,
                            semantic_dim, vocab_size, encoder_dim)
    sh: Dict[str, Tuple[int, ...]] = {}
    if kind != PURE_SCN:
        sh.update({"attention.encoder_att.weight": (A, E), "attention.encoder_att.bias": (A,),
                   "attention.decoder_att.weight": (A, D), "attention.decoder_att.bias": (A,),
                   "attention.full_att.weight": (1, A), "attention.full_att.bias": (1,)})
    sh["embedding.weight"] = (V, M)
    X = M if kind == PURE_SCN else M + E
    if kind == PURE_ATTENTION:
        sh.update({"decode_step.weight_ih": (4 * D, X), "decode_step.weight_hh": (4 * D, D),
                   "decode_step.bias_ih": (4 * D,), "decode_step.bias_hh": (4 * D,)})
    else:
        sh.update({"decode_step.weight_ia": (X, 4 * Fd), "decode_step.weight_ib": (S, 4 * Fd),
                   "decode_step.weight_ic": (D, 4 * Fd), "decode_step.weight_ha": (D, 4 * Fd),
                   "decode_step.weight_hb": (S, 4 * Fd), "decode_step.weight_hc": (D, 4 * Fd),
                   "decode_step.bias_ih": (4 * D,), "decode_step.bias_hh": (4 * D,)})
    sh.update({"init_h.weight": (D, E), "init_h.bias": (D,),
               "init_c.weight": (D, E), "init_c.bias": (D,)})
    if kind != PURE_SCN:
        sh.update({"f_beta.weight": (E, D), "f_beta.bias": (E,)})
    sh.update({"fc.weight": (V, D), "fc.bias": (V,)})
    return sh


def random_params(kind: str, seed: int = 0, dtype=torch.float32, **dims) -> Params:
    """Random-init parameters with the reference's init *distributions*
    (scn_cell.py:156-159 U(+-1/sqrt(D)); attention_scn.py:58-63 U(+-0.1), fc.bias 0;
    nn.Linear defaults U(+-1/sqrt(fan_in))).  Not bit-identical to constructing
    the reference module (different RNG consumption order) -- parity tests
    always load ONE state_dict into both sides."""
    g = torch.Generator().manual_seed(seed)
    sh = param_shapes(kind, **dims)
    D = sh["init_h.weight"][0]
    out: Params = {}
    for name, shape in sh.items():
        if name in ("embedding.weight", "fc.weight"):
            bound = 0.1
        elif name == "fc.bias":
            out[name] = torch.zeros(shape, dtype=dtype)
            continue
        elif name.startswith("decode_step."):
            bound = 1.0 / (D ** 0.5)
        else:
            fan_in = shape[1] if len(shape) == 2 else sh[name.replace(".bias", ".weight")][1]
            bound = 1.0 / (fan_in ** 0.5)
        out[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out


def synthetic_batch(B: int, V: int, *, seed: int = 1234, side: int = 14, E: int = 2048,
                    S: int = 1000, max_len: int = 52, lengths: Optional[Sequence[int]] = None,
                    dtype=torch.float32):
    """Synthetic inputs of SURVEY.md §8d: post-ReLU-like features, sigmoid-range
    tags, captions laid out like utils/dataset.py:302-306,388-392
    (<pad>=0, words 1.., <unk>=V-3, <start>=V-2, <end>=V-1)."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, side, side, E, generator=g, dtype=dtype).relu_()
    tags = torch.rand(B, S, generator=g, dtype=dtype)
    if lengths is None:
        lengths = [max_len - 1] * B          # throughput case: 51 -> T = 50
    caps = torch.zeros(B, max_len, dtype=torch.long)
    for b, L in enumerate(lengths):
        caps[b, 0] = V - 2
        if L > 2:
            caps[b, 1:L - 1] = torch.randint(1, V - 3, (L - 2,), generator=g)
        caps[b, L - 1] = V - 1
    caplens = torch.tensor(list(lengths), dtype=torch.long).unsqueeze(1)
    return enc, tags, caps, caplens


def tie_free_lengths(B: int, lo: int = 3, hi: int = 52, seed: int = 7) -> List[int]:
    """Distinct caption lengths (needs B <= hi-lo+1) so the unstable sort of
    attention_scn.py:117-118 is deterministic."""
    g = torch.Generator().manual_seed(seed)
    assert B <= hi - lo + 1
    return (torch.randperm(hi - lo + 1, generator=g)[:B] + lo).tolist()
