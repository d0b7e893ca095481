"""Autograd wrappers around the C ABI (teacher-forced decoder, fused loss, unit ops).

All tensors must live on a CUDA device; nothing here computes on the CPU.
"""
import ctypes as C

import torch

from . import _lib
from .config import get_precision, precision_code

# state_dict name -> CapdecParams field, per decoder kind
_COMMON_TAIL = [("init_h.weight", "init_h_w"), ("init_h.bias", "init_h_b"),
                ("init_c.weight", "init_c_w"), ("init_c.bias", "init_c_b")]
_ATT = [("attention.encoder_att.weight", "enc_att_w"), ("attention.encoder_att.bias", "enc_att_b"),
        ("attention.decoder_att.weight", "dec_att_w"), ("attention.decoder_att.bias", "dec_att_b"),
        ("attention.full_att.weight", "full_att_w"), ("attention.full_att.bias", "full_att_b")]
_SCN = [("decode_step.weight_ia", "w_ia"), ("decode_step.weight_ib", "w_ib"),
        ("decode_step.weight_ic", "w_ic"), ("decode_step.weight_ha", "w_ha"),
        ("decode_step.weight_hb", "w_hb"), ("decode_step.weight_hc", "w_hc"),
        ("decode_step.bias_ih", "b_ih"), ("decode_step.bias_hh", "b_hh")]
_LSTM = [("decode_step.weight_ih", "w_ia"), ("decode_step.weight_hh", "w_ha"),
         ("decode_step.bias_ih", "b_ih"), ("decode_step.bias_hh", "b_hh")]
_BETA = [("f_beta.weight", "f_beta_w"), ("f_beta.bias", "f_beta_b")]
_FC = [("fc.weight", "fc_w"), ("fc.bias", "fc_b")]
PARAM_MAP = {
    "attention_scn": _ATT + [("embedding.weight", "emb")] + _SCN + _COMMON_TAIL + _BETA + _FC,
    "pure_scn": [("embedding.weight", "emb")] + _SCN + _COMMON_TAIL + _FC,
    "pure_attention": _ATT + [("embedding.weight", "emb")] + _LSTM + _COMMON_TAIL + _BETA + _FC,
}


def param_names(kind):
    return [n for n, _ in PARAM_MAP[kind]]


def _params_struct(kind, tensors):
    s = _lib.Params()
    for (name, field), t in zip(PARAM_MAP[kind], tensors):
        setattr(s, field, None if t is None else t.data_ptr())
    return s


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CapdecError("capdec ops need CUDA tensors (got a %s tensor); there is no CPU path"
                                   % t.device)


def make_dims(kind, precision, B, T, P, E, A, M, D, F, S, V, L):
    return _lib.Dims(_lib.KIND[kind], precision_code(precision), B, T, P, E, A, M, D, F, S, V, L)


class DecoderTrainFn(torch.autograd.Function):
    """predictions, alphas = decoder(enc, tags, caps_sorted, sort_ind; params) -- one C call each way."""

    @staticmethod
    def forward(ctx, meta, enc, tags, caps_sorted, sort_ind, *params):
        lib = _lib.load()
        kind = meta["kind"]
        _require_cuda(enc, tags, caps_sorted, sort_ind, *params)
        dims = meta["dims"]
        B, T, P, V = dims.B, dims.T, dims.P, dims.V
        dev = enc.device
        params = [p.detach().contiguous() for p in params]
        for p in params:
            if p.dtype != torch.float32:
                raise _lib.CapdecError("decoder parameters must be float32 master weights")
        need_bwd = meta["need_bwd"]
        ws_bytes = lib.capdec_workspace_bytes(C.byref(dims), 1 if need_bwd else 0)
        if ws_bytes == 0:
            _lib.check(-1, "capdec_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        predictions = torch.empty(B, T, V, dtype=torch.float32, device=dev)
        alphas = None if kind == "pure_scn" else torch.empty(B, T, P, dtype=torch.float32, device=dev)
        len_h = (C.c_int32 * B)(*meta["decode_lengths"])
        pstruct = _params_struct(kind, params)
        sb, sp, se = enc.stride()
        with torch.cuda.device(dev):
            rc = lib.capdec_forward_train(
                C.byref(dims), C.byref(pstruct), _lib.ptr(enc), sb, sp, se, _lib.ptr(sort_ind),
                _lib.ptr(tags), _lib.ptr(caps_sorted), len_h, meta["dropout_p"], meta["seed"],
                1 if need_bwd else 0, _lib.ptr(predictions), _lib.ptr(alphas), _lib.ptr(ws), ws_bytes,
                _stream())
        _lib.check(rc, "capdec_forward_train")
        ctx.meta = meta
        ctx.len_h = len_h
        ctx.ws = ws if need_bwd else None
        ctx.param_tensors = params
        ctx.save_for_backward(tags if tags is not None else torch.empty(0, device=dev), caps_sorted,
                              alphas if alphas is not None else torch.empty(0, device=dev))
        if alphas is None:
            return predictions
        return predictions, alphas

    @staticmethod
    def backward(ctx, d_pred, d_alphas=None):
        lib = _lib.load()
        meta = ctx.meta
        kind = meta["kind"]
        dims = meta["dims"]
        if ctx.ws is None:
            raise _lib.CapdecError("backward called but forward ran without save_for_backward")
        tags, caps_sorted, alphas = ctx.saved_tensors
        dev = caps_sorted.device
        params = ctx.param_tensors
        # one flat buffer, per-parameter views: the data-parallel helper all-reduces it in place
        flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for p in params:
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        meta["flat_grads"] = flat
        fused = meta.get("fused_dlogits")      # set by FusedLossFn: gradient already in feature type
        d_logits_ft = None
        if fused is not None and fused.get("buf") is not None:
            d_logits_ft = fused["buf"]
            d_pred_c = None
            d_alphas = fused.get("d_alphas")
        else:
            d_pred_c = torch.zeros(dims.B, dims.T, dims.V, device=dev) if d_pred is None \
                else d_pred.contiguous().float()
        d_alphas_c = None if (d_alphas is None or kind == "pure_scn") else d_alphas.contiguous().float()
        pstruct = _params_struct(kind, params)
        gstruct = _params_struct(kind, grads)
        with torch.cuda.device(dev):
            rc = lib.capdec_backward(
                C.byref(dims), C.byref(pstruct), _lib.ptr(tags if tags.numel() else None),
                _lib.ptr(caps_sorted), ctx.len_h, meta["dropout_p"], meta["seed"], _lib.ptr(d_pred_c),
                _lib.ptr(d_logits_ft), _lib.ptr(d_alphas_c), _lib.ptr(alphas if alphas.numel() else None),
                C.byref(gstruct), _lib.ptr(ctx.ws), ctx.ws.numel(), _stream())
        _lib.check(rc, "capdec_backward")
        ctx.ws = None
        return (None, None, None, None, None) + tuple(grads)


def decoder_forward(kind, module_params, enc, tags, caps_sorted, sort_ind, decode_lengths, *,
                    dims_kw, dropout_p=0.0, seed=0, precision=None):
    """Run the teacher-forced decoder.  `module_params`: tensors in PARAM_MAP[kind] order."""
    B, P, E = enc.shape
    T = max(decode_lengths)
    dims = make_dims(kind, precision or get_precision(), B, T, P, E, dims_kw.get("A", 0), dims_kw["M"],
                     dims_kw["D"], dims_kw.get("F", 0), dims_kw.get("S", 0), dims_kw["V"],
                     caps_sorted.shape[1])
    need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in module_params)
    meta = {"kind": kind, "dims": dims, "decode_lengths": [int(x) for x in decode_lengths],
            "dropout_p": float(dropout_p), "seed": int(seed) & ((1 << 63) - 1), "need_bwd": need_bwd}
    out = DecoderTrainFn.apply(meta, enc, tags, caps_sorted, sort_ind, *module_params)
    return out, meta


class FusedLossFn(torch.autograd.Function):
    """loss = packed CE(scores, caps_sorted[:,1:]) + alpha_c * mean((1 - sum_t alphas)^2)
    (trains/attention_scn.py:219-235) with both terms and their gradients computed by
    capdec_loss_fwd/_bwd.  When `meta` of the producing DecoderTrainFn is given, the logits
    gradient is handed to capdec_backward directly in the GEMM feature type (no fp32
    (B,T,V) gradient tensor is materialised)."""

    @staticmethod
    def forward(ctx, scores, alphas, caps_sorted, len_d, n_tokens, alpha_c, dims, meta):
        lib = _lib.load()
        _require_cuda(scores, alphas, caps_sorted, len_d)
        dev = scores.device
        B, T = dims.B, dims.T
        scores_c = scores.contiguous()
        alphas_c = None if alphas is None else alphas.contiguous()
        loss3 = torch.empty(3, dtype=torch.float32, device=dev)
        lse = torch.empty(2 * B * T + B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.capdec_loss_fwd(C.byref(dims), _lib.ptr(scores_c), _lib.ptr(alphas_c),
                                     _lib.ptr(caps_sorted), _lib.ptr(len_d), n_tokens, alpha_c,
                                     _lib.ptr(loss3), _lib.ptr(lse), _stream())
        _lib.check(rc, "capdec_loss_fwd")
        ctx.dims, ctx.meta, ctx.n_tokens, ctx.alpha_c = dims, meta, n_tokens, alpha_c
        ctx.has_alphas = alphas is not None
        ctx.save_for_backward(scores_c, alphas_c if alphas_c is not None else torch.empty(0, device=dev),
                              caps_sorted, len_d, lse)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, g_loss, _g3=None):
        lib = _lib.load()
        scores, alphas, caps_sorted, len_d, lse = ctx.saved_tensors
        dims, meta = ctx.dims, ctx.meta
        dev = scores.device
        alphas_p = alphas if ctx.has_alphas else None
        # upstream gradient stays on the device: no host sync in the training step
        g_dev = g_loss.detach().reshape(1).float().contiguous()
        d_alphas = torch.empty_like(alphas) if ctx.has_alphas else None
        if meta is not None and meta.get("need_bwd"):
            ldq = (dims.V + 7) // 8 * 8
            esz = 2 if dims.precision == 1 else 4
            buf = torch.empty(dims.B * dims.T * ldq * esz, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                rc = lib.capdec_loss_bwd(C.byref(dims), _lib.ptr(scores), _lib.ptr(alphas_p),
                                         _lib.ptr(caps_sorted), _lib.ptr(len_d), ctx.n_tokens,
                                         ctx.alpha_c, 1.0, _lib.ptr(g_dev), _lib.ptr(lse), None,
                                         _lib.ptr(buf), _lib.ptr(d_alphas), _stream())
            _lib.check(rc, "capdec_loss_bwd")
            meta["fused_dlogits"] = {"buf": buf, "d_alphas": d_alphas}
            # the decoder Function picks the buffers up from meta; autograd still needs
            # tensors of the right shape to route the call, so hand it cheap expanded zeros
            d_scores = torch.zeros((), device=dev).expand(scores.shape)
            d_al = None if not ctx.has_alphas else torch.zeros((), device=dev).expand(alphas.shape)
            return d_scores, d_al, None, None, None, None, None, None
        d_scores = torch.empty_like(scores)
        with torch.cuda.device(dev):
            rc = lib.capdec_loss_bwd(C.byref(dims), _lib.ptr(scores), _lib.ptr(alphas_p),
                                     _lib.ptr(caps_sorted), _lib.ptr(len_d), ctx.n_tokens, ctx.alpha_c,
                                     1.0, _lib.ptr(g_dev), _lib.ptr(lse), _lib.ptr(d_scores), None,
                                     _lib.ptr(d_alphas), _stream())
        _lib.check(rc, "capdec_loss_bwd")
        return d_scores, d_alphas, None, None, None, None, None, None


def caption_loss(scores, caps_sorted, decode_lengths, alphas=None, alpha_c=1.0, *, meta=None,
                 precision=None, n_tokens=None):
    """Fused replacement of the loss glue in trains/attention_scn.py:219-235.
    Returns (loss, (total, ce, reg) tensor).  `n_tokens` overrides the CE denominator
    (data-parallel training divides by the GLOBAL token count so that summing the ranks'
    gradients reproduces the single-process mean)."""
    B, T, V = scores.shape
    P = alphas.shape[2] if alphas is not None else 1
    if meta is not None:
        dims = meta["dims"]
    else:
        dims = make_dims("attention_scn" if alphas is not None else "pure_scn",
                         precision or get_precision(), B, T, P, 8, 8, 8, 8, 8, 8, V, caps_sorted.shape[1])
    len_d = torch.tensor(list(decode_lengths), dtype=torch.int32, device=scores.device)
    n_tokens = int(sum(decode_lengths)) if n_tokens is None else int(n_tokens)
    return FusedLossFn.apply(scores, alphas, caps_sorted, len_d, n_tokens, float(alpha_c), dims, meta)


# ---------------------------------------------------------------------------------
# unit ops (used by models/scn_cell.py, models/attention.py and the parity tests)
# ---------------------------------------------------------------------------------
def _ft_dtype(precision):
    return torch.bfloat16 if precision_code(precision) == 1 else torch.float32


def gemm(X, W, bias=None, addm=None, out_ft=False, precision=None, splitk=0, out=None):
    """out[r,n] = sum_k X[r,k] W[n,k] (+bias[n]) (+addm[r,n]) through the selected engine.
    X (rows,K) / W (N,K) must already be in the engine's operand type with 16-byte pitch."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(X, W, bias, addm)
    ft = _ft_dtype(prec)
    assert X.dtype == ft and W.dtype == ft and X.stride(-1) == 1 and W.stride(-1) == 1
    batch = X.shape[0] if X.dim() == 3 else 1
    rows, K = X.shape[-2], X.shape[-1]
    N = W.shape[-2]
    if out is None:
        make = torch.zeros if splitk else torch.empty      # split-K accumulates into the output
        out = make((batch, rows, N) if X.dim() == 3 else (rows, N),
                   dtype=ft if out_ft else torch.float32, device=X.device)
    sX = X.stride(0) if X.dim() == 3 else 0
    sW = W.stride(0) if W.dim() == 3 else 0
    sO = out.stride(0) if X.dim() == 3 else 0
    with torch.cuda.device(X.device):
        rc = lib.capdec_gemm(precision_code(prec), _lib.ptr(X), X.stride(-2), _lib.ptr(W), W.stride(-2),
                             _lib.ptr(out), out.stride(-2), 1 if out_ft else 0, _lib.ptr(bias),
                             _lib.ptr(addm), addm.stride(-2) if addm is not None else 0, rows, N, K,
                             batch, sX, sW, sO, splitk, _stream())
    _lib.check(rc, "capdec_gemm")
    return out


def attention_step(att1, enc, g1, beta_col, w_f, b_f, rows_per_map=1, precision=None, want_awe=True):
    """One soft-attention step on prepared features; returns (z, alpha, awe)."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(att1, enc, g1, w_f, b_f)
    G_, P, A = att1.shape
    E = enc.shape[2]
    rows = g1.shape[0]
    ft = _ft_dtype(prec)
    assert att1.dtype == ft and enc.dtype == ft and att1.is_contiguous() and enc.is_contiguous()
    alpha = torch.empty(rows, P, dtype=torch.float32, device=enc.device)
    z = torch.empty(rows, E, dtype=ft, device=enc.device)
    awe = torch.empty(rows, E, dtype=torch.float32, device=enc.device) if want_awe else None
    with torch.cuda.device(enc.device):
        rc = lib.capdec_attention_step(precision_code(prec), _lib.ptr(att1), _lib.ptr(enc), _lib.ptr(g1),
                                       g1.stride(0), beta_col, _lib.ptr(w_f), _lib.ptr(b_f),
                                       _lib.ptr(alpha), P, _lib.ptr(z), _lib.ptr(awe), rows, rows_per_map,
                                       P, E, A, _stream())
    _lib.check(rc, "capdec_attention_step")
    return z, alpha, awe


def scn_cell_step(weights, x, s, h, c, precision=None):
    """SCNCell.forward (models/scn_cell.py:52-154) through the C ABI.  weights = (w_ia, w_ib, w_ic,
    w_ha, w_hb, w_hc, b_ih, b_hh) fp32 CUDA tensors."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(x, s, h, c, *weights)
    rows, X = x.shape
    D = h.shape[1]
    F = weights[0].shape[1] // 4
    S = s.shape[1]
    pc = precision_code(prec)
    nbytes = lib.capdec_scn_cell_workspace_bytes(pc, rows, X, D, F, S)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
    h_out = torch.empty(rows, D, dtype=torch.float32, device=x.device)
    c_out = torch.empty(rows, D, dtype=torch.float32, device=x.device)
    args = [w.detach().contiguous().float() for w in weights] + \
           [t.detach().contiguous().float() for t in (x, s, h, c)]
    with torch.cuda.device(x.device):
        rc = lib.capdec_scn_cell_step(pc, rows, X, D, F, S, *[_lib.ptr(a) for a in args],
                                      _lib.ptr(h_out), _lib.ptr(c_out), _lib.ptr(ws), ws.numel(),
                                      _stream())
    _lib.check(rc, "capdec_scn_cell_step")
    return h_out, c_out
