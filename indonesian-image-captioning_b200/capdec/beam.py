"""Batched beam search through capdec_beam_search (reference `sample`:
attention_scn.py:160-296, pure_scn.py:142-249, pure_attention.py:153-281).

G independent searches run in ONE C call; the loop, the log-softmax + top-k selection, the beam
re-ordering and the back-tracking of the winning caption all stay on the device, and nothing is
copied to the host until the caller reads the result tensors.
"""
import ctypes as C

import torch

from . import _lib
from . import functional as CF
from .config import get_precision


def beam_search_batch(dec, beam_size, start_id, end_id, encoder_out, tag_out=None, max_steps=50,
                      want_alphas=True, want_trace=False, precision=None):
    """Returns a dict of device tensors:
      seq (G, n+1) int32 incl. <start> (zero padded), len (G), score (G) fp32,
      completed (G) int32 -- 0 where no beam emitted <end> (the reference raises ValueError there,
      App. C-4; the defined fallback is the best live beam), alpha (G, n+1, P) or None,
      trace = (parent, word, score) each (G, n, k) or None;  n = max_steps + 1 decode steps, the
      reference's `step > 50` rule (:288)."""
    lib = _lib.load()
    kind = dec.kind
    G_ = encoder_out.size(0)
    E = encoder_out.size(-1)
    enc = encoder_out.reshape(G_, -1, E)
    CF._require_cuda(enc, tag_out)
    enc = enc.detach()
    if enc.dtype != torch.float32:
        enc = enc.float()
    # strided views (the encoder's permuted NCHW output) go to the library as they are: its gather reorders and
    # casts in one pass (never .contiguous() in Python)
    P = enc.size(1)
    tags = None
    if kind != "pure_attention":
        if tag_out is None:
            raise ValueError("tag_out is required for %s" % kind)
        tags = tag_out.detach().float().contiguous()
        if tags.size(0) != G_:
            raise RuntimeError("tag_out batch size %d doesn't match encoder_out batch size %d"
                               % (tags.size(0), G_))
    kw = dec._dims_kw()
    n_steps = int(max_steps) + 1
    k = int(beam_size)
    dims = CF.make_dims(kind, precision or get_precision(), 1, 1, P, E, kw.get("A", 0), kw["M"], kw["D"],
                        kw.get("F", 0), kw.get("S", 0), kw["V"], 2)
    params = [p.detach() for p in dec._param_list()]
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
            raise _lib.CapdecError("decoder parameters must be contiguous float32 CUDA tensors")
    pstruct = CF._params_struct(kind, params)
    dev = enc.device
    ws_bytes = lib.capdec_beam_workspace_bytes(C.byref(dims), G_, k, n_steps)
    if ws_bytes == 0:
        _lib.check(-1, "capdec_beam_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    seq = torch.empty(G_, n_steps + 1, dtype=torch.int32, device=dev)
    length = torch.empty(G_, dtype=torch.int32, device=dev)
    score = torch.empty(G_, dtype=torch.float32, device=dev)
    completed = torch.empty(G_, dtype=torch.int32, device=dev)
    alpha = None
    if want_alphas and kind != "pure_scn":
        alpha = torch.empty(G_, n_steps + 1, P, dtype=torch.float32, device=dev)
    tr = (None, None, None)
    if want_trace:
        tr = (torch.empty(G_, n_steps, k, dtype=torch.int32, device=dev),
              torch.empty(G_, n_steps, k, dtype=torch.int32, device=dev),
              torch.empty(G_, n_steps, k, dtype=torch.float32, device=dev))
    with torch.cuda.device(dev):
        rc = lib.capdec_beam_search_strided(C.byref(dims), C.byref(pstruct), _lib.ptr(enc), enc.stride(0),
                                            enc.stride(1), enc.stride(2), _lib.ptr(tags), G_, k, n_steps,
                                            int(start_id), int(end_id), _lib.ptr(seq), _lib.ptr(length),
                                            _lib.ptr(score), _lib.ptr(completed), _lib.ptr(alpha), _lib.ptr(tr[0]),
                                            _lib.ptr(tr[1]), _lib.ptr(tr[2]), _lib.ptr(ws), ws_bytes, CF._stream())
    _lib.check(rc, "capdec_beam_search")
    return {"seq": seq, "len": length, "score": score, "completed": completed, "alpha": alpha,
            "trace": tr if want_trace else None}
