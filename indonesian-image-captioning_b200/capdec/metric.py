"""Top-k accuracy on the device.

Reference: `utils/metric.py:25-39` `accuracy(scores, targets, k)` -- `scores.topk(k)` over the packed logits
and a `.item()` every iteration (`trains/attention_scn.py:255, 338`).  Here one pass over the logits counts the
rows whose target ranks among the k largest; `accuracy` keeps the reference's signature and return value,
`topk_hits_unpacked` works directly on the decoder's (B,T,V) output and leaves the count on the device."""
import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def topk_hits(scores, targets, k):
    """Device int32 tensor (1,) = #rows of `scores` (N,V) whose `targets` (N,) entry is in the top k."""
    if not (scores.is_cuda and targets.is_cuda):
        raise _lib.CapdecError("capdec.metric needs CUDA tensors; there is no CPU path")
    scores = scores.detach()
    if scores.dtype != torch.float32 or scores.stride(-1) != 1:
        scores = scores.float().contiguous()
    targets = targets.detach().to(torch.int64).contiguous().view(-1)
    N, V = scores.shape
    hits = torch.empty(1, dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().capdec_topk_hits(_lib.ptr(scores), scores.stride(0), _lib.ptr(targets), None, None, N, 1, 0,
                                          V, int(k), _lib.ptr(hits), _stream())
    _lib.check(rc, "capdec_topk_hits")
    return hits


def accuracy(scores, targets, k):
    """Same contract as utils/metric.py:25-39: top-k accuracy in percent as a python float (host sync)."""
    batch_size = targets.size(0)
    return topk_hits(scores, targets, k).item() * (100.0 / batch_size)


def topk_hits_unpacked(scores, caps_sorted, decode_lengths, k):
    """Device int32 tensor (1,): hits over the rows (b,t), t < decode_lengths[b], of the decoder's (B,T,V) output
    against caps_sorted[b, t+1] -- what `accuracy(pack(scores), pack(targets), k)` counts, without the packing."""
    if not (scores.is_cuda and caps_sorted.is_cuda):
        raise _lib.CapdecError("capdec.metric needs CUDA tensors; there is no CPU path")
    scores = scores.detach()
    if scores.dtype != torch.float32 or not scores.is_contiguous():
        scores = scores.float().contiguous()
    caps = caps_sorted.to(torch.int64).contiguous()
    B, T, V = scores.shape
    meta = getattr(scores, "_capdec_meta", None)
    plan = meta.get("plan") if isinstance(meta, dict) else None
    if plan is not None and plan.len_list == [int(x) for x in decode_lengths]:
        len_d = plan.len_d
    else:
        len_d = torch.tensor([int(x) for x in decode_lengths], dtype=torch.int32).pin_memory().to(scores.device,
                                                                                                  non_blocking=True)
    hits = torch.empty(1, dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = _lib.load().capdec_topk_hits(_lib.ptr(scores), V, None, _lib.ptr(caps), _lib.ptr(len_d), B * T, T,
                                          caps.shape[1], V, int(k), _lib.ptr(hits), _stream())
    _lib.check(rc, "capdec_topk_hits")
    return hits
