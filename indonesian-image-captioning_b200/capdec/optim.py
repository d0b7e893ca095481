"""Fused gradient clip + Adam for the decoder parameters.

Reference: `trains/attention_scn.py:91-92` builds `torch.optim.Adam(decoder.parameters(), lr)`; every
iteration calls `clip_gradient(decoder_optimizer, grad_clip)` (`utils/optimizer.py:1-11`: in-place clamp
of each `.grad` to [-clip, clip]) and then `decoder_optimizer.step()` (`:244-252`).  `ClipAdam` is a
`torch.optim.Optimizer` with torch.optim.Adam's hyper-parameters, `param_groups` and per-parameter state
(`step`, `exp_avg`, `exp_avg_sq` -- state_dicts move both ways, `utils.optimizer.adjust_learning_rate` and
the checkpoint pickling of the optimizer object work unchanged), whose `step()` is ONE kernel launch over
all parameters (`capdec_clip_adam_step`): clamp + moment update + parameter update read p, g, m, v once.

    opt = capdec.optim.ClipAdam(filter(lambda p: p.requires_grad, decoder.parameters()), lr=4e-4, grad_clip=5.)
    loss.backward(); opt.step()        # no separate clip_gradient call needed (calling it anyway is harmless)
"""
import ctypes as C

import torch

from . import _lib


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clip=None,
                 write_clipped=True):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, grad_clip=grad_clip,
                        write_clipped=bool(write_clipped))
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = None
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if not (p.is_cuda and g.is_cuda):
                    raise _lib.CapdecError("capdec.optim.ClipAdam needs CUDA parameters; there is no CPU path")
                if p.dtype != torch.float32 or g.dtype != torch.float32 or g.is_sparse:
                    raise _lib.CapdecError("ClipAdam: parameters and gradients must be dense float32")
                if not p.is_contiguous():
                    raise _lib.CapdecError("ClipAdam: parameters must be contiguous")
                if not g.is_contiguous():
                    g = g.contiguous()
                    p.grad = g
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)          # as torch.optim.Adam (host scalar)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append((p, g, st))
            for step, items in by_step.items():
                if lib is None:
                    lib = _lib.load()
                segs = (_lib.AdamSeg * len(items))()
                for i, (p, g, st) in enumerate(items):
                    segs[i] = _lib.AdamSeg(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                           st["exp_avg_sq"].data_ptr(), p.numel())
                dev = items[0][0].device
                with torch.cuda.device(dev):
                    rc = lib.capdec_clip_adam_step(
                        segs, len(items), float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                        float(group["eps"]), float(group["weight_decay"]),
                        float(group["grad_clip"]) if group.get("grad_clip") else 0.0, step,
                        1 if group.get("write_clipped", True) else 0,
                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
                _lib.check(rc, "capdec_clip_adam_step")
        return loss
