"""Fused clip + Adam (capdec.optim.ClipAdam -> capdec_clip_adam_step) against the reference's
`clip_gradient` + `torch.optim.Adam.step()` (utils/optimizer.py:1-11, trains/attention_scn.py:244-252)."""
import copy

import pytest
import torch

from capdec.optim import ClipAdam

pytestmark = pytest.mark.gpu

SHAPES = [(2560, 2048), (1000, 2048), (512,), (1,), (10000, 512), (7, 3), (2048,), (3, 5, 7)]


def _params(seed, n_extra=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = SHAPES + [(17 + i,) for i in range(n_extra)]
    return [torch.nn.Parameter(torch.randn(*s, device="cuda", generator=g)) for s in shapes]


def _set_grads(params, seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for p in params:
        p.grad = torch.randn(p.shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("clip", [5.0, None])
@pytest.mark.parametrize("n_extra", [0, 60])       # 68 tensors: more than one segment table
def test_clip_adam_matches_torch(clip, n_extra):
    ours = _params(0, n_extra)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o1 = ClipAdam(ours, lr=4e-4, grad_clip=clip)
    o2 = torch.optim.Adam(ref, lr=4e-4)
    for it in range(4):
        _set_grads(ours, 10 + it, 20.0)       # many elements beyond the clip value
        for a, b in zip(ours, ref):
            b.grad = a.grad.clone()
        o1.step()
        if clip is not None:
            for b in ref:
                b.grad.data.clamp_(-clip, clip)          # utils/optimizer.py:10
        o2.step()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (it, a.shape, (a - b).abs().max().item())
            assert torch.allclose(o1.state[a]["exp_avg"], o2.state[b]["exp_avg"], rtol=1e-5, atol=1e-8)
            assert torch.allclose(o1.state[a]["exp_avg_sq"], o2.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-10)
            if clip is not None:
                assert torch.equal(a.grad, b.grad)       # the clamped gradient is left in .grad, as upstream


def test_state_dict_moves_both_ways_and_lr_decay():
    ours = _params(1)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o2 = torch.optim.Adam(ref, lr=1e-3)
    _set_grads(ref, 3, 1.0)
    o2.step()
    o1 = ClipAdam(ours, lr=1e-3, grad_clip=5.0)
    with torch.no_grad():
        for a, b in zip(ours, ref):
            a.copy_(b)
    o1.load_state_dict(copy.deepcopy(o2.state_dict()))       # a torch.optim.Adam checkpoint continues here
    for group in o1.param_groups:                             # utils/optimizer.py adjust_learning_rate
        group["lr"] = group["lr"] * 0.8
    for group in o2.param_groups:
        group["lr"] = group["lr"] * 0.8
    _set_grads(ours, 4, 1.0)
    for a, b in zip(ours, ref):
        b.grad = a.grad.clone()
    o1.step()
    o2.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    sd = o1.state_dict()
    assert sd["param_groups"][0]["lr"] == pytest.approx(8e-4)
    assert float(sd["state"][0]["step"]) == 2.0
