"""Helpers shared by the GPU parity tests."""
import torch

from oracle import capdec_oracle as O


def build_decoder(kind, dims, device="cuda"):
    """Instantiate the drop-in module for a golden/oracle dims dict (keys A,M,D,F,S,V,E)."""
    if kind == O.ATTENTION_SCN:
        from models.decoders.attention_scn import AttentionSCN
        m = AttentionSCN(dims["A"], dims["M"], dims["D"], dims["F"], dims["S"], dims["V"],
                         encoder_dim=dims["E"], dropout=0.5)
    elif kind == O.PURE_SCN:
        from models.decoders.pure_scn import PureSCN
        m = PureSCN(dims["M"], dims["D"], dims["F"], dims["S"], dims["V"], encoder_dim=dims["E"],
                    dropout=0.5)
    else:
        from models.decoders.pure_attention import PureAttention
        m = PureAttention(dims["A"], dims["M"], dims["D"], dims["V"], encoder_dim=dims["E"], dropout=0.5)
    return m.to(device)


def call_forward(dec, kind, enc, tags, caps, caplens):
    if kind == O.PURE_ATTENTION:
        out = dec(enc, caps, caplens)
    else:
        out = dec(enc, tags, caps, caplens)
    if kind == O.PURE_SCN:
        scores, caps_sorted, dl, sort_ind = out
        return scores, caps_sorted, dl, None, sort_ind
    return out


def torch_loss_glue(scores, caps_sorted, decode_lengths, alphas, alpha_c=1.0):
    """The reference's loss glue (trains/attention_scn.py:219-235) with stock torch ops,
    used to drive the generic autograd contract of the drop-in modules."""
    from torch.nn.utils.rnn import pack_padded_sequence
    targets = caps_sorted[:, 1:]
    s = pack_padded_sequence(scores, decode_lengths, batch_first=True).data
    t = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    loss = torch.nn.functional.cross_entropy(s, t)
    if alphas is not None:
        loss = loss + alpha_c * ((1. - alphas.sum(dim=1)) ** 2).mean()
    return loss


def rel_err(got, ref):
    ref = ref.detach().double().cpu()
    got = got.detach().double().cpu()
    denom = max(ref.abs().max().item(), 1e-30)
    return (got - ref).abs().max().item() / denom


def rel_err_fro(got, ref):
    """Frobenius-norm relative error: robust to single-element jumps (relu-kink flips)."""
    ref = ref.detach().double().cpu()
    got = got.detach().double().cpu()
    return (got - ref).norm().item() / max(ref.norm().item(), 1e-30)


def oracle_run(kind, sd, enc, tags, caps, caplens, sort_ind=None, dtype=torch.float64, alpha_c=1.0,
               dropout_masks=None, hoist=False):
    """Forward + loss + backward through the oracle (CPU, fp64 by default).  `dropout_masks`: (B,T,D) keep
    factors in SORTED row order (None = eval mode); `hoist`: att1 computed once (same arithmetic, see
    tests/test_oracle_golden.py::test_hoisted_att1_is_the_same_arithmetic)."""
    p = {k: v.detach().cpu().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    t = None if kind == O.PURE_ATTENTION else tags.cpu().to(dtype)
    out = O.decoder_forward(kind, p, enc.cpu().to(dtype), t, caps.cpu(), caplens.cpu(),
                            sort_ind=None if sort_ind is None else sort_ind.cpu(),
                            dropout_masks=None if dropout_masks is None else dropout_masks.cpu().to(dtype),
                            hoist=hoist)
    if kind == O.PURE_SCN:
        scores, caps_sorted, dl, si = out
        alphas = None
    else:
        scores, caps_sorted, dl, alphas, si = out
    loss = O.caption_loss(scores, caps_sorted, dl, alphas, alpha_c)
    loss.backward()
    return {"scores": scores.detach(), "alphas": None if alphas is None else alphas.detach(),
            "loss": loss.detach(), "grads": {k: v.grad for k, v in p.items()},
            "decode_lengths": dl, "sort_ind": si, "caps_sorted": caps_sorted}
