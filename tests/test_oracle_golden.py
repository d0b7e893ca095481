"""Pin the oracle against vectors produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import capdec_oracle as O
from conftest import load_golden

TRAIN_CASES = ["train_attention_scn_small", "train_pure_scn_small", "train_pure_attention_small",
               "train_attention_scn_medium", "train_pure_scn_medium", "train_pure_attention_medium",
               "train_attention_scn_hot"]


def _run(blob, dtype=torch.float32):
    kind = blob["kind"]
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in blob["state_dict"].items()}
    tags = None if kind == O.PURE_ATTENTION else blob["tags"].to(dtype)
    out = O.decoder_forward(kind, p, blob["encoder_out"].to(dtype), tags, blob["captions"],
                            blob["caption_lengths"], sort_ind=blob["sort_ind"])
    if kind == O.PURE_SCN:
        scores, caps_sorted, dl, sort_ind = out
        alphas = None
    else:
        scores, caps_sorted, dl, alphas, sort_ind = out
    loss = O.caption_loss(scores, caps_sorted, dl, alphas)
    loss.backward()
    return p, scores, caps_sorted, dl, alphas, sort_ind, loss


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_forward_loss_backward_match_reference(name):
    torch.set_num_threads(1)
    blob = load_golden(name)
    p, scores, caps_sorted, dl, alphas, sort_ind, loss = _run(blob)
    assert dl == blob["decode_lengths"]
    assert torch.equal(caps_sorted, blob["caps_sorted"])
    # forward: same ops in the same order on the same machine class -> tight
    ref = blob["predictions"]
    scale = ref.abs().max().item()
    assert (scores.detach() - ref).abs().max().item() <= 2e-6 * scale
    if alphas is not None:
        assert (alphas.detach() - blob["alphas"]).abs().max().item() <= 1e-6
    assert abs(loss.item() - blob["loss"].item()) <= 1e-6 * max(1.0, abs(blob["loss"].item()))
    for n, g in blob["grads"].items():
        got = p[n].grad
        assert got is not None, n
        tol = 1e-5 * max(g.abs().max().item(), 1e-6) + 1e-9
        assert (got - g).abs().max().item() <= tol, n


def test_sort_without_forced_index_matches_on_tie_free_lengths():
    blob = load_golden("train_attention_scn_small")
    p = blob["state_dict"]
    out = O.decoder_forward(blob["kind"], p, blob["encoder_out"], blob["tags"], blob["captions"],
                            blob["caption_lengths"])
    assert torch.equal(out[4], blob["sort_ind"])


def test_reference_invariants():
    """SURVEY.md §4 invariants 1,2,3,6,7."""
    blob = load_golden("train_attention_scn_small")
    p, scores, caps_sorted, dl, alphas, sort_ind, loss = _run(blob)
    for i, L in enumerate(dl):
        assert scores[i, L:].abs().max().item() == 0 if L < scores.shape[1] else True
        assert abs(alphas[i, :L].sum(-1) - 1).max().item() < 1e-5
    assert torch.equal(p["decode_step.bias_ih"].grad, p["decode_step.bias_hh"].grad)
    assert p["attention.full_att.bias"].grad.abs().item() < 1e-7
    # tags are NOT permuted by sort_ind (App. C-1)
    q = {k: v.detach() for k, v in p.items()}
    o2 = O.decoder_forward(blob["kind"], q, blob["encoder_out"], blob["tags"][sort_ind],
                           blob["captions"], blob["caption_lengths"])
    assert (o2[0] - scores.detach()).abs().max().item() > 1e-6


@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_beam_search_matches_reference(kind):
    torch.set_num_threads(1)
    blob = load_golden("beam_" + kind)
    p = blob["state_dict"]
    n_done = n_fail = 0
    for im in blob["images"]:
        for k, ref in im["results"].items():
            tags = None if kind == O.PURE_ATTENTION else im["tags"]
            with torch.no_grad():
                got = O.beam_search(kind, p, im["encoder_out"], tags, k,
                                    blob["start_id"], blob["end_id"])
            assert got["completed"] == ref["completed"]
            if ref["completed"]:
                n_done += 1
                assert got["seq"] == ref["seq"]
                if ref["alphas"] is not None:
                    a = torch.tensor(got["alphas"])
                    assert (a - ref["alphas"]).abs().max().item() <= 1e-6
            else:
                n_fail += 1
                assert len(got["trace"]) == 51      # App. C-5: up to 51 decode steps
                assert len(got["seq"]) == 52
    assert n_done > 0 and n_fail > 0


@pytest.mark.parametrize("name", ["train_attention_scn_medium", "train_pure_attention_medium", "train_attention_scn_hot"])
def test_hoisted_att1_is_the_same_arithmetic(name):
    """decoder_forward(hoist=True) (att1 computed once, batched weighted sum: what the full-size GPU parity
    tests use so that the checker finishes in seconds) against the as-written structure and the live
    reference's vectors, with and without an injected dropout mask."""
    blob = load_golden(name)
    kind = blob["kind"]
    tags = None if kind == O.PURE_ATTENTION else blob["tags"].double()
    B, T = blob["predictions"].shape[:2]
    D = blob["state_dict"]["init_h.weight"].shape[0]
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(B, T, D, generator=g) >= 0.5).double() * 2.0
    for masks in (None, mask):
        res = []
        for hoist in (False, True):
            p = {k: v.double().clone().requires_grad_(True) for k, v in blob["state_dict"].items()}
            out = O.decoder_forward(kind, p, blob["encoder_out"].double(), tags, blob["captions"],
                                    blob["caption_lengths"], sort_ind=blob["sort_ind"], dropout_masks=masks,
                                    hoist=hoist)
            loss = O.caption_loss(out[0], out[1], out[2], out[3])
            loss.backward()
            res.append((out[0].detach(), out[3].detach(), loss.item(), {k: v.grad for k, v in p.items()}))
        a, b = res
        assert (a[0] - b[0]).abs().max().item() <= 1e-12 * a[0].abs().max().item()
        assert (a[1] - b[1]).abs().max().item() <= 1e-13
        assert abs(a[2] - b[2]) <= 1e-12 * abs(a[2])
        for n in a[3]:
            if n == "attention.full_att.bias":      # mathematically zero (softmax shift invariance): rounding noise
                assert a[3][n].abs().max().item() < 1e-15 and b[3][n].abs().max().item() < 1e-15
                continue
            assert (a[3][n] - b[3][n]).abs().max().item() <= 1e-11 * max(a[3][n].abs().max().item(), 1e-9), n
        if masks is None:       # and both agree with what the live reference produced (fp32 goldens)
            assert (b[0].float() - blob["predictions"]).abs().max().item() <= 2e-5 * blob["predictions"].abs().max().item()
