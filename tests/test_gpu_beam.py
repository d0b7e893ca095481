"""GPU parity tests of the batched beam search (capdec_beam_search) against the golden captions
produced by the live reference `sample` and against the oracle's step-by-step trace.

fp32 mode: token indices exactly identical (BASELINE.json north_star)."""
import pytest
import torch

import capdec
from oracle import capdec_oracle as O
from conftest import load_golden
from gpu_util import build_decoder

pytestmark = pytest.mark.gpu

KINDS = [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION]


def _decoder(blob):
    dec = build_decoder(blob["kind"], blob["dims"])
    dec.load_state_dict(blob["state_dict"], strict=True)
    return dec.eval()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("k", [1, 3, 5])
def test_beam_batch_matches_reference_and_oracle_trace(kind, k):
    blob = load_golden("beam_" + kind)
    start, end = blob["start_id"], blob["end_id"]
    with capdec.precision_scope("fp32"):
        dec = _decoder(blob)
        enc = torch.cat([im["encoder_out"] for im in blob["images"]]).cuda()
        tags = torch.cat([im["tags"] for im in blob["images"]]).cuda()
        with torch.no_grad():
            res = dec.sample_batch(k, start, end, enc, None if kind == O.PURE_ATTENTION else tags,
                                   want_alphas=True, want_trace=True)
        seqs, lens = res["seq"].cpu(), res["len"].cpu()
        done = res["completed"].cpu()
        tr_parent, tr_word, tr_score = [t.cpu() for t in res["trace"]]
        n_done = n_fail = 0
        for g, im in enumerate(blob["images"]):
            ref = im["results"][k]
            assert bool(done[g]) == ref["completed"]
            got_seq = seqs[g, :lens[g]].tolist()
            with torch.no_grad():
                orc = O.beam_search(kind, blob["state_dict"], im["encoder_out"],
                                    None if kind == O.PURE_ATTENTION else im["tags"], k, start, end)
            # every step's top-k picks (parents, words) are identical, scores within fp32 rounding
            for t, (parents, words, scores) in enumerate(orc["trace"]):
                n = len(words)
                assert tr_parent[g, t, :n].tolist() == parents, (g, t)
                assert tr_word[g, t, :n].tolist() == words, (g, t)
                assert (tr_score[g, t, :n] - torch.tensor(scores)).abs().max().item() < 1e-3, (g, t)
                assert (tr_parent[g, t, n:] == -1).all()
            assert (tr_parent[g, len(orc["trace"]):] == -1).all()      # finished images stay idle
            assert got_seq == orc["seq"]
            assert abs(res["score"][g].item() - orc["score"]) < 1e-3
            if ref["completed"]:
                n_done += 1
                assert got_seq == ref["seq"]
                assert got_seq[0] == start and got_seq[-1] == end
                if ref["alphas"] is not None:
                    a = res["alpha"][g, :lens[g]].cpu().view(ref["alphas"].shape)
                    assert (a - ref["alphas"]).abs().max().item() < 1e-5
                    assert (a[0] == 1).all()
            else:
                n_fail += 1
                assert lens[g].item() == 52
                if kind != O.PURE_SCN:
                    a = res["alpha"][g].cpu().view(52, -1)
                    ref_a = torch.tensor(orc["alphas"]).view(52, -1)
                    assert (a - ref_a).abs().max().item() < 1e-5
        assert n_done > 0


@pytest.mark.parametrize("kind", KINDS)
def test_sample_single_image_contract(kind):
    """`sample(beam, word_map, enc[, tags])` keeps the reference's return types and its ValueError."""
    blob = load_golden("beam_" + kind)
    V = blob["dims"]["V"]
    word_map = {"w%d" % i: i for i in range(1, V - 3)}
    word_map.update({"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1})
    with capdec.precision_scope("fp32"):
        dec = _decoder(blob)
        for im in blob["images"]:
            ref = im["results"][3]
            args = (im["encoder_out"].cuda(),) if kind == O.PURE_ATTENTION else \
                (im["encoder_out"].cuda(), im["tags"].cuda())
            if not ref["completed"]:
                with pytest.raises(ValueError):
                    dec.sample(3, word_map, *args)
                dec.raise_on_incomplete = False
                out = dec.sample(3, word_map, *args)
                dec.raise_on_incomplete = True
                seq = out if kind == O.PURE_SCN else out[0]
                assert len(seq) == 52 and not dec.last_sample_completed
                continue
            out = dec.sample(3, word_map, *args)
            if kind == O.PURE_SCN:
                assert out == ref["seq"]
            else:
                seq, alphas = out
                assert seq == ref["seq"]
                assert isinstance(alphas, list) and len(alphas) == len(seq)
                side = im["encoder_out"].size(1)
                assert len(alphas[0]) == side and len(alphas[0][0]) == side


def test_beam_bf16_runs_and_full_width_properties():
    """Reference dims, beam 3, random-init weights: no beam ever terminates, every
    image decodes the full 51 steps, scores are finite and non-increasing along the trace."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    G = 5
    g = torch.Generator().manual_seed(3)
    enc = torch.randn(G, 14, 14, dims["E"], generator=g).relu_().cuda()
    tags = torch.rand(G, dims["S"], generator=g).cuda()
    for prec in ("fp32", "bf16"):
        with capdec.precision_scope(prec):
            torch.manual_seed(0)
            dec = build_decoder(O.ATTENTION_SCN, dims).eval()
            with torch.no_grad():
                res = dec.sample_batch(3, dims["V"] - 2, dims["V"] - 1, enc, tags, want_trace=True)
            assert (res["completed"] == 0).all()
            assert (res["len"] == 52).all()
            sc = res["trace"][2]
            assert torch.isfinite(sc).all()
            assert (sc[:, 1:, 0] <= sc[:, :-1, 0] + 1e-6).all()      # best cumulative log-prob decreases
            assert (res["seq"][:, 0] == dims["V"] - 2).all()
            assert ((res["alpha"][:, 1:].sum(-1) - 1).abs() < 1e-3).all()
            # the same image twice in one batch decodes identically (searches are independent)
            enc2 = torch.cat([enc[:1], enc[:1]])
            tags2 = torch.cat([tags[:1], tags[:1]])
            with torch.no_grad():
                r2 = dec.sample_batch(3, dims["V"] - 2, dims["V"] - 1, enc2, tags2)
            assert torch.equal(r2["seq"][0], r2["seq"][1])
            if prec == "fp32":
                assert torch.equal(r2["seq"][0], res["seq"][0])


@pytest.mark.parametrize("k", [1, 3, 5])
def test_beam_streaming_weighted_sum_matches_register_kernel(k, monkeypatch):
    """bf16 beam search at the reference dims: the shared-memory-ring weighted sum (chunk-major feature copy,
    one bulk copy per stage; the large-batch path) and the register-streaming kernel are two schedules of the
    same sums -- identical tokens, alphas and scores to rounding."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    G = 7
    g = torch.Generator().manual_seed(11)
    enc = torch.randn(G, 14, 14, dims["E"], generator=g).relu_().cuda()
    tags = torch.rand(G, dims["S"], generator=g).cuda()
    out = {}
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).eval()
        with torch.no_grad():
            dec.fc.weight.mul_(30.0)          # peaked distributions: a rounding difference must not flip a near-tie
        for mode in ("1", "0"):
            monkeypatch.setenv("CAPDEC_WSUM_STREAM", mode)
            with torch.no_grad():
                out[mode] = dec.sample_batch(k, dims["V"] - 2, dims["V"] - 1, enc, tags, max_steps=12, want_trace=True)
    a, b = out["1"], out["0"]
    assert torch.equal(a["seq"], b["seq"])
    assert (a["alpha"] - b["alpha"]).abs().max().item() < 1e-5
    assert ((a["score"] - b["score"]).abs() <= 5e-4 * b["score"].abs() + 1e-3).all()     # fc x 30 scales the scores
    assert ((a["alpha"][:, 1:].sum(-1) - 1).abs() < 1e-3).all()


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("k", [1, 3, 5])
def test_beam_single_pass_selection_matches_exact_selection(k, fused, monkeypatch):
    """bf16 mode either fuses the vocabulary projection with the log-softmax statistics and the per-tile top-k
    candidates (gemm_tc_vocab_topk: the logits are never written; CAPDEC_BEAM_FUSED unset / 1) or selects with ONE
    pass per row over materialised logits (CAPDEC_BEAM_FUSED=0); the parity mode's three-pass arithmetic on the same
    logits must give the same picks and scores to rounding."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    G = 6
    g = torch.Generator().manual_seed(5)
    enc = torch.randn(G, 14, 14, dims["E"], generator=g).relu_().cuda()
    tags = torch.rand(G, dims["S"], generator=g).cuda()
    out = {}
    monkeypatch.setenv("CAPDEC_BEAM_FUSED", fused)
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).eval()
        with torch.no_grad():
            dec.fc.weight.mul_(30.0)          # peaked distributions: distinct candidates, well-separated scores
        for mode in ("0", "1"):
            monkeypatch.setenv("CAPDEC_BEAM_EXACT", mode)
            with torch.no_grad():
                out[mode] = dec.sample_batch(k, dims["V"] - 2, dims["V"] - 1, enc, tags, max_steps=10, want_trace=True)
    a, b = out["0"], out["1"]
    assert torch.equal(a["trace"][0], b["trace"][0]) and torch.equal(a["trace"][1], b["trace"][1])
    assert torch.equal(a["seq"], b["seq"])
    assert (a["trace"][2] - b["trace"][2]).abs().max().item() < 1e-3


def test_beam_fused_vocabulary_kernel_large_batch(monkeypatch):
    """The fused vocabulary kernel against separate GEMM + single-pass selection at a batch that spans several
    128-row tiles (300 rows, ragged last tile; V = 10 000 has a ragged last vocabulary tile too)."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    G = 100
    g = torch.Generator().manual_seed(21)
    enc = torch.randn(G, 14, 14, dims["E"], generator=g).relu_().cuda()
    tags = torch.rand(G, dims["S"], generator=g).cuda()
    out = {}
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).eval()
        with torch.no_grad():
            dec.fc.weight.mul_(30.0)
        for mode in ("1", "0"):
            monkeypatch.setenv("CAPDEC_BEAM_FUSED", mode)
            with torch.no_grad():
                out[mode] = dec.sample_batch(3, dims["V"] - 2, dims["V"] - 1, enc, tags, max_steps=8, want_trace=True)
    a, b = out["1"], out["0"]
    assert torch.equal(a["trace"][0], b["trace"][0]) and torch.equal(a["trace"][1], b["trace"][1])
    assert torch.equal(a["seq"], b["seq"])
    assert (a["trace"][2] - b["trace"][2]).abs().max().item() < 1e-3


@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_ATTENTION])
def test_beam_strided_encoder_features(kind):
    """The reference encoder hands `sample` a permuted view that is physically NCHW (models/encoders/caption.py:43,
    SURVEY App. C-22): capdec_beam_search_strided takes it as it is -- same tokens, parents and scores as with the
    dense copy, in the token-exact fp32 mode."""
    blob = load_golden("beam_" + kind)
    start, end = blob["start_id"], blob["end_id"]
    with capdec.precision_scope("fp32"):
        dec = _decoder(blob)
        enc = torch.cat([im["encoder_out"] for im in blob["images"]]).cuda()
        G = enc.size(0)
        assert enc.dim() == 4                                                          # (G, side, side, E)
        enc4 = enc
        view = enc4.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)              # NHWC view of NCHW memory
        assert not view.is_contiguous() and torch.equal(view, enc4)
        tags = None if kind == O.PURE_ATTENTION else torch.cat([im["tags"] for im in blob["images"]]).cuda()
        with torch.no_grad():
            a = dec.sample_batch(3, start, end, enc4, tags, want_trace=True)
            b = dec.sample_batch(3, start, end, view, tags, want_trace=True)
        assert torch.equal(a["seq"], b["seq"]) and torch.equal(a["len"], b["len"])
        assert torch.equal(a["trace"][0], b["trace"][0]) and torch.equal(a["trace"][1], b["trace"][1])
        assert torch.equal(a["score"], b["score"])
