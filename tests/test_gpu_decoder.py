"""GPU parity tests of the drop-in decoder modules through the C ABI.

fp32 mode: logits / alphas / loss / every parameter gradient within 1e-4 (max-norm relative,
BASELINE.json north_star) of the golden vectors produced by the live reference and of the
fp64 oracle.  bf16 mode: logits within 2e-2.
"""
import pytest
import torch

import capdec
from oracle import capdec_oracle as O
from conftest import load_golden
from gpu_util import build_decoder, call_forward, torch_loss_glue, rel_err, rel_err_fro, oracle_run

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
GOLDEN = ["train_attention_scn_small", "train_pure_scn_small", "train_pure_attention_small",
          "train_attention_scn_medium", "train_pure_scn_medium", "train_pure_attention_medium",
          "train_attention_scn_hot"]


def _load(blob):
    kind = blob["kind"]
    dec = build_decoder(kind, blob["dims"])
    dec.load_state_dict(blob["state_dict"], strict=True)
    dec.eval()
    args = [blob[k].cuda() for k in ("encoder_out", "tags", "captions", "caption_lengths")]
    return kind, dec, args


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("loss_path", ["torch_glue", "fused"])
def test_fp32_matches_reference_golden(name, loss_path):
    blob = load_golden(name)
    with capdec.precision_scope("fp32"):
        kind, dec, (enc, tags, caps, caplens) = _load(blob)
        scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, enc, tags, caps, caplens)
        assert dl == blob["decode_lengths"]
        assert torch.equal(sort_ind.cpu(), blob["sort_ind"])
        assert torch.equal(caps_sorted.cpu(), blob["caps_sorted"])
        assert rel_err(scores, blob["predictions"]) < FP32_TOL
        if alphas is not None:
            assert rel_err(alphas, blob["alphas"]) < FP32_TOL
        # rows beyond each caption's length stay exactly zero (reference zero-inits the buffers)
        for i, L in enumerate(dl):
            assert scores[i, L:].abs().max().item() == 0 if L < scores.shape[1] else True
        if loss_path == "fused":
            loss, parts = dec.loss(scores, caps_sorted, dl, alphas, alpha_c=1.0)
        else:
            loss = torch_loss_glue(scores, caps_sorted, dl, alphas)
        assert abs(loss.item() - blob["loss"].item()) < FP32_TOL * max(1.0, abs(blob["loss"].item()))
        dec.zero_grad()
        loss.backward()
        worst = ("", 0.0)
        for n, p in dec.named_parameters():
            ref = blob["grads"][n]
            assert p.grad is not None, n
            if ref.abs().max().item() < 1e-7:      # full_att.bias: mathematically zero
                assert p.grad.abs().max().item() < 1e-5, n
                continue
            e = rel_err(p.grad, ref)
            if e > worst[1]:
                worst = (n, e)
        assert worst[1] < 2 * FP32_TOL, worst
        assert torch.equal(dec.decode_step.bias_ih.grad, dec.decode_step.bias_hh.grad)


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("name", ["train_attention_scn_medium", "train_pure_scn_medium",
                                  "train_pure_attention_medium"])
def test_bf16_within_tolerance(name, fused, monkeypatch):
    """fused=1 also drives the opt-in fused GEMM epilogues (CAPDEC_FUSED_EPILOGUE, gemm_tc.cu)."""
    monkeypatch.setenv("CAPDEC_FUSED_EPILOGUE", fused)
    blob = load_golden(name)
    with capdec.precision_scope("bf16"):
        kind, dec, (enc, tags, caps, caplens) = _load(blob)
        scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, enc, tags, caps, caplens)
        assert rel_err(scores, blob["predictions"]) < BF16_TOL
        if alphas is not None:
            assert rel_err(alphas, blob["alphas"]) < BF16_TOL
        loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
        assert abs(loss.item() - blob["loss"].item()) < BF16_TOL * abs(blob["loss"].item())
        dec.zero_grad()
        loss.backward()
        for n, p in dec.named_parameters():
            ref = blob["grads"][n]
            if ref.abs().max().item() < 1e-7:
                continue
            assert rel_err(p.grad, ref) < 0.1, n     # loose: bf16 operands through T steps


@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_fp32_full_width_matches_oracle(kind):
    """Reference dims (512/2048/1000, V=10k) at a batch the CPU oracle finishes in seconds."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    B = 4
    lengths = [9, 5, 12, 3]
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=3, lengths=lengths)
    with capdec.precision_scope("fp32"):
        torch.manual_seed(0)
        dec = build_decoder(kind, dims).eval()
        sd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
        scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, enc.cuda(), tags.cuda(),
                                                                 caps.cuda(), caplens.cuda())
        ref = oracle_run(kind, sd, enc, tags, caps, caplens, sort_ind=sort_ind)
        # the reference's own arithmetic (fp32) measured against the same fp64 truth: the relu
        # kink of the attention scores makes a few gradients jump when a pre-activation changes
        # sign in the last ulp, so every fp32 implementation sits this far from fp64
        ref32 = oracle_run(kind, sd, enc, tags, caps, caplens, sort_ind=sort_ind, dtype=torch.float32)
        assert dl == ref["decode_lengths"]
        assert rel_err(scores, ref["scores"]) < FP32_TOL
        if alphas is not None:
            assert rel_err(alphas, ref["alphas"]) < FP32_TOL
        loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
        assert abs(loss.item() - ref["loss"].item()) < FP32_TOL * abs(ref["loss"].item())
        loss.backward()
        bad = []
        for n, p in dec.named_parameters():
            g = ref["grads"][n]
            if g.abs().max().item() < 1e-9:
                continue
            tol = max(2 * FP32_TOL, 4 * rel_err(ref32["grads"][n], g))
            e = rel_err(p.grad, g)
            if n.startswith("attention.encoder_att") or n.startswith("attention.decoder_att"):
                # downstream of the relu mask 1[att1+att2 > 0]: ONE mask that flips in the last ulp
                # moves single entries of these sums by a full term (1e-2 of the max with only 29
                # (b,t) rows here), so they are judged in the Frobenius norm.  The kernel itself is
                # pinned to 2e-5 by test_attention_bwd_step_matches_autograd on kink-free inputs.
                tol = max(tol, 1e-2)
                e = rel_err_fro(p.grad, g)
            if e >= tol:
                bad.append((n, e, tol))
        assert not bad, bad


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16-fused"])
def test_full_size_properties(precision, monkeypatch):
    """BASELINE config 3 per-GPU shape (B=32, T=50, V=10k): size-independent properties."""
    if precision == "bf16-fused":
        monkeypatch.setenv("CAPDEC_FUSED_EPILOGUE", "1")
        precision = "bf16"
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    B = 32
    lengths = O.tie_free_lengths(B)
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=5, lengths=lengths)
    with capdec.precision_scope(precision):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).eval()
        args = [t.cuda() for t in (enc, tags, caps, caplens)]
        scores, caps_sorted, dl, alphas, sort_ind = dec(*args)
        assert dl == sorted([l - 1 for l in lengths], reverse=True)
        for i, L in enumerate(dl):
            if L < scores.shape[1]:
                assert scores[i, L:].abs().max().item() == 0
                assert alphas[i, L:].abs().max().item() == 0
            assert (alphas[i, :L].sum(-1) - 1).abs().max().item() < 1e-4
        assert torch.isfinite(scores).all()
        # tags are NOT permuted with the captions (reference quirk)
        scores2 = dec(args[0], args[1][sort_ind], args[2], args[3])[0]
        assert (scores2 - scores).abs().max().item() > 0
        # determinism of the whole forward
        scores3 = dec(*args)[0]
        if precision == "fp32":
            assert torch.equal(scores3, scores)
        else:       # split-K slices are reduced with fp32 atomics: order-dependent in the last bits
            assert rel_err(scores3, scores) < BF16_TOL
        loss, parts = dec.loss(scores, caps_sorted, dl, alphas)
        loss.backward()
        assert torch.equal(dec.decode_step.bias_ih.grad, dec.decode_step.bias_hh.grad)
        emb_g = dec.embedding.weight.grad
        used = torch.zeros(dims["V"], dtype=torch.bool, device="cuda")
        for i, L in enumerate(dl):
            used[caps_sorted[i, :L]] = True
        assert emb_g[~used].abs().max().item() == 0      # SURVEY §4 invariant 5
        assert emb_g[used].abs().sum().item() > 0
        for n, p in dec.named_parameters():
            assert torch.isfinite(p.grad).all(), n


def test_dropout_training_mode_runs_and_is_seeded():
    dims = dict(A=64, M=48, D=64, F=56, S=100, V=203, E=128)
    enc, tags, caps, caplens = O.synthetic_batch(8, dims["V"], seed=9, side=14, E=dims["E"], S=dims["S"],
                                                 max_len=20, lengths=[20, 7, 13, 5, 18, 9, 3, 16])
    with capdec.precision_scope("fp32"):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).train()
        args = [t.cuda() for t in (enc, tags, caps, caplens)]
        torch.manual_seed(1)
        s1 = dec(*args)[0]
        torch.manual_seed(1)
        s2 = dec(*args)[0]
        torch.manual_seed(2)
        s3 = dec(*args)[0]
        assert torch.equal(s1, s2) and not torch.equal(s1, s3)
        dec.eval()
        s_eval = dec(*args)[0]
        assert not torch.equal(s_eval, s1)
        dec.train()
        out = dec(*args)
        loss, _ = dec.loss(out[0], out[1], out[2], out[3])
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in dec.parameters())


def test_state_dict_round_trip_and_strided_encoder_features():
    """state_dict keys/shapes are the reference's; NCHW-physical encoder views are accepted (App. C-22)."""
    blob = load_golden("train_attention_scn_medium")
    with capdec.precision_scope("fp32"):
        kind, dec, (enc, tags, caps, caplens) = _load(blob)
        assert list(dec.state_dict().keys()) == list(blob["state_dict"].keys())
        ref_scores = dec(enc, tags, caps, caplens)[0]
        nchw = enc.permute(0, 3, 1, 2).contiguous()          # what the ResNet trunk really produces
        view = nchw.permute(0, 2, 3, 1)                      # the reference encoder's permuted view
        assert not view.is_contiguous()
        got = dec(view, tags, caps, caplens)[0]
        assert torch.equal(got, ref_scores)


@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_graph_replay_matches_eager(kind):
    """CUDA-graph replay of the compute phases gives the eager result bit for bit (fp32 mode is
    deterministic), also when inputs, weights and the dropout seed change between replays."""
    blob = load_golden("train_%s_medium" % kind)
    with capdec.precision_scope("fp32"):
        k, dec, (enc, tags, caps, caplens) = _load(blob)
        dec.train()

        def step(scale, seed):
            torch.manual_seed(seed)
            dec.zero_grad(set_to_none=True)
            out = call_forward(dec, kind, enc * scale, tags, caps, caplens)
            loss, _ = dec.loss(out[0], out[1], out[2], out[3])
            loss.backward()
            return out[0].clone(), loss.item(), [p.grad.clone() for p in dec.parameters()]

        eager = [step(1.0, 1), step(0.5, 2), step(0.25, 3)]
        capdec.set_graphs(True)
        try:
            for i, (scale, seed) in enumerate([(1.0, 1), (0.5, 2), (0.25, 3)]):   # eager, capture, replay
                s, l, g = step(scale, seed)
                assert torch.equal(s, eager[i][0])
                assert l == eager[i][1]
                for (n, _), a, b in zip(dec.named_parameters(), g, eager[i][2]):
                    assert torch.equal(a, b), (i, n, (a - b).abs().max().item())
            with torch.no_grad():          # weights change in place: the graph reads the new values
                for p in dec.parameters():
                    p.mul_(1.01)
            s_graph, l_graph, _ = step(1.0, 4)
        finally:
            capdec.set_graphs(False)
        s_eager, l_eager, _ = step(1.0, 4)
        assert torch.equal(s_graph, s_eager) and l_graph == l_eager


@pytest.mark.parametrize("graphs", [False, True])
def test_gradients_live_in_the_flat_buffer_and_accumulate(graphs):
    """`.grad` of every parameter is a view of the ONE flat gradient buffer the backward writes (what
    capdec.parallel.GradReducer all-reduces in place, no copies), and a second backward WITHOUT zeroing the
    gradients accumulates correctly although the static buffer is reused (eval mode: deterministic)."""
    blob = load_golden("train_%s_medium" % O.ATTENTION_SCN)
    with capdec.precision_scope("fp32"):
        k, dec, (enc, tags, caps, caplens) = _load(blob)
        dec.eval()

        def fwd_bwd():
            out = call_forward(dec, k, enc, tags, caps, caplens)
            loss, _ = dec.loss(out[0], out[1], out[2], out[3])
            loss.backward()
            return out[0]

        capdec.set_graphs(graphs)
        try:
            for _ in range(3 if graphs else 1):              # eager, capture, replay
                dec.zero_grad(set_to_none=True)
                scores = fwd_bwd()
            flat = scores._capdec_meta["flat_grads"]
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            params = [p for p in dec.parameters() if p.requires_grad]
            assert all(lo <= p.grad.data_ptr() < hi for p in params)
            assert sum(p.grad.numel() for p in params) <= flat.numel()
            assert all(p.grad.data_ptr() % 256 == 0 for p in params)      # aligned slots: vector loads in ClipAdam
            once = [p.grad.clone() for p in params]
            fwd_bwd()                                        # no zero_grad: accumulate
            for p, g in zip(params, once):
                assert torch.allclose(p.grad, 2 * g, rtol=1e-6, atol=1e-12)
        finally:
            capdec.set_graphs(False)


@pytest.mark.parametrize("persistent", ["1", "0"])
@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_bf16_full_width_matches_oracle(kind, persistent, monkeypatch):
    """bf16 fast path at the reference dims against the fp64 oracle; persistent=1 runs the SCN
    recurrences as one cooperative kernel each way (recur.cu), 0 the per-step kernel chains."""
    monkeypatch.setenv("CAPDEC_PERSISTENT", persistent)
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    lengths = [9, 5, 12, 3]
    enc, tags, caps, caplens = O.synthetic_batch(4, dims["V"], seed=3, lengths=lengths)
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(kind, dims).eval()
        sd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
        scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, enc.cuda(), tags.cuda(), caps.cuda(),
                                                                 caplens.cuda())
        ref = oracle_run(kind, sd, enc, tags, caps, caplens, sort_ind=sort_ind)
        assert rel_err(scores, ref["scores"]) < BF16_TOL
        if alphas is not None:
            assert rel_err(alphas, ref["alphas"]) < BF16_TOL
        loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
        assert abs(loss.item() - ref["loss"].item()) < BF16_TOL * abs(ref["loss"].item())
        loss.backward()
        bad = []
        for n, p in dec.named_parameters():
            g = ref["grads"][n]
            if g.abs().max().item() < 1e-9:
                continue
            e = rel_err(p.grad, g)
            if n.startswith("attention.encoder_att") or n.startswith("attention.decoder_att"):
                e = rel_err_fro(p.grad, g)       # relu-kink flips move single entries (see fp32 test)
            if e > 0.08:
                bad.append((n, e))
        assert not bad, bad


@pytest.mark.parametrize("train_mode", [False, True])
@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_persistent_recurrence_matches_step_kernels(kind, train_mode, monkeypatch):
    """Config-3 per-GPU shape (B=32, ragged lengths): the persistent cooperative kernels (recur.cu) and
    the per-step kernel chains are two schedules of the same bf16 arithmetic -- outputs, saved state and
    gradients agree to summation-order noise, the dropout masks are the same counter-based stream, and
    the persistent forward is bit-reproducible (no atomics)."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    B = 32
    lengths = O.tie_free_lengths(B)
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=7, lengths=lengths)
    args = [t.cuda() for t in (enc, tags, caps, caplens)]
    res = {}
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(kind, dims)
        dec.train(train_mode)
        for mode in ("1", "0", "1"):
            monkeypatch.setenv("CAPDEC_PERSISTENT", mode)
            torch.manual_seed(11)
            dec.zero_grad(set_to_none=True)
            scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, *args)
            loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
            loss.backward()
            cur = (scores.clone(), None if alphas is None else alphas.clone(), loss.item(),
                   {n: p.grad.clone() for n, p in dec.named_parameters()})
            if mode == "1" and "1" in res:
                assert torch.equal(cur[0], res["1"][0])          # deterministic forward
            res[mode] = cur
    a, b = res["1"], res["0"]
    assert torch.isfinite(a[0]).all()
    assert rel_err(a[0], b[0]) < 5e-3
    if a[1] is not None:
        assert rel_err(a[1], b[1]) < 5e-3
    assert abs(a[2] - b[2]) < 1e-3 * abs(b[2])
    for n in a[3]:
        if n == "attention.full_att.bias":      # mathematically zero (softmax shift invariance): rounding noise
            continue
        e = rel_err_fro(a[3][n], b[3][n])
        assert e < 3e-2, (n, e)


@pytest.mark.parametrize("lengths_kind", ["fixed51", "ragged"])
@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_benchmarked_mode_matches_oracle(kind, lengths_kind, monkeypatch):
    """The exact mode bench.py times -- BASELINE config-3 per-GPU shape (B=32, T=50, V=10k, dims 512), TRAIN mode
    with dropout p=0.5, bf16 features, persistent recurrence kernels -- against the fp64 oracle, and the same step
    in fp32 mode.  torch's dropout RNG cannot be matched, so the oracle is fed the decoder's own
    keep factors (capdec_dropout_mask: the counter-based hash of (seed, b, t, d) that the forward AND the backward
    kernels evaluate in place) as `dropout_masks` at the reference's dropout site (attention_scn.py:154): a wrong
    keep scale, a mask that differs between forward and backward, or a mis-indexed mask all show up here.
    Tolerances: logits / alphas / loss 2e-2 (bf16, BASELINE north_star) and 1e-4 (fp32); gradients of all
    parameters 8e-2 (bf16) and 2e-4 (fp32) max-norm relative, the two parameters downstream of the relu mask of
    the attention scores (a mask that flips in the last ulp moves single entries by a whole term) in the
    Frobenius norm (fp32: 1e-3)."""
    from capdec import functional as CF
    monkeypatch.setenv("CAPDEC_PERSISTENT", "1")
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    B, p_drop = 32, 0.5
    lengths = [51] * B if lengths_kind == "fixed51" else O.tie_free_lengths(B)
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=21, lengths=lengths)
    args = [t.cuda() for t in (enc, tags, caps, caplens)]
    torch.manual_seed(0)
    dec = build_decoder(kind, dims).train()
    assert dec.dropout.p == p_drop
    sd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    ref = None
    for prec, tol, gtol, gtol_fro in (("bf16", BF16_TOL, 8e-2, 8e-2), ("fp32", FP32_TOL, 2 * FP32_TOL, 1e-3)):
        with capdec.precision_scope(prec):
            torch.manual_seed(123)                       # same dropout seed in both modes
            dec.zero_grad(set_to_none=True)
            scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, *args)
            T = max(dl)
            assert T == 50 and dl == sorted([l - 1 for l in lengths], reverse=True)
            seed = scores._capdec_meta["seed"]
            mask = CF.dropout_mask(seed, p_drop, B, T, dims["D"])
            assert set(mask.unique().tolist()) == {0.0, 1.0 / (1.0 - p_drop)}
            assert abs((mask > 0).float().mean().item() - (1.0 - p_drop)) < 5e-3
            if ref is None:
                ref = oracle_run(kind, sd, enc, tags, caps, caplens, sort_ind=sort_ind, dropout_masks=mask,
                                 hoist=True)
            assert dl == ref["decode_lengths"]
            assert rel_err(scores, ref["scores"]) < tol, prec
            if alphas is not None:
                assert rel_err(alphas, ref["alphas"]) < tol, prec
            loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
            assert abs(loss.item() - ref["loss"].item()) < tol * abs(ref["loss"].item()), prec
            loss.backward()
            bad = []
            for n, p in dec.named_parameters():
                g = ref["grads"][n]
                if g.abs().max().item() < 1e-9:          # full_att.bias: mathematically zero
                    continue
                if n.startswith("attention.encoder_att") or n.startswith("attention.decoder_att"):
                    e, lim = rel_err_fro(p.grad, g), gtol_fro
                else:
                    e, lim = rel_err(p.grad, g), gtol
                if e >= lim:
                    bad.append((prec, n, e, lim))
            assert not bad, bad
    # the mask matters: the eval-mode logits are far from the training-mode ones
    dec.eval()
    with capdec.precision_scope("bf16"):
        s_eval = call_forward(dec, kind, *args)[0]
    assert rel_err(s_eval, ref["scores"]) > 10 * BF16_TOL


@pytest.mark.parametrize("precision,tol,gtol", [("bf16", BF16_TOL, 8e-2), ("fp32", FP32_TOL, 4 * FP32_TOL)])
def test_scaled_shape_matches_oracle(precision, tol, gtol):
    """BASELINE config 5 dims (decoder / factor 1024, vocab 30k; attention / embedding 512) at a batch the CPU
    oracle finishes in seconds, train mode with the decoder's own dropout mask fed to the oracle."""
    from capdec import functional as CF
    dims = dict(A=512, M=512, D=1024, F=1024, S=1000, V=30000, E=2048)
    lengths = [9, 5, 12, 3]
    enc, tags, caps, caplens = O.synthetic_batch(4, dims["V"], seed=31, lengths=lengths)
    with capdec.precision_scope(precision):
        torch.manual_seed(0)
        dec = build_decoder(O.ATTENTION_SCN, dims).train()
        sd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
        torch.manual_seed(5)
        scores, caps_sorted, dl, alphas, sort_ind = dec(enc.cuda(), tags.cuda(), caps.cuda(), caplens.cuda())
        mask = CF.dropout_mask(scores._capdec_meta["seed"], 0.5, 4, max(dl), dims["D"])
        ref = oracle_run(O.ATTENTION_SCN, sd, enc, tags, caps, caplens, sort_ind=sort_ind, dropout_masks=mask, hoist=True)
        assert rel_err(scores, ref["scores"]) < tol
        assert rel_err(alphas, ref["alphas"]) < tol
        loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
        assert abs(loss.item() - ref["loss"].item()) < tol * abs(ref["loss"].item())
        loss.backward()
        bad = []
        for n, p in dec.named_parameters():
            g = ref["grads"][n]
            if g.abs().max().item() < 1e-9:
                continue
            e = rel_err(p.grad, g)
            lim = gtol
            if n.startswith("attention.encoder_att") or n.startswith("attention.decoder_att"):
                e, lim = rel_err_fro(p.grad, g), max(gtol, 1e-2)     # relu-kink flips (see the fp32 full-width test)
            if e >= lim:
                bad.append((n, e, lim))
        assert not bad, bad


@pytest.mark.xfail(strict=False, reason="written after the round's GPU minutes were spent: never run on a B200 yet "
                                        "(expected to pass: every role mapping of recur_bwd_kernel is grid-size generic)")
def test_persistent_backward_on_fewer_sms_gives_the_same_gradients(monkeypatch):
    """Data-parallel runs launch recur_bwd_kernel on 144 instead of 148 SMs (CAPDEC_RECUR_BWD_CTAS, set by
    capdec.parallel.GradReducer so that the fc bucket's all-reduce runs next to the reverse loop): the work items
    are dealt to other CTAs but every item's arithmetic is unchanged, so the gradients agree with the full-grid
    launch.  Eager launches: a captured graph keeps the grid it was captured with."""
    dims = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
    B = 32
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=9, lengths=O.tie_free_lengths(B))
    args = [t.cuda() for t in (enc, tags, caps, caplens)]
    was = capdec.graphs_enabled() if hasattr(capdec, "graphs_enabled") else True
    capdec.set_graphs(False)
    try:
        with capdec.precision_scope("bf16"):
            torch.manual_seed(0)
            dec = build_decoder(O.ATTENTION_SCN, dims).train()
            grads = {}
            for ctas in (None, "144"):
                if ctas is None:
                    monkeypatch.delenv("CAPDEC_RECUR_BWD_CTAS", raising=False)
                else:
                    monkeypatch.setenv("CAPDEC_RECUR_BWD_CTAS", ctas)
                torch.manual_seed(11)
                dec.zero_grad(set_to_none=True)
                scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, O.ATTENTION_SCN, *args)
                loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
                loss.backward()
                grads[ctas] = {n: p.grad.clone() for n, p in dec.named_parameters()}
        for n, g in grads[None].items():
            assert torch.isfinite(grads["144"][n]).all(), n
            if g.abs().max().item() < 1e-9:
                continue
            assert rel_err_fro(grads["144"][n], g) < 1e-5, n
    finally:
        capdec.set_graphs(was)
