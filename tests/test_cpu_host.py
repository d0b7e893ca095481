"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the drop-in
modules expose the reference's state_dict layout, and nothing silently computes on the CPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import capdec_oracle as O
from conftest import ROOT, PKG


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "capdec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(capdec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from capdec import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for " + n
    assert lib.capdec_version() == 100


def test_workspace_query_is_host_only():
    from capdec import _lib, functional as CF
    lib = _lib.load()
    d = CF.make_dims("attention_scn", "bf16", 32, 50, 196, 2048, 512, 512, 512, 512, 1000, 10000, 52)
    fwd = lib.capdec_workspace_bytes(ctypes.byref(d), 0)
    both = lib.capdec_workspace_bytes(ctypes.byref(d), 1)
    assert 0 < fwd < both < 4 << 30
    bad = CF.make_dims("attention_scn", "bf16", 32, 50, 196, 2048, 512, 512, 510, 512, 1000, 10000, 52)
    assert lib.capdec_workspace_bytes(ctypes.byref(bad), 0) == 0
    assert b"multiples of 8" in lib.capdec_last_error()


@pytest.mark.parametrize("kind", O.KINDS)
def test_state_dict_layout_matches_reference(kind):
    from gpu_util import build_decoder
    dims = dict(A=24, M=16, D=32, F=24, S=12, V=37, E=40)
    dec = build_decoder(kind, dims, device="cpu")
    want = O.param_shapes(kind, attention_dim=24, embed_dim=16, decoder_dim=32, factored_dim=24,
                          semantic_dim=12, vocab_size=37, encoder_dim=40)
    got = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert got == want
    assert list(got) == list(want)           # same order as the reference modules register them
    from capdec import functional as CF
    assert sorted(CF.param_names(kind)) == sorted(want)


def test_no_cpu_fallback():
    from gpu_util import build_decoder
    from capdec._lib import CapdecError
    dims = dict(A=24, M=16, D=32, F=24, S=12, V=37, E=40)
    dec = build_decoder(O.ATTENTION_SCN, dims, device="cpu").eval()
    enc, tags, caps, caplens = O.synthetic_batch(3, 37, side=3, E=40, S=12, max_len=14,
                                                 lengths=[5, 9, 3])
    with pytest.raises(CapdecError):
        dec(enc, tags, caps, caplens)


def test_product_never_imports_oracle():
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
@pytest.mark.parametrize("kind,cls", [("attention_scn", "AttentionSCN"), ("pure_scn", "PureSCN"),
                                      ("pure_attention", "PureAttention")])
def test_same_seed_same_init_as_reference(kind, cls, tmp_path):
    """Construction order mirrors the reference, so torch.manual_seed(k) gives identical weights."""
    out = tmp_path / "ref_sd.pt"
    code = (
        "import sys, torch; sys.path.insert(0, '/root/reference');"
        "from models.decoders.%s import %s as C;"
        "torch.manual_seed(7);"
        "m = C(24,16,32,24,12,37,encoder_dim=40) if '%s'=='attention_scn' else "
        "(C(16,32,24,12,37,encoder_dim=40) if '%s'=='pure_scn' else C(24,16,32,37,encoder_dim=40));"
        "torch.save(m.state_dict(), r'%s')" % (kind, cls, kind, kind, out))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=str(tmp_path),
                   env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    ref = torch.load(out)
    from gpu_util import build_decoder
    torch.manual_seed(7)
    dec = build_decoder(kind, dict(A=24, M=16, D=32, F=24, S=12, V=37, E=40), device="cpu")
    mine = dec.state_dict()
    assert list(mine) == list(ref)
    for k in ref:
        assert torch.equal(mine[k], ref[k]), k


def test_clip_adam_is_an_optimizer_and_has_no_cpu_path():
    """ClipAdam keeps torch.optim.Adam's param_groups / state layout (reference: trains/attention_scn.py:91-92,
    utils/optimizer.py adjust_learning_rate) and refuses CPU tensors."""
    import torch
    from capdec.optim import ClipAdam
    from capdec._lib import CapdecError
    p = torch.nn.Parameter(torch.zeros(4, 3))
    opt = ClipAdam([p], lr=4e-4, grad_clip=5.0)
    assert isinstance(opt, torch.optim.Optimizer)
    assert opt.param_groups[0]["lr"] == 4e-4 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    for group in opt.param_groups:
        group["lr"] *= 0.8
    assert abs(opt.state_dict()["param_groups"][0]["lr"] - 3.2e-4) < 1e-12
    opt.step()                      # no gradients: nothing to do, no library call
    p.grad = torch.ones_like(p)
    try:
        opt.step()
        assert False, "CPU tensors must be refused"
    except CapdecError:
        pass


def test_evalcap_reference_layout():
    """capdec.evalcap.transpose_references = the re-shaping of eval_caption.py:135-141 ([image][caption] ->
    [caption][image]); compute_metrics needs nlg-eval, which this image does not have."""
    from capdec import evalcap
    refs = [["a b", "c"], ["d", "e f"], ["g", "h"]]
    assert evalcap.transpose_references(refs) == [["a b", "d", "g"], ["c", "e f", "h"]]
    assert evalcap.transpose_references([]) == []
    try:
        import nlgeval  # noqa: F401
    except ImportError:
        import pytest
        with pytest.raises(ImportError):
            evalcap.compute_metrics(refs, ["x", "y", "z"])


def test_evalcap_driver_with_a_stub_decoder():
    """Host logic of capdec.evalcap.generate_captions (string formatting of eval_caption.py:121-129, batching,
    train/eval mode restore) with a stub in place of the CUDA decoder."""
    import torch
    from capdec import evalcap

    word_map = {"<pad>": 0, "a": 1, "b": 2, "c": 3, "<unk>": 4, "<start>": 5, "<end>": 6}

    class Stub(torch.nn.Module):
        kind = "attention_scn"

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.calls = []

        def sample_batch(self, beam, start, end, enc, tags, max_steps=50, want_alphas=True):
            assert not self.training and tags is not None and start == 5 and end == 6
            self.calls.append(enc.size(0))
            G = enc.size(0)
            seq = torch.zeros(G, 6, dtype=torch.int32)
            seq[:, 0] = 5
            seq[:, 1] = 1 + (enc.view(G, -1)[:, 0].long() % 3).int()
            seq[:, 2] = 2
            seq[:, 3] = 6
            return {"seq": seq, "len": torch.full((G,), 4, dtype=torch.int32),
                    "completed": torch.tensor([1] * (G - 1) + [0], dtype=torch.int32)}

    dec = Stub().train()
    enc = torch.arange(5, dtype=torch.float32).view(5, 1, 1, 1)
    tags = torch.zeros(5, 3)
    allcaps = torch.tensor([[[5, 1, 2, 6, 0], [5, 3, 6, 0, 0]]] * 5)
    batches = [(enc[:3], tags[:3], allcaps[:3]), (enc[3:], tags[3:], allcaps[3:])]
    refs, hyps, done = evalcap.generate_captions(dec, batches, word_map, beam_size=3)
    assert dec.training and dec.calls == [3, 2]
    assert hyps == ["a b", "b b", "c b", "a b", "b b"]
    assert refs == [["a b", "c"]] * 5
    assert done == [True, True, False, True, False]
    try:
        evalcap.generate_captions(dec, [(enc, allcaps)], word_map)      # SCN decoder without tags
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_eval_caption_overlay_runs_the_reference_loop_in_batches(tmp_path):
    """The overlay eval_caption.py: the reference's `evaluate(args)` with its per-image loop
    (eval_caption.py:96-131) replaced by the batched driver -- same CLI arguments, same reference / hypothesis
    strings, same nlg-eval layout and JSON files.  Stubs stand in for the dataset, the encoders and the CUDA decoder."""
    import importlib.util
    import json
    import torch

    spec = importlib.util.spec_from_file_location(
        "capdec_eval_caption", os.path.join(ROOT, "indonesian-image-captioning_b200", "eval_caption.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    # every option of the reference's parser (eval_caption.py:171-185) is still there
    args = mod.build_parser().parse_args(["-t", "attention_scn", "-mc", "c.pth", "-mt", "t.pth", "-df", "d", "-dn", "n",
                                          "-tm", "tm.json", "-wm", "wm.json", "-bs", "3", "--output_dir", str(tmp_path)])
    assert (args.type, args.model_caption, args.beam_size, args.batch_size) == ("attention_scn", "c.pth", 3, 32)

    word_map = {"<pad>": 0, "a": 1, "b": 2, "c": 3, "<unk>": 4, "<start>": 5, "<end>": 6}

    class Enc(torch.nn.Module):
        def forward(self, image):                      # (G, 3, H, W) -> (G, 1, 1, 1) "features"
            return image[:, :1, :1, :1].permute(0, 2, 3, 1)

    class Tagger(torch.nn.Module):
        def forward(self, image):
            return torch.zeros(image.size(0), 4)

    class Dec(torch.nn.Module):
        kind = "attention_scn"

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.calls = []

        def sample_batch(self, beam, start, end, enc, tags, max_steps=50, want_alphas=True):
            G = enc.size(0)
            self.calls.append((G, beam, tuple(tags.shape)))
            seq = torch.zeros(G, 5, dtype=torch.int32)
            seq[:, 0] = start
            seq[:, 1] = 1 + (enc.reshape(G, -1)[:, 0].long() % 3).int()
            seq[:, 2] = end
            return {"seq": seq, "len": torch.full((G,), 3, dtype=torch.int32),
                    "completed": torch.tensor([1] * (G - 1) + [0], dtype=torch.int32)}

    images = torch.arange(5, dtype=torch.float32).view(5, 1, 1, 1).expand(5, 3, 2, 2).contiguous()
    allcaps = torch.tensor([[[5, 1, 2, 6, 0], [5, 3, 6, 0, 0]]] * 5)
    loader = [(images[:3], None, None, allcaps[:3]), (images[3:], None, None, allcaps[3:])]
    dec = Dec()
    seen = {}

    def metrics(references, hypotheses):
        seen["refs"], seen["hyps"] = references, hypotheses
        return {"Bleu_4": 0.5}

    scores = mod.evaluate(args, loader=loader, models=(Enc(), Tagger(), dec), word_map=word_map, metrics=metrics)
    assert scores == {"Bleu_4": 0.5}
    assert dec.calls == [(3, 3, (3, 4)), (2, 3, (2, 4))]
    assert seen["hyps"] == ["a", "b", "c", "a", "b"]
    assert seen["refs"] == [["a b"] * 5, ["c"] * 5]                  # [caption][image], eval_caption.py:135-141
    out = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path) for f in fs]
    names = sorted(os.path.basename(f) for f in out)
    assert names == ["attention_scn_beam_3_hypotheses.json", "attention_scn_beam_3_incomplete.json",
                     "attention_scn_beam_3_references.json", "attention_scn_beam_3_scores.json"]
    by = {os.path.basename(f): json.load(open(f)) for f in out}
    assert by["attention_scn_beam_3_hypotheses.json"] == seen["hyps"]
    assert by["attention_scn_beam_3_incomplete.json"] == [2, 4]


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: ONE JSON line with the B200 arm's metric / unit / config, `impl: reference`, a
    cpu_baseline describing the run (kind = the unmodified reference modules of oracle/_ref) and an e2e block with no
    copies; ranks other than 0 of a torchrun launch exit 0 without printing."""
    import json
    import subprocess
    import sys
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")) and not os.path.isdir("/root/reference"):
        import pytest
        pytest.skip("neither oracle/_ref nor the reference tree is present")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--batch", "2"]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "captions/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("captions/sec") and d["config"]["workload"] == "attention_scn_train"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-9 * max(1.0, d["value"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0
    env.update(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
