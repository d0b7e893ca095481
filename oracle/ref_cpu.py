#!/usr/bin/env python
"""CPU timing of the reference decoder path for bench.py (`--impl reference` and the `cpu_baseline` leg).
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the product never imports it.

    python oracle/ref_cpu.py --kind attention_scn --mode train --batch 32 --steps 20 --warmup 5 [--budget 240]

Runs in its OWN process (bench.py spawns it) so that the reference's `models` / `utils` packages never meet the
drop-in overlay of the same names, and prints ONE JSON object.

kind "reference": the UNMODIFIED reference modules from `oracle/_ref/` (oracle/make_ref.py), with the two
harness-side restatements of SURVEY.md §8c -- the module-global `device` (attention_scn.py:11) forced to CPU, and,
for `sample` only, `top_k_words / vocab_size` spelled `//` (attention_scn.py:252) -- and the loss glue of
trains/attention_scn.py:219-235 restated with the same stock torch ops.  Training steps run the reference's own
`forward` in `train()` mode (its nn.Dropout(0.5)), `CrossEntropyLoss` on the packed scores (+ the alpha
regulariser) and autograd's backward.
kind "port": the same through oracle/capdec_oracle.py, when `oracle/_ref/` is not there.

A step is `batch` captions of length 51 (T = 50) at the given dims; when the projected time of warmup + steps
exceeds `--budget` seconds the number of timed steps is cut (never below 2) and the JSON says how many ran.
"""
import argparse
import importlib
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, os.path.dirname(HERE))

CLS = {"attention_scn": "AttentionSCN", "pure_scn": "PureSCN", "pure_attention": "PureAttention"}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def load_reference_class(kind, fix_division):
    """The reference class from oracle/_ref (None when absent)."""
    from oracle import make_ref
    if not make_ref.available():
        return None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    m = importlib.import_module("models.decoders." + kind)
    assert os.path.abspath(m.__file__).startswith(REF), m.__file__
    m.device = torch.device("cpu")
    if not fix_division:
        return getattr(m, CLS[kind])
    with open(m.__file__) as fh:
        src = fh.read()
    assert "top_k_words / vocab_size" in src
    ns = dict(m.__dict__)
    exec(compile(src.replace("top_k_words / vocab_size", "top_k_words // vocab_size"), m.__file__, "exec"), ns)
    ns["device"] = torch.device("cpu")
    return ns[CLS[kind]]


def build_reference(cls, kind, d):
    torch.manual_seed(0)
    if kind == "attention_scn":
        return cls(d["A"], d["M"], d["D"], d["F"], d["S"], d["V"], encoder_dim=d["E"], dropout=0.5)
    if kind == "pure_scn":
        return cls(d["M"], d["D"], d["F"], d["S"], d["V"], encoder_dim=d["E"], dropout=0.5)
    return cls(d["A"], d["M"], d["D"], d["V"], encoder_dim=d["E"], dropout=0.5)


def loss_glue(scores, caps_sorted, decode_lengths, alphas, alpha_c=1.0):
    """trains/attention_scn.py:219-235 with the same stock torch ops."""
    from torch.nn.utils.rnn import pack_padded_sequence
    targets = caps_sorted[:, 1:]
    s = pack_padded_sequence(scores, decode_lengths, batch_first=True).data
    t = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    loss = torch.nn.CrossEntropyLoss()(s, t)
    if alphas is not None:
        loss = loss + alpha_c * ((1. - alphas.sum(dim=1)) ** 2).mean()
    return loss


def time_train(kind, d, batch, steps, warmup, budget, cap_len=51):
    from oracle import capdec_oracle as O
    enc, tags, caps, caplens = O.synthetic_batch(batch, d["V"], seed=1234, lengths=[cap_len] * batch)
    cls = load_reference_class(kind, False)
    if cls is not None:
        how = "reference"
        dec = build_reference(cls, kind, d).train()
        params = list(dec.parameters())

        def step():
            out = dec(enc, caps, caplens) if kind == "pure_attention" else dec(enc, tags, caps, caplens)
            alphas = None if kind == "pure_scn" else out[3]
            loss = loss_glue(out[0], out[1], out[2], alphas)
            for p in params:
                p.grad = None
            loss.backward()
    else:
        how = "port"
        dkw = dict(attention_dim=d["A"], embed_dim=d["M"], decoder_dim=d["D"], factored_dim=d["F"],
                   semantic_dim=d["S"], vocab_size=d["V"], encoder_dim=d["E"])
        pd = O.random_params(kind, seed=0, **dkw)
        for v in pd.values():
            v.requires_grad_(True)
        g = torch.Generator().manual_seed(1)

        def step():
            masks = (torch.rand(batch, cap_len - 1, d["D"], generator=g) >= 0.5).float() * 2.0
            out = O.decoder_forward(kind, pd, enc, None if kind == O.PURE_ATTENTION else tags, caps, caplens,
                                    dropout_masks=masks)
            loss = O.caption_loss(out[0], out[1], out[2], None if kind == O.PURE_SCN else out[3])
            for v in pd.values():
                v.grad = None
            loss.backward()

    t_start = time.perf_counter()
    times, done_warm = [], 0
    planned = steps
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it < warmup:
            done_warm += 1
        else:
            times.append(dt)
        if it == 0 and budget > 0:
            # bound the whole run: keep at least 1 warm-up + 2 timed steps
            fit = int(budget / max(dt, 1e-9))
            if fit < warmup + steps:
                warmup = min(warmup, max(1, fit // 4))
                steps = max(2, fit - warmup)
        if it + 1 >= warmup + steps:
            break
    ms = 1e3 * sum(times) / len(times)
    return {"kind": how, "value": batch / (ms / 1e3), "unit": "captions/s", "ms_per_step": ms,
            "steps": len(times), "steps_requested": planned, "warmup": done_warm, "batch": batch,
            "wall_s": time.perf_counter() - t_start}


def time_decode(kind, d, n_images, beam):
    from oracle import capdec_oracle as O
    V = d["V"]
    cls = load_reference_class(kind, True)
    if cls is not None:
        how = "reference"
        dec = build_reference(cls, kind, d).eval()
        word_map = {i: i for i in range(V)}
        word_map["<start>"] = V - 2           # utils/token.py: the only two keys `sample` reads
        word_map["<end>"] = V - 1
        del word_map[V - 2], word_map[V - 1]  # keep len(word_map) == V (attention_scn.py:174)

        def one(enc, tags):
            try:
                if kind == "pure_attention":
                    dec.sample(beam, word_map, enc)
                else:
                    dec.sample(beam, word_map, enc, tags)
            except ValueError:
                pass        # no beam emitted <end> within 51 steps: the reference's own error
    else:
        how = "port"
        dkw = dict(attention_dim=d["A"], embed_dim=d["M"], decoder_dim=d["D"], factored_dim=d["F"],
                   semantic_dim=d["S"], vocab_size=V, encoder_dim=d["E"])
        pd = O.random_params(kind, seed=0, **dkw)

        def one(enc, tags):
            O.beam_search(kind, pd, enc, None if kind == O.PURE_ATTENTION else tags, beam, V - 2, V - 1)

    inputs = [O.synthetic_batch(1, V, seed=100 + i)[:2] for i in range(n_images + 1)]
    with torch.no_grad():
        one(*inputs[0])                         # warm-up image
        t0 = time.perf_counter()
        for enc, tags in inputs[1:]:
            one(enc, tags)
        dt = time.perf_counter() - t0
    return {"kind": how, "value": n_images / dt, "unit": "captions/s", "ms_per_step": 1e3 * dt,
            "steps": 1, "steps_requested": 1, "warmup": 1, "batch": n_images, "wall_s": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", required=True, choices=sorted(CLS))
    ap.add_argument("--mode", default="train", choices=["train", "decode"])
    ap.add_argument("--dims", required=True, help="json dict with A,M,D,F,S,V,E")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--budget", type=float, default=0.0, help="seconds; 0 = run all steps")
    ap.add_argument("--beam", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    n_thr = a.threads or (os.cpu_count() or 1)
    torch.set_num_threads(n_thr)
    d = json.loads(a.dims)
    if a.mode == "train":
        r = time_train(a.kind, d, a.batch, a.steps, a.warmup, a.budget)
        what = ("%d timed steps (+%d warm-up) x %d captions (length 51, T=50, dims %s) of %s, train() mode with "
                "dropout 0.5, fwd + loss glue + autograd bwd, torch %s CPU fp32"
                % (r["steps"], r["warmup"], r["batch"], json.dumps(d, sort_keys=True),
                   "the unmodified reference modules (oracle/_ref)" if r["kind"] == "reference"
                   else "the oracle port of the reference", torch.__version__))
    else:
        r = time_decode(a.kind, d, a.batch, a.beam)
        what = ("%d images one after the other (+1 warm-up), beam=%d, 51 steps each (no beam terminates with "
                "random-init weights), %s, torch %s CPU fp32"
                % (r["batch"], a.beam, "the reference's own sample() from oracle/_ref ('/' -> '//')"
                   if r["kind"] == "reference" else "the oracle port of sample()", torch.__version__))
    r["cores"] = torch.get_num_threads()
    r["sample"] = "%s; cpu=%s, os.cpu_count=%s, torch threads=%d" % (what, cpu_model(), os.cpu_count(),
                                                                     torch.get_num_threads())
    print(json.dumps(r))
    return 0


if __name__ == "__main__":
    sys.exit(main())
