#!/usr/bin/env python
"""bench.py -- captions/sec of the caption-decoder hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ...] [--precision bf16|fp32]
    python bench.py --impl reference ...      # the reference algorithm's CPU port on the host cores

A "step" is one pass of the hot path over one batch of synthetic input:
  * train workloads: decoder forward + loss glue + reverse-time backward (+ NCCL all-reduce of the
    decoder gradients when N > 1).  Default: BASELINE.json config 3 at its per-GPU shape --
    attention_scn, bf16, 32 captions per GPU (global batch 32*N, 256 at N = 8), caption length 51
    (T = 50 decode steps), vocab 10k, dims 512, 14x14x2048 features, 1000 tags.
  * decode workload (config 4): beam=3 search, <= 51 steps, images sharded over the ranks.
`value` times the step with its inputs already resident in HBM; `e2e` times the same step through
the public module API starting from pinned HOST buffers: every timed step issues one batch of H2D
copies (features / tags / captions) on a copy stream -- double buffered, so the copy of the next
batch overlaps the current step's compute -- and reads the loss back to the host.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "indonesian-image-captioning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, dims, per-GPU batch, mode)
    "attention_scn_train": ("attention_scn", dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048), 32, "train"),
    "pure_attention_train": ("pure_attention", dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048), 32, "train"),
    "pure_scn_train": ("pure_scn", dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048), 32, "train"),
    "attention_scn_train_scaled": ("attention_scn", dict(A=512, M=512, D=1024, F=1024, S=1000, V=30000, E=2048), 128, "train"),
    "attention_scn_decode": ("attention_scn", dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048), 625, "decode"),
}
CAP_LEN = 51          # caption length incl. <start>/<end> -> T = 50 decode steps
MAX_LEN = 52


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# --------------------------------------------------------------------------- CPU reference arm
def cpu_train_sample(kind, dims, sample_b, steps, warmup):
    """Time the oracle port (the reference algorithm as written, torch CPU fp32, autograd backward)
    on a bounded sample: `sample_b` captions of the same shape per step."""
    from oracle import capdec_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dkw = dict(attention_dim=dims["A"], embed_dim=dims["M"], decoder_dim=dims["D"], factored_dim=dims["F"],
               semantic_dim=dims["S"], vocab_size=dims["V"], encoder_dim=dims["E"])
    params = O.random_params(kind, seed=0, **dkw)
    for v in params.values():
        v.requires_grad_(True)
    enc, tags, caps, caplens = O.synthetic_batch(sample_b, dims["V"], seed=1234, lengths=[CAP_LEN] * sample_b)
    g = torch.Generator().manual_seed(1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        masks = (torch.rand(sample_b, CAP_LEN - 1, dims["D"], generator=g) >= 0.5).float() * 2.0   # dropout 0.5
        out = O.decoder_forward(kind, params, enc, None if kind == O.PURE_ATTENTION else tags, caps, caplens,
                                dropout_masks=masks)
        alphas = None if kind == O.PURE_SCN else out[3]
        loss = O.caption_loss(out[0], out[1], out[2], alphas)
        for v in params.values():
            v.grad = None
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return {"value": sample_b / (ms / 1e3), "unit": "captions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps x %d captions (T=50, full dims) of the oracle port, fwd+loss+bwd, torch CPU fp32; "
                      "cpu=%s, os.cpu_count=%s" % (steps, sample_b, cpu_model(), os.cpu_count()),
            "ms_per_step": ms}


def cpu_decode_sample(kind, dims, n_images, beam):
    from oracle import capdec_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dkw = dict(attention_dim=dims["A"], embed_dim=dims["M"], decoder_dim=dims["D"], factored_dim=dims["F"],
               semantic_dim=dims["S"], vocab_size=dims["V"], encoder_dim=dims["E"])
    params = O.random_params(kind, seed=0, **dkw)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_images):
            enc, tags, _, _ = O.synthetic_batch(1, dims["V"], seed=100 + i)
            O.beam_search(kind, params, enc, None if kind == O.PURE_ATTENTION else tags, beam, dims["V"] - 2,
                          dims["V"] - 1)
    dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": "captions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d images, beam=%d, 51 steps (no beam terminates with random weights) of the oracle "
                      "port; cpu=%s, os.cpu_count=%s" % (n_images, beam, cpu_model(), os.cpu_count()),
            "ms_per_step": 1e3 * dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, dims, per_gpu_b, mode = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if mode == "train":
        # bounded sample so that K+W steps finish within minutes: 4 captions/step of the same shape
        steps_b = min(steps, 6)
        cb = cpu_train_sample(kind, dims, 4, steps_b, min(warmup, 1))
        metric = "captions/sec (train fwd+loss+bwd, %s)" % kind          # the B200 arm's metric string
    else:
        cb = cpu_decode_sample(kind, dims, min(8, max(2, steps)), 3)
        metric = "captions/sec (beam=3 decode, %s)" % kind
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "captions/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, kind, dims, per_gpu_b, mode),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, kind, dims, per_gpu_b, mode):
    return {"workload": args.workload, "decoder": kind, "mode": mode, "per_gpu_batch": per_gpu_b,
            "global_batch": per_gpu_b * args.gpus, "caption_len": CAP_LEN, "decode_steps": CAP_LEN - 1,
            "vocab": dims["V"], "dims": dims, "features": "14x14x%d" % dims["E"], "parallelism": "dp%d" % args.gpus,
            "precision": args.precision, "dropout": 0.5 if mode == "train" else 0.0,
            "cuda_graphs": not getattr(args, "no_graphs", False),
            "l2": "flushed between timed steps by writing a 256 MiB buffer (outside the per-step event pairs)"}


# --------------------------------------------------------------------------- B200 arm
def make_decoder(kind, dims):
    if kind == "attention_scn":
        from models.decoders.attention_scn import AttentionSCN
        return AttentionSCN(dims["A"], dims["M"], dims["D"], dims["F"], dims["S"], dims["V"], encoder_dim=dims["E"])
    if kind == "pure_scn":
        from models.decoders.pure_scn import PureSCN
        return PureSCN(dims["M"], dims["D"], dims["F"], dims["S"], dims["V"], encoder_dim=dims["E"])
    from models.decoders.pure_attention import PureAttention
    return PureAttention(dims["A"], dims["M"], dims["D"], dims["V"], encoder_dim=dims["E"])


def run_b200(args):
    import capdec
    from capdec import _lib
    from capdec import parallel as cpar
    from capdec import functional as CFm
    from oracle import capdec_oracle as O      # synthetic-input generator + CPU baseline only

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    capdec.set_precision(args.precision)
    capdec.set_graphs(not args.no_graphs)      # public switch: replay the step from CUDA graphs
    kind, dims, per_gpu_b, mode = WORKLOADS[args.workload]
    if args.batch:
        per_gpu_b = args.batch
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    peaks = load_peaks()

    torch.manual_seed(0)
    dec = make_decoder(kind, dims).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    result = {}
    if mode == "train":
        dec.train()
        enc_h, tags_h, caps_h, caplens_h = O.synthetic_batch(per_gpu_b, dims["V"], seed=1234 + rank,
                                                             lengths=[CAP_LEN] * per_gpu_b)
        pinned = [t.pin_memory() for t in (enc_h, tags_h, caps_h, caplens_h)]
        resident = [t.to(dev) for t in pinned]
        n_tok_global = per_gpu_b * (CAP_LEN - 1) * world
        reducer = cpar.GradReducer(dec, dist) if dist is not None else None

        def step(inputs):
            enc, tags, caps, caplens = inputs
            if kind == "pure_attention":
                out = dec(enc, caps, caplens)
            else:
                out = dec(enc, tags, caps, caplens)
            scores, caps_sorted, dl = out[0], out[1], out[2]
            alphas = None if kind == "pure_scn" else out[3]
            loss, parts = dec.loss(scores, caps_sorted, dl, alphas, alpha_c=1.0 / world, n_tokens=n_tok_global)
            for p in dec.parameters():
                p.grad = None
            loss.backward()
            if reducer is not None:
                reducer.allreduce(getattr(scores, "_capdec_meta", None))
            return loss

        def timed(fn, n):
            evs = []
            for _ in range(n):
                flush.fill_(1)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                evs.append((s, e))
            torch.cuda.synchronize()
            return [s.elapsed_time(e) for s, e in evs]

        def step_resident():
            step(resident)

        # e2e input pipeline: every step copies ONE batch host -> device from pinned memory on a copy stream
        # (double buffered: the copy of the next batch overlaps this step's compute, the way a DataLoader
        # with pin_memory + non_blocking feeds the reference's training loop) and reads the loss back
        copy_stream = torch.cuda.Stream(device=dev)
        pending = []

        def issue_copy():
            with torch.cuda.stream(copy_stream):
                bufs = [t.to(dev, non_blocking=True) for t in pinned]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            pending.append((bufs, ev))

        def step_e2e():
            if not pending:
                issue_copy()
            bufs, ev = pending.pop(0)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for b in bufs:
                b.record_stream(cur)
            loss = step(bufs)
            issue_copy()              # next step's inputs start moving while this step computes (host cost hidden too)
            return loss.item()        # D2H read of the step's result

        for _ in range(warmup):
            step_resident()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = CFm.launch_count()
        times = timed(step_resident, steps)
        launches = CFm.launch_count() - l0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        total_ms = sum(times)
        for _ in range(2):
            step_e2e()
        barrier()
        e2e_ms = sum(timed(step_e2e, steps))
        barrier()
        h2d = sum(t.numel() * t.element_size() for t in pinned)
        result.update(total_ms=total_ms, e2e_ms=e2e_ms, launches=launches, clocks=clocks, h2d=h2d, d2h=4,
                      units_per_step=per_gpu_b)
        if rank == 0:
            result["optimizer_step"] = measure_optimizer(dec, peaks)
        metric = "captions/sec (train fwd+loss+bwd, %s)" % kind
    else:
        dec.eval()
        n_img = per_gpu_b
        g = torch.Generator().manual_seed(4321 + rank)
        enc_h = torch.randn(n_img, 14, 14, dims["E"], generator=g).relu_().pin_memory()
        tags_h = torch.rand(n_img, dims["S"], generator=g).pin_memory()
        enc_d, tags_d = enc_h.to(dev), tags_h.to(dev)
        V = dims["V"]

        def decode(enc, tags):
            with torch.no_grad():
                return dec.sample_batch(3, V - 2, V - 1, enc, None if kind == "pure_attention" else tags,
                                        max_steps=50, want_alphas=False)

        def timed(fn, n):
            evs = []
            for _ in range(n):
                flush.fill_(1)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                evs.append((s, e))
            torch.cuda.synchronize()
            return [s.elapsed_time(e) for s, e in evs]

        def step_resident():
            decode(enc_d, tags_d)

        copy_stream = torch.cuda.Stream(device=dev)
        pending = []

        def issue_copy():
            with torch.cuda.stream(copy_stream):
                bufs = [enc_h.to(dev, non_blocking=True), tags_h.to(dev, non_blocking=True)]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            pending.append((bufs, ev))

        def step_e2e():
            if not pending:
                issue_copy()
            bufs, ev = pending.pop(0)
            issue_copy()                                   # the next shard of images starts moving now
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for b in bufs:
                b.record_stream(cur)
            r = decode(bufs[0], bufs[1])
            return r["seq"].cpu(), r["len"].cpu()

        for _ in range(warmup):
            step_resident()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = CFm.launch_count()
        times = timed(step_resident, steps)
        launches = CFm.launch_count() - l0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        total_ms = sum(times)
        step_e2e()
        barrier()
        e2e_ms = sum(timed(step_e2e, steps))
        barrier()
        result.update(total_ms=total_ms, e2e_ms=e2e_ms, launches=launches, clocks=clocks,
                      h2d=enc_h.numel() * 4 + tags_h.numel() * 4, d2h=n_img * 53 * 4, units_per_step=n_img)
        metric = "captions/sec (beam=3 decode, %s)" % kind

    # max over ranks of the timed totals
    t = torch.tensor([result["total_ms"], result["e2e_ms"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    units = result["units_per_step"] * world * steps
    value = units / (total_ms / 1e3)
    e2e_value = units / (e2e_ms / 1e3)

    roof = cpu_b = None
    if rank == 0:
        roof = None
        if mode == "train" and kind != "pure_attention" and args.precision == "bf16":
            roof = measure_recurrence_roofline(lib, dec, kind, dims, per_gpu_b, resident, peaks)
        pair = measure_roofline(dev, kind, dims, per_gpu_b if mode == "train" else 3 * 64, peaks, args.precision)
        if roof is None:
            roof = pair
        elif pair is not None:
            roof["attention_step_kernels"] = {k: pair[k] for k in ("kernel", "achieved", "frac", "us_per_launch",
                                                                   "algorithmic_bytes_per_launch", "traffic", "note")}
        if not args.no_cpu_baseline:
            if mode == "train":
                cpu_b = cpu_train_sample(kind, dims, 4, 3, 1)
            else:
                cpu_b = cpu_decode_sample(kind, dims, 4, 3)
            cpu_b = {k: cpu_b[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "captions/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": workload_config(args, kind, dims, per_gpu_b, mode),
                "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": result["h2d"],
                        "d2h_bytes_per_step": result["d2h"], "ms_per_step": e2e_ms / steps},
                "gpu_launches": int(result["launches"]), "clocks": result["clocks"], "roofline": roof,
                "cpu_baseline": cpu_b, "peaks": peaks}
        if result.get("optimizer_step"):
            line["optimizer_step"] = result["optimizer_step"]       # reported separately
        print(json.dumps(line))
    return 0


def load_traffic(kernel, rows=32):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), or None when there is no capture of it.  The captures were
    taken at the config-3 shape (32 rows per GPU): other shapes get None."""
    if rows != 32:
        return None
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measure_recurrence_roofline(lib, dec, kind, dims, rows, inputs, peaks):
    """Dominant kernels of the training step: recur_bwd_kernel and recur_fwd_kernel (csrc/recur.cu), ONE
    cooperative launch each for all T decode steps.  Algorithmic bytes per launch = T * rows * P * (A + E) * 2
    in both directions (every caption-step streams its att1 and enc rows, 1.004 MB in bf16; the
    backward re-reads both and keeps no per-step dAtt1 read-modify-write); for pure_scn (no feature stream)
    the recurrent weights touched per step instead.  Duration: CUDA events recorded by the library on the
    launching stream around each kernel (capdec_recur_timing), eager launches of the full step, mean of 5
    after 2 warm-ups.  The roofline entry is the slower of the two; the other one rides along."""
    import capdec
    was = capdec.graphs_enabled()
    capdec.set_graphs(False)
    lib.capdec_recur_timing(1)
    enc, tags, caps, caplens = inputs
    ms = {0: [], 1: []}
    try:
        for it in range(7):
            out = dec(enc, tags, caps, caplens)      # grad mode: the kernel also saves awe for the backward
            alphas = None if kind == "pure_scn" else out[3]
            loss, _ = dec.loss(out[0], out[1], out[2], alphas)
            for p in dec.parameters():
                p.grad = None
            loss.backward()
            for which in (0, 1):
                t = float(lib.capdec_recur_last_ms(which))
                if it >= 2 and t > 0:
                    ms[which].append(t)
    finally:
        lib.capdec_recur_timing(0)
        capdec.set_graphs(was)
    if not ms[0] or not ms[1]:
        return None
    P, E, A, D, F, T = 196, dims["E"], dims["A"], dims["D"], dims["F"], CAP_LEN - 1
    if kind == "attention_scn":
        bytes_alg = T * rows * P * (A + E) * 2
        what = "T*rows*P*(A+E)*2 B of attention features"
    else:
        bytes_alg = T * (4 * F * D + 4 * D * 2 * F) * 2
        what = ("T * (W_ha + [W_ic|W_hc]) bf16 recurrent weights (resident in shared memory, so the HBM "
                "figure is an upper bound of need)")

    def entry(which, name):
        t_ms = sum(ms[which]) / len(ms[which])
        achieved = bytes_alg / (t_ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic(name, rows), "us_per_launch": 1e3 * t_ms,
                "rows": rows, "steps_per_launch": T, "algorithmic_bytes_per_launch": bytes_alg,
                "peak_source": peaks["source"]}

    fwd, bwd = entry(0, "recur_fwd_kernel"), entry(1, "recur_bwd_kernel")
    main, other = (bwd, fwd) if bwd["us_per_launch"] >= fwd["us_per_launch"] else (fwd, bwd)
    nbar = 6 if kind == "attention_scn" else 3
    main["note"] = ("one cooperative launch = all %d decode steps, two independent 16-row groups per CTA; algorithmic "
                    "bytes = %s; at %d rows the kernel is bound by its %d grid barriers per step (>= 1.3 us each: "
                    "tools/barrier_bench.cu) and dependent L2 round trips, not by HBM: the features stay L2-resident "
                    "across steps, so the measured DRAM traffic is far BELOW the algorithmic stream -- see DESIGN.md"
                    % (T, what, rows, nbar))
    main["other_direction"] = other
    return main


def measure_optimizer(dec, peaks):
    """The reference's per-iteration optimizer work (trains/attention_scn.py:244-252: clip_gradient + Adam.step)
    on the decoder parameters, outside the headline metric: the fused kernel of capdec.optim.ClipAdam next to
    the stock torch ops.  CUDA-event timed, mean of 10 after 3 warm-ups."""
    from capdec.optim import ClipAdam
    params = [p for p in dec.parameters() if p.requires_grad and p.grad is not None]
    if not params:
        return None
    n = sum(p.numel() for p in params)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / 10

    saved = [p.detach().clone() for p in params]
    fused = ClipAdam(params, lr=4e-4, grad_clip=5.0)
    t_fused = timed(fused.step)
    stock = torch.optim.Adam(params, lr=4e-4)

    def stock_step():
        for p in params:
            p.grad.data.clamp_(-5.0, 5.0)
        stock.step()
    t_stock = timed(stock_step)
    with torch.no_grad():
        for p, q in zip(params, saved):
            p.copy_(q)
    nbytes = n * 4 * 8                 # read p, g, m, v ; write p, m, v, g
    return {"parameters": n, "fused_clip_adam_ms": t_fused, "torch_clamp_plus_adam_ms": t_stock,
            "fused_gbs": nbytes / (t_fused * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (t_fused * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "algorithmic_bytes": nbytes}


def measure_roofline(dev, kind, dims, rows, peaks, precision):
    """The attention step as two stand-alone kernels (csrc/attention.cu attn_scores_kernel + attn_wsum_kernel),
    the per-step path of pure_attention, fp32 mode, beam search and of shapes the persistent kernel does not
    cover.  Algorithmic bytes per launch pair = rows * P * (A + E) * sizeof(feature) (1.004 MB
    per caption-step in bf16), divided by its average duration measured with CUDA events on the launching
    stream."""
    if kind == "pure_scn":
        return None
    from capdec import functional as CF
    P, E, A = 196, dims["E"], dims["A"]
    ft = torch.bfloat16 if precision == "bf16" else torch.float32
    g = torch.Generator(device=dev).manual_seed(1)
    att1 = torch.randn(rows, P, A, device=dev, generator=g).to(ft)
    enc = torch.randn(rows, P, E, device=dev, generator=g).relu_().to(ft)
    g1 = torch.randn(rows, A + E, device=dev, generator=g)
    w_f = torch.randn(A, device=dev, generator=g) * 0.05
    b_f = torch.zeros(1, device=dev)
    def launch():
        CF.attention_step(att1, enc, g1, A, w_f, b_f, precision=precision, want_awe=True)

    # n launches captured into a CUDA graph and replayed: device time per launch, no Python overhead
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    n = 20
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n):
            launch()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    graph.replay()
    e.record()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / n
    bytes_alg = rows * P * (A + E) * (2 if precision == "bf16" else 4)
    achieved = bytes_alg / (us * 1e-6) / 1e9
    return {"kernel": "attn_scores_kernel+attn_wsum_kernel", "bound": "hbm", "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic("attn_scores_kernel+attn_wsum_kernel", rows),
            "us_per_launch": us, "rows": rows,
            "algorithmic_bytes_per_launch": bytes_alg, "peak_source": peaks["source"],
            "note": "features of one step (%.1f MB) are L2-resident across back-to-back launches, as in the "
                    "decode loop; %d launch pairs replayed from a CUDA graph, CUDA-event timed" % (bytes_alg / 1e6, n)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="attention_scn_train", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
