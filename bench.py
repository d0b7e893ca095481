#!/usr/bin/env python
"""bench.py -- captions/sec of the caption-decoder hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ...] [--precision bf16|fp32] [--ragged]
    python bench.py --impl reference ...      # the reference's own CPU code on the host cores (oracle/_ref)

A "step" is one pass of the hot path over one batch of synthetic input:
  * train workloads: decoder forward + loss glue + reverse-time backward (+ the bucketed, overlapped NCCL
    all-reduce of the decoder gradients when N > 1).  Default: BASELINE.json config 3 at its per-GPU shape --
    attention_scn, bf16, 32 captions per GPU (global batch 32*N, 256 at N = 8), caption length 51
    (T = 50 decode steps), vocab 10k, dims 512, 14x14x2048 features, 1000 tags.  `--ragged`: tie-free caption
    lengths 3..52 instead (what a real batch looks like; same CUDA graphs, the kernels read the lengths on the
    device).
  * decode workload (config 4): beam=3 search, <= 51 steps, images sharded over the ranks.
`value` times the step with its inputs already resident in HBM; `e2e` times the same step through the public
module API starting from pinned HOST buffers: every timed step issues one batch of H2D copies (features / tags /
captions) on a copy stream -- double buffered, so the copy of the next batch overlaps the current step's compute
-- and reads the loss back to the host.  `secondary` carries short runs of the other BASELINE configs (and of the
ragged / eager variants of the headline) at the same N.  Prints ONE JSON line (rank 0).
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

# a process-wide NCCL_DEBUG=VERSION would put NCCL's banner on stdout next to the ONE JSON line of the contract
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

DIMS512 = dict(A=512, M=512, D=512, F=512, S=1000, V=10000, E=2048)
WORKLOADS = {
    # name: (kind, dims, per-GPU batch, mode)
    "attention_scn_train": ("attention_scn", DIMS512, 32, "train"),                 # BASELINE config 3 (per GPU)
    "pure_attention_train": ("pure_attention", DIMS512, 32, "train"),               # config 2
    "pure_scn_train": ("pure_scn", DIMS512, 32, "train"),                           # config 1
    "attention_scn_train_scaled": ("attention_scn", dict(A=512, M=512, D=1024, F=1024, S=1000, V=30000, E=2048),
                                   128, "train"),                                   # config 5
    "attention_scn_decode": ("attention_scn", DIMS512, 625, "decode"),              # config 4 (5000 images / 8)
}
# short runs reported next to the headline: (label, workload, ragged, graphs)
SECONDARY = [
    ("attention_scn_train_ragged", "attention_scn_train", True, True),
    ("attention_scn_train_ragged_eager", "attention_scn_train", True, False),
    ("pure_scn_train", "pure_scn_train", False, True),
    ("pure_attention_train", "pure_attention_train", False, True),
    ("attention_scn_train_scaled", "attention_scn_train_scaled", False, True),
    ("attention_scn_decode", "attention_scn_decode", False, True),
]
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
    """SM clock / throttle-reason samples DURING the timed region: an NVML polling thread (one sample every ~2 ms;
    `nvidia-smi -lms` cannot go below ~100 ms, longer than the whole timed region of a 20-step run).  Samples carry
    host timestamps; stop(t0, t1) keeps those inside the timed window."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:       # noqa: BLE001
            self.err = "nvml unavailable: %s" % e
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:    # noqa: BLE001
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, why))
            except Exception as e:   # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(0.002)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "no sampler"]}
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
                 "hw_power_brake_slowdown": nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown}
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = sorted(r[1] for r in inside)
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in inside))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "samples": len(sm),
                "samples_total": len(self.rows), "reasons": reasons,
                "how": "NVML polled every ~2 ms by a host thread; only samples inside the timed region count"}


# --------------------------------------------------------------------------- CPU reference arm
def run_ref_cpu(kind, dims, mode, batch, steps, warmup, budget):
    """The reference's own decoder code on the host cores, in its own process (oracle/ref_cpu.py: the unmodified
    modules from oracle/_ref when the recipe oracle/make_ref.py has run, else the oracle port)."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_cpu.py"), "--kind", kind, "--mode", mode,
           "--dims", json.dumps(dims), "--batch", str(batch), "--steps", str(steps), "--warmup", str(warmup),
           "--budget", str(budget)]
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""          # the reference allocates on `device`; keep it on the CPU
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):      # torchrun sets OMP_NUM_THREADS=1 for its workers
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("oracle/ref_cpu.py failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, dims, per_gpu_b, mode = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if mode == "train":
        # the SAME per-GPU batch as the B200 arm; the number of steps is cut only if K+W would not end within minutes
        cb = run_ref_cpu(kind, dims, "train", per_gpu_b, steps, warmup, 240.0)
        metric = "captions/sec (train fwd+loss+bwd, %s)" % kind          # the B200 arm's metric string
        batch = cb["batch"]
    else:
        cb = run_ref_cpu(kind, dims, "decode", min(24, max(4, steps)), 1, 1, 0)
        metric = "captions/sec (beam=3 decode, %s)" % kind
        batch = cb["batch"]
    cfg = workload_config(args, kind, dims, batch, mode, False, False)
    cfg["note"] = ("CPU arm: %s; `steps`/`warmup` of this line are the ones that RAN (requested %d/%d)"
                   % (cb["sample"], steps, warmup))
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "captions/s",
            "n_gpus": args.gpus, "steps": cb["steps"], "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, kind, dims, per_gpu_b, mode, ragged, graphs, workload=None):
    cfg = {"workload": workload or args.workload, "decoder": kind, "mode": mode, "per_gpu_batch": per_gpu_b,
           "global_batch": per_gpu_b * args.gpus, "vocab": dims["V"], "dims": dims, "features": "14x14x%d" % dims["E"],
           "parallelism": "dp%d" % args.gpus, "precision": args.precision,
           "dropout": 0.5 if mode == "train" else 0.0, "cuda_graphs": bool(graphs),
           "l2": "flushed between timed steps by writing a 256 MiB buffer (outside the per-step event pairs)"}
    if mode == "train":
        cfg.update({"caption_len": "tie-free 3..52 (ragged)" if ragged else CAP_LEN,
                    "decode_steps": "max over the batch" if ragged else CAP_LEN - 1})
    else:
        cfg.update({"beam": 3, "decode_steps": 51})
    return cfg


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


class Env:
    pass


def ragged_lengths(O, B, rank):
    """Tie-free caption lengths 3..52, a different draw per rank; batches above 50 rows tile it."""
    out = []
    i = 0
    while len(out) < B:
        out += O.tie_free_lengths(min(B - len(out), 50), seed=7 + 101 * rank + i)
        i += 1
    return out


def timed(env, fn, n):
    evs = []
    for _ in range(n):
        env.flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in evs]


def run_workload(env, name, steps, warmup, ragged=False, graphs=True, full=False):
    """Time one workload; returns a dict with per-rank totals (ms) and bookkeeping.  full: also the e2e leg, the
    launch count, the clock samples and (train) the optimizer microbenchmark."""
    import capdec
    from capdec import parallel as cpar
    from capdec import functional as CFm
    O = env.O
    kind, dims, per_gpu_b, mode = WORKLOADS[name]
    if env.args.batch and full:
        per_gpu_b = env.args.batch
    dev, dist, rank, world = env.dev, env.dist, env.rank, env.world
    capdec.set_graphs(False)              # drops the static buffers of the previous workload
    capdec.set_graphs(graphs)
    torch.manual_seed(0)
    dec = make_decoder(kind, dims).to(dev)
    res = {"kind": kind, "dims": dims, "mode": mode, "per_gpu_b": per_gpu_b}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    reducer = None
    if mode == "train":
        dec.train()
        lengths = ragged_lengths(O, per_gpu_b, rank) if ragged else [CAP_LEN] * per_gpu_b
        enc_h, tags_h, caps_h, caplens_h = O.synthetic_batch(per_gpu_b, dims["V"], seed=1234 + rank, lengths=lengths)
        pinned = [t.pin_memory() for t in (enc_h, tags_h, caps_h, caplens_h)]
        resident = [t.to(dev) for t in pinned]
        if ragged:
            n_tok_global = sum(sum(l - 1 for l in ragged_lengths(O, per_gpu_b, r)) for r in range(world))
        else:
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

        def step_resident():
            step(resident)

        h2d = sum(t.numel() * t.element_size() for t in pinned)
        d2h = 4

        def copy_in():
            return [t.to(dev, non_blocking=True) for t in pinned]

        def run_e2e(bufs):
            return step(bufs).item()          # D2H read of the step's result
        res["tokens_per_step"] = sum(l - 1 for l in lengths)
        metric = "captions/sec (train fwd+loss+bwd, %s)" % kind
    else:
        dec.eval()
        n_img = per_gpu_b
        g = torch.Generator().manual_seed(4321 + rank)
        enc_h = torch.randn(n_img, 14, 14, dims["E"], generator=g).relu_().pin_memory()
        tags_h = torch.rand(n_img, dims["S"], generator=g).pin_memory()
        resident = [enc_h.to(dev), tags_h.to(dev)]
        V = dims["V"]

        def decode(enc, tags):
            with torch.no_grad():
                return dec.sample_batch(3, V - 2, V - 1, enc, None if kind == "pure_attention" else tags,
                                        max_steps=50, want_alphas=False)

        def step_resident():
            decode(*resident)

        h2d = enc_h.numel() * 4 + tags_h.numel() * 4
        d2h = n_img * 53 * 4

        def copy_in():
            return [enc_h.to(dev, non_blocking=True), tags_h.to(dev, non_blocking=True)]

        def run_e2e(bufs):
            r = decode(bufs[0], bufs[1])
            return r["seq"].cpu(), r["len"].cpu()
        metric = "captions/sec (beam=3 decode, %s)" % kind

    # e2e input pipeline: every step copies ONE batch host -> device from pinned memory on a copy stream (double
    # buffered: the copy of the next batch overlaps this step's compute, the way a DataLoader with pin_memory +
    # non_blocking feeds the reference's training loop) and reads the step's result back
    copy_stream = torch.cuda.Stream(device=dev)
    pending = []

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            bufs = copy_in()
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
        if mode == "decode":
            issue_copy()                  # the next shard of images starts moving now
            return run_e2e(bufs)
        # train: the step is launched first, then the next batch's copy (the host cost of the copy is hidden too)
        enc_loss = step(bufs)
        issue_copy()
        return enc_loss.item()

    sampler = ClockSampler(env.local_rank) if (full and rank == 0) else None
    if sampler is not None:
        sampler.start()                   # before the warm-up: the polling thread is up when the timed region starts
    for _ in range(warmup):
        step_resident()
    barrier()
    l0 = CFm.launch_count()
    t0 = time.perf_counter()
    times = timed(env, step_resident, steps)
    t1 = time.perf_counter()
    res["launches"] = CFm.launch_count() - l0
    barrier()
    res["clocks"] = sampler.stop(t0, t1) if sampler is not None else None
    res["total_ms"] = sum(times)
    res["e2e_ms"] = None
    if full:
        for _ in range(2):
            step_e2e()
        barrier()
        res["e2e_ms"] = sum(timed(env, step_e2e, steps))
        barrier()
        if mode == "train" and rank == 0:
            res["optimizer_step"] = measure_optimizer(dec, env.peaks)
    res.update(h2d=h2d, d2h=d2h, units_per_step=per_gpu_b, metric=metric, dec=dec, resident=resident)
    if reducer is not None:
        reducer.close()
    return res


def reduce_max(env, vals):
    t = torch.tensor([v if v is not None else 0.0 for v in vals], dtype=torch.float64, device=env.dev)
    if env.dist is not None:
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
    return t.tolist()


def run_b200(args):
    import capdec
    from capdec import _lib
    from oracle import capdec_oracle as O      # synthetic-input generator only (the CPU legs run in their own process)

    env = Env()
    env.args, env.O = args, O
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference)"
    torch.cuda.set_device(env.local_rank)
    env.dev = dev = torch.device("cuda", env.local_rank)
    env.dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        env.dist = dist
    lib = _lib.load()
    capdec.set_precision(args.precision)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    env.peaks = peaks = load_peaks()
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    kind, dims, per_gpu_b, mode = WORKLOADS[args.workload]
    graphs = not args.no_graphs
    res = run_workload(env, args.workload, steps, warmup, ragged=args.ragged, graphs=graphs, full=True)
    per_gpu_b = res["per_gpu_b"]
    total_ms, e2e_ms = reduce_max(env, [res["total_ms"], res["e2e_ms"]])
    units = res["units_per_step"] * world * steps
    value = units / (total_ms / 1e3)
    e2e_value = units / (e2e_ms / 1e3)

    roof = cpu_b = None
    if rank == 0:
        if mode == "train" and args.precision == "bf16" and not args.ragged:
            roof = measure_recurrence_roofline(lib, res["dec"], kind, dims, per_gpu_b, res["resident"], peaks)
        pair = measure_roofline(dev, kind, dims, per_gpu_b if mode == "train" else 3 * 64, peaks, args.precision)
        if roof is None:
            roof = pair
        elif pair is not None:
            roof["attention_step_kernels"] = {k: pair[k] for k in ("kernel", "achieved", "frac", "us_per_launch",
                                                                   "algorithmic_bytes_per_launch", "traffic", "note")}
    res.pop("dec")
    res.pop("resident")
    torch.cuda.empty_cache()

    secondary = {}
    if not args.no_secondary:
        s_steps, s_warm = min(steps, 8), 3
        for label, wl, ragged, gr in SECONDARY:
            if (wl, ragged, gr) == (args.workload, args.ragged, graphs):
                continue
            try:
                r = run_workload(env, wl, s_steps, s_warm, ragged=ragged, graphs=gr, full=False)
                (tot,) = reduce_max(env, [r["total_ms"]])
                secondary[label] = {
                    "metric": r["metric"], "value": r["units_per_step"] * world * s_steps / (tot / 1e3),
                    "unit": "captions/s", "ms_per_step": tot / s_steps, "steps": s_steps, "warmup": s_warm,
                    "n_gpus": world, "gpu_launches": int(r["launches"]),
                    "config": workload_config(args, r["kind"], r["dims"], r["per_gpu_b"], r["mode"], ragged, gr, wl)}
                if r["mode"] == "train":
                    secondary[label]["tokens_per_step_rank0"] = r["tokens_per_step"]
            except Exception as e:      # noqa: BLE001  (a secondary line must never take the headline down)
                secondary[label] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            r = None
            torch.cuda.empty_cache()

    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # rank 0 at N = 1 only (under torchrun the other ranks would spin in a barrier while this runs)
        try:
            if mode == "train":
                cb = run_ref_cpu(kind, dims, "train", per_gpu_b, 3, 1, 25.0)
            else:
                cb = run_ref_cpu(kind, dims, "decode", 8, 1, 1, 0)
            cpu_b = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:          # noqa: BLE001
            cpu_b = {"error": str(e)[:500]}
    if rank == 0:
        line = {"metric": res["metric"], "value": value, "unit": "captions/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": workload_config(args, kind, dims, per_gpu_b, mode, args.ragged, graphs),
                "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"], "ms_per_step": e2e_ms / steps},
                "gpu_launches": int(res["launches"]), "clocks": res["clocks"], "roofline": roof,
                "cpu_baseline": cpu_b, "peaks": peaks, "secondary": secondary}
        if res.get("optimizer_step"):
            line["optimizer_step"] = res["optimizer_step"]       # reported separately
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
    """Dominant kernels of the training step: the two persistent recurrence kernels (csrc/recur.cu), ONE
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
            out = dec(enc, caps, caplens) if kind == "pure_attention" else dec(enc, tags, caps, caplens)
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
    if kind != "pure_scn":
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
                "us_per_decode_step": 1e3 * t_ms / T, "rows": rows, "steps_per_launch": T,
                "algorithmic_bytes_per_launch": bytes_alg, "peak_source": peaks["source"]}

    fwd, bwd = entry(0, "recur_fwd_kernel"), entry(1, "recur_bwd_kernel")
    main, other = (bwd, fwd) if bwd["us_per_launch"] >= fwd["us_per_launch"] else (fwd, bwd)
    main["note"] = ("one cooperative launch = all %d decode steps; algorithmic bytes = %s; at %d rows per GPU the "
                    "features stay L2-resident across steps, so the measured DRAM traffic is far BELOW the algorithmic "
                    "stream and the kernel is bound by the dependent L2 round trips of the recurrence, not by HBM -- see "
                    "DESIGN.md" % (T, what, rows))
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

    def timed_(fn):
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
    t_fused = timed_(fused.step)
    stock = torch.optim.Adam(params, lr=4e-4)

    def stock_step():
        for p in params:
            p.grad.data.clamp_(-5.0, 5.0)
        stock.step()
    t_stock = timed_(stock_step)
    with torch.no_grad():
        for p, q in zip(params, saved):
            p.copy_(q)
    nbytes = n * 4 * 8                 # read p, g, m, v ; write p, m, v, g
    return {"parameters": n, "fused_clip_adam_ms": t_fused, "torch_clamp_plus_adam_ms": t_stock,
            "fused_gbs": nbytes / (t_fused * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (t_fused * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "algorithmic_bytes": nbytes}


def measure_roofline(dev, kind, dims, rows, peaks, precision):
    """The attention step as two stand-alone kernels (csrc/attention.cu attn_scores_kernel + attn_wsum_kernel),
    the per-step path of fp32 mode, beam search and of shapes the persistent kernels do not cover.  Algorithmic
    bytes per launch pair = rows * P * (A + E) * sizeof(feature) (1.004 MB per caption-step in
    bf16), divided by its average duration measured with CUDA events on the launching stream."""
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
    ap.add_argument("--ragged", action="store_true", help="tie-free caption lengths 3..52 instead of 51 everywhere")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other workloads")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
