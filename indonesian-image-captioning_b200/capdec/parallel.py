"""Data-parallel training glue: one process per GPU, torch.distributed (NCCL over NVLink on the
B200 box, gloo in the CPU tests), ONE exchange step per iteration -- the sum all-reduce of the decoder
gradients.  The reference has no multi-GPU path; this is the new work BASELINE
configs 3/5 ask for.

The exchange is BUCKETED AND OVERLAPPED: capdec_backward produces the gradients in
stages (fc | reverse loop -> weight_ia + embedding -> other cell weights + init -> f_beta / decoder_att ->
encoder_att) into one flat buffer laid out in that order; after each stage has been queued, `GradReducer` records an event on the compute
stream and issues the all-reduce of that bucket on a side stream, so it runs while the next stage's GEMMs are
still computing.  Only the last (smallest) bucket's all-reduce is exposed.  `allreduce()` after
`loss.backward()` then merely makes the compute stream wait for the side stream.

Loss scaling contract (see CaptionDecoderBase.loss): every rank divides its cross-entropy sum by
the GLOBAL token count and its alpha regulariser by world_size, so the SUM of the per-rank
gradients equals the gradient of the single-process mean loss over the global batch.
"""
import os

import torch


class GradReducer:
    """early_fc (default: on for NCCL): the first bucket (fc.weight / fc.bias, 19 % of the bytes) is complete BEFORE
    the reverse-time loop.  Its all-reduce goes to a second communicator limited to `SMALL_COMM_CTAS` CTAs, and the
    persistent backward kernel is launched on that many fewer SMs (CAPDEC_RECUR_BWD_CTAS), so the two run side by
    side for the 1.2 ms of the loop instead of the exchange waiting for the loop to end."""
    SMALL_COMM_CTAS = 4

    def __init__(self, module, dist, overlap=True, early_fc=None):
        self.module = module
        self.dist = dist
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._flat = None
        self._side = None
        self._side_small = None
        self._pg_small = None
        self._reduced = None          # flat buffer whose buckets were all-reduced by the hook in this backward
        self.overlap = bool(overlap)
        self.early_fc = False
        self._saved_env = None
        if self.overlap:
            on_cuda = bool(self.params) and self.params[0].is_cuda
            if early_fc is None:
                early_fc = on_cuda and dist.get_backend() == "nccl" and os.environ.get("CAPDEC_EARLY_FC", "1") != "0"
            if early_fc:
                self._make_small_comm()
            from . import functional as CF
            CF.set_grad_bucket_hook(self)

    def _make_small_comm(self):
        """A second NCCL communicator capped at a few CTAs (collective call: every rank constructs its reducer at
        the same point).  Any failure leaves early_fc off."""
        try:
            opts = self.dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = self.SMALL_COMM_CTAS
            opts.config.min_ctas = 1
            self._pg_small = self.dist.new_group(backend="nccl", pg_options=opts)
            sms = torch.cuda.get_device_properties(self.params[0].device).multi_processor_count
            self._saved_env = os.environ.get("CAPDEC_RECUR_BWD_CTAS")
            os.environ["CAPDEC_RECUR_BWD_CTAS"] = str(sms - self.SMALL_COMM_CTAS)
            self.early_fc = True
        except Exception:       # noqa: BLE001
            self._pg_small = None
            self.early_fc = False

    def close(self):
        if self.overlap:
            from . import functional as CF
            CF.set_grad_bucket_hook(None)
            self.overlap = False
        if self.early_fc:
            if self._saved_env is None:
                os.environ.pop("CAPDEC_RECUR_BWD_CTAS", None)
            else:
                os.environ["CAPDEC_RECUR_BWD_CTAS"] = self._saved_env
            self.early_fc = False

    # ------------------------------------------------------------------ overlapped path
    def __call__(self, index, n_buckets, flat_slice):
        """Called by DecoderTrainFn.backward after the launches that fill bucket `index` have been queued."""
        if index < 0:                 # "wait for what has been issued" (gradient accumulation needs the sums now)
            self._join()
            return
        if flat_slice.numel() > 0:
            if flat_slice.is_cuda:
                small = index == 0 and self.early_fc
                if self._side is None:
                    # high-priority streams: the all-reduce kernels get free SM slots ahead of the queued GEMM tiles
                    # of the compute stream (the exchange, not the GEMMs, is what the step ends on)
                    prio = -1 if os.environ.get("CAPDEC_AR_PRIORITY", "1") != "0" else 0
                    self._side = torch.cuda.Stream(device=flat_slice.device, priority=prio)
                    self._side_small = torch.cuda.Stream(device=flat_slice.device, priority=prio)
                side = self._side_small if small else self._side
                ev = torch.cuda.Event()
                ev.record()               # on the compute stream, after this bucket's last kernel
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    if small:
                        self.dist.all_reduce(flat_slice, op=self.dist.ReduceOp.SUM, group=self._pg_small)
                    else:
                        self.dist.all_reduce(flat_slice, op=self.dist.ReduceOp.SUM)
            else:
                self.dist.all_reduce(flat_slice, op=self.dist.ReduceOp.SUM)
        if index == n_buckets - 1:
            self._reduced = flat_slice.untyped_storage().data_ptr()

    _on_bucket = __call__

    def _join(self):
        cur = torch.cuda.current_stream() if self._side is not None else None
        if cur is not None:
            cur.wait_stream(self._side)
            cur.wait_stream(self._side_small)

    # ------------------------------------------------------------------ after loss.backward()
    @staticmethod
    def _views_of(flat, grads):
        lo = flat.data_ptr()
        hi = lo + flat.numel() * flat.element_size()
        total = 0
        for g in grads:
            if g is None or not (lo <= g.data_ptr() < hi) or not g.is_contiguous():
                return False
            total += g.numel()
        return 0 < total <= flat.numel()          # the buffer may pad every gradient to an aligned slot

    def allreduce(self, meta=None):
        """Sum the gradients over all ranks.  Overlapped mode: the buckets were already all-reduced on the side
        stream during the backward; the compute stream waits for them.  Otherwise: zero-copy when the backward
        wrote the gradients into the flat buffer of capdec.functional.DecoderTrainFn (meta['flat_grads']), else
        one coalesced copy in / copy out."""
        grads = [p.grad for p in self.params]
        flat = meta.get("flat_grads") if isinstance(meta, dict) else None
        if flat is not None and self._reduced is not None and \
                self._reduced == flat.untyped_storage().data_ptr():
            self._reduced = None
            self._join()
            return flat
        if flat is not None and self._views_of(flat, grads):
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM)
            return flat
        live = [g for g in grads if g is not None]
        if not live:
            return None
        n = sum(g.numel() for g in live)
        if self._flat is None or self._flat.numel() != n or self._flat.device != live[0].device:
            self._flat = torch.empty(n, dtype=live[0].dtype, device=live[0].device)
        off = 0
        for g in live:
            self._flat[off:off + g.numel()].copy_(g.reshape(-1))
            off += g.numel()
        self.dist.all_reduce(self._flat, op=self.dist.ReduceOp.SUM)
        off = 0
        for g in live:
            g.copy_(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return self._flat


def shard_range(n_items, rank, world):
    """Contiguous block of `ceil(n/world)` items for this rank (inference shards images with no
    communication)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
