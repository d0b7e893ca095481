"""Data-parallel training glue: one process per GPU, torch.distributed (NCCL over NVLink on the
B200 box, gloo in the CPU tests), ONE exchange per iteration -- the sum all-reduce of the decoder
gradients.  The reference has no multi-GPU path; this is the new work BASELINE
configs 3/5 ask for.

Loss scaling contract (see CaptionDecoderBase.loss): every rank divides its cross-entropy sum by
the GLOBAL token count and its alpha regulariser by world_size, so the SUM of the per-rank
gradients equals the gradient of the single-process mean loss over the global batch.
"""
import torch


class GradReducer:
    def __init__(self, module, dist, async_op=False):
        self.module = module
        self.dist = dist
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._flat = None

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
        """Sum the gradients over all ranks.  Zero-copy when the backward wrote them into the flat
        buffer of capdec.functional.DecoderTrainFn (meta['flat_grads']); otherwise one coalesced
        copy in / copy out."""
        grads = [p.grad for p in self.params]
        flat = meta.get("flat_grads") if isinstance(meta, dict) else None
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
