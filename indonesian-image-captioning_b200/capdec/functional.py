"""Autograd wrappers around the C ABI (teacher-forced decoder, fused loss, unit ops).

All tensors must live on a CUDA device; nothing here computes on the CPU.

Two execution modes for the training step:
  * graph replay (default; `capdec.set_graphs(False)` or CAPDEC_GRAPHS=0 turns it off): per problem FAMILY
    (dims but T, parameter storage, device) one set of static workspace / output / gradient buffers sized for the
    longest caption, and per T one plan whose compute phase of the forward and whose backward are captured once
    into CUDA graphs and replayed, so a step costs a handful of launches on the host.  On the bf16 path with the
    persistent recurrence kernels the graphs are LENGTH-INDEPENDENT: the kernels read the decode lengths from
    the device, so ragged batches with the same longest caption replay the same graph (fp32 mode and shapes
    outside recur.cu key their plans on the exact lengths instead).  Outputs are views of the family's static
    buffers and stay valid until the next forward of the same family; a forward issued while an earlier one of
    the family still waits for its backward runs eagerly on buffers of its own.
  * eager: one C call per forward and per backward on fresh buffers; each issues its launches on the current
    stream.
"""
import collections
import ctypes as C
import os
import warnings
import weakref

import torch

from . import _lib
from .config import get_precision, precision_code

# state_dict name -> CapdecParams field, per decoder kind
_COMMON_TAIL = [("init_h.weight", "init_h_w"), ("init_h.bias", "init_h_b"),
                ("init_c.weight", "init_c_w"), ("init_c.bias", "init_c_b")]
_ATT = [("attention.encoder_att.weight", "enc_att_w"), ("attention.encoder_att.bias", "enc_att_b"),
        ("attention.decoder_att.weight", "dec_att_w"), ("attention.decoder_att.bias", "dec_att_b"),
        ("attention.full_att.weight", "full_att_w"), ("attention.full_att.bias", "full_att_b")]
_SCN = [("decode_step.weight_ia", "w_ia"), ("decode_step.weight_ib", "w_ib"),
        ("decode_step.weight_ic", "w_ic"), ("decode_step.weight_ha", "w_ha"),
        ("decode_step.weight_hb", "w_hb"), ("decode_step.weight_hc", "w_hc"),
        ("decode_step.bias_ih", "b_ih"), ("decode_step.bias_hh", "b_hh")]
_LSTM = [("decode_step.weight_ih", "w_ia"), ("decode_step.weight_hh", "w_ha"),
         ("decode_step.bias_ih", "b_ih"), ("decode_step.bias_hh", "b_hh")]
_BETA = [("f_beta.weight", "f_beta_w"), ("f_beta.bias", "f_beta_b")]
_FC = [("fc.weight", "fc_w"), ("fc.bias", "fc_b")]
PARAM_MAP = {
    "attention_scn": _ATT + [("embedding.weight", "emb")] + _SCN + _COMMON_TAIL + _BETA + _FC,
    "pure_scn": [("embedding.weight", "emb")] + _SCN + _COMMON_TAIL + _FC,
    "pure_attention": _ATT + [("embedding.weight", "emb")] + _LSTM + _COMMON_TAIL + _BETA + _FC,
}

# gradient buckets in PRODUCTION ORDER of capdec_backward (its `phases` 1|2, 4, 8, 16, 32): the flat gradient buffer is
# laid out in this order, so every bucket is one contiguous slice that a data-parallel caller can all-reduce as
# soon as its stage has been launched (capdec/parallel.py)
_BUCKET_NAMES = [
    ("fc.weight", "fc.bias"),
    ("decode_step.weight_ia", "decode_step.weight_ih", "embedding.weight"),
    ("decode_step.weight_ic", "decode_step.weight_hc", "decode_step.weight_ha", "decode_step.weight_hh",
     "decode_step.weight_ib", "decode_step.weight_hb", "decode_step.bias_ih", "decode_step.bias_hh",
     "init_h.weight", "init_h.bias", "init_c.weight", "init_c.bias"),
    ("f_beta.weight", "f_beta.bias", "attention.decoder_att.weight", "attention.decoder_att.bias",
     "attention.full_att.weight", "attention.full_att.bias"),
    ("attention.encoder_att.weight", "attention.encoder_att.bias"),
]
# stage groups of capdec_backward after which bucket i is complete.  Default: the fc bucket is handed over after
# the reverse loop (an all-reduce kernel issued before it would fight the persistent kernel for SMs); a hook
# with `early_fc` (a communicator limited to a few CTAs, see capdec/parallel.py) gets it BEFORE the loop and
# runs next to it.
BUCKET_PHASES = (1 | 2, 4, 8, 16, 32)
BUCKET_PHASES_EARLY_FC = (1, 2 | 4, 8, 16, 32)


def grad_buckets(kind):
    """Per bucket, the indices into PARAM_MAP[kind] of the parameters whose gradients it holds."""
    names = param_names(kind)
    out = [[names.index(n) for n in bucket if n in names] for bucket in _BUCKET_NAMES]
    assert sorted(i for b in out for i in b) == list(range(len(names)))
    return out


_graphs_enabled = os.environ.get("CAPDEC_GRAPHS", "1") != "0"
_MAX_FAMILIES = 4
_replayed_launches = 0
_bucket_hook = None


def set_grad_bucket_hook(fn):
    """fn(bucket_index, n_buckets, flat_slice) is called inside the decoder's backward right after the launches
    that produce that bucket of the flat gradient buffer have been queued (None removes the hook).  Used by
    capdec.parallel.GradReducer to overlap the gradient all-reduce with the rest of the backward."""
    global _bucket_hook
    _bucket_hook = fn


def launch_count():
    """Kernels of libcapdec launched by this process: direct launches counted by the library plus
    the kernel nodes of every graph replay."""
    return int(_lib.load().capdec_launch_count()) + _replayed_launches


def set_graphs(flag):
    """Enable / disable CUDA-graph replay of the training step (see module docstring)."""
    global _graphs_enabled
    _graphs_enabled = bool(flag)
    if not flag:
        _families.clear()


def graphs_enabled():
    return _graphs_enabled


def param_names(kind):
    return [n for n, _ in PARAM_MAP[kind]]


def _params_struct(kind, tensors):
    s = _lib.Params()
    for (name, field), t in zip(PARAM_MAP[kind], tensors):
        setattr(s, field, None if t is None else t.data_ptr())
    return s


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CapdecError("capdec ops need CUDA tensors (got a %s tensor); there is no CPU path"
                                   % t.device)


def make_dims(kind, precision, B, T, P, E, A, M, D, F, S, V, L):
    return _lib.Dims(_lib.KIND[kind], precision_code(precision), B, T, P, E, A, M, D, F, S, V, L)


def _dims_key(d):
    return tuple(getattr(d, n) for n, _ in _lib.Dims._fields_)


def _with_T(d, T):
    return _lib.Dims(*[T if n == "T" else getattr(d, n) for n, _ in _lib.Dims._fields_])


class _GradStore:
    """ONE flat fp32 buffer for all gradients of a decoder, in bucket (production) order; every gradient starts
    on a 256-byte boundary (vector loads in the fused optimizer; the padding stays zero)."""

    def __init__(self, kind, params, dev):
        pad = lambda n: (n + 63) // 64 * 64
        self.flat = torch.zeros(sum(pad(p.numel()) for p in params), dtype=torch.float32, device=dev)
        self.grads = [None] * len(params)
        self.bucket_slices = []
        off = 0
        for bucket in grad_buckets(kind):
            lo = off
            for i in bucket:
                p = params[i]
                self.grads[i] = self.flat[off:off + p.numel()].view_as(p)
                off += pad(p.numel())
            self.bucket_slices.append((lo, off))
        self.gstruct = _params_struct(kind, self.grads)


class _Buffers:
    """Device buffers of the training step, sized for `T_cap` decode steps."""

    def __init__(self, kind, dims_cap, need_bwd, dev):
        lib = _lib.load()
        B, T, P, V = dims_cap.B, dims_cap.T, dims_cap.P, dims_cap.V
        self.kind, self.dev, self.T_cap, self.need_bwd = kind, dev, T, need_bwd
        self.ws_bytes = lib.capdec_workspace_bytes(C.byref(dims_cap), 1 if need_bwd else 0)
        if self.ws_bytes == 0:
            _lib.check(-1, "capdec_workspace_bytes")
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.pred = torch.empty(B * T * V, dtype=torch.float32, device=dev)
        self.alph = None if kind == "pure_scn" else torch.empty(B * T * P, dtype=torch.float32, device=dev)
        self.dlog = None          # d logits in the GEMM feature type (fused-loss path)
        self.d_alph = None
        self.d_pred = None        # static copy of an autograd-provided gradient (generic path)
        self.len_pin = torch.empty(B, dtype=torch.int32, pin_memory=True)
        self.len_d = torch.empty(B, dtype=torch.int32, device=dev)


class _Plan:
    """One decoder problem (dims incl. T) on a set of buffers: views of the outputs, the decode lengths as the C
    ABI wants them and, in graph mode, the captured graphs."""

    def __init__(self, kind, dims, buf, params, gstore, len_free):
        self.kind, self.dims, self.buf, self.dev = kind, dims, buf, buf.dev
        self.need_bwd = buf.need_bwd
        self.len_free = len_free  # graphs do not depend on the decode lengths (only on B and T)
        B, T, P, V = dims.B, dims.T, dims.P, dims.V
        self.ws, self.ws_bytes = buf.ws, buf.ws_bytes
        self.predictions = buf.pred[:B * T * V].view(B, T, V)
        self.alphas = None if buf.alph is None else buf.alph[:B * T * P].view(B, T, P)
        self.len_h = (C.c_int32 * B)()
        self.len_list = None
        self.len_d = buf.len_d
        self.params = params
        self.pstruct = _params_struct(kind, params)
        self.gstore = gstore
        self.dlog = self.d_alphas = None
        self.graphs = {}          # slot -> torch.cuda.CUDAGraph
        self.graph_nodes = {}     # slot -> kernels in the captured graph
        self.calls = {}           # slot -> number of launches so far

    def set_lengths(self, decode_lengths):
        """Host copy for the C ABI + device copy for the fused loss: pinned + non_blocking, so that nothing on the
        host waits for the stream (a pageable torch.tensor(..., device=) copy would block until the forward has
        drained and expose the whole host cost of loss + backward)."""
        self.len_list = [int(x) for x in decode_lengths]
        self.len_h[:] = self.len_list
        self.buf.len_pin.copy_(torch.tensor(self.len_list, dtype=torch.int32))
        self.len_d.copy_(self.buf.len_pin, non_blocking=True)

    # gradient storage is shared by every plan of a family
    @property
    def flat_grads(self):
        return self.gstore.flat

    @property
    def grads(self):
        return self.gstore.grads

    @property
    def gstruct(self):
        return self.gstore.gstruct

    def ensure_dlog(self):
        buf, d = self.buf, self.dims
        ldq = (d.V + 7) // 8 * 8
        esz = 2 if d.precision == 1 else 4
        if buf.dlog is None:
            buf.dlog = torch.empty(d.B * buf.T_cap * ldq * esz, dtype=torch.uint8, device=self.dev)
        if buf.d_alph is None and buf.alph is not None:
            buf.d_alph = torch.zeros_like(buf.alph)
        self.dlog = buf.dlog
        self.d_alphas = None if buf.d_alph is None else buf.d_alph[:d.B * d.T * d.P].view(d.B, d.T, d.P)

    def dlog_matrix(self):
        """The (B*T, V) logits-gradient matrix inside the feature-type buffer."""
        d = self.dims
        ldq = (d.V + 7) // 8 * 8
        ft = torch.bfloat16 if d.precision == 1 else torch.float32
        return self.dlog.view(ft)[:d.B * d.T * ldq].view(d.B * d.T, ldq)[:, :d.V]

    def run(self, slot, launch):
        """Graph mode: call #1 eager (also warms lazy state: tensor maps, function attributes),
        call #2 capture + replay, later calls replay."""
        global _replayed_launches
        g = self.graphs.get(slot)
        n = self.calls.get(slot, 0)
        if g is not None:
            g.replay()
            _replayed_launches += self.graph_nodes[slot]
        elif n == 0:
            launch()
        else:
            lib = _lib.load()
            before = lib.capdec_launch_count()
            g = torch.cuda.CUDAGraph()
            with warnings.catch_warnings():
                # a backward stage that queues nothing for this decoder (pure_scn has no attention gradients) is a
                # legitimately empty graph: replaying it is a no-op, torch's warning about it is noise on every rank
                warnings.filterwarnings("ignore", message="The CUDA Graph is empty")
                with torch.cuda.graph(g):
                    launch()
            # the capture call counted the kernels once without running them; this replay runs them
            self.graph_nodes[slot] = lib.capdec_launch_count() - before
            g.replay()
            self.graphs[slot] = g
        self.calls[slot] = n + 1


class _Family:
    """Everything the graph mode keeps for one decoder problem family = (kind, dims but T, backward wanted, dropout
    on/off, device, parameter storage, library switches): ONE set of buffers sized for the longest caption
    (T = L - 1), the flat gradient buffer, and one plan -- i.e. one set of captured graphs -- per T.  The graphs are
    keyed on (B, T) only: the kernels read the decode lengths from the device, so ragged batches replay them."""

    def __init__(self, kind, dims, need_bwd, params, dev):
        self.kind, self.need_bwd, self.params, self.dev = kind, need_bwd, params, dev
        self.buf = _Buffers(kind, _with_T(dims, dims.L - 1), need_bwd, dev)
        self.gstore = None
        self.plans = {}           # T (length-independent) or (T, lengths) -> _Plan
        self.len_free = None      # None = not tried yet
        self.pending = []         # weak references to outputs whose backward has not run yet
        self.last_T = self.prev_T = None

    def grad_store(self):
        if self.gstore is None:
            self.gstore = _GradStore(self.kind, self.params, self.dev)
        return self.gstore

    def backward_pending(self):
        self.pending = [r for r in self.pending if r() is not None]
        return bool(self.pending)

    def plan(self, dims, decode_lengths):
        key = dims.T if self.len_free is not False else (dims.T, tuple(decode_lengths))
        plan = self.plans.get(key)
        if plan is None:
            plan = _Plan(self.kind, dims, self.buf, self.params, self.grad_store() if self.need_bwd else None,
                         self.len_free is not False)
            self.plans[key] = plan
            if self.len_free is False and len(self.plans) > 8:      # tuple-keyed plans of a ragged stream
                self.plans.pop(next(iter(self.plans)))
        return plan


_families = collections.OrderedDict()


def _family_key(kind, dims, need_bwd, dropout_on, dev, params):
    return (kind, dims.precision, dims.B, dims.P, dims.E, dims.A, dims.M, dims.D, dims.F, dims.S, dims.V, dims.L,
            need_bwd, dropout_on, dev.index, tuple(p.data_ptr() for p in params),
            os.environ.get("CAPDEC_PERSISTENT", ""), os.environ.get("CAPDEC_FUSED_EPILOGUE", ""))


def _get_family(kind, dims, need_bwd, dropout_on, dev, params, create=True):
    key = _family_key(kind, dims, need_bwd, dropout_on, dev, params)
    fam = _families.get(key)
    if fam is None:
        if not create:
            return None
        fam = _Family(kind, dims, need_bwd, params, dev)
        _families[key] = fam
        while len(_families) > _MAX_FAMILIES:
            _families.popitem(last=False)
    else:
        _families.move_to_end(key)
    return fam


def _fwd_call(plan, enc, tags, caps_sorted, sort_ind, dropout_p, seed, phases):
    lib = _lib.load()
    if enc is not None:
        sb, sp, se = enc.stride()
    else:
        sb = sp = se = 0
    rc = lib.capdec_forward_train(
        C.byref(plan.dims), C.byref(plan.pstruct), _lib.ptr(enc), sb, sp, se, _lib.ptr(sort_ind),
        _lib.ptr(tags), _lib.ptr(caps_sorted), plan.len_h, float(dropout_p), int(seed) & ((1 << 63) - 1),
        1 if plan.need_bwd else 0, phases, _lib.ptr(plan.predictions), _lib.ptr(plan.alphas),
        _lib.ptr(plan.ws), plan.ws_bytes, _stream())
    return rc


def speculate(kind, module_params, enc, tags, caps_sorted, sort_ind, *, dims_kw, dropout_p=0.0, seed=0,
              precision=None):
    """Graph mode only.  The reference API hands the decode lengths back as a python list, so every forward
    has a host sync; the host work between that sync and the first launch would leave the GPU idle.  When the
    last two forwards of this family had the same T (fixed-length batches, or a stable longest caption), the
    INPUT phase and the PROLOGUE of the compute phase for that T are queued here, BEFORE the caller waits for the
    lengths -- nothing in them depends on the lengths beyond T.  decoder_forward() then checks the real T against
    the plan: equal -> the lengths are re-staged and only the rest of the compute phase is launched; different ->
    the normal path runs and this launch was wasted work."""
    if not _graphs_enabled:
        return None
    B, P, E = enc.shape
    prec = precision or get_precision()
    params = [p.detach() for p in module_params]
    need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in module_params)
    probe = make_dims(kind, prec, B, 1, P, E, dims_kw.get("A", 0), dims_kw["M"], dims_kw["D"],
                      dims_kw.get("F", 0), dims_kw.get("S", 0), dims_kw["V"], caps_sorted.shape[1])
    fam = _get_family(kind, probe, need_bwd, dropout_p > 0, enc.device, params, create=False)
    if fam is None or not fam.len_free or fam.last_T is None or fam.last_T != fam.prev_T or fam.backward_pending():
        return None
    plan = fam.plans.get(fam.last_T)
    if plan is None or plan.len_list is None or not (enc.is_cuda and enc.dtype == torch.float32):
        return None
    with torch.cuda.device(enc.device):
        _lib.check(_fwd_call(plan, enc, tags, caps_sorted, sort_ind, dropout_p, seed, 1), "capdec_forward_train")
        plan.run("pre", lambda: _lib.check(_fwd_call(plan, None, None, None, None, dropout_p, seed, 4 | 16),
                                           "capdec_forward_train"))
    return plan


class DecoderTrainFn(torch.autograd.Function):
    """predictions, alphas = decoder(enc, tags, caps_sorted, sort_ind; params) -- one C call each way
    (or one graph replay each way)."""

    @staticmethod
    def forward(ctx, meta, enc, tags, caps_sorted, sort_ind, *params):
        kind = meta["kind"]
        _require_cuda(enc, tags, caps_sorted, sort_ind, *params)
        dims = meta["dims"]
        dev = enc.device
        params = [p.detach() for p in params]
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.CapdecError("decoder parameters must be contiguous float32 master weights")
        need_bwd = meta["need_bwd"]
        lengths = meta["decode_lengths"]
        dp, seed = meta["dropout_p"], meta["seed"]
        fam = None
        if _graphs_enabled:
            fam = _get_family(kind, dims, need_bwd, dp > 0, dev, params)
            if need_bwd and fam.backward_pending():
                # an earlier forward of this family still waits for its backward (two forwards before a backward,
                # several losses, checkpointing): its saved activations live in the family's buffers, so this
                # call runs eagerly on buffers of its own
                fam = None
        use_graph = fam is not None

        def call(plan, phases, inputs=True):
            if inputs:
                rc = _fwd_call(plan, enc, tags, caps_sorted, sort_ind, dp, seed, phases)
            else:
                rc = _fwd_call(plan, None, None, None, None, dp, seed, phases)
            return rc

        with torch.cuda.device(dev):
            if not use_graph:
                buf = _Buffers(kind, dims, need_bwd, dev)
                plan = _Plan(kind, dims, buf, params, _GradStore(kind, params, dev) if need_bwd else None, False)
                plan.set_lengths(lengths)
                _lib.check(call(plan, 3), "capdec_forward_train")
            else:
                plan = fam.plan(dims, lengths)
                spec = meta.get("spec")
                if spec is plan and plan.len_free:
                    # inputs + prologue were queued before the host waited for the lengths (speculate()): the
                    # lengths are re-staged (32) and the rest of the compute phase follows
                    plan.set_lengths(lengths)
                    _lib.check(call(plan, 32), "capdec_forward_train")
                    plan.run("rest", lambda: _lib.check(call(plan, 2 | 8 | 16, False), "capdec_forward_train"))
                else:
                    plan.set_lengths(lengths)
                    _lib.check(call(plan, 1), "capdec_forward_train")       # input phase: reads the caller's tensors
                    if fam.len_free is None:
                        # first forward of the family: does the shape take the length-independent launch?
                        rc = call(plan, 2 | 16, False)
                        if rc == -5:                                        # CAPDEC_ERR_UNSUPPORTED
                            fam.len_free = False
                            fam.plans.clear()
                            plan = fam.plan(dims, lengths)
                            plan.set_lengths(lengths)
                            plan.run("fwd", lambda: _lib.check(call(plan, 2, False), "capdec_forward_train"))
                        else:
                            _lib.check(rc, "capdec_forward_train")
                            fam.len_free = True
                            plan.calls["fwd"] = 1
                    elif plan.len_free:
                        plan.run("fwd", lambda: _lib.check(call(plan, 2 | 16, False), "capdec_forward_train"))
                    else:
                        plan.run("fwd", lambda: _lib.check(call(plan, 2, False), "capdec_forward_train"))
                fam.prev_T, fam.last_T = fam.last_T, dims.T
        meta.pop("spec", None)
        ctx.meta = meta
        ctx.plan = plan
        ctx.family = fam
        ctx.use_graph = use_graph
        ctx.param_refs = meta.pop("param_refs", None)
        meta["plan"] = plan
        predictions = plan.predictions.detach() if use_graph else plan.predictions
        alphas = None
        if plan.alphas is not None:
            alphas = plan.alphas.detach() if use_graph else plan.alphas
        if use_graph and need_bwd:
            ctx.token = _Token()
            fam.pending.append(weakref.ref(ctx.token))
        if alphas is None:
            return predictions
        return predictions, alphas

    @staticmethod
    def backward(ctx, d_pred, d_alphas=None):
        lib = _lib.load()
        meta, plan = ctx.meta, ctx.plan
        kind, dims, dev = plan.kind, plan.dims, plan.dev
        if not plan.need_bwd:
            raise _lib.CapdecError("backward called but forward ran without save_for_backward")
        # a .grad left over from an earlier step may ALIAS the static gradient buffer (autograd adopts the
        # views returned below without copying): this call is about to overwrite that buffer, so such a
        # gradient is detached into its own storage first (gradient accumulation stays correct)
        if ctx.param_refs is not None:
            lo = plan.flat_grads.data_ptr()
            hi = lo + plan.flat_grads.numel() * 4
            for p in ctx.param_refs:
                if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                    p.grad = p.grad.clone()
        buf = plan.buf
        fused = meta.pop("fused_dlogits", None)    # set by FusedLossFn: gradient already in plan.dlog
        if fused:
            slot, d_pred_p, dlog_p = "bwd_fused", None, plan.dlog
            d_alphas_p = plan.d_alphas if kind != "pure_scn" else None
            # FusedLossFn hands autograd zero-stride zeros; anything else means a SECOND consumer of the scores /
            # alphas contributed a gradient: it is added to what the fused loss left in the buffers
            if d_pred is not None and any(d_pred.stride()):
                plan.dlog_matrix().add_(d_pred.reshape(dims.B * dims.T, dims.V).to(plan.dlog_matrix().dtype))
            if d_alphas is not None and d_alphas_p is not None and any(d_alphas.stride()):
                d_alphas_p.add_(d_alphas)
        else:
            slot, dlog_p = "bwd_generic", None
            if buf.d_pred is None:
                buf.d_pred = torch.zeros(dims.B * buf.T_cap * dims.V, dtype=torch.float32, device=dev)
            d_pred_p = buf.d_pred[:dims.B * dims.T * dims.V].view(dims.B, dims.T, dims.V)
            if d_pred is None:
                d_pred_p.zero_()
            else:
                d_pred_p.copy_(d_pred)
            d_alphas_p = None
            if kind != "pure_scn":
                if buf.d_alph is None:
                    buf.d_alph = torch.zeros_like(buf.alph)
                d_alphas_p = buf.d_alph[:dims.B * dims.T * dims.P].view(dims.B, dims.T, dims.P)
                if d_alphas is None:
                    d_alphas_p.zero_()
                else:
                    d_alphas_p.copy_(d_alphas)

        def call(phases):
            rc = lib.capdec_backward(
                C.byref(dims), C.byref(plan.pstruct), None if plan.len_free else plan.len_h, meta["dropout_p"],
                _lib.ptr(d_pred_p), _lib.ptr(dlog_p), _lib.ptr(d_alphas_p), _lib.ptr(plan.alphas),
                C.byref(plan.gstruct), _lib.ptr(plan.ws), plan.ws_bytes, phases, _stream())
            _lib.check(rc, "capdec_backward")

        hook = _bucket_hook
        with torch.cuda.device(dev):
            if hook is None:
                if ctx.use_graph:
                    plan.run(slot, lambda: call(0))
                else:
                    call(0)
            else:
                # stage by stage: the hook starts the all-reduce of a bucket while the next stage is being computed
                flat = plan.flat_grads
                stages = BUCKET_PHASES_EARLY_FC if getattr(hook, "early_fc", False) else BUCKET_PHASES
                nb = len(stages)
                for i, ph in enumerate(stages):
                    if ctx.use_graph:
                        plan.run("%s_%d" % (slot, ph), lambda ph=ph: call(ph))
                    else:
                        call(ph)
                    lo, hi = plan.gstore.bucket_slices[i]
                    hook(i, nb, flat[lo:hi])            # an empty slice (pure_scn: no attention) still closes the bucket
        meta["flat_grads"] = plan.flat_grads
        if ctx.family is not None:
            ctx.token = None                                   # this forward's saved state is no longer needed
        # FRESH view objects: autograd's AccumulateGrad adopts an incoming gradient without a copy only when
        # nobody else holds the tensor object (returning the stored views made it clone all 23 gradients --
        # 108.7 MB of device copies per step -- and made GradReducer fall back to copy-in / copy-out around
        # the all-reduce, because .grad no longer lived in the flat buffer)
        grads = [g.view(g.shape) for g in plan.grads]
        if ctx.use_graph and ctx.param_refs is not None and \
                any(p.grad is not None for p in ctx.param_refs):
            # gradient accumulation into an existing .grad: hand out copies, the static buffer is reused
            if hook is not None:
                hook(-1, len(BUCKET_PHASES), plan.flat_grads)   # the reduced values are what must be copied
            grads = [g.clone() for g in grads]
        return (None, None, None, None, None) + tuple(grads)


class _Token:
    """Lives as long as the autograd node of a forward whose backward has not run."""
    __slots__ = ("__weakref__",)


def decoder_forward(kind, module_params, enc, tags, caps_sorted, sort_ind, decode_lengths, *,
                    dims_kw, dropout_p=0.0, seed=0, precision=None, spec=None):
    """Run the teacher-forced decoder.  `module_params`: tensors in PARAM_MAP[kind] order."""
    B, P, E = enc.shape
    T = max(decode_lengths)
    dims = make_dims(kind, precision or get_precision(), B, T, P, E, dims_kw.get("A", 0), dims_kw["M"],
                     dims_kw["D"], dims_kw.get("F", 0), dims_kw.get("S", 0), dims_kw["V"],
                     caps_sorted.shape[1])
    need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in module_params)
    meta = {"kind": kind, "dims": dims, "decode_lengths": [int(x) for x in decode_lengths],
            "dropout_p": float(dropout_p), "seed": int(seed) & ((1 << 63) - 1), "need_bwd": need_bwd,
            "param_refs": list(module_params), "spec": spec}
    out = DecoderTrainFn.apply(meta, enc, tags, caps_sorted, sort_ind, *module_params)
    return out, meta


class FusedLossFn(torch.autograd.Function):
    """loss = packed CE(scores, caps_sorted[:,1:]) + alpha_c * mean((1 - sum_t alphas)^2)
    (trains/attention_scn.py:219-235) with both terms and their gradients computed by
    capdec_loss_fwd/_bwd.  When `meta` of the producing DecoderTrainFn is given, the logits
    gradient is handed to capdec_backward directly in the GEMM feature type (no fp32
    (B,T,V) gradient tensor is materialised)."""

    @staticmethod
    def forward(ctx, scores, alphas, caps_sorted, len_d, n_tokens, alpha_c, dims, meta):
        lib = _lib.load()
        _require_cuda(scores, alphas, caps_sorted, len_d)
        dev = scores.device
        B, T = dims.B, dims.T
        scores_c = scores.contiguous()
        alphas_c = None if alphas is None else alphas.contiguous()
        loss3 = torch.empty(3, dtype=torch.float32, device=dev)
        lse = torch.empty(2 * B * T + B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.capdec_loss_fwd(C.byref(dims), _lib.ptr(scores_c), _lib.ptr(alphas_c),
                                     _lib.ptr(caps_sorted), _lib.ptr(len_d), n_tokens, alpha_c,
                                     _lib.ptr(loss3), _lib.ptr(lse), _stream())
        _lib.check(rc, "capdec_loss_fwd")
        ctx.dims, ctx.meta, ctx.n_tokens, ctx.alpha_c = dims, meta, n_tokens, alpha_c
        ctx.has_alphas = alphas is not None
        ctx.save_for_backward(scores_c, alphas_c if alphas_c is not None else torch.empty(0, device=dev),
                              caps_sorted, len_d, lse)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, g_loss, _g3=None):
        lib = _lib.load()
        scores, alphas, caps_sorted, len_d, lse = ctx.saved_tensors
        dims, meta = ctx.dims, ctx.meta
        dev = scores.device
        alphas_p = alphas if ctx.has_alphas else None
        # upstream gradient stays on the device: no host sync in the training step
        g_dev = g_loss.detach().reshape(1).float().contiguous()
        plan = meta.get("plan") if meta is not None else None
        if plan is not None and plan.need_bwd:
            if meta.get("fused_dlogits"):
                # the logits gradient of ONE loss lives in the plan's buffer until the decoder's backward consumes it
                raise _lib.CapdecError("decoder.loss() was applied twice to the outputs of one forward; use the torch "
                                       "loss glue (pack_padded_sequence + CrossEntropyLoss) for additional losses")
            plan.ensure_dlog()
            d_alphas = plan.d_alphas if ctx.has_alphas else None
            with torch.cuda.device(dev):
                rc = lib.capdec_loss_bwd(C.byref(dims), _lib.ptr(scores), _lib.ptr(alphas_p),
                                         _lib.ptr(caps_sorted), _lib.ptr(len_d), ctx.n_tokens,
                                         ctx.alpha_c, 1.0, _lib.ptr(g_dev), _lib.ptr(lse), None,
                                         _lib.ptr(plan.dlog), _lib.ptr(d_alphas), _stream())
            _lib.check(rc, "capdec_loss_bwd")
            meta["fused_dlogits"] = True
            # the decoder Function picks the buffers up from its plan; autograd still needs tensors of
            # the right shape to route the call, so hand it cheap expanded zeros
            d_scores = torch.zeros((), device=dev).expand(scores.shape)
            d_al = None if not ctx.has_alphas else torch.zeros((), device=dev).expand(alphas.shape)
            return d_scores, d_al, None, None, None, None, None, None
        d_scores = torch.empty_like(scores)
        d_alphas = torch.empty_like(alphas) if ctx.has_alphas else None
        with torch.cuda.device(dev):
            rc = lib.capdec_loss_bwd(C.byref(dims), _lib.ptr(scores), _lib.ptr(alphas_p),
                                     _lib.ptr(caps_sorted), _lib.ptr(len_d), ctx.n_tokens, ctx.alpha_c,
                                     1.0, _lib.ptr(g_dev), _lib.ptr(lse), _lib.ptr(d_scores), None,
                                     _lib.ptr(d_alphas), _stream())
        _lib.check(rc, "capdec_loss_bwd")
        return d_scores, d_alphas, None, None, None, None, None, None


def caption_loss(scores, caps_sorted, decode_lengths, alphas=None, alpha_c=1.0, *, meta=None,
                 precision=None, n_tokens=None):
    """Fused replacement of the loss glue in trains/attention_scn.py:219-235.
    Returns (loss, (total, ce, reg) tensor).  `n_tokens` overrides the CE denominator
    (data-parallel training divides by the GLOBAL token count so that summing the ranks'
    gradients reproduces the single-process mean)."""
    B, T, V = scores.shape
    P = alphas.shape[2] if alphas is not None else 1
    if meta is not None:
        dims = meta["dims"]
    else:
        dims = make_dims("attention_scn" if alphas is not None else "pure_scn",
                         precision or get_precision(), B, T, P, 8, 8, 8, 8, 8, 8, V, caps_sorted.shape[1])
    plan = meta.get("plan") if meta is not None else None
    if plan is not None and plan.len_list == [int(x) for x in decode_lengths]:
        len_d = plan.len_d                      # staged by the forward: no blocking host-to-device copy here
    else:
        len_d = torch.tensor(list(decode_lengths), dtype=torch.int32, device=scores.device)
    n_tokens = int(sum(decode_lengths)) if n_tokens is None else int(n_tokens)
    return FusedLossFn.apply(scores, alphas, caps_sorted, len_d, n_tokens, float(alpha_c), dims, meta)


def dropout_mask(seed, p, B, T, D, device="cuda"):
    """(B, T, D) keep factors (0 or 1/(1-p)) of the dropout between h_t and fc (attention_scn.py:154) that a
    forward/backward with this seed applies -- the mask-injection hook of the parity tests: an independent
    implementation fed with this mask reproduces the training-mode arithmetic."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.CapdecError("capdec ops need a CUDA device; there is no CPU path")
    out = torch.empty(B, T, D, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.capdec_dropout_mask(int(seed) & ((1 << 63) - 1), float(p), out.numel(), _lib.ptr(out), _stream())
    _lib.check(rc, "capdec_dropout_mask")
    return out


# ---------------------------------------------------------------------------------
# unit ops (used by models/scn_cell.py, models/attention.py and the parity tests)
# ---------------------------------------------------------------------------------
def _ft_dtype(precision):
    return torch.bfloat16 if precision_code(precision) == 1 else torch.float32


def gemm(X, W, bias=None, addm=None, out_ft=False, precision=None, splitk=0, out=None):
    """out[r,n] = sum_k X[r,k] W[n,k] (+bias[n]) (+addm[r,n]) through the selected engine.
    X (rows,K) / W (N,K) must already be in the engine's operand type with 16-byte pitch."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(X, W, bias, addm)
    ft = _ft_dtype(prec)
    assert X.dtype == ft and W.dtype == ft and X.stride(-1) == 1 and W.stride(-1) == 1
    batch = X.shape[0] if X.dim() == 3 else 1
    rows, K = X.shape[-2], X.shape[-1]
    N = W.shape[-2]
    if out is None:
        make = torch.zeros if splitk else torch.empty      # split-K accumulates into the output
        out = make((batch, rows, N) if X.dim() == 3 else (rows, N),
                   dtype=ft if out_ft else torch.float32, device=X.device)
    sX = X.stride(0) if X.dim() == 3 else 0
    sW = W.stride(0) if W.dim() == 3 else 0
    sO = out.stride(0) if X.dim() == 3 else 0
    with torch.cuda.device(X.device):
        rc = lib.capdec_gemm(precision_code(prec), _lib.ptr(X), X.stride(-2), _lib.ptr(W), W.stride(-2),
                             _lib.ptr(out), out.stride(-2), 1 if out_ft else 0, _lib.ptr(bias),
                             _lib.ptr(addm), addm.stride(-2) if addm is not None else 0, rows, N, K,
                             batch, sX, sW, sO, splitk, _stream())
    _lib.check(rc, "capdec_gemm")
    return out


def gemm_tn(XT, WT, out=None):
    """out[r,n] = sum_k XT[k,r] WT[k,n] on transposed bf16 operands (weight-gradient products): the tcgen05
    engine with MN-major operand descriptors, no transposition pass.  XT (K,rows) / WT (K,N), last dim
    contiguous, 16-byte pitch."""
    lib = _lib.load()
    _require_cuda(XT, WT)
    assert XT.dtype == torch.bfloat16 and WT.dtype == torch.bfloat16 and XT.stride(-1) == 1 and WT.stride(-1) == 1
    K, rows = XT.shape
    N = WT.shape[1]
    assert WT.shape[0] == K
    if out is None:
        out = torch.empty(rows, N, dtype=torch.float32, device=XT.device)
    with torch.cuda.device(XT.device):
        rc = lib.capdec_gemm_tn(_lib.ptr(XT), XT.stride(0), _lib.ptr(WT), WT.stride(0), _lib.ptr(out), out.stride(0),
                                rows, N, K, 1, 0, 0, 0, _stream())
    _lib.check(rc, "capdec_gemm_tn")
    return out


def attention_step(att1, enc, g1, beta_col, w_f, b_f, rows_per_map=1, precision=None, want_awe=True):
    """One soft-attention step on prepared features; returns (z, alpha, awe)."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(att1, enc, g1, w_f, b_f)
    G_, P, A = att1.shape
    E = enc.shape[2]
    rows = g1.shape[0]
    ft = _ft_dtype(prec)
    assert att1.dtype == ft and enc.dtype == ft and att1.is_contiguous() and enc.is_contiguous()
    alpha = torch.empty(rows, P, dtype=torch.float32, device=enc.device)
    z = torch.empty(rows, E, dtype=ft, device=enc.device)
    awe = torch.empty(rows, E, dtype=torch.float32, device=enc.device) if want_awe else None
    scratch = torch.empty(max(1, lib.capdec_attention_scratch_floats(precision_code(prec), rows, P, E)),
                          dtype=torch.float32, device=enc.device)
    with torch.cuda.device(enc.device):
        rc = lib.capdec_attention_step(precision_code(prec), _lib.ptr(att1), _lib.ptr(enc), _lib.ptr(g1),
                                       g1.stride(0), beta_col, _lib.ptr(w_f), _lib.ptr(b_f),
                                       _lib.ptr(alpha), P, _lib.ptr(z), _lib.ptr(awe), rows, rows_per_map,
                                       P, E, A, _lib.ptr(scratch), _stream())
    _lib.check(rc, "capdec_attention_step")
    return z, alpha, awe


def attention_bwd_step(att1, enc, g1, beta_col, w_f, alpha, dz, awe, dalpha_ext=None, dAtt1=None,
                       precision=None):
    """Backward of attention_step (one feature map per row).  Returns (dbeta_pre, datt2, dAtt1,
    dwf_part, dbf_part); dAtt1 is accumulated into when given."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(att1, enc, g1, w_f, alpha, dz, awe)
    rows, P, A = att1.shape
    E = enc.shape[2]
    ft = _ft_dtype(prec)
    dev = enc.device
    ld = (E + A + 7) // 8 * 8
    dba = torch.zeros(rows, ld, dtype=ft, device=dev)
    if dAtt1 is None:
        dAtt1 = torch.zeros(rows, P, A, dtype=torch.float32, device=dev)
    dwf = torch.empty(rows, A, dtype=torch.float32, device=dev)
    dbf = torch.empty(rows, dtype=torch.float32, device=dev)
    n_scr = lib.capdec_attention_scratch_floats(precision_code(prec), rows, P, E) + rows * ((P + 3) // 4 * 4)
    scratch = torch.empty(max(1, n_scr), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.capdec_attention_bwd_step(
            precision_code(prec), _lib.ptr(att1), _lib.ptr(enc), _lib.ptr(g1), g1.stride(0), beta_col,
            _lib.ptr(w_f), _lib.ptr(alpha), alpha.stride(0), _lib.ptr(dalpha_ext),
            dalpha_ext.stride(0) if dalpha_ext is not None else 0, _lib.ptr(dz), _lib.ptr(awe),
            _lib.ptr(dba), ld, _lib.ptr(dAtt1), _lib.ptr(dwf), _lib.ptr(dbf), rows, P, E, A,
            _lib.ptr(scratch), _stream())
    _lib.check(rc, "capdec_attention_bwd_step")
    return dba[:, :E], dba[:, E:E + A], dAtt1, dwf, dbf


def scn_cell_step(weights, x, s, h, c, precision=None):
    """SCNCell.forward (models/scn_cell.py:52-154) through the C ABI.  weights = (w_ia, w_ib, w_ic,
    w_ha, w_hb, w_hc, b_ih, b_hh) fp32 CUDA tensors."""
    lib = _lib.load()
    prec = precision or get_precision()
    _require_cuda(x, s, h, c, *weights)
    rows, X = x.shape
    D = h.shape[1]
    F = weights[0].shape[1] // 4
    S = s.shape[1]
    pc = precision_code(prec)
    nbytes = lib.capdec_scn_cell_workspace_bytes(pc, rows, X, D, F, S)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
    h_out = torch.empty(rows, D, dtype=torch.float32, device=x.device)
    c_out = torch.empty(rows, D, dtype=torch.float32, device=x.device)
    args = [w.detach().contiguous().float() for w in weights] + \
           [t.detach().contiguous().float() for t in (x, s, h, c)]
    with torch.cuda.device(x.device):
        rc = lib.capdec_scn_cell_step(pc, rows, X, D, F, S, *[_lib.ptr(a) for a in args],
                                      _lib.ptr(h_out), _lib.ptr(c_out), _lib.ptr(ws), ws.numel(),
                                      _stream())
    _lib.check(rc, "capdec_scn_cell_step")
    return h_out, c_out
