"""ctypes binding of libcapdec.so (include/capdec.h).  Fails loudly when the library is
missing -- there is no fallback path."""
import ctypes as C
import os

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# CAPDEC_LIB: another build of the same library (e.g. the -DCAPDEC_RECUR_FINE debug build of tools/recur_prof.py)
LIB_PATH = os.environ.get("CAPDEC_LIB") or os.path.join(PKG_ROOT, "libcapdec.so")

KIND = {"attention_scn": 0, "pure_scn": 1, "pure_attention": 2}

ERR_NAMES = {0: "OK", -1: "BAD_SHAPE", -2: "BAD_ARG", -3: "WORKSPACE", -4: "CUDA", -5: "UNSUPPORTED"}


class CapdecError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "precision", "B", "T", "P", "E", "A", "M", "D", "F", "S", "V", "L")]


PARAM_FIELDS = ("enc_att_w", "enc_att_b", "dec_att_w", "dec_att_b", "full_att_w", "full_att_b",
                "emb", "w_ia", "w_ib", "w_ic", "w_ha", "w_hb", "w_hc", "b_ih", "b_hh",
                "init_h_w", "init_h_b", "init_c_w", "init_c_b", "f_beta_w", "f_beta_b",
                "fc_w", "fc_b")


class AdamSeg(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


# every symbol include/capdec.h declares: (restype, argtypes)
_vp, _i, _i64, _sz, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_uint64
SIGNATURES = {
    "capdec_version": (_i, []),
    "capdec_last_error": (C.c_char_p, []),
    "capdec_init": (_i, []),
    "capdec_launch_count": (C.c_ulonglong, []),
    "capdec_topk_hits": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "capdec_clip_adam_step": (_i, [C.POINTER(AdamSeg), _i] + [C.c_double] * 6 + [_i, _i, _vp]),
    "capdec_dropout_mask": (_i, [_u64, _f, _i64, _vp, _vp]),
    "capdec_recur_timing": (None, [_i]),
    "capdec_recur_last_ms": (_f, [_i]),
    "capdec_workspace_bytes": (_sz, [C.POINTER(Dims), _i]),
    "capdec_forward_train": (_i, [C.POINTER(Dims), C.POINTER(Params), _vp, _i64, _i64, _i64, _vp, _vp,
                                  _vp, _vp, _f, _u64, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "capdec_backward": (_i, [C.POINTER(Dims), C.POINTER(Params), _vp, _f, _vp, _vp, _vp,
                             _vp, C.POINTER(Params), _vp, _sz, _i, _vp]),
    "capdec_loss_fwd": (_i, [C.POINTER(Dims), _vp, _vp, _vp, _vp, C.c_int32, _f, _vp, _vp, _vp]),
    "capdec_loss_bwd": (_i, [C.POINTER(Dims), _vp, _vp, _vp, _vp, C.c_int32, _f, _f, _vp, _vp, _vp, _vp,
                             _vp, _vp]),
    "capdec_beam_workspace_bytes": (_sz, [C.POINTER(Dims), _i, _i, _i]),
    "capdec_beam_search": (_i, [C.POINTER(Dims), C.POINTER(Params), _vp, _vp, _i, _i, _i, C.c_int32,
                                C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "capdec_beam_search_strided": (_i, [C.POINTER(Dims), C.POINTER(Params), _vp, _i64, _i64, _i64, _vp, _i, _i, _i,
                                        C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "capdec_gemm": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i64,
                         _i64, _i64, _i, _vp]),
    "capdec_gemm_tn": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i64, _i64, _i64, _vp]),
    "capdec_attention_scratch_floats": (_sz, [_i, _i, _i, _i]),
    "capdec_attention_step": (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i,
                                   _i, _i, _i, _vp, _vp]),
    "capdec_attention_bwd_step": (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp, _vp, _i64, _vp, _i64, _vp, _vp,
                                       _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "capdec_scn_cell_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "capdec_scn_cell_step": (_i, [_i, _i, _i, _i, _i, _i] + [_vp] * 14 + [_vp, _sz, _vp]),
}

_lib = None


def load():
    """Load libcapdec.so once.  Raises CapdecError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CapdecError(
            "libcapdec.so not found at %s -- build it with `python __graft_entry__.py` or "
            "`python indonesian-image-captioning_b200/capdec/build.py`; there is no fallback path"
            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().capdec_last_error()
        raise CapdecError("%s failed: %s (%s)" % (what, ERR_NAMES.get(rc, rc),
                                                  msg.decode() if msg else ""))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
