"""Global precision switch: 'bf16' (tcgen05 fast path, default) or 'fp32' (parity mode)."""
import contextlib
import os

_VALID = ("fp32", "bf16")
_precision = os.environ.get("CAPDEC_PRECISION", "bf16").lower()
if _precision not in _VALID:
    raise ValueError("CAPDEC_PRECISION must be one of %s" % (_VALID,))


def get_precision():
    return _precision


def set_precision(p):
    global _precision
    p = str(p).lower()
    if p not in _VALID:
        raise ValueError("precision must be one of %s, got %r" % (_VALID, p))
    _precision = p


@contextlib.contextmanager
def precision_scope(p):
    old = get_precision()
    set_precision(p)
    try:
        yield
    finally:
        set_precision(old)


def precision_code(p=None):
    return 1 if (p or _precision) == "bf16" else 0
