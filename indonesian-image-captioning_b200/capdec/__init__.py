"""capdec -- Python host side of the B200 caption-decoder hot path.

Thin glue over the C ABI of ``libcapdec.so`` (``include/capdec.h``): ctypes
binding, ``torch.autograd.Function`` wrappers and the data-parallel helper.
There is no CPU or pure-PyTorch fallback: importing :mod:`capdec._lib` fails
loudly if the shared library has not been built, and every op raises if the
tensors are not on a CUDA device.
"""
from .config import get_precision, set_precision, precision_scope  # noqa: F401


def set_graphs(flag):
    """Replay the training step from CUDA graphs (see capdec.functional)."""
    from . import functional
    functional.set_graphs(flag)


def graphs_enabled():
    from . import functional
    return functional.graphs_enabled()
