"""Overlay package: mirrors the reference's `models/` tree for the decoder hot path only.

Put this directory on sys.path BEFORE the reference checkout; `extend_path` merges
both `models` packages so `models.encoders.*` (out of scope, kept as the reference has
them) still resolves to the reference while `models.decoders.*`, `models.scn_cell` and
`models.attention` resolve to the B200 implementations here.  Checkpoints that pickle
whole modules (reference utils/checkpoint.py:20-28) keep working because the class
import paths are unchanged.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
