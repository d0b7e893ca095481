#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference decoder modules, for the CPU baseline only.

    python oracle/make_ref.py [--ref /root/reference]

The reference (rayandrew/indonesian-image-captioning) is pure Python: there is nothing to compile and
`pip install /root/reference` fails ("neither setup.py nor pyproject.toml"), so the "build" of the reference
arm is a byte-for-byte copy of the ten files on the decoder path from the reference tree,
where they lie, into `oracle/_ref/` -- git-ignored (never part of this repo's history or product), NOT
gpurun-ignored, so that `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's OWN
code on the GPU box's host cores (`cpu_baseline.kind == "reference"`).  `__graft_entry__.build()` runs this
when `/root/reference` is present; on the GPU box only the files it left behind are used.  Without
`oracle/_ref/` the bench falls back to the oracle port and says `kind: "port"`.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under `indonesian-image-captioning_b200/` imports it.
A MANIFEST with the sha256 of every file is written next to the copies.
"""
import argparse
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "models/__init__.py", "models/attention.py", "models/scn_cell.py",
    "models/decoders/__init__.py", "models/decoders/attention_scn.py", "models/decoders/pure_scn.py",
    "models/decoders/pure_attention.py",
    "utils/__init__.py", "utils/tensor.py", "utils/token.py", "utils/device.py",
]


def make_ref(ref_root="/root/reference", dest=DEST):
    """Copy the decoder-path files; returns the destination, or None when the reference tree is absent."""
    if not os.path.isdir(ref_root):
        return None
    lines = []
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            lines.append("%s  %s" % (hashlib.sha256(fh.read()).hexdigest(), rel))
    with open(os.path.join(dest, "MANIFEST"), "w") as fh:
        fh.write("# unmodified copies from %s made by oracle/make_ref.py (sha256  path)\n" % ref_root)
        fh.write("\n".join(lines) + "\n")
    return dest


def available(dest=DEST):
    return all(os.path.exists(os.path.join(dest, rel)) for rel in FILES)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    out = make_ref(ap.parse_args().ref)
    print(out if out else "reference tree not found; oracle/_ref not (re)built")
    sys.exit(0)
