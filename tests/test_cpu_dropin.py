"""The overlay against the reference's OWN callers (CPU, build container only: needs /root/reference).

Two interpreter configurations are driven as subprocesses so that `sys.path` is exactly what a user would set:
  * "reference":  PYTHONPATH = /root/reference                      (the stock code)
  * "overlay":    PYTHONPATH = indonesian-image-captioning_b200 : /root/reference   
and exchange checkpoints through a temp directory:
  1. the reference builds its three decoders, saves `state_dict`s and pickles WHOLE modules the way
     `utils/checkpoint.py:20-28` does;
  2. under the overlay, the reference's `utils/loader.load_decoder` (`utils/loader.py:9-68`, resolved from the
     reference tree) returns the B200 classes with identical parameters for all three model types, the
     reference-made whole-module pickles unpickle as the B200 classes, `models.encoders.*` still resolves to the
     reference, and an overlay module is pickled through the reference's `save_checkpoint`;
  3. back under the stock reference, that overlay-made checkpoint unpickles as the reference classes with the
     same parameters (checkpoints move both ways).
"""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import PKG

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")

DIMS = dict(vocab_size=37, embed_dim=16, attention_dim=24, decoder_dim=32, factored_dim=24)

STAGE1 = """
import sys, torch
sys.path.insert(0, %(ref)r)
out = sys.argv[1]
from models.decoders.attention_scn import AttentionSCN
from models.decoders.pure_scn import PureSCN
from models.decoders.pure_attention import PureAttention
import models.decoders.attention_scn as m
assert m.__file__.startswith(%(ref)r), m.__file__
torch.manual_seed(0)
D = %(dims)r
decs = {
    "attention_scn": AttentionSCN(attention_dim=D["attention_dim"], embed_dim=D["embed_dim"], decoder_dim=D["decoder_dim"],
                                  factored_dim=D["factored_dim"], semantic_dim=1000, vocab_size=D["vocab_size"]),
    "pure_scn": PureSCN(embed_dim=D["embed_dim"], decoder_dim=D["decoder_dim"], factored_dim=D["factored_dim"],
                        semantic_dim=1000, vocab_size=D["vocab_size"]),
    "pure_attention": PureAttention(attention_dim=D["attention_dim"], embed_dim=D["embed_dim"],
                                    decoder_dim=D["decoder_dim"], vocab_size=D["vocab_size"]),
}
for k, d in decs.items():
    torch.save(d.state_dict(), "%%s/sd_%%s.pt" %% (out, k))
    torch.save({"decoder": d, "epoch": 3}, "%%s/whole_ref_%%s.pth.tar" %% (out, k))     # utils/checkpoint.py:20-28
print("STAGE1_OK")
"""

STAGE2 = """
import os, sys, torch
sys.path[:0] = [%(pkg)r, %(ref)r]
out = sys.argv[1]
os.chdir(out)
import utils.loader, utils.checkpoint
assert utils.loader.__file__.startswith(%(ref)r) and utils.checkpoint.__file__.startswith(%(ref)r)
from utils.loader import load_decoder
import models.decoders.attention_scn as m
assert m.__file__.startswith(%(pkg)r), m.__file__
import models.encoders.caption as enc_mod, models.encoders.tagger as tag_mod      # out of scope: the reference's
assert enc_mod.__file__.startswith(%(ref)r) and tag_mod.__file__.startswith(%(ref)r)
from capdec.decoder_base import CaptionDecoderBase
D = %(dims)r
for kind in ("attention_scn", "pure_scn", "pure_attention"):
    sd = torch.load("sd_%%s.pt" %% kind)
    dec = load_decoder(kind, sd, **D)                       # strict load_state_dict inside (utils/loader.py:65)
    assert isinstance(dec, CaptionDecoderBase) and dec.kind == kind, type(dec)
    assert type(dec).__module__ == "models.decoders." + kind
    got = dec.state_dict()
    assert list(got.keys()) == list(sd.keys())
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    whole = torch.load("whole_ref_%%s.pth.tar" %% kind, weights_only=False)["decoder"]
    assert isinstance(whole, CaptionDecoderBase), type(whole)          # reference-made pickle -> B200 class
    assert all(torch.equal(whole.state_dict()[k], sd[k]) for k in sd)
    assert len(whole._param_list()) == len(sd)                        # the host glue works on an unpickled module
    # the reference's own checkpoint writer on an overlay module (file name from model / data name)
    opt = torch.optim.Adam(dec.parameters(), lr=4e-4)
    utils.checkpoint.save_checkpoint(kind, "dropin", 1, 0, None, dec, None, opt, 0.5, False)
    assert os.path.exists("checkpoint_%%s_dropin.pth.tar" %% kind)
try:
    load_decoder("no_such_model", {}, 10)
    raise SystemExit("load_decoder accepted a bad type")
except ValueError:
    pass
print("STAGE2_OK")
"""

STAGE3 = """
import sys, torch
sys.path.insert(0, %(ref)r)
out = sys.argv[1]
import models.decoders.attention_scn as m
assert m.__file__.startswith(%(ref)r)
for kind in ("attention_scn", "pure_scn", "pure_attention"):
    sd = torch.load("%%s/sd_%%s.pt" %% (out, kind))
    ck = torch.load("%%s/checkpoint_%%s_dropin.pth.tar" %% (out, kind), weights_only=False)
    dec = ck["decoder"]
    assert type(dec).__module__ == "models.decoders." + kind
    assert sys.modules[type(dec).__module__].__file__.startswith(%(ref)r)      # the stock class
    got = dec.state_dict()
    assert list(got.keys()) == list(sd.keys()) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert ck["decoder_optimizer"].param_groups[0]["lr"] == 4e-4
print("STAGE3_OK")
"""


def _run(code, tmp):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code), str(tmp)], capture_output=True, text=True,
                       env=env, cwd=str(tmp))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_reference_loader_and_checkpoints_through_the_overlay(tmp_path):
    fmt = {"ref": REF, "pkg": PKG, "dims": DIMS}
    assert "STAGE1_OK" in _run(STAGE1 % fmt, tmp_path)
    assert "STAGE2_OK" in _run(STAGE2 % fmt, tmp_path)
    assert "STAGE3_OK" in _run(STAGE3 % fmt, tmp_path)
