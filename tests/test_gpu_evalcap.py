"""the batched evaluation driver (capdec.evalcap) produces, for every image, the hypothesis
string the reference's per-image loop (eval_caption.py:96-131: decoder.sample + join) produces."""
import pytest
import torch

import capdec
from capdec import evalcap
from oracle import capdec_oracle as O
from conftest import load_golden
from gpu_util import build_decoder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION])
def test_batched_eval_driver_matches_per_image_sample(kind):
    blob = load_golden("beam_" + kind)
    V = blob["dims"]["V"]
    word_map = {"w%d" % i: i for i in range(1, V - 3)}
    word_map.update({"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1})
    rev = {v: k for k, v in word_map.items()}
    drop = {word_map["<start>"], word_map["<end>"], word_map["<pad>"]}
    with capdec.precision_scope("fp32"):
        dec = build_decoder(kind, blob["dims"])
        dec.load_state_dict(blob["state_dict"], strict=True)
        dec.train()                                       # the driver switches to eval and restores the mode
        enc = torch.cat([im["encoder_out"] for im in blob["images"]])
        tags = torch.cat([im["tags"] for im in blob["images"]])
        G = enc.size(0)
        allcaps = torch.randint(1, V - 3, (G, 2, 7))
        allcaps[:, :, 0] = word_map["<start>"]
        allcaps[:, :, 5] = word_map["<end>"]
        allcaps[:, :, 6] = word_map["<pad>"]
        # two ragged batches
        cut = G // 2 + 1
        if kind == O.PURE_ATTENTION:
            batches = [(enc[:cut], allcaps[:cut]), (enc[cut:], allcaps[cut:])]
        else:
            batches = [(enc[:cut], tags[:cut], allcaps[:cut]), (enc[cut:], tags[cut:], allcaps[cut:])]
        refs, hyps, done = evalcap.generate_captions(dec, batches, word_map, beam_size=3)
        assert dec.training
        assert len(refs) == len(hyps) == len(done) == G
        dec.eval()
        dec.raise_on_incomplete = False
        for g, im in enumerate(blob["images"]):
            args = (im["encoder_out"].cuda(),) if kind == O.PURE_ATTENTION else \
                (im["encoder_out"].cuda(), im["tags"].cuda())
            out = dec.sample(3, word_map, *args)
            seq = out if kind == O.PURE_SCN else out[0]
            assert hyps[g] == " ".join(rev[w] for w in seq if w not in drop), g
            assert done[g] == im["results"][3]["completed"]
            if im["results"][3]["completed"]:
                assert hyps[g] == " ".join(rev[w] for w in im["results"][3]["seq"] if w not in drop)
            assert refs[g] == [" ".join(rev[w] for w in c if w not in drop) for c in allcaps[g].tolist()]
        t = evalcap.transpose_references(refs)
        assert len(t) == 2 and len(t[0]) == G and t[1][3] == refs[3][1]
