import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "indonesian-image-captioning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, capdec
from oracle import capdec_oracle as O
from gpu_util import build_decoder, call_forward
dims = dict(A=512, M=512, D=512, F=512, S=1000, V=1000, E=2048)
B = 6
lengths = [7, 5, 4, 6, 3, 2]
for kind in (O.ATTENTION_SCN, O.PURE_SCN, O.PURE_ATTENTION):
    enc, tags, caps, caplens = O.synthetic_batch(B, dims["V"], seed=7, lengths=lengths)
    args = [t.cuda() for t in (enc, tags, caps, caplens)]
    with capdec.precision_scope("bf16"):
        torch.manual_seed(0)
        dec = build_decoder(kind, dims).train()
        scores, caps_sorted, dl, alphas, sort_ind = call_forward(dec, kind, *args)
        loss, _ = dec.loss(scores, caps_sorted, dl, alphas)
        loss.backward()
        torch.cuda.synchronize()
        print(kind, float(loss), bool(torch.isfinite(scores).all()))
# beam search small
with capdec.precision_scope("bf16"):
    dec = build_decoder(O.ATTENTION_SCN, dims).eval()
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(3, 14, 14, dims["E"], generator=g).relu_().cuda()
    tags = torch.rand(3, dims["S"], generator=g).cuda()
    os.environ["CAPDEC_WSUM_STREAM"] = "1"
    with torch.no_grad():
        r = dec.sample_batch(3, dims["V"] - 2, dims["V"] - 1, enc, tags, max_steps=4)
    torch.cuda.synchronize()
    print("beam ok", r["len"].tolist())
