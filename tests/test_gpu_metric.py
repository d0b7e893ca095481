"""capdec.metric against the reference's utils/metric.py accuracy (topk + eq) restated with stock torch ops."""
import pytest
import torch
from torch.nn.utils.rnn import pack_padded_sequence

from capdec import metric

pytestmark = pytest.mark.gpu


def ref_accuracy(scores, targets, k):          # utils/metric.py:25-39
    batch_size = targets.size(0)
    _, ind = scores.topk(k, 1, True, True)
    correct = ind.eq(targets.view(-1, 1).expand_as(ind))
    return correct.view(-1).float().sum().item() * (100.0 / batch_size)


@pytest.mark.parametrize("k", [1, 5])
@pytest.mark.parametrize("N,V", [(1600, 10000), (37, 203), (5, 8)])
def test_accuracy_matches_reference(N, V, k):
    g = torch.Generator(device="cuda").manual_seed(N + V + k)
    scores = torch.randn(N, V, device="cuda", generator=g)
    targets = torch.randint(0, V, (N,), device="cuda", generator=g)
    top = scores.topk(3, 1).indices                        # make a good share of the targets actual hits
    pick = torch.rand(N, device="cuda", generator=g) < 0.5
    targets = torch.where(pick, top[torch.arange(N, device="cuda"), torch.randint(0, 3, (N,), device="cuda", generator=g)], targets)
    assert metric.accuracy(scores, targets, k) == pytest.approx(ref_accuracy(scores, targets, k), abs=1e-9)


def test_unpacked_counts_what_the_packed_glue_counts():
    B, T, V, L = 6, 9, 301, 12
    g = torch.Generator(device="cuda").manual_seed(3)
    scores = torch.randn(B, T, V, device="cuda", generator=g)
    caps = torch.randint(0, V, (B, L), device="cuda", generator=g)
    dl = [9, 7, 7, 4, 2, 1]
    for b in range(B):                                      # plant hits
        for t in range(0, dl[b], 2):
            caps[b, t + 1] = scores[b, t].argmax()
    s = pack_padded_sequence(scores, dl, batch_first=True).data
    t = pack_padded_sequence(caps[:, 1:], dl, batch_first=True).data
    for k in (1, 5):
        hits = metric.topk_hits_unpacked(scores, caps, dl, k).item()
        assert hits * (100.0 / sum(dl)) == pytest.approx(ref_accuracy(s, t, k), abs=1e-9)
