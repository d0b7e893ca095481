"""world_size-2 gloo tests of the data-parallel host logic (CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _worker(rank, world, port, ret):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from capdec.parallel import GradReducer, shard_range
        from oracle import capdec_oracle as O
        torch.manual_seed(0)
        dims = dict(attention_dim=24, embed_dim=16, decoder_dim=32, factored_dim=24, semantic_dim=12,
                    vocab_size=37, encoder_dim=40)
        params = O.random_params(O.ATTENTION_SCN, seed=0, **dims)
        # two shards of a global batch of 6 captions
        lengths = [[9, 3, 14], [6, 11, 4]]
        shards = [O.synthetic_batch(3, 37, seed=50 + r, side=3, E=40, S=12, max_len=14, lengths=lengths[r])
                  for r in range(world)]
        n_global = sum(l - 1 for ls in lengths for l in ls)

        def shard_loss(p, r):
            enc, tags, caps, caplens = shards[r]
            out = O.decoder_forward(O.ATTENTION_SCN, p, enc, tags, caps, caplens)
            n_local = sum(out[2])
            # global-denominator scaling used by CaptionDecoderBase.loss(n_tokens=..., alpha_c=1/world)
            ce = O.caption_loss(out[0], out[1], out[2], None) * n_local / n_global
            reg = ((1.0 - out[3].sum(dim=1)) ** 2).mean() / world
            return ce + reg

        class Holder(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in params.values()])
        h = Holder()
        p_local = dict(zip(params.keys(), h.ps))
        shard_loss(p_local, rank).backward()
        red = GradReducer(h, dist)
        red.allreduce(None)
        # single-process truth: sum of both shard losses
        p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        (shard_loss(p_ref, 0) + shard_loss(p_ref, 1)).backward()
        err = max((p_local[k].grad - p_ref[k].grad).abs().max().item() for k in params)
        # zero-copy path: grads that are views of one flat buffer
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        class Two(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.zeros(4))
                self.b = torch.nn.Parameter(torch.zeros(2, 3))
        m = Two()
        m.a.grad = flat[:4]
        m.b.grad = flat[4:].view(2, 3)
        out = GradReducer(m, dist).allreduce({"flat_grads": flat})
        ok_zero_copy = out.data_ptr() == flat.data_ptr() and torch.equal(flat, torch.arange(10.) * 3)
        # padded slots (every gradient starts on an aligned boundary of the flat buffer, as DecoderTrainFn lays them
        # out): still one in-place all-reduce of the whole buffer
        padded = torch.zeros(128)
        m2 = Two()
        m2.a.grad = padded[0:4]
        m2.b.grad = padded[64:70].view(2, 3)
        with torch.no_grad():
            m2.a.grad.fill_(float(rank + 1))
            m2.b.grad.fill_(10.0 * (rank + 1))
        out2 = GradReducer(m2, dist).allreduce({"flat_grads": padded})
        ok_zero_copy = ok_zero_copy and out2.data_ptr() == padded.data_ptr() and \
            bool((m2.a.grad == 3).all()) and bool((m2.b.grad == 30).all()) and float(padded[4:64].abs().sum()) == 0.0
        # overlapped path: the decoder's backward hands the reducer one bucket of the flat buffer at a time (production
        # order); allreduce() afterwards must NOT reduce a second time
        from capdec import functional as CF
        flatb = torch.zeros(256)
        m3 = Two()
        m3.a.grad = flatb[0:4]
        m3.b.grad = flatb[64:70].view(2, 3)
        with torch.no_grad():
            m3.a.grad.fill_(float(rank + 1))
            m3.b.grad.fill_(10.0 * (rank + 1))
        red3 = GradReducer(m3, dist, overlap=True)
        assert CF._bucket_hook is not None
        CF._bucket_hook(0, 2, flatb[0:64])
        mid = bool((m3.a.grad == 3).all()) and bool((m3.b.grad == 10.0 * (rank + 1)).all())   # bucket 1 not yet
        CF._bucket_hook(1, 2, flatb[64:256])
        out3 = red3.allreduce({"flat_grads": flatb})
        ok_overlap = mid and out3.data_ptr() == flatb.data_ptr() and bool((m3.a.grad == 3).all()) and \
            bool((m3.b.grad == 30).all())
        out4 = red3.allreduce({"flat_grads": flatb})          # a later call without hook activity reduces again
        ok_overlap = ok_overlap and bool((m3.a.grad == 6).all())
        red3.close()
        assert CF._bucket_hook is None
        ret[rank] = (err, ok_zero_copy and ok_overlap, shard_range(5000, rank, world))
    finally:
        dist.destroy_process_group()


def test_dp_gradients_sum_to_global_gradient():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err, ok, rng = ret[r]
        assert err < 1e-6
        assert ok
    assert ret[0][2] == (0, 2500) and ret[1][2] == (2500, 5000)


def test_gradient_buckets_cover_every_parameter_in_production_order():
    """The flat gradient buffer is laid out in the order capdec_backward produces the gradients (its `phases`):
    every parameter of every decoder kind sits in exactly one bucket, fc first."""
    from capdec import functional as CF
    for kind in ("attention_scn", "pure_scn", "pure_attention"):
        names = CF.param_names(kind)
        buckets = CF.grad_buckets(kind)
        assert len(buckets) == len(CF.BUCKET_PHASES) == len(CF.BUCKET_PHASES_EARLY_FC) == 5
        flat = [i for b in buckets for i in b]
        assert sorted(flat) == list(range(len(names)))
        assert [names[i] for i in buckets[0]] == ["fc.weight", "fc.bias"]
        assert "embedding.weight" in [names[i] for i in buckets[1]]
        assert all(names[i].startswith(("decode_step.", "init_h.", "init_c.")) for i in buckets[2])
        assert [names[i] for i in buckets[4]] == (["attention.encoder_att.weight", "attention.encoder_att.bias"]
                                                  if kind != "pure_scn" else [])
