"""Shared host logic of the three drop-in decoder modules (models/decoders/*.py).

Mirrors the reference call contract (attention_scn.py:95-158 forward, :160-296 sample):
sorting by caption length, the returned tuple, the un-permuted tag quirk (App. C-1, handled
inside capdec_forward_train by passing tags as given) -- while the tensor math is ONE C call
per forward and one per backward.
"""
import torch
from torch import nn

import os

from . import functional as CF
from .config import get_precision

_CHECK_IDS = os.environ.get("CAPDEC_CHECK_IDS", "1") != "0"


class CaptionDecoderBase(nn.Module):
    kind = None          # "attention_scn" | "pure_scn" | "pure_attention"

    # ------------------------------------------------------------------ helpers
    def _param_list(self):
        # (owning module, attribute) pairs are resolved once; the Parameter objects are looked up on every
        # call, so re-assigned parameters (load_pretrained_embeddings) are picked up
        slots = self.__dict__.get("_capdec_param_slots")
        if slots is None:
            slots = []
            for n in CF.param_names(self.kind):
                mod_path, _, attr = n.rpartition(".")
                slots.append((self.get_submodule(mod_path) if mod_path else self, attr))
            self.__dict__["_capdec_param_slots"] = slots
        return [m._parameters[a] for m, a in slots]

    def _dims_kw(self):
        return {"A": getattr(self, "attention_dim", 0), "M": self.embed_dim, "D": self.decoder_dim,
                "F": getattr(self, "factored_dim", 0), "S": getattr(self, "semantic_dim", 0),
                "V": self.vocab_size}

    def init_weights(self):
        r"""Uniform init of embedding / fc as the reference (attention_scn.py:58-63)."""
        self.embedding.weight.data.uniform_(-0.1, 0.1)
        self.fc.bias.data.fill_(0)
        self.fc.weight.data.uniform_(-0.1, 0.1)

    def load_pretrained_embeddings(self, embeddings):
        self.embedding.weight = nn.Parameter(embeddings)

    def fine_tune_embeddings(self, fine_tune=True):
        for p in self.embedding.parameters():
            p.requires_grad = fine_tune

    def init_hidden_state(self, encoder_out):
        r"""h0, c0 from the mean encoder feature (attention_scn.py:82-93), on the GEMM engine."""
        prec = get_precision()
        ft = torch.bfloat16 if prec == "bf16" else torch.float32
        m = encoder_out.detach().float().mean(dim=1).to(ft).contiguous()
        h = CF.gemm(m, self.init_h.weight.detach().to(ft).contiguous(), bias=self.init_h.bias.detach())
        c = CF.gemm(m, self.init_c.weight.detach().to(ft).contiguous(), bias=self.init_c.bias.detach())
        return h, c

    # ------------------------------------------------------------------ teacher-forced forward
    def _forward_impl(self, encoder_out, semantic_input, encoded_captions, caption_lengths):
        batch_size = encoder_out.size(0)
        encoder_dim = encoder_out.size(-1)
        if torch.is_grad_enabled() and (encoder_out.requires_grad or
                                        (semantic_input is not None and semantic_input.requires_grad)):
            # the reference can fine-tune the encoders through the decoder (fine_tune_encoder=True with an
            # encoder_optimizer, trains/attention_scn.py:94-109, 141); capdec_backward produces no gradient
            # with respect to encoder_out / the tags, so that configuration must not train silently
            raise RuntimeError("capdec decoders do not back-propagate into encoder_out / semantic_input "
                               "(fine_tune_encoder=True is not supported): detach the encoder outputs")
        if encoded_captions.dtype != torch.int64 or caption_lengths.dtype != torch.int64:
            raise RuntimeError("encoded_captions and caption_lengths must be int64 tensors (got %s, %s)"
                               % (encoded_captions.dtype, caption_lengths.dtype))
        if self.kind != "pure_attention" and tuple(semantic_input.shape) != (batch_size, self.semantic_dim):
            raise RuntimeError("semantic_input must have shape (%d, %d), got %s"
                               % (batch_size, self.semantic_dim, tuple(semantic_input.shape)))
        enc = encoder_out.view(batch_size, -1, encoder_dim)        # strided views are fine (App. C-22)
        if enc.dtype != torch.float32:
            enc = enc.float()
        lens, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True)   # same op as :117-118
        # the lengths travel to the host through a pinned buffer + event instead of a blocking .tolist(): the
        # launches queued below run while the host waits for them
        lens_host = torch.empty(lens.shape, dtype=lens.dtype, pin_memory=True) if lens.is_cuda else None
        lens_ev = None
        if lens_host is not None:
            lens_host.copy_(lens, non_blocking=True)
            lens_ev = torch.cuda.Event()
            lens_ev.record()
        caps_sorted = encoded_captions[sort_ind].contiguous()
        if _CHECK_IDS and caps_sorted.is_cuda:
            # nn.Embedding / CrossEntropyLoss raise on ids outside the vocabulary; the kernels would clamp them.
            # Asynchronous device-side check (no host sync), CAPDEC_CHECK_IDS=0 removes it.
            torch._assert_async(((caps_sorted >= 0) & (caps_sorted < self.vocab_size)).all(),
                                "caption token id outside [0, vocab_size)")
        # everything that does not need the lengths on the host comes BEFORE the sync below: after it the
        # GPU is idle until the first launch
        tags = None
        if self.kind != "pure_attention":
            tags = semantic_input.detach().float().contiguous()    # NOT permuted (App. C-1)
        p = self.dropout.p if self.training else 0.0
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if p > 0 else 0
        params, dims_kw, enc_d = self._param_list(), self._dims_kw(), enc.detach()
        spec = CF.speculate(self.kind, params, enc_d, tags, caps_sorted, sort_ind, dims_kw=dims_kw, dropout_p=p,
                            seed=seed) if lens_ev is not None else None
        if lens_ev is not None:                                    # host sync, as upstream (:131)
            lens_ev.synchronize()
            decode_lengths = (lens_host - 1).tolist()
        else:
            decode_lengths = (lens - 1).tolist()
        out, meta = CF.decoder_forward(self.kind, params, enc_d, tags, caps_sorted,
                                       sort_ind, decode_lengths, dims_kw=dims_kw, dropout_p=p,
                                       seed=seed, spec=spec)
        if self.kind == "pure_scn":
            predictions, alphas = out, None
        else:
            predictions, alphas = out
        predictions._capdec_meta = meta
        if alphas is None:
            return predictions, caps_sorted, decode_lengths, sort_ind
        return predictions, caps_sorted, decode_lengths, alphas, sort_ind

    def loss(self, scores, caps_sorted, decode_lengths, alphas=None, alpha_c=1.0, n_tokens=None):
        r"""Fused loss glue of trains/attention_scn.py:219-235; returns (loss, [total, ce, reg])."""
        return CF.caption_loss(scores, caps_sorted, decode_lengths, alphas, alpha_c,
                               meta=getattr(scores, "_capdec_meta", None), n_tokens=n_tokens)

    # ------------------------------------------------------------------ beam search
    def sample_batch(self, beam_size, start_id, end_id, encoder_out, tag_out=None, max_steps=50,
                     want_alphas=True, want_trace=False):
        r"""Independent beam searches for G images in one device-side loop (no per-step host sync)."""
        from . import beam
        return beam.beam_search_batch(self, beam_size, start_id, end_id, encoder_out, tag_out,
                                      max_steps=max_steps, want_alphas=want_alphas,
                                      want_trace=want_trace)

    # The reference's `sample` dies with ValueError("max() arg is an empty sequence") when no beam
    # emits <end> within 51 steps (attention_scn.py:292).  True keeps that error
    # behaviour; False returns the best live beam instead (see `last_sample_completed`).
    raise_on_incomplete = True

    def _sample_one(self, beam_size, word_map, encoder_out, tag_out):
        if len(word_map) != self.vocab_size:
            raise ValueError("len(word_map)=%d != vocab_size=%d" % (len(word_map), self.vocab_size))
        res = self.sample_batch(beam_size, word_map['<start>'], word_map['<end>'], encoder_out, tag_out,
                                want_alphas=self.kind != "pure_scn")
        n = int(res["len"][0])
        seq = res["seq"][0, :n].tolist()
        self.last_sample_completed = bool(res["completed"][0])
        self.last_sample_score = float(res["score"][0])
        if not self.last_sample_completed and self.raise_on_incomplete:
            raise ValueError("max() arg is an empty sequence")
        if self.kind == "pure_scn":
            return seq
        side = encoder_out.size(1)
        alphas = res["alpha"][0, :n].view(n, side, side).tolist()
        return seq, alphas
