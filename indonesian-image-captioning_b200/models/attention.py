"""Soft attention, drop-in for the reference `models/attention.py` (class Attention :6-44).

Parameters live in three `nn.Linear` sub-modules with the reference names
(`encoder_att`, `decoder_att`, `full_att`); `forward(encoder_out, decoder_hidden)` returns
`(attention_weighted_encoding, alpha)` computed by libcapdec: the encoder projection is a
GEMM on the selected engine, the score/softmax/weighted-sum step is the cluster kernel of
csrc/attention.cu.
"""
import torch
from torch import nn

from capdec import functional as CF
from capdec.config import get_precision


class Attention(nn.Module):
    r"""Soft Attention Network.

    Arguments
        encoder_dim (int): feature size of encoded images
        decoder_dim (int): size of decoder's RNN
        attention_dim (int): size of the attention network
    """

    def __init__(self, encoder_dim, decoder_dim, attention_dim):
        super(Attention, self).__init__()
        self.encoder_att = nn.Linear(encoder_dim, attention_dim)
        self.decoder_att = nn.Linear(decoder_dim, attention_dim)
        self.full_att = nn.Linear(attention_dim, 1)
        self.relu = nn.ReLU()
        self.softmax = nn.Softmax(dim=1)

    def forward(self, encoder_out, decoder_hidden):
        if torch.is_grad_enabled() and (encoder_out.requires_grad or decoder_hidden.requires_grad):
            raise RuntimeError("Attention.forward is inference-only when driven step by step; "
                               "gradients flow through the fused decoder forward (capdec_backward)")
        prec = get_precision()
        ft = torch.bfloat16 if prec == "bf16" else torch.float32
        B, P, E = encoder_out.shape
        enc = encoder_out.detach().to(ft).contiguous()
        att1 = CF.gemm(enc.view(B * P, E), self.encoder_att.weight.detach().to(ft).contiguous(),
                       bias=self.encoder_att.bias.detach().float(), out_ft=True).view(B, P, -1)
        att2 = CF.gemm(decoder_hidden.detach().to(ft).contiguous(),
                       self.decoder_att.weight.detach().to(ft).contiguous(),
                       bias=self.decoder_att.bias.detach().float())
        z, alpha, awe = CF.attention_step(att1, enc, att2, -1,
                                          self.full_att.weight.detach().float().reshape(-1).contiguous(),
                                          self.full_att.bias.detach().float())
        return awe, alpha
