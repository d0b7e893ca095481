"""SCN-LSTM cell, drop-in for the reference `models/scn_cell.py` (class SCNCell :10-184).

Same constructor, parameter names/shapes/initialisation and `forward(x, s, hx)` contract;
the arithmetic runs in libcapdec (capdec_scn_cell_step).  Inside the decoders the cell is
never called step by step from Python -- the whole recurrence is one C call -- so this
`forward` exists for code that drives the cell directly (reference notebooks).
"""
import math

import torch
from torch import nn

from capdec import functional as CF


class SCNCell(nn.Module):
    r"""Semantic Compositional Network LSTM cell.

    Arguments
        input_size (int): size of input
        hidden_size (int): size of hidden state
        semantic_size (int): size of the tag (semantic concept) vector
        factor_size (int): size of the factorisation
        bias (boolean, optional): use the two bias vectors
    """

    def __init__(self, input_size, hidden_size, semantic_size, factor_size, bias=True):
        super(SCNCell, self).__init__()
        self.factor_size = factor_size
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.semantic_size = semantic_size
        # registration order fixes the RNG consumption order of reset_parameters
        # (reference scn_cell.py:29-48) so equal seeds give equal weights
        for name, rows in (("weight_ia", input_size), ("weight_ib", semantic_size),
                           ("weight_ic", hidden_size), ("weight_ha", hidden_size),
                           ("weight_hb", semantic_size), ("weight_hc", hidden_size)):
            setattr(self, name, nn.Parameter(torch.empty(rows, 4 * factor_size)))
        if bias:
            self.bias_ih = nn.Parameter(torch.empty(4 * hidden_size))
            self.bias_hh = nn.Parameter(torch.empty(4 * hidden_size))
        else:
            self.register_parameter('bias_ih', None)
            self.register_parameter('bias_hh', None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.hidden_size)      # reference scn_cell.py:156-159
        for w in self.parameters():
            nn.init.uniform_(w, -bound, bound)

    def extra_repr(self):
        return '{}, {}, semantic={}, factor={}'.format(self.input_size, self.hidden_size,
                                                       self.semantic_size, self.factor_size)

    def _check(self, x, h, label):
        # same messages as the reference guards (scn_cell.py:169-184)
        if x.size(0) != h.size(0):
            raise RuntimeError("Input batch size {} doesn't match hidden{} batch size {}".format(
                x.size(0), label, h.size(0)))
        if h.size(1) != self.hidden_size:
            raise RuntimeError("hidden{} has inconsistent hidden_size: got {}, expected {}".format(
                label, h.size(1), self.hidden_size))

    def forward(self, wemb_input, semantic_input, hx=None):
        if wemb_input.size(1) != self.input_size:
            raise RuntimeError("input has inconsistent input_size: got {}, expected {}".format(
                wemb_input.size(1), self.input_size))
        if hx is None:
            z = wemb_input.new_zeros(wemb_input.size(0), self.hidden_size)
            hx = (z, z)
        self._check(wemb_input, hx[0], '[0]')
        self._check(wemb_input, hx[1], '[1]')
        if torch.is_grad_enabled() and any(t.requires_grad for t in (wemb_input, semantic_input, *hx)):
            raise RuntimeError("SCNCell.forward is inference-only when driven step by step; "
                               "gradients flow through the fused decoder forward (capdec_backward)")
        b_ih = self.bias_ih if self.bias_ih is not None else wemb_input.new_zeros(4 * self.hidden_size)
        b_hh = self.bias_hh if self.bias_hh is not None else wemb_input.new_zeros(4 * self.hidden_size)
        weights = (self.weight_ia, self.weight_ib, self.weight_ic, self.weight_ha, self.weight_hb,
                   self.weight_hc, b_ih, b_hh)
        return CF.scn_cell_step(weights, wemb_input, semantic_input, hx[0], hx[1])
