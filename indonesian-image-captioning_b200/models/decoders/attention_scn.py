"""AttentionSCN decoder, drop-in for the reference `models/decoders/attention_scn.py`.

Same constructor (reference :28), sub-module names and state_dict layout,
`forward(encoder_out, semantic_input, encoded_captions, caption_lengths)` 5-tuple (:95-158) and
`sample(beam_size, word_map, encoder_out, tag_out)` (:160-296); the math runs in libcapdec.
"""
from torch import nn

from capdec.decoder_base import CaptionDecoderBase
from models.attention import Attention
from models.scn_cell import SCNCell


class AttentionSCN(CaptionDecoderBase):
    kind = "attention_scn"

    def __init__(self, attention_dim, embed_dim, decoder_dim, factored_dim, semantic_dim, vocab_size,
                 encoder_dim=2048, dropout=0.5):
        super(AttentionSCN, self).__init__()
        self.attention_dim = attention_dim
        self.embed_dim = embed_dim
        self.encoder_dim = encoder_dim
        self.decoder_dim = decoder_dim
        self.factored_dim = factored_dim
        self.semantic_dim = semantic_dim
        self.vocab_size = vocab_size
        # construction order == reference order, so equal torch seeds give equal weights
        self.attention = Attention(encoder_dim, decoder_dim, attention_dim)
        self.embedding = nn.Embedding(vocab_size, embed_dim)
        self.dropout = nn.Dropout(p=dropout)
        self.decode_step = SCNCell(embed_dim + encoder_dim, decoder_dim, semantic_dim, factored_dim,
                                   bias=True)
        self.init_h = nn.Linear(encoder_dim, decoder_dim)
        self.init_c = nn.Linear(encoder_dim, decoder_dim)
        self.f_beta = nn.Linear(decoder_dim, encoder_dim)
        self.sigmoid = nn.Sigmoid()
        self.fc = nn.Linear(decoder_dim, vocab_size)
        self.init_weights()

    def forward(self, encoder_out, semantic_input, encoded_captions, caption_lengths):
        r"""Returns (scores, sorted captions, decode lengths, alphas, sort indices)."""
        return self._forward_impl(encoder_out, semantic_input, encoded_captions, caption_lengths)

    def sample(self, beam_size, word_map, encoder_out, tag_out):
        r"""Beam search for one image: (token list incl. <start>/<end>, alphas nested list)."""
        return self._sample_one(beam_size, word_map, encoder_out, tag_out)
