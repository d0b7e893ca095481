"""PureSCN decoder, drop-in for the reference `models/decoders/pure_scn.py`
(constructor :26, forward :87-140 -> 4-tuple without alphas, sample :142-249 -> token list)."""
from torch import nn

from capdec.decoder_base import CaptionDecoderBase
from models.scn_cell import SCNCell


class PureSCN(CaptionDecoderBase):
    kind = "pure_scn"

    def __init__(self, embed_dim, decoder_dim, factored_dim, semantic_dim, vocab_size,
                 encoder_dim=2048, dropout=0.5):
        super(PureSCN, self).__init__()
        self.embed_dim = embed_dim
        self.encoder_dim = encoder_dim
        self.decoder_dim = decoder_dim
        self.factored_dim = factored_dim
        self.semantic_dim = semantic_dim
        self.vocab_size = vocab_size
        self.embedding = nn.Embedding(vocab_size, embed_dim)
        self.dropout = nn.Dropout(p=dropout)
        self.decode_step = SCNCell(embed_dim, decoder_dim, semantic_dim, factored_dim, bias=True)
        self.init_h = nn.Linear(encoder_dim, decoder_dim)
        self.init_c = nn.Linear(encoder_dim, decoder_dim)
        self.fc = nn.Linear(decoder_dim, vocab_size)
        self.init_weights()

    def forward(self, encoder_out, semantic_input, encoded_captions, caption_lengths):
        r"""Returns (scores, sorted captions, decode lengths, sort indices)."""
        return self._forward_impl(encoder_out, semantic_input, encoded_captions, caption_lengths)

    def sample(self, beam_size, word_map, encoder_out, tag_out):
        return self._sample_one(beam_size, word_map, encoder_out, tag_out)
