"""PureAttention (Show-Attend-Tell) decoder, drop-in for the reference
`models/decoders/pure_attention.py` (constructor :25, 3-argument forward :90-151,
sample(beam_size, word_map, encoder_out) :153-281).  `decode_step` is a stock nn.LSTMCell
holder for the parameters (gate order i,f,g,o); the recurrence itself runs in libcapdec."""
from torch import nn

from capdec.decoder_base import CaptionDecoderBase
from models.attention import Attention


class PureAttention(CaptionDecoderBase):
    kind = "pure_attention"

    def __init__(self, attention_dim, embed_dim, decoder_dim, vocab_size, encoder_dim=2048,
                 dropout=0.5):
        super(PureAttention, self).__init__()
        self.encoder_dim = encoder_dim
        self.attention_dim = attention_dim
        self.embed_dim = embed_dim
        self.decoder_dim = decoder_dim
        self.vocab_size = vocab_size
        self.attention = Attention(encoder_dim, decoder_dim, attention_dim)
        self.embedding = nn.Embedding(vocab_size, embed_dim)
        self.dropout = nn.Dropout(p=dropout)
        self.decode_step = nn.LSTMCell(embed_dim + encoder_dim, decoder_dim, bias=True)
        self.init_h = nn.Linear(encoder_dim, decoder_dim)
        self.init_c = nn.Linear(encoder_dim, decoder_dim)
        self.f_beta = nn.Linear(decoder_dim, encoder_dim)
        self.sigmoid = nn.Sigmoid()
        self.fc = nn.Linear(decoder_dim, vocab_size)
        self.init_weights()

    def forward(self, encoder_out, encoded_captions, caption_lengths):
        r"""Returns (scores, sorted captions, decode lengths, alphas, sort indices)."""
        return self._forward_impl(encoder_out, None, encoded_captions, caption_lengths)

    def sample(self, beam_size, word_map, encoder_out):
        return self._sample_one(beam_size, word_map, encoder_out, None)
