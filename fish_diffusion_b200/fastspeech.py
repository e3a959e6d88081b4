"""FastSpeech2 feed-forward-transformer text encoder of the SVS recipes (BASELINE config #5
`svs_baseline.py`): `ENCODERS["FastSpeech2Encoder"]`, reference `fish_diffusion/modules/encoders/fast_speech.py:892-947`
on `FFTBlocks` (:798-889), `EncSALayer` (:698-764), `TransformerFFNLayer` (:230-276), `RelPositionalEncoding` (:94-119).

It sits BEFORE the hot path, acts on [B, T_phonemes, 256] and stays ordinary torch (as SURVEY.md scopes it): what the
native sampler needs from it is only its output layout, `[B, T, E]` channels-last.  Same parameter names as the
reference, so `model.text_encoder.*` keys of a Lightning checkpoint load unchanged.

Behaviour kept, including the quirks:
  * the input is scaled twice (`embed_scale` then the positional module's own sqrt(H)): x = H * proj(contents) + pe;
  * the "relative" positional table is built reversed over max_len = 5000 positions at construction and sliced from the
    front, so frame t gets the sinusoid of position 4999 - t (T - 1 - t once T exceeds 5000);
  * self-attention has no biases, pre-norm LayerNorm eps = 1e-12 inside the blocks and torch's default 1e-5 at the end;
  * padded frames are zeroed after every sub-layer.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .diffsinger import ENCODERS


def _reversed_sinusoid_table(length: int, dim: int) -> torch.Tensor:
    """[1, length, dim]: row i = interleaved sin/cos of position length-1-i (fast_speech.py:29-50 with reverse=True)."""
    pos = torch.arange(length - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(length, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


class _SelfAttention(nn.Module):
    """Parameter holder with the reference's names (`in_proj_weight`, `out_proj.weight`; no biases)."""

    def __init__(self, dim, heads):
        super().__init__()
        assert dim % heads == 0, "embed_dim must be divisible by num_heads"
        self.dim, self.heads = dim, heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.register_parameter("in_proj_bias", None)
        self.out_proj = nn.Linear(dim, dim, bias=False)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)

    def forward(self, x_tbc, key_padding_mask):
        y, _ = F.multi_head_attention_forward(
            x_tbc, x_tbc, x_tbc, self.dim, self.heads, self.in_proj_weight, None, None, None, False, 0.0,
            self.out_proj.weight, None, training=self.training, key_padding_mask=key_padding_mask, need_weights=False)
        return y


class _FFN(nn.Module):
    def __init__(self, dim, kernel_size, dropout, act):
        super().__init__()
        self.kernel_size, self.dropout, self.act = kernel_size, dropout, act
        self.ffn_1 = nn.Conv1d(dim, 4 * dim, kernel_size, padding=kernel_size // 2)
        self.ffn_2 = nn.Linear(4 * dim, dim)

    def forward(self, x_tbc):
        x = self.ffn_1(x_tbc.permute(1, 2, 0)).permute(2, 0, 1) * self.kernel_size ** -0.5
        if self.act == "gelu":
            x = F.gelu(x)
        elif self.act == "relu":
            x = F.relu(x)
        elif self.act == "swish":
            x = x * torch.sigmoid(x)
        x = F.dropout(x, self.dropout, training=self.training)
        return self.ffn_2(x)


class _Block(nn.Module):
    """`layers.{i}.op.*` of the reference: pre-norm self-attention and conv feed-forward, both residual."""

    def __init__(self, dim, heads, kernel_size, dropout, act):
        super().__init__()
        self.dropout = dropout
        self.layer_norm1 = nn.LayerNorm(dim, eps=1e-12)
        self.self_attn = _SelfAttention(dim, heads)
        self.layer_norm2 = nn.LayerNorm(dim, eps=1e-12)
        self.ffn = _FFN(dim, kernel_size, dropout, act)

    def forward(self, x, padding_mask, keep):
        y = self.self_attn(self.layer_norm1(x), padding_mask)
        x = (x + F.dropout(y, self.dropout, training=self.training)) * keep
        y = self.ffn(self.layer_norm2(x))
        return (x + F.dropout(y, self.dropout, training=self.training)) * keep


class _Layer(nn.Module):
    def __init__(self, *a):
        super().__init__()
        self.op = _Block(*a)

    def forward(self, x, padding_mask, keep):
        return self.op(x, padding_mask, keep)


@ENCODERS.register_module(name="FastSpeech2Encoder", force=True)
class FastSpeech2Encoder(nn.Module):
    def __init__(self, input_size=1024, max_seq_len=4096, num_layers=4, hidden_size=256, ffn_kernel_size=9, dropout=0.1,
                 num_heads=2, ffn_padding="SAME", ffn_act="gelu", padding_idx=0, use_embedding_to_input=False):
        super().__init__()
        if ffn_padding != "SAME":
            raise NotImplementedError("FastSpeech2Encoder: only ffn_padding='SAME' (what every shipped config uses)")
        self.hidden_size, self.num_layers, self.dropout = hidden_size, num_layers, dropout
        self.embed_scale = math.sqrt(hidden_size)
        self.layers = nn.ModuleList([_Layer(hidden_size, num_heads, ffn_kernel_size, dropout, ffn_act)
                                     for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(hidden_size)
        self.proj = nn.Embedding(input_size, hidden_size) if use_embedding_to_input else nn.Linear(input_size, hidden_size)
        self._pe = _reversed_sinusoid_table(5000, hidden_size)       # plain attribute: not part of the state_dict

    def _positions(self, x):
        if self._pe.shape[1] < x.shape[1]:
            self._pe = _reversed_sinusoid_table(x.shape[1], self.hidden_size)
        if self._pe.device != x.device or self._pe.dtype != x.dtype:
            self._pe = self._pe.to(device=x.device, dtype=x.dtype)
        return self._pe[:, :x.shape[1]]

    def forward(self, contents, encoder_padding_mask):
        """contents [B,T,N] float (or [B,T] ids with use_embedding_to_input), encoder_padding_mask [B,T] (True = pad)
        -> [B,T,hidden]."""
        x = self.embed_scale * self.proj(contents)
        x = x * self.embed_scale + self._positions(x)
        keep = 1 - encoder_padding_mask.transpose(0, 1).float()[:, :, None]            # [T,B,1]
        x = x.transpose(0, 1) * keep
        for layer in self.layers:
            x = layer(x, encoder_padding_mask, keep) * keep
        x = self.layer_norm(x) * keep
        return x.transpose(0, 1)
