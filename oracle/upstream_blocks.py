"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the third-party blocks the reference model instantiates but
does not vendor:

  * ``linear-attention-transformer==0.19.1``  (``LinearAttentionTransformer``)
  * ``axial-positional-embedding==0.2.1``     (``AxialPositionalEmbedding``)
  * ``local-attention`` (transitive, unpinned) (``LocalAttention``)

pinned at /root/reference/requirements.txt:28-29 and called from
/root/reference/Stage3_source/cond_diff_transformer_layer.py:108-113 (axial table),
:123-143 (one ``LinearAttentionTransformer(depth=1)`` per layer) and :158, :171.

PARITY UNPINNED for this file: none of those packages is installed here, there is
no network, and the reference ships no tests or golden vectors, so these blocks are
written from the packages' published behaviour
and cannot be checked against the real wheels offline.  The module/attribute names
mirror upstream so that ``state_dict()`` keys equal the reference checkpoint schema,
which also lets tests/golden/make_golden.py import the *real* in-tree reference
files with these classes standing in for the missing wheels.

Known ambiguity A1: ``local-attention`` is unpinned; older
releases compute ``(q . k^T) * scale``, newer ones ``(q * scale) . k^T``.  The two
differ only in fp32 rounding; ``LocalAttention(scale_q_first=...)`` selects, default
is the newer behaviour.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class AxialPositionalEmbedding(nn.Module):
    """Summed axial table: pos[p] = weights_0[0, p // W, 0] + weights_1[0, 0, p % W]."""

    def __init__(self, dim, axial_shape, axial_dims=None):
        super().__init__()
        assert axial_dims is None, "only the summed mode is on the hot path"
        self.dim = dim
        self.shape = tuple(axial_shape)
        self.max_seq_len = 1
        for s in self.shape:
            self.max_seq_len *= s
        for ind, s in enumerate(self.shape):
            ax_shape = [1] * len(self.shape)
            ax_shape[ind] = s
            p = nn.Parameter(torch.zeros(1, *ax_shape, dim).normal_(0, 1))
            setattr(self, f'weights_{ind}', p)

    def forward(self, x):
        b, t, _ = x.shape
        assert t <= self.max_seq_len
        total = 0
        for ind in range(len(self.shape)):
            w = getattr(self, f'weights_{ind}')
            total = total + w.expand(b, *self.shape, self.dim).reshape(b, self.max_seq_len, self.dim)
        return total[:, :t].to(x)


def linear_attention(q, k, v):
    """Non-causal linear attention on [B, H, N, dh] tensors, no mask."""
    dh = q.shape[-1]
    q = q.softmax(dim=-1)          # over the dh features of each token
    k = k.softmax(dim=-2)          # over the N tokens, per feature
    q = q * dh ** -0.5
    ctx = torch.einsum('bhnd,bhne->bhde', k, v)
    out = torch.einsum('bhnd,bhde->bhne', q, ctx)
    return out.reshape(*q.shape)


def _look_around(x, backward, forward, pad_value):
    """x: [b, windows, W, ...] -> [b, windows, (backward+1+forward)*W, ...]."""
    windows = x.shape[1]
    dims = (len(x.shape) - 2) * (0, 0)
    padded = F.pad(x, (*dims, backward, forward), value=pad_value)
    parts = [padded[:, i:i + windows] for i in range(backward + forward + 1)]
    return torch.cat(parts, dim=2)


class LocalAttention(nn.Module):
    """Non-causal windowed softmax attention, look one window back and one forward."""

    def __init__(self, window_size, causal=False, look_backward=1, look_forward=None,
                 dropout=0., scale_q_first=True):
        super().__init__()
        assert not causal, "the sampling path is non-causal"
        self.window_size = window_size
        self.look_backward = look_backward
        self.look_forward = 1 if look_forward is None else look_forward
        self.scale_q_first = scale_q_first
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, input_mask=None):
        assert input_mask is None
        shape = q.shape
        q, k, v = (t.reshape(-1, *t.shape[-2:]) for t in (q, k, v))
        b, n, e = q.shape
        W = self.window_size
        assert n % W == 0
        windows = n // W
        pos = torch.arange(n, device=q.device, dtype=q.dtype).reshape(1, windows, W)
        bq, bk, bv = (t.reshape(b, windows, W, e) for t in (q, k, v))
        bk = _look_around(bk, self.look_backward, self.look_forward, -1.)
        bv = _look_around(bv, self.look_backward, self.look_forward, -1.)
        key_pos = _look_around(pos, self.look_backward, self.look_forward, -1.)
        scale = e ** -0.5
        if self.scale_q_first:
            dots = torch.einsum('bhie,bhje->bhij', bq * scale, bk)
        else:
            dots = torch.einsum('bhie,bhje->bhij', bq, bk) * scale
        mask = key_pos[:, :, None, :] == -1
        dots = dots.masked_fill(mask, -torch.finfo(dots.dtype).max)
        attn = self.dropout(dots.softmax(dim=-1))
        out = torch.einsum('bhij,bhje->bhie', attn, bv)
        return out.reshape(-1, n, e).reshape(*shape)


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)


class Chunk(nn.Module):
    """ff_chunks == 1 on the hot path: a pass-through wrapper (keeps the key ``fn.fn``)."""

    def __init__(self, chunks, fn, along_dim=-1):
        super().__init__()
        assert chunks == 1
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(x, **kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.):
        super().__init__()
        self.w1 = nn.Linear(dim, dim * mult)
        self.act = nn.GELU()          # approximate='none' (erf)
        self.dropout = nn.Dropout(dropout)
        self.w2 = nn.Linear(dim * mult, dim)

    def forward(self, x, **kwargs):
        return self.w2(self.dropout(self.act(self.w1(x))))


class SelfAttention(nn.Module):
    """Heads [0, n_local) use LocalAttention, heads [n_local, heads) linear attention;
    outputs are concatenated in head order before ``to_out``."""

    def __init__(self, dim, heads, n_local_attn_heads, local_attn_window_size,
                 dropout=0., attn_dropout=0., scale_q_first=True):
        super().__init__()
        assert dim % heads == 0
        self.heads = heads
        self.d_heads = dim // heads
        self.local_attn_heads = n_local_attn_heads
        self.local_attn = LocalAttention(local_attn_window_size, causal=False,
                                         dropout=attn_dropout, scale_q_first=scale_q_first)
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, **kwargs):
        b, t, _ = x.shape
        h, dh, nl = self.heads, self.d_heads, self.local_attn_heads
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        q, k, v = (z.reshape(b, t, h, dh).transpose(1, 2) for z in (q, k, v))
        outs = []
        if nl > 0:
            outs.append(self.local_attn(q[:, :nl], k[:, :nl], v[:, :nl]))
        if h - nl > 0:
            outs.append(linear_attention(q[:, nl:], k[:, nl:], v[:, nl:]))
        attn = torch.cat(outs, dim=1).transpose(1, 2).reshape(b, t, -1)
        return self.dropout(self.to_out(attn))


class SequentialSequence(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = layers

    def forward(self, x, **kwargs):
        for f, g in self.layers:
            x = x + f(x)
            x = x + g(x)
        return x


class LinearAttentionTransformer(nn.Module):
    """The constructor signature matches the call at
    /root/reference/Stage3_source/cond_diff_transformer_layer.py:123-143; every option
    the sampling config does not use is asserted to hold its hot-path value."""

    def __init__(self, dim, depth, max_seq_len, heads=8, dim_head=None, bucket_size=64,
                 causal=False, ff_chunks=1, ff_glu=False, ff_dropout=0.,
                 attn_layer_dropout=0., attn_dropout=0., reversible=False, blindspot_size=1,
                 n_local_attn_heads=0, local_attn_window_size=128, receives_context=False,
                 attend_axially=False, pkm_layers=tuple(), pkm_num_keys=128,
                 linformer_settings=None, context_linformer_settings=None,
                 shift_tokens=False, scale_q_first=True):
        super().__init__()
        assert dim_head is None and not causal and not ff_glu and not reversible
        assert not receives_context and not attend_axially and not shift_tokens
        assert linformer_settings is None and context_linformer_settings is None
        assert len(tuple(pkm_layers)) == 0
        if not isinstance(n_local_attn_heads, tuple):
            n_local_attn_heads = (n_local_attn_heads,) * depth
        assert len(n_local_attn_heads) == depth
        self.max_seq_len = max_seq_len
        self.pad_multiple = local_attn_window_size if any(n_local_attn_heads) else 1
        layers = nn.ModuleList([])
        for nl in n_local_attn_heads:
            attn = SelfAttention(dim, heads, nl, local_attn_window_size,
                                 dropout=attn_layer_dropout, attn_dropout=attn_dropout,
                                 scale_q_first=scale_q_first)
            ff = Chunk(ff_chunks, FeedForward(dim, dropout=ff_dropout), along_dim=1)
            layers.append(nn.ModuleList([PreNorm(dim, attn), PreNorm(dim, ff)]))
        self.layers = SequentialSequence(layers)

    def forward(self, x, **kwargs):
        t = x.shape[1]
        assert t % self.pad_multiple == 0, "sequence length must be a multiple of the window"
        return self.layers(x)[:, :t]
