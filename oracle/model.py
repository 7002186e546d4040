"""ORACLE (test infrastructure, never on the product path).

CPU fp32 restatement of the ProteoScribe per-step forward, written functionally over
a flat ``state_dict`` with the reference key schema.  Follows

  /root/reference/Stage3_source/cond_diff_transformer_layer.py:30-42   (sinusoidal time embedding)
  /root/reference/Stage3_source/cond_diff_transformer_layer.py:93-105  (y_mlp / mlp conditioning MLPs)
  /root/reference/Stage3_source/cond_diff_transformer_layer.py:149-176 (forward: embed, axial pos,
        per-layer additive conditioning with the (d*depth + j) flat layout, final LN + head, permute)
  /root/reference/Stage3_source/cond_diff_transformer_layer.py:198-256 (get_model shapes)

and, for the un-vendored transformer block, oracle/upstream_blocks.py (PARITY UNPINNED
there, see its header).  The in-tree part IS pinned: tests/golden/make_golden.py runs
the real reference files (with upstream_blocks standing in for the missing wheels) and
tests/test_oracle_vs_reference.py checks this file against those outputs bit for bit.
"""
from __future__ import annotations

import math
from argparse import Namespace
from typing import Dict

import torch
import torch.nn.functional as F

from .upstream_blocks import LocalAttention, linear_attention


def time_embedding(t: torch.Tensor, dim: int, num_steps: int, rescale: float = 4000.0) -> torch.Tensor:
    """SinusoidalPosEmb (cond_diff_transformer_layer.py:30-42). t: int64 [B] -> fp32 [B, dim]."""
    x = t / float(num_steps) * float(rescale)
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -e)
    arg = x[:, None] * freqs[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def conditioning(sd: Dict[str, torch.Tensor], name: str, inp: torch.Tensor) -> torch.Tensor:
    """Linear -> Softplus -> Linear (cond_diff_transformer_layer.py:93-105). -> [B, D*nb*depth]."""
    p = 'transformer.' + name
    h = F.linear(inp, sd[p + '.0.weight'], sd[p + '.0.bias'])
    h = F.softplus(h)
    return F.linear(h, sd[p + '.2.weight'], sd[p + '.2.bias'])


class OracleModel:
    """forward(x, t, y_c) -> logits [B, C, L], fp32, same semantics as the reference
    ``DiffTransformer.forward`` (cond_diff_transformer_layer.py:249-251)."""

    def __init__(self, args: Namespace, state_dict: Dict[str, torch.Tensor], scale_q_first: bool = True):
        assert args.transformer_blocks == 1 and not args.transformer_reversible
        self.args = args
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.D = args.transformer_dim
        self.H = args.transformer_heads
        self.depth = args.transformer_depth
        self.NL = args.transformer_local_heads
        self.W = args.transformer_local_size
        self.L = args.diffusion_steps
        self.local = LocalAttention(self.W, scale_q_first=scale_q_first)

    # -- pieces (exposed so kernel unit tests can compare stage by stage) ---------
    def embed(self, x: torch.Tensor) -> torch.Tensor:
        sd, D, W = self.sd, self.D, self.W
        B, L = x.shape
        tok = F.embedding(x.long(), sd['transformer.x_emb_NN.weight'])
        w0 = sd['transformer.axial_pos_emb.weights_0']   # [1, L/W, 1, D]
        w1 = sd['transformer.axial_pos_emb.weights_1']   # [1, 1,   W, D]
        pos = (0 + w0.expand(B, self.L // W, W, D).reshape(B, self.L, D)) \
            + w1.expand(B, self.L // W, W, D).reshape(B, self.L, D)
        pos = pos[:, :L]
        xe = tok + pos
        return torch.zeros_like(xe) + xe

    def cond_vectors(self, t: torch.Tensor, y_c: torch.Tensor):
        """-> (T, Y), each [B, D, depth]; layer j uses [..., j] (flat index d*depth + j)."""
        B = t.shape[0]
        te = time_embedding(t, self.D, self.L).float()
        T = conditioning(self.sd, 'mlp', te).reshape(B, self.D, self.depth)
        Y = conditioning(self.sd, 'y_mlp', y_c.float()).reshape(B, self.D, self.depth)
        return T, Y

    def block(self, j: int, u: torch.Tensor) -> torch.Tensor:
        sd, D, H, NL = self.sd, self.D, self.H, self.NL
        dh = D // H
        B, L, _ = u.shape
        p = f'transformer.transformer_blocks.0.{j}.layers.layers.0.'
        a = F.layer_norm(u, (D,), sd[p + '0.norm.weight'], sd[p + '0.norm.bias'], 1e-5)
        q = F.linear(a, sd[p + '0.fn.to_q.weight'])
        k = F.linear(a, sd[p + '0.fn.to_k.weight'])
        v = F.linear(a, sd[p + '0.fn.to_v.weight'])
        q, k, v = (z.reshape(B, L, H, dh).transpose(1, 2) for z in (q, k, v))
        outs = []
        if NL > 0:
            outs.append(self.local(q[:, :NL], k[:, :NL], v[:, :NL]))
        if H - NL > 0:
            outs.append(linear_attention(q[:, NL:], k[:, NL:], v[:, NL:]))
        o = torch.cat(outs, dim=1).transpose(1, 2).reshape(B, L, D)
        u = u + F.linear(o, sd[p + '0.fn.to_out.weight'], sd[p + '0.fn.to_out.bias'])
        m = F.layer_norm(u, (D,), sd[p + '1.norm.weight'], sd[p + '1.norm.bias'], 1e-5)
        f = F.linear(m, sd[p + '1.fn.fn.w1.weight'], sd[p + '1.fn.fn.w1.bias'])
        f = F.gelu(f)
        f = F.linear(f, sd[p + '1.fn.fn.w2.weight'], sd[p + '1.fn.fn.w2.bias'])
        return u + f

    def head(self, h: torch.Tensor) -> torch.Tensor:
        sd, D = self.sd, self.D
        h = F.layer_norm(h, (D,), sd['transformer.norm.weight'], sd['transformer.norm.bias'], 1e-5)
        return F.linear(h, sd['transformer.out.weight'], sd['transformer.out.bias'])   # [B, L, C]

    # -- whole forward --------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, t: torch.Tensor, y_c: torch.Tensor, return_hidden: bool = False):
        T, Y = self.cond_vectors(t.reshape(-1), y_c)
        h = self.embed(x)
        for j in range(self.depth):
            h = self.block(j, h + T[:, None, :, j] + Y[:, None, :, j])
        logits = self.head(h)
        out = logits.permute(0, 2, 1)
        return (out, h) if return_hidden else out

    __call__ = forward

    def eval(self):
        return self

    def to(self, *a, **k):
        return self
