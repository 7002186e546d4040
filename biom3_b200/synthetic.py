"""Synthetic weights and inputs for the ProteoScribe sampling path.

There are no checkpoints or datasets in this environment (no network), so every
test and benchmark runs on random-init weights with the *reference state-dict key
schema* and on synthetic ``z_c`` / sampling paths / Exp(1) noise.

Key schema and shapes follow what ``get_model`` builds in the reference
(/root/reference/Stage3_source/cond_diff_transformer_layer.py:86-146, 198-256) plus
the parameter names of the pinned third-party blocks it instantiates
(linear-attention-transformer==0.19.1, axial-positional-embedding==0.2.1,
/root/reference/requirements.txt:28-29).  Initialisers are the PyTorch defaults
the reference would get without a checkpoint: ``nn.Linear`` weight and bias
~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); ``nn.Embedding`` and the axial tables
~ N(0,1); ``nn.LayerNorm`` weight 1, bias 0.
"""
from __future__ import annotations

import math
from argparse import Namespace
from typing import Dict, List, Tuple

import torch

# id -> symbol table used by the reference CLI
# (/root/reference/run_ProteoScribe_sample.py:88-92)
TOKENS = [
    '-', '<START>', 'A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M',
    'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y', '<END>', '<PAD>',
    'X', 'U', 'Z', 'B', 'O',
]


def stage3_args(**overrides) -> Namespace:
    """Namespace with the stage3_config.json keys the sampling path reads
    (/root/reference/stage3_config.json; SURVEY.md section 5 'Config / flags')."""
    d = dict(
        device='cpu', num_replicas=5, batch_size_sample=32, diffusion_steps=1024,
        image_size=32, num_classes=29, text_emb_dim=512, transformer_dim=512,
        transformer_heads=16, transformer_depth=16, transformer_blocks=1,
        transformer_local_heads=8, transformer_local_size=128,
        transformer_reversible=False, input_dp_rate=0.0, seed=42,
    )
    d.update(overrides)
    return Namespace(**d)


def state_dict_schema(args: Namespace) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, init-kind) for every tensor of the reference model, in the
    order ``model.state_dict()`` would list them."""
    D = args.transformer_dim
    depth = args.transformer_depth
    nb = args.transformer_blocks
    C = args.num_classes
    L = args.diffusion_steps
    W = args.transformer_local_size
    T = args.text_emb_dim
    p = 'transformer.'
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    out.append((p + 'x_emb_NN.weight', (C, D), 'normal'))
    for name, fin in (('y_mlp', T), ('mlp', D)):
        out.append((p + f'{name}.0.weight', (4 * D, fin), f'lin{fin}'))
        out.append((p + f'{name}.0.bias', (4 * D,), f'lin{fin}'))
        out.append((p + f'{name}.2.weight', (D * nb * depth, 4 * D), f'lin{4 * D}'))
        out.append((p + f'{name}.2.bias', (D * nb * depth,), f'lin{4 * D}'))
    out.append((p + 'axial_pos_emb.weights_0', (1, L // W, 1, D), 'normal'))
    out.append((p + 'axial_pos_emb.weights_1', (1, 1, W, D), 'normal'))
    for i in range(nb):
        for j in range(depth):
            q = p + f'transformer_blocks.{i}.{j}.layers.layers.0.'
            out.append((q + '0.norm.weight', (D,), 'ones'))
            out.append((q + '0.norm.bias', (D,), 'zeros'))
            out.append((q + '0.fn.to_q.weight', (D, D), f'lin{D}'))
            out.append((q + '0.fn.to_k.weight', (D, D), f'lin{D}'))
            out.append((q + '0.fn.to_v.weight', (D, D), f'lin{D}'))
            out.append((q + '0.fn.to_out.weight', (D, D), f'lin{D}'))
            out.append((q + '0.fn.to_out.bias', (D,), f'lin{D}'))
            out.append((q + '1.norm.weight', (D,), 'ones'))
            out.append((q + '1.norm.bias', (D,), 'zeros'))
            out.append((q + '1.fn.fn.w1.weight', (4 * D, D), f'lin{D}'))
            out.append((q + '1.fn.fn.w1.bias', (4 * D,), f'lin{D}'))
            out.append((q + '1.fn.fn.w2.weight', (D, 4 * D), f'lin{4 * D}'))
            out.append((q + '1.fn.fn.w2.bias', (D,), f'lin{4 * D}'))
    out.append((p + 'norm.weight', (D,), 'ones'))
    out.append((p + 'norm.bias', (D,), 'zeros'))
    out.append((p + 'out.weight', (C, D), f'lin{D}'))
    out.append((p + 'out.bias', (C,), f'lin{D}'))
    return out


def random_state_dict(args: Namespace, seed: int = 0,
                      perturb_norm: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded random-init fp32 state dict with the reference key schema.

    ``perturb_norm`` moves the LayerNorm affine parameters off their 1/0 defaults
    (tests use it so that a kernel ignoring gamma/beta cannot pass)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_schema(args):
        if kind == 'normal':
            t = torch.randn(shape, generator=g)
        elif kind == 'ones':
            t = torch.ones(shape)
            if perturb_norm:
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif kind == 'zeros':
            t = torch.zeros(shape)
            if perturb_norm:
                t = 0.1 * torch.randn(shape, generator=g)
        else:
            bound = 1.0 / math.sqrt(int(kind[3:]))
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = t.float().contiguous()
    return sd


def synthetic_z_c(num_prompts: int, dim: int = 512, seed: int = 1) -> torch.Tensor:
    """z_c ~ N(0, 0.18^2): the reference README reports |z_c|_2 ~ 3.98 at dim 512
    (/root/reference/README.md:318) => sigma ~ 0.176."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(num_prompts, dim, generator=g) * 0.18


def synthetic_paths(batch: int, steps: int, seed: int = 2) -> torch.Tensor:
    """One random permutation of the positions per sample, int64 [B, L]
    (/root/reference/run_ProteoScribe_sample.py:108)."""
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randperm(steps, generator=g) for _ in range(batch)])


def synthetic_noise(steps: int, batch: int, seq_len: int, num_classes: int,
                    seed: int = 3) -> torch.Tensor:
    """Exp(1) race noise q[step, b*L + l, c], fp32 — what ``torch.multinomial``'s
    single-draw fast path consumes inside ``OneHotCategorical.sample()``
    (/root/reference/Stage3_source/sampling_analysis.py:251)."""
    g = torch.Generator().manual_seed(seed)
    return torch.empty(steps, batch * seq_len, num_classes).exponential_(1.0, generator=g)


def facilitator_state_dict(emb_dim: int = 512, hid_dim: int = 1024, seed: int = 31) -> Dict[str, torch.Tensor]:
    """Seeded Facilitator weights with the reference keys (weight_norm(Linear, dim=None):
    main.{0,3}.{bias, weight_g (scalar), weight_v}); weight_g is deliberately != ||weight_v|| so the fold matters."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, (o, i) in (('main.0', (hid_dim, emb_dim)), ('main.3', (emb_dim, hid_dim))):
        sd[name + '.bias'] = torch.randn(o, generator=g) * 0.3
        sd[name + '.weight_g'] = torch.randn((), generator=g) * 0.3 + 2.0
        sd[name + '.weight_v'] = torch.randn(o, i, generator=g) * 0.05
    return sd
