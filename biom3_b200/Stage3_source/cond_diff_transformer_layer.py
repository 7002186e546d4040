"""Mirror of /root/reference/Stage3_source/cond_diff_transformer_layer.py for sampling.

``get_model(args, data_shape, num_classes)`` (:198-256) returns an ``nn.Module`` whose
``state_dict()`` keys and shapes equal the reference's (so ``load_state_dict(torch.load(...))``,
``.eval()``, ``.to(device)`` work unchanged) and whose ``forward(x, t, y_c)`` returns logits
``[B, num_classes, L]`` (:149-176, :249-251) computed by the sm_100a engine.  There is no PyTorch
implementation of the forward here: on a machine without the CUDA library it raises."""
from __future__ import annotations

from argparse import Namespace
from typing import Optional

import torch
import torch.nn as nn

from .. import synthetic
from ..engine import Engine


def add_model_args(parser):
    """Same flags as the reference (:179-196); unused by sampling."""
    parser.add_argument('--num_steps', type=int, default=1)
    parser.add_argument('--actnorm', type=eval, default=False)
    parser.add_argument('--perm_channel', type=str, default='none', choices={'conv', 'shuffle', 'none'})
    parser.add_argument('--perm_length', type=str, default='reverse', choices={'reverse', 'none'})
    parser.add_argument('--input_dp_rate', type=float, default=0.0)
    parser.add_argument('--transformer_dim', type=int, default=512)
    parser.add_argument('--transformer_heads', type=int, default=16)
    parser.add_argument('--transformer_depth', type=int, default=16)
    parser.add_argument('--transformer_blocks', type=int, default=1)
    parser.add_argument('--transformer_dropout', type=float, default=0.1)
    parser.add_argument('--transformer_reversible', type=eval, default=False)
    parser.add_argument('--transformer_local_heads', type=int, default=8)
    parser.add_argument('--transformer_local_size', type=int, default=128)


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    parts = dotted.split('.')
    mod = root
    for name in parts[:-1]:
        if name not in mod._modules:
            mod.add_module(name, nn.Module())
        mod = mod._modules[name]
    mod.register_parameter(parts[-1], param)


class DiffTransformer(nn.Module):
    """Parameter container with the reference key schema + the CUDA engine behind ``forward``."""

    def __init__(self, args: Namespace, num_classes: int):
        super().__init__()
        a = Namespace(**vars(args))
        a.num_classes = num_classes
        self.sampler_args = a
        sd = synthetic.random_state_dict(a, seed=torch.initial_seed() % (2 ** 31))
        for key, t in sd.items():
            _attach(self, key, nn.Parameter(t, requires_grad=False))
        self._engine: Optional[Engine] = None
        self._engine_batch = 0

    # weights changed -> the engine copy is stale
    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._drop_engine()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._drop_engine()
        return out

    def _drop_engine(self):
        if getattr(self, '_engine', None) is not None:
            self._engine.close()
        self._engine = None
        self._engine_batch = 0

    def engine(self, batch: int) -> Engine:
        """The CUDA engine, (re)built for at least ``batch`` sequences of workspace."""
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError(
                'biom3_b200 has no CPU path: move the model to a CUDA device (model.to("cuda")) before '
                'calling forward / sampling')
        if self._engine is None or batch > self._engine_batch:
            self._drop_engine()
            want = max(batch, int(getattr(self.sampler_args, 'batch_size_sample', 1) or 1))
            self._engine = Engine(self.sampler_args, self.state_dict(), dev, want)
            self._engine_batch = want
        return self._engine

    @torch.no_grad()
    def forward(self, x, t, y_c):
        return self.engine(x.shape[0]).forward(x, t, y_c)


def get_model(args, data_shape, num_classes):
    """Reference signature (:198); ``data_shape`` is ignored for L exactly as in the reference
    (L = args.diffusion_steps, :212-213)."""
    L = args.diffusion_steps
    print('Data shape index 0:', L)
    return DiffTransformer(args, num_classes)
