"""Mirror of ``Facilitator`` (/root/reference/Stage1_source/model.py:473-493): same constructor, same
state-dict keys (``main.0.{bias,weight_g,weight_v}``, ``main.3.{...}`` as produced by
``weight_norm(nn.Linear(...), dim=None)``), ``forward(z_t) -> z_c`` on the CUDA engine
(``biom3_facilitator``).  The MMD/MSE losses of the reference class are training-only and out of scope."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib


class _WNLinear(nn.Module):
    """Parameter holder with the key names of weight_norm(nn.Linear(i, o), dim=None)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        bound = 1.0 / math.sqrt(in_dim)
        v = (torch.rand(out_dim, in_dim) * 2 - 1) * bound
        self.bias = nn.Parameter((torch.rand(out_dim) * 2 - 1) * bound, requires_grad=False)
        self.weight_g = nn.Parameter(v.norm().clone(), requires_grad=False)       # scalar: W == V at init
        self.weight_v = nn.Parameter(v, requires_grad=False)


class Facilitator(nn.Module):
    def __init__(self, in_dim: int, hid_dim: int, out_dim: int, dropout: float = 0.):
        super().__init__()
        self.in_dim, self.hid_dim, self.out_dim = in_dim, hid_dim, out_dim
        self.main = nn.Module()
        self.main.add_module('0', _WNLinear(in_dim, hid_dim))
        self.main.add_module('3', _WNLinear(hid_dim, out_dim))
        self._handle = None          # biom3_facilitator_t*: folded weights resident on one device
        self._handle_key = None      # (device index, parameter versions) the handle was built from

    def _params(self):
        l0, l1 = self.main._modules['0'], self.main._modules['3']
        return (l0.weight_v, l0.weight_g, l0.bias, l1.weight_v, l1.weight_g, l1.bias)

    def _release(self):
        if self._handle is not None:
            _lib.load().biom3_facilitator_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _engine(self, dev: torch.device):
        """Fold weight_norm and upload once per (device, parameter version); load_state_dict / in-place edits bump
        the tensors' version counters and so rebuild the handle."""
        ps = self._params()
        key = (dev.index, tuple((p.data_ptr(), p._version) for p in ps))
        if self._handle is not None and self._handle_key == key:
            return self._handle
        self._release()
        lib = _lib.load()
        host = [p.detach().to('cpu', torch.float32).contiguous() for p in ps]
        h = C.c_void_p()
        _lib.check(lib.biom3_facilitator_create(
            self.in_dim, self.hid_dim, self.out_dim,
            C.c_void_p(host[0].data_ptr()), C.c_float(float(host[1])), C.c_void_p(host[2].data_ptr()),
            C.c_void_p(host[3].data_ptr()), C.c_float(float(host[4])), C.c_void_p(host[5].data_ptr()),
            dev.index, C.byref(h)))
        self._handle, self._handle_key = h, key
        return h

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not torch.cuda.is_available():
            raise RuntimeError('biom3_b200 has no CPU path: Facilitator.forward needs a CUDA device')
        lib = _lib.load()
        dev = x.device if x.is_cuda else torch.device('cuda', torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        z_t = x.to(dev, torch.float32).contiguous()
        if z_t.dim() != 2 or z_t.shape[1] != self.in_dim:
            raise ValueError(f'expected z_t [P, {self.in_dim}], got {tuple(z_t.shape)}')
        P = z_t.shape[0]
        z_c = torch.empty(P, self.out_dim, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.biom3_facilitator_forward(self._engine(dev), C.c_void_p(z_t.data_ptr()), P,
                                                     C.c_void_p(z_c.data_ptr()),
                                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return z_c if x.is_cuda else z_c.cpu()
