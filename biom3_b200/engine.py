"""Host-side handle on the CUDA engine: owns a biom3_model*, feeds it weights by reference
state-dict key, and exposes forward / decode on torch CUDA tensors (torch is plumbing here: device
memory, streams)."""
from __future__ import annotations

import ctypes as C
import os
from argparse import Namespace
from typing import Dict, Optional

import torch

from . import _lib


def config_from_args(args: Namespace) -> _lib.Config:
    """stage3_config.json Namespace -> biom3_config (include/biom3_b200.h)."""
    return _lib.Config(
        seq_len=int(args.diffusion_steps), dim=int(args.transformer_dim), heads=int(args.transformer_heads),
        depth=int(args.transformer_depth), n_blocks=int(args.transformer_blocks),
        local_heads=int(args.transformer_local_heads), local_window=int(args.transformer_local_size),
        num_classes=int(args.num_classes), text_emb_dim=int(args.text_emb_dim),
        reversible=int(bool(getattr(args, 'transformer_reversible', False))))


_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16,
           torch.float64: _lib.DTYPE_F64}


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    PRECISIONS = {'bf16': 0, 'fp32': 1}

    def __init__(self, args: Namespace, state_dict: Dict[str, torch.Tensor], device: torch.device, max_batch: int,
                 precision: Optional[str] = None):
        """precision: 'bf16' (default) or 'fp32' (fp32-class arithmetic, see biom3_set_precision); when None it is
        taken from ``args.b200_precision`` or the BIOM3_PRECISION environment variable."""
        if precision is None:
            precision = getattr(args, 'b200_precision', None) or os.environ.get('BIOM3_PRECISION', 'bf16')
        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(self.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        if not torch.cuda.is_available():
            raise RuntimeError('biom3_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.cfg = config_from_args(args)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError(f'biom3_b200 engine cannot run on device {device}')
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.max_batch = int(max_batch)
        self.L = self.cfg.seq_len
        self.C = self.cfg.num_classes
        self.E = self.cfg.text_emb_dim
        h = C.c_void_p()
        _lib.check(self.lib.biom3_create(C.byref(self.cfg), self.dev_index, self.max_batch, C.byref(h)))
        self.handle = h
        try:
            _lib.check(self.lib.biom3_set_precision(self.handle, self.PRECISIONS[precision]))
            for key, t in state_dict.items():
                # host or device memory, fp32 / bf16 / fp16 / fp64 as stored: the library stages and converts
                t = t.detach()
                if t.dtype not in _DTYPES:
                    t = t.float()
                t = t.contiguous()
                shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
                _lib.check(self.lib.biom3_set_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), _DTYPES[t.dtype],
                                                     shape, t.dim()))
            _lib.check(self.lib.biom3_finalize_weights(self.handle))
        except Exception:
            self.close()
            raise

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle:
            self.lib.biom3_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches_per_step(self) -> int:
        return int(self.lib.biom3_launches_per_step(self.handle))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def check_inputs(self) -> None:
        """Raises IndexError if a forward / decode since the last check was given a token id >= num_classes (what the
        reference's nn.Embedding raises for), a time index outside [0, L) or a path entry outside [0, L).  The kernels
        clamp such values instead of indexing out of bounds and record them; this call synchronises the device."""
        flags = C.c_int32(0)
        _lib.check(self.lib.biom3_input_errors(self.handle, C.byref(flags)))
        if flags.value:
            what = [msg for bit, msg in ((1, f'token id outside [0, {self.C})'), (2, f'time index outside [0, {self.L})'),
                                         (4, f'sampling path entry outside [0, {self.L})')) if flags.value & bit]
            raise IndexError('biom3_b200: ' + ', '.join(what))

    def forward(self, x: torch.Tensor, t: torch.Tensor, y_c: torch.Tensor) -> torch.Tensor:
        """x int [B, L], t int [B], y_c [B, E] -> logits fp32 [B, C, L] (reference layout)."""
        B = x.shape[0]
        assert x.shape == (B, self.L) and t.numel() == B and y_c.shape == (B, self.E)
        x = x.to(self.device, torch.int64).contiguous()
        t = t.reshape(-1).to(self.device, torch.int64).contiguous()
        y_c = y_c.to(self.device, torch.float32).contiguous()
        out = torch.empty(B, self.C, self.L, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            torch.cuda.nvtx.range_push(f'biom3_forward B={B}')           # NVTX: one range per C-ABI call
            try:
                _lib.check(self.lib.biom3_forward(self.handle, _ptr(x), _ptr(t), _ptr(y_c), B, _ptr(out), self._stream()))
            finally:
                torch.cuda.nvtx.range_pop()
        return out

    def decode(self, y_c: torch.Tensor, path: torch.Tensor, state0: Optional[torch.Tensor] = None,
               start_step: int = 0, num_steps: Optional[int] = None, group: Optional[int] = None,
               noise: Optional[torch.Tensor] = None, seed: int = 0, want_traj: bool = False,
               group_seeds: Optional[torch.Tensor] = None):
        """All tensors already on the device.  Returns (tokens int64 [B, L], traj uint8 [T, B, L] | None).
        group_seeds: int64 [B / group] (cuda), one Philox seed per reference batch of a fused launch."""
        B = path.shape[0]
        L = self.L
        if num_steps is None:
            num_steps = L - start_step
        if group is None:
            group = B
        assert path.shape == (B, L) and path.dtype == torch.int64 and path.is_cuda and path.is_contiguous()
        assert y_c.shape == (B, self.E) and y_c.dtype == torch.float32 and y_c.is_cuda and y_c.is_contiguous()
        if state0 is not None:
            assert state0.shape == (B, L) and state0.dtype == torch.int64 and state0.is_cuda and state0.is_contiguous()
        if noise is not None:
            assert noise.dtype == torch.float32 and noise.is_cuda and noise.is_contiguous()
            assert noise.shape == (num_steps, B * L, self.C), (noise.shape, (num_steps, B * L, self.C))
        if group_seeds is not None:
            assert group_seeds.shape == (B // group,) and group_seeds.dtype == torch.int64 and group_seeds.is_cuda
            group_seeds = group_seeds.contiguous()
        tokens = torch.empty(B, L, device=self.device, dtype=torch.int64)
        traj = torch.empty(num_steps, B, L, device=self.device, dtype=torch.uint8) if want_traj else None
        with torch.cuda.device(self.device):
            torch.cuda.nvtx.range_push(f'biom3_decode B={B} group={group} steps={num_steps}')
            try:
                _lib.check(self.lib.biom3_decode(
                    self.handle, _ptr(y_c), _ptr(path), _ptr(state0), int(start_step), int(num_steps), int(group),
                    _ptr(noise), C.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(group_seeds), _ptr(tokens), _ptr(traj), B,
                    self._stream()))
            finally:
                torch.cuda.nvtx.range_pop()
        return tokens, traj

    def debug_buffer(self, name: str, shape, dtype: torch.dtype) -> torch.Tensor:
        """Test hook: host copy of an internal buffer (leading part, sized by shape/dtype)."""
        out = torch.empty(shape, dtype=dtype)
        _lib.check(self.lib.biom3_debug_copy(self.handle, name.encode(), C.c_void_p(out.data_ptr()),
                                             out.numel() * out.element_size()))
        return out

    def profile_step(self, B: int, group: int) -> Dict[str, float]:
        prof = _lib.StepProfile()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.biom3_profile_step(self.handle, B, group, C.byref(prof)))
        return {n: getattr(prof, n) for n, _ in _lib.StepProfile._fields_}


def random_paths(B: int, L: int, seed: int, device) -> torch.Tensor:
    """One uniformly random permutation of 0..L-1 per row, drawn on the device -> int64 [B, L] (cuda)."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('biom3_b200 has no CPU path: random_paths needs a CUDA device')
    path = torch.empty(B, L, device=device, dtype=torch.int64)
    with torch.cuda.device(device):
        _lib.check(lib.biom3_random_paths(C.c_uint64(int(seed) & (2 ** 64 - 1)), B, L, _ptr(path),
                                          C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return path


def device_noise(seed: int, step: int, B: int, L: int, C_: int, device) -> torch.Tensor:
    """Test hook (biom3_debug_noise): the Exp(1) draws a decode with ``noise=None`` consumes at time index ``step`` under
    ``seed`` -> fp32 [B*L, C] on the device."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('biom3_b200 has no CPU path: device_noise needs a CUDA device')
    out = torch.empty(B * L, C_, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(lib.biom3_debug_noise(C.c_uint64(int(seed) & (2 ** 64 - 1)), int(step), B, L, C_, _ptr(out),
                                         C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return out


def sample_all(logits: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """K11: token at every position.  logits fp32 [B, C, L], noise fp32 [B*L, C] (cuda) -> int64 [B, L]."""
    lib = _lib.load()
    B, Cc, L = logits.shape
    assert logits.is_cuda and noise.is_cuda and noise.shape == (B * L, Cc)
    logits = logits.contiguous().float()
    noise = noise.contiguous().float()
    tok = torch.empty(B, L, device=logits.device, dtype=torch.int64)
    with torch.cuda.device(logits.device):
        _lib.check(lib.biom3_sample_all(_ptr(logits), _ptr(noise), _ptr(tok), B, L, Cc,
                                        C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)))
    return tok


def unmask_(state: torch.Tensor, tok: torch.Tensor, path: torch.Tensor, step: int, group: Optional[int] = None) -> None:
    """K12 in place: state[b', loc[b]] = tok[b', loc[b]] for all b', b of a group (int64 [B, L], cuda)."""
    lib = _lib.load()
    B, L = state.shape
    group = B if group is None else group
    assert state.is_cuda and state.dtype == torch.int64 and state.is_contiguous()
    tok = tok.to(torch.int64).contiguous()
    path = path.to(state.device, torch.int64).contiguous()
    with torch.cuda.device(state.device):
        _lib.check(lib.biom3_unmask(_ptr(tok), _ptr(path), _ptr(state), B, L, group, int(step),
                                    C.c_void_p(torch.cuda.current_stream(state.device).cuda_stream)))


def gemm_test(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], epi: int, block_n: int = 256,
              out: Optional[torch.Tensor] = None, pair: bool = False, split3: bool = False) -> torch.Tensor:
    """Unit-test hook: A bf16 [M, K], W bf16 [N, K] -> out per `epi` (see include/biom3_b200.h).
    split3: A and W are [hi | lo] bf16 halves, [M, 2K] and [N, 2K] (the fp32-class K schedule)."""
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    if split3:
        K //= 2
    if out is None:
        if epi in (5, 6):
            raise ValueError('epilogues 5 and 6 update `out` in place: bf16 [2, M, N] (hi plane, lo plane)')
        out = torch.empty(M, 2 * N if epi == 7 else N, device=A.device, dtype=torch.float32 if epi in (3, 4) else torch.bfloat16)
    with torch.cuda.device(A.device):
        _lib.check(lib.biom3_gemm_test(_ptr(A), _ptr(W), _ptr(bias), _ptr(out), M, N, K, epi, block_n, int(pair) | (2 if split3 else 0),
                                       C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)))
    return out


def attention_test(qkv: torch.Tensor, NL: int, variant: int = 0) -> torch.Tensor:
    """Unit-test hook: qkv bf16 [3, B, H, L, 32] (cuda) -> attention output bf16 [B*L, H*32]; variant bit 0 also records
    the local-attention kernel's clock64 timeline (biom3_debug_trace), bit 1: q of the linear heads is already
    softmax(q) over the features (the form the QKV GEMM epilogue writes in the decode)."""
    lib = _lib.load()
    _, B, H, L, dh = qkv.shape
    assert dh == 32 and qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    out = torch.zeros(B * L, H * 32, device=qkv.device, dtype=torch.bfloat16)
    with torch.cuda.device(qkv.device):
        _lib.check(lib.biom3_attention_test(_ptr(qkv), _ptr(out), B, H, L, NL, variant,
                                            C.c_void_p(torch.cuda.current_stream(qkv.device).cuda_stream)))
    return out
