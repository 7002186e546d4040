"""Mirror of the on-path part of /root/reference/Stage3_source/sampling_analysis.py.

``batch_generate_denoised_sampled`` (:204-265) keeps the reference signature and return contract
but runs the whole step loop on the device (one C-ABI call, no host round trip per step); the
per-step lists the reference builds with ``.cpu().numpy()`` every step (:259-260) are materialised
lazily from one device trajectory buffer."""
from __future__ import annotations

from collections.abc import Sequence
from typing import Optional

import numpy as np
import torch

from . import transformer_training_helper as train_helper


@torch.no_grad()
def predict_next_index(model, args, mask_realization, y_c, idx):
    """Reference signature (:122-147): returns (OneHotCategorical, probs.cpu())."""
    model.eval()
    conditional_prob, probs = train_helper.cond_predict_conditional_prob(
        model, mask_realization.squeeze(1), y_c, idx, args)
    return conditional_prob, probs.cpu()


class _LazyStates(Sequence):
    """``mask_realization_list``: entry i is np.int64 [B, 1, L], the state after step start+i."""

    def __init__(self, traj_u8: np.ndarray):
        self._t = traj_u8            # [T, B, L] uint8 (host)

    def __len__(self):
        return self._t.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return self._t[i].astype(np.int64)[:, None, :]


class _LazyTimes(Sequence):
    """``time_idx_list``: entry i is np.int64 [B, 1] filled with start+i."""

    def __init__(self, start: int, n: int, batch: int):
        self._s, self._n, self._b = start, n, batch

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return np.full((self._b, 1), self._s + i, dtype=np.int64)


def _pinned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if t.is_cuda or t.is_pinned() else t.pin_memory()


@torch.no_grad()
def batch_generate_denoised_sampled(args, model, extract_digit_samples, extract_time, extract_digit_label,
                                    sampling_path, noise: Optional[torch.Tensor] = None,
                                    seed: Optional[int] = None):
    """Reference signature (:204-211) plus two optional keywords:

    noise  explicit Exp(1) draws [T, B*L, C] (parity runs); default: drawn on the device.
    seed   Philox seed for the on-device draws; default: taken from torch's global CPU generator,
           so ``torch.manual_seed`` makes a run repeatable just as it does for the reference.
    """
    assert extract_digit_samples.size(0) == extract_digit_label.size(0) == sampling_path.size(0) == \
        extract_time.size(0), "Mismatched batch dimensions"
    if not hasattr(model, 'engine'):
        raise TypeError('batch_generate_denoised_sampled needs the biom3_b200 model returned by get_model(); '
                        'there is no PyTorch fallback')
    B = extract_digit_samples.size(0)
    eng = model.engine(B)
    dev = eng.device
    L = eng.L
    start = int(extract_time.reshape(-1)[0].item())          # the reference assumes one shared start step
    steps = max(0, int(args.diffusion_steps) - start)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    y_c = _pinned(extract_digit_label.float()).to(dev, non_blocking=True)
    path = _pinned(sampling_path.long()).to(dev, non_blocking=True)
    x0 = extract_digit_samples
    state0 = None
    if start > 0 or bool((x0 != 0).any()):
        state0 = _pinned(x0.long()).to(dev, non_blocking=True)
    if noise is not None:
        noise = _pinned(noise.float()).to(dev, non_blocking=True)
    tokens, traj = eng.decode(y_c, path, state0=state0, start_step=start, num_steps=steps, group=B,
                              noise=noise, seed=seed, want_traj=True)
    host = torch.empty(traj.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(traj, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return _LazyStates(host.numpy()), _LazyTimes(start, steps, B)


def convert_num_to_chars(tokenizer, num_seq):
    """Reference helper (:270-276)."""
    return ''.join(tokenizer[num] for num in num_seq)
