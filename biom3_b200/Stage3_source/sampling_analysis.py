"""Mirror of /root/reference/Stage3_source/sampling_analysis.py: the batch sampler on the hot path plus the
partial-start / inpainting entry points around it.

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
def cond_autocomplete_real_samples(model, args, realization, y_c, idx):
    """One-shot fill-in of the masked positions (:21-61): draws a random path per sample, masks every position
    whose path index is >= idx, runs ONE forward and scores the true tokens.  Returns, like the reference,
    (OneHotCategorical, probs.cpu() [B, C, L], masked tokens, true tokens, log_prob [B, L], paths, mask)."""
    model.eval()
    bs, channel, seq_length = realization.size()
    sampled_random_path = train_helper.sample_random_path(bs, seq_length, device=args.device)
    idx = idx.to(sampled_random_path.device)
    random_path_mask = train_helper.create_mask_at_random_path_index(sampled_random_path, idx, bs, seq_length)
    real_tokens, bs, seq_length = train_helper.create_token_labels(args, realization.to(sampled_random_path.device))
    real_token_masked = train_helper.mask_realizations(real_tokens, random_path_mask)
    conditional_prob, probs = train_helper.cond_predict_conditional_prob(model, real_token_masked, y_c, idx, args)
    log_prob = train_helper.log_prob_of_realization(args, conditional_prob, real_tokens.to(probs.device))
    return (conditional_prob, probs.cpu(), real_token_masked.cpu(), real_tokens.cpu(), log_prob.cpu(),
            sampled_random_path.cpu(), random_path_mask.cpu())


def extract_samples_with_labels(dataloader, target_labels: int, total_num: int, pad_included: bool = False) -> dict:
    """First ``total_num`` samples of a (data, labels) loader whose label equals ``target_labels`` (:65-93); ids
    shift by one to make room for the absorbing state unless the data already includes it."""
    extracted = {'sample': [], 'label': []}
    for data, labels in dataloader:
        for i, label in enumerate(labels):
            if label.item() == target_labels:
                if not pad_included:
                    data[i] += 1
                extracted['sample'].append(data[i])
                extracted['label'].append(label)
                if len(extracted['label']) == total_num:
                    return extracted
    return extracted


def corrupt_samples(args, realization, perc: float):
    """Mask all but the first ``perc`` of a random path (:96-119): returns (masked tokens [B, L], paths [B, L],
    idx = int(diffusion_steps * perc) as a 1-element tensor) — the (state, path, start step) triple a resumed
    decode takes."""
    bs, channels, seq_length = realization.size()
    idx = (args.diffusion_steps * torch.Tensor([perc])).to(int).to(args.device)
    sampled_random_path = train_helper.sample_random_path(bs, seq_length, device=args.device)
    random_path_mask = train_helper.create_mask_at_random_path_index(sampled_random_path, idx, bs, seq_length)
    real_tokens, bs, seq_length = train_helper.create_token_labels(args, realization.to(sampled_random_path.device))
    real_token_masked = train_helper.mask_realizations(real_tokens, random_path_mask)
    return real_token_masked, sampled_random_path, idx


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


class _FinalState(Sequence):
    """``mask_realization_list`` of a ``final_only`` call: it has the reference's length, but only the last entry (the
    one run_ProteoScribe_sample.py:121 reads) was brought back from the device."""

    def __init__(self, tokens_u8: np.ndarray, n: int):
        self._tok, self._n = tokens_u8, n          # [B, L] uint8 (host)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice) or i not in (-1, self._n - 1):
            raise IndexError('final_only=True keeps only the last state; call without it for the whole trajectory')
        return self._tok.astype(np.int64)[:, None, :]


def _pinned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if t.is_cuda or t.is_pinned() else t.pin_memory()


def default_noise_seed(device) -> int:
    """Philox seed for the on-device Exp(1) draws when the caller gives none: one draw from the CUDA generator of
    ``device`` — the generator the reference's ``OneHotCategorical.sample()`` consumes on its CUDA device
    (sampling_analysis.py:251), so ``torch.manual_seed`` makes a run repeatable as it does for the reference, and
    torch's CPU generator (the reference's ``torch.randperm`` paths, run_ProteoScribe_sample.py:108) is left alone."""
    return int(torch.randint(0, 2 ** 62, (1,), device=device).item())


@torch.no_grad()
def batch_generate_denoised_sampled(args, model, extract_digit_samples, extract_time, extract_digit_label,
                                    sampling_path, noise: Optional[torch.Tensor] = None,
                                    seed: Optional[int] = None, final_only: bool = False,
                                    group: Optional[int] = None, group_seeds=None):
    """Reference signature (:204-211) plus optional keywords:

    noise       explicit Exp(1) draws [T, B*L, C] (parity runs); default: drawn on the device.
    seed        Philox seed for the on-device draws; default: ``default_noise_seed`` (CUDA generator).
    final_only  bring back only the final state (64 KB instead of the 64 MB trajectory at B = 64): the returned
                list has the reference's length but only ``[-1]`` can be read.
    group, group_seeds  several independent reference batches of ``group`` samples fused into one launch (the
                cross-sample unmask write, :254-256, stays inside each batch); ``group_seeds[g]`` is batch g's Philox
                seed, and its tokens are exactly those of a separate call with ``seed=group_seeds[g]``.
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
    if group is None:
        group = B
    gs = None
    if group_seeds is not None:
        gs = torch.as_tensor([int(v) & (2 ** 63 - 1) for v in group_seeds], dtype=torch.int64).to(dev)
    elif seed is None and noise is None:
        seed = default_noise_seed(dev)
    y_c = _pinned(extract_digit_label.float()).to(dev, non_blocking=True)
    path = _pinned(sampling_path.long()).to(dev, non_blocking=True)
    x0 = extract_digit_samples
    state0 = None
    if start > 0 or bool((x0 != 0).any()):
        state0 = _pinned(x0.long()).to(dev, non_blocking=True)
    if noise is not None:
        noise = _pinned(noise.float()).to(dev, non_blocking=True)
    tokens, traj = eng.decode(y_c, path, state0=state0, start_step=start, num_steps=steps, group=group,
                              noise=noise, seed=seed or 0, want_traj=not final_only, group_seeds=gs)
    if final_only:
        host = torch.empty((B, L), dtype=torch.uint8, pin_memory=True)
        host.copy_(tokens.to(torch.uint8), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        eng.check_inputs()
        return _FinalState(host.numpy(), steps), _LazyTimes(start, steps, B)
    host = torch.empty(traj.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(traj, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    eng.check_inputs()
    return _LazyStates(host.numpy()), _LazyTimes(start, steps, B)


@torch.no_grad()
def generate_denoised_sampled(args, model, extract_digit_samples, extract_time, extract_digit_label, sampling_path,
                              noise: Optional[torch.Tensor] = None, seed: Optional[int] = None):
    """Single-sequence resume loop (:152-201).  The reference's indexing (``state[0, current_location]`` with a
    [B, L] mask) only works for one sequence, and so does this: inputs [1, L] / [1] / [1, E] / [1, L].  Returns
    (list of np.int64 [1, 1, L] states, list of 0-d float32 step indices), device loop as in the batch call."""
    if extract_digit_samples.size(0) != 1 or sampling_path.size(0) != 1:
        raise IndexError('generate_denoised_sampled handles one sequence (the reference mask indexing is [1, L])')
    states, times = batch_generate_denoised_sampled(
        args, model, extract_digit_samples, torch.as_tensor(extract_time).reshape(1).long(), extract_digit_label,
        sampling_path, noise=noise, seed=seed)
    start = int(torch.as_tensor(extract_time).reshape(-1)[0].item())
    return states, [np.array(start + i, dtype=np.float32) for i in range(len(states))]


def convert_num_to_chars(tokenizer, num_seq):
    """Reference helper (:270-276)."""
    return ''.join(tokenizer[num] for num in num_seq)
