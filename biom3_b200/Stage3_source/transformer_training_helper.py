"""Mirror of the sampling-side part of /root/reference/Stage3_source/transformer_training_helper.py:
``cond_predict_conditional_prob`` (:432-455) and the path / mask / token helpers the inpainting entry
points of sampling_analysis call (:16-69, 126-135, 187-232).  The ELBO loss and the training metrics
in that file are out of scope."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.distributions import OneHotCategorical


def sample_random_path(batch_size: int, seq_length: int, device='cpu') -> torch.Tensor:
    """One ``torch.randperm`` per sample, stacked [B, L] (:16-33).  Drawn on the host generator whatever the
    target device, so a seeded run yields the same paths as the reference does on CPU."""
    paths = torch.stack([torch.randperm(seq_length) for _ in range(batch_size)], dim=0)
    return paths.to(device)


def create_mask_at_random_path_index(sample_random_path: torch.Tensor, idx, batch_size: int, seq_length: int):
    """True where the position was already sampled: path < idx (:35-44)."""
    return sample_random_path < idx


def create_sampling_location_mask(sampled_random_path: torch.Tensor, idx, batch_size: int, seq_length: int):
    """1 at the position being sampled now: path == idx (:47-56)."""
    return (sampled_random_path == idx).long()


def create_mask_at_future_path_index(sampled_random_path: torch.Tensor, idx, batch_size: int, seq_length: int):
    """1 where the position is sampled later: path > idx (:59-69)."""
    return (sampled_random_path > idx).long()


def sample_from_conditional(conditional_prob) -> torch.Tensor:
    """One-hot draw, [B, C, L] (:266-269)."""
    return conditional_prob.sample().permute(0, 2, 1)


def compute_entropy(conditional_prob) -> torch.Tensor:
    return conditional_prob.entropy()


def log_prob_of_realization(args, conditional_prob, real_tokens: torch.Tensor) -> torch.Tensor:
    """log p(token) at every position, [B, L] (:126-135)."""
    return conditional_prob._categorical.log_prob(real_tokens)


def log_prob_of_unsampled_locations(log_prob: torch.Tensor, token_mask: torch.Tensor) -> torch.Tensor:
    """Sum of the log-probs over the still-masked positions (:160-168)."""
    return ((token_mask == 0) * 1 * log_prob).sum(1)


def create_token_labels(args, realization: torch.Tensor):
    """[B, 1, L] raw ids -> token ids [B, L]; 0 is the absorbing (mask) state, so protein ids shift by one and
    MNIST pixels map {0, 1} -> {1, 2} (:187-208).  Returns (tokens, B, L)."""
    bs, channel, seq_length = realization.size()
    temp_real = realization.reshape(bs, channel, seq_length) * 1
    if args.task == 'MNIST':
        real_tokens = (temp_real == 1) * 2 + (temp_real == 0) * 1
    elif args.task == 'proteins':
        real_tokens = temp_real + 1
    else:
        raise UnboundLocalError(f"unknown task {args.task!r}")       # what the reference's if/elif falls into
    return real_tokens.squeeze(1), bs, seq_length


def mask_realizations(real_tokens: torch.Tensor, random_path_mask: torch.Tensor) -> torch.Tensor:
    """Copy of ``real_tokens`` with every not-yet-sampled position set to the mask token 0 (:211-232; the
    reference loops over the batch and index-assigns, the result is this select)."""
    keep = random_path_mask.to(dtype=torch.bool)
    if keep.dim() == 3:
        keep = keep.squeeze(1)
    return torch.where(keep, real_tokens, torch.zeros_like(real_tokens))


def cond_predict_conditional_prob(model, real_token_masked, y_c, idx, args):
    """logits = model(x, t, y_c) on the CUDA engine; softmax over the class axis (dim=1);
    OneHotCategorical over [B, L, C] (:443-449).  Returns (distribution, probs [B, C, L])."""
    logits = model(x=real_token_masked, t=idx.view(-1,), y_c=y_c)
    probs = F.softmax(logits, dim=1)
    conditional_prob = OneHotCategorical(probs=probs.permute(0, 2, 1))
    return conditional_prob, probs
