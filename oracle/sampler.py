"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the ProteoScribe sampling loop with the Exp(1) race noise made an
explicit argument.  Follows

  /root/reference/Stage3_source/transformer_training_helper.py:432-455
        (cond_predict_conditional_prob: softmax over dim=1, OneHotCategorical(probs=permute(0,2,1)))
  /root/reference/Stage3_source/sampling_analysis.py:122-147  (predict_next_index)
  /root/reference/Stage3_source/sampling_analysis.py:204-265  (batch_generate_denoised_sampled:
        step loop, sample every position, current_location = argmax(path == t), the
        ``state[:, 0, loc] = tok[:, loc]`` write that touches every sample of the batch at
        every sample's location, per-step host lists, t += 1)

``torch.argmax(OneHotCategorical(probs=p).sample(), -1)`` is restated as
``argmax(p_norm / q)`` with ``p_norm = p / p.sum(-1)`` (the Categorical constructor's
renormalisation) and ``q ~ Exp(1)`` of shape [B*L, C]: ATen's single-draw multinomial.
tests/test_oracle_vs_reference.py pins this against the real in-tree reference sampler
(run with a seeded global generator) through tests/golden/.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def global_generator_noise(steps: int, B: int, L: int, C: int) -> torch.Tensor:
    """The next ``steps`` Exp(1) draws of the reference loop, taken from torch's global CPU generator in its CURRENT
    state, as logical q[step, b*L + l, c] (one ``.sample()`` per step).

    For B > 1 ``probs.reshape(-1, C)`` inside ``OneHotCategorical.sample`` copies to a contiguous
    [B*L, C] tensor, so the stream fills q row-major.  For B == 1 the permuted [1, L, C] view
    reshapes WITHOUT a copy (strides (1, L)); ``empty_like`` keeps those strides and
    ``exponential_`` fills in memory order, so the stream fills q class-major: q[l, c] =
    stream[c*L + l].  (Probed on torch 2.11 CPU; pinned by tests/test_oracle_vs_reference.py.)"""
    out = []
    for _ in range(steps):
        s = torch.empty(B * L * C).exponential_(1)
        out.append(s.reshape(C, L).t().contiguous() if B == 1 else s.reshape(B * L, C))
    return torch.stack(out)


def reference_noise_stream(seed: int, steps: int, B: int, L: int, C: int) -> torch.Tensor:
    """Rebuild the Exp(1) draws the reference loop consumes after ``torch.manual_seed(seed)``."""
    torch.manual_seed(seed)
    return global_generator_noise(steps, B, L, C)


def race_margins(logits: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Relative gap between the winner and the runner-up of the argmax(p / q) race at every position:
    logits [B, C, L], q [B*L, C] -> [B, L].  A forward within tolerance eps of this one can only flip a token
    where the margin is of order eps."""
    B, C, L = logits.shape
    p = F.softmax(logits, dim=1).permute(0, 2, 1)
    p = p / p.sum(-1, keepdim=True)
    top2 = (p.reshape(B * L, C) / q).topk(2, -1).values
    return ((top2[:, 0] - top2[:, 1]) / top2[:, 0]).reshape(B, L)


def sample_tokens(logits: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """logits [B, C, L] fp32, q [B*L, C] fp32 Exp(1) -> tokens int64 [B, L]."""
    B, C, L = logits.shape
    probs = F.softmax(logits, dim=1)
    p = probs.permute(0, 2, 1)
    p = p / p.sum(-1, keepdim=True)
    return torch.argmax(p.reshape(B * L, C) / q, dim=-1).reshape(B, L)


def unmask(state: torch.Tensor, tok: torch.Tensor, path: torch.Tensor, t: torch.Tensor) -> None:
    """In-place reference write (sampling_analysis.py:254-256). state [B,1,L], tok [B,L],
    path [B,L], t [B,1]. Every sample receives its own token at every sample's location."""
    loc = torch.argmax((path == t).cpu() * 1, dim=-1)
    state[:, 0, loc] = tok[:, loc]


@torch.no_grad()
def decode(model: Callable, state0: torch.Tensor, start_time: torch.Tensor, y_c: torch.Tensor,
           path: torch.Tensor, noise: torch.Tensor, num_steps: int,
           max_iters: Optional[int] = None, logits_hook: Optional[Callable] = None
           ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Reference loop with explicit noise.

    state0 [B, L], start_time int64 [B], y_c [B, E], path int64 [B, L] (a permutation per
    row), noise [T, B*L, C] where noise[i] is consumed at loop iteration i.  Returns the
    same two lists as the reference (np.int64 [B,1,L] states, np.int64 [B,1] step ids).
    """
    assert state0.size(0) == y_c.size(0) == path.size(0) == start_time.size(0)
    states: List[np.ndarray] = []
    times: List[np.ndarray] = []
    state = state0.unsqueeze(1).long().clone()
    t = start_time.unsqueeze(-1).clone()
    start = int(t[0].item())
    it = 0
    for _ in range(start, num_steps):
        if torch.any(t >= num_steps):
            break
        if max_iters is not None and it >= max_iters:
            break
        logits = model(state.squeeze(1), t.view(-1), y_c)
        if logits_hook is not None:
            logits_hook(it, logits)
        tok = sample_tokens(logits, noise[it])
        unmask(state, tok, path, t)
        states.append(state.cpu().numpy().copy())
        times.append(t.cpu().numpy().copy())
        t += 1
        it += 1
    return states, times
