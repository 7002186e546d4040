"""Multi-GPU plumbing for ProteoScribe sampling: one process per GPU (torchrun), independent
(prompt, replica-batch) units sharded round-robin over ranks, ONE all-gather of the generated
token ids at the end of a decode.  There is no per-step communication: units share only the
read-only weights (the reference has no sampling-time parallelism at all, it loops serially over
prompts and replica batches, /root/reference/run_ProteoScribe_sample.py:98-118).

A reference batch is never split across ranks: its samples are coupled through the unmask write
(/root/reference/Stage3_source/sampling_analysis.py:254-256)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def units_for_rank(n_units: int, rank: int, world: int) -> List[int]:
    """Round-robin unit ids owned by `rank`."""
    return list(range(rank, n_units, world))


def plan_units(num_prompts: int, num_replicas: int, batch_size_sample: int) -> List[Tuple[int, int, int]]:
    """The reference's serial loop nest as a flat unit list: (prompt index, batch_start, batch size)
    (run_ProteoScribe_sample.py:98-102)."""
    units = []
    for p in range(num_prompts):
        for start in range(0, num_replicas, batch_size_sample):
            units.append((p, start, min(batch_size_sample, num_replicas - start)))
    return units


def gather_unit_tokens(local_tokens: torch.Tensor, local_unit_ids: Sequence[int], n_units: int,
                       unit_rows: int) -> torch.Tensor:
    """All-gather generated token ids.

    local_tokens: uint8 [n_local_units, unit_rows, L] on this rank's device (rows of short units are
    zero padded).  Returns uint8 [n_units, unit_rows, L] in unit order on every rank.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        out = torch.zeros(n_units, unit_rows, local_tokens.shape[-1], dtype=torch.uint8, device=local_tokens.device)
        out[list(local_unit_ids)] = local_tokens
        return out
    rank = dist.get_rank()
    per_rank = (n_units + world - 1) // world
    L = local_tokens.shape[-1]
    send = torch.zeros(per_rank, unit_rows, L, dtype=torch.uint8, device=local_tokens.device)
    send[:local_tokens.shape[0]] = local_tokens
    flat = torch.empty(world * per_rank, unit_rows, L, dtype=torch.uint8, device=local_tokens.device)
    dist.all_gather_into_tensor(flat, send)
    recv = flat.view(world, per_rank, unit_rows, L)
    out = torch.zeros(n_units, unit_rows, L, dtype=torch.uint8, device=local_tokens.device)
    for r in range(world):
        ids = units_for_rank(n_units, r, world)
        if ids:
            out[ids] = recv[r, :len(ids)]
    del rank
    return out
