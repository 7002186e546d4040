"""world_size-2 gloo test of the multi-GPU host logic (unit sharding + all-gather of token ids)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from biom3_b200 import distributed as bd


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, rows, L, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    r, w, _ = bd.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    mine = bd.units_for_rank(n_units, rank, world)
    # a unit's "generated tokens" are a pure function of its id -> result must not depend on the sharding
    local = torch.stack([torch.full((rows, L), uid % 29, dtype=torch.uint8) + torch.arange(L, dtype=torch.uint8) % 3
                         for uid in mine]) if mine else torch.zeros(0, rows, L, dtype=torch.uint8)
    allt = bd.gather_unit_tokens(local, mine, n_units, rows)
    if rank == 0:
        torch.save(allt, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    n_units, rows, L = 5, 4, 64                          # odd unit count: ranks own 3 and 2 units
    out = str(tmp_path / 'gathered.pt')
    mp.spawn(_worker, args=(2, _free_port(), n_units, rows, L, out), nprocs=2, join=True)
    got = torch.load(out)
    ref = torch.stack([torch.full((rows, L), uid % 29, dtype=torch.uint8) + torch.arange(L, dtype=torch.uint8) % 3
                       for uid in range(n_units)])
    assert torch.equal(got, ref)
    # single-process path of the same function
    one = bd.gather_unit_tokens(ref, list(range(n_units)), n_units, rows)
    assert torch.equal(one, ref)
