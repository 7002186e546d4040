"""The oracle against fixtures produced by the REAL in-tree reference
(tests/golden/make_golden.py).  CPU only; pins the oracle before it is trusted."""
import ast
import os

import numpy as np
import pytest
import torch

from biom3_b200 import synthetic
from oracle.model import OracleModel
from oracle import sampler as osamp
from conftest import GOLDEN

CASES = ['tiny_b3', 'tiny_b1', 'tiny_resume_b2', 'mid_b2', 'gpu_small_b3', 'gpu_resume_b2']


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    over = ast.literal_eval(str(z['overrides']))
    args = synthetic.stage3_args(**over)
    sd = synthetic.random_state_dict(args, seed=int(z['weight_seed']), perturb_norm=True)
    return z, args, sd


@pytest.mark.parametrize('name', CASES)
def test_forward_bit_exact(name):
    z, args, sd = load_case(name)
    model = OracleModel(args, sd)
    logits = model(torch.from_numpy(z['x'].astype(np.int64)), torch.from_numpy(z['t'].astype(np.int64)),
                   torch.from_numpy(z['z_c']))
    assert logits.shape == z['logits'].shape
    assert torch.equal(logits, torch.from_numpy(z['logits'])), \
        f"max diff {(logits - torch.from_numpy(z['logits'])).abs().max().item()}"


@pytest.mark.parametrize('name', CASES)
def test_sampler_trajectory_bit_exact(name):
    """Explicit-noise restatement == reference loop drawing from the seeded global generator,
    including the B x B unmask write and the per-step order."""
    z, args, sd = load_case(name)
    model = OracleModel(args, sd)
    B, L = z['path'].shape
    C = args.num_classes
    T = z['traj'].shape[0]
    start = int(z['start'])
    noise = osamp.reference_noise_stream(int(z['noise_seed']), T, B, L, C)
    states, times = osamp.decode(model, torch.from_numpy(z['state0'].astype(np.float32)),
                                 torch.full((B,), start).long(), torch.from_numpy(z['z_c']),
                                 torch.from_numpy(z['path'].astype(np.int64)), noise, args.diffusion_steps)
    assert len(states) == T
    got = np.stack(states)
    assert got.dtype == np.int64 and got.shape == (T, B, 1, L)
    np.testing.assert_array_equal(got, z['traj'].astype(np.int64))
    np.testing.assert_array_equal(np.stack(times), z['times'].astype(np.int64))
    assert times[0].shape == (B, 1) and int(times[0][0, 0]) == start and int(times[-1][0, 0]) == args.diffusion_steps - 1


def test_full_config_forward_fixture():
    z = np.load(os.path.join(GOLDEN, 'full_forward_b2.npz'))
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=int(z['weight_seed']))
    model = OracleModel(args, sd)
    logits = model(torch.from_numpy(z['x'].astype(np.int64)), torch.from_numpy(z['t'].astype(np.int64)),
                   torch.from_numpy(z['z_c']))
    ref = torch.from_numpy(z['logits'])
    # same ops, same order, same machine class: bit exact here; allow 1e-5 for other BLAS builds
    assert (logits - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_schema_matches_reference_count():
    args = synthetic.stage3_args()
    schema = synthetic.state_dict_schema(args)
    assert len(schema) == 223      # 13 tensors x 16 layers + 15 (the real get_model accepts it strictly, make_golden.py)
    n = sum(int(np.prod(s)) for _, s, _ in schema)
    assert n == 86_186_013


def test_convert_num_to_char_fixture():
    with open(os.path.join(GOLDEN, 'convert_num_to_char.txt')) as f:
        assert f.read() == ''.join(synthetic.TOKENS)
