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


def test_full_b64_fixture_forward_and_sampler_step():
    """BASELINE configs[1] fixture (stage3 shape, one batch of 64, perturbed LayerNorms, the REAL loop): the oracle forward
    on four of its sequences reproduces the recorded step-1 logits (the forward has no cross-sample op), and the
    fixture's own bookkeeping holds: a step changes a row only at the 64 current locations of its batch."""
    z = np.load(os.path.join(GOLDEN, 'full_b64_g64.npz'))
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=int(z['weight_seed']), perturb_norm=True)
    B, L = int(z['B']), 1024
    y = synthetic.synthetic_z_c(1, 512, seed=int(z['z_seed'])).repeat(4, 1)
    x = torch.from_numpy(z['early_traj'][0, :4].astype(np.int64))
    logits = OracleModel(args, sd)(x, torch.full((4,), 1), y)[:, :, ::4]
    ref = torch.from_numpy(z['early_logits1'])
    assert (logits - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    path = synthetic.synthetic_paths(B, L, seed=int(z['path_seed']))
    inv = torch.argsort(path, dim=1).numpy()
    for tag in ('early', 'late'):
        start = int(z[f'{tag}_start'])
        prev = z[f'{tag}_state0']
        assert z[f'{tag}_margins'].shape == (3, B, B) and float(z[f'{tag}_margins'].min()) >= 0
        for s in range(3):
            cur = z[f'{tag}_traj'][s]
            changed = np.nonzero((cur != prev).any(0))[0]
            assert set(changed.tolist()) <= set(inv[:, start + s].tolist())
            prev = cur


def test_cli_units_fixture_replays_through_the_oracle():
    """a1 fixture: the REAL batch_stage3_generate_sequences result is what the oracle decodes when it is fed the same
    per-unit draws of the seeded global generator (paths first, then the unit's noise)."""
    import json
    z = np.load(os.path.join(GOLDEN, 'cli_units.npz'))
    over = ast.literal_eval(str(z['overrides']))
    args = synthetic.stage3_args(**over)
    sd = synthetic.random_state_dict(args, seed=int(z['weight_seed']), perturb_norm=True)
    model = OracleModel(args, sd)
    z_c = synthetic.synthetic_z_c(2, args.text_emb_dim, seed=int(z['z_seed']))
    L, C = args.diffusion_steps, args.num_classes
    want = json.loads(str(z['result']))
    torch.manual_seed(int(z['global_seed']))
    for p in range(2):
        for start, bs in ((0, 2), (2, 1)):
            path = torch.stack([torch.randperm(L) for _ in range(bs)])
            noise = osamp.global_generator_noise(L, bs, L, C)
            states, _ = osamp.decode(model, torch.zeros(bs, L), torch.zeros(bs).long(), z_c[p].repeat(bs, 1), path, noise, L)
            for i in range(bs):
                s = ''.join(synthetic.TOKENS[t] for t in states[-1][i, 0])
                s = s.replace('<START>', '').replace('<END>', '').replace('<PAD>', '')
                assert s == want[f'replica_{start + i}'][p]
