"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares (no compute
calls without a GPU), and the host-side mirror of the reference interface behaves like it."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import biom3_b200
from biom3_b200 import _lib, synthetic
from conftest import GOLDEN, ROOT


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__
    __graft_entry__.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'biom3_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(biom3_[a-z_0-9]+)\s*\(', hdr)))
    assert declared, 'no declarations parsed'
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f'{name} not exported by libbiom3_b200.so'


def test_no_cpu_path(lib):
    """Without a CUDA device create() must fail loudly, not fall back."""
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    cfg = _lib.Config(1024, 512, 16, 16, 1, 8, 128, 29, 512, 0)
    h = ctypes.c_void_p()
    rc = lib.biom3_create(ctypes.byref(cfg), 0, 1, ctypes.byref(h))
    assert rc != 0 and b'CUDA' in lib.biom3_last_error()


def test_config_rejected_before_cuda(lib):
    for bad in (dict(reversible=1), dict(n_blocks=2), dict(heads=8), dict(local_window=64), dict(seq_len=1000),
                dict(num_classes=40)):
        d = dict(seq_len=1024, dim=512, heads=16, depth=16, n_blocks=1, local_heads=8, local_window=128,
                 num_classes=29, text_emb_dim=512, reversible=0)
        d.update(bad)
        cfg = _lib.Config(**d)
        h = ctypes.c_void_p()
        assert lib.biom3_create(ctypes.byref(cfg), 0, 1, ctypes.byref(h)) == -1, bad


def test_model_mirror_keys_match_reference():
    """state_dict keys/shapes of the mirror == the REAL reference get_model (fixture)."""
    from biom3_b200.Stage3_source import cond_diff_transformer_layer as mod
    with open(os.path.join(GOLDEN, 'state_dict_keys.json')) as f:
        ref = json.load(f)
    model = mod.get_model(synthetic.stage3_args(), (32, 32), 29)
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == ref          # same key set and shapes (load_state_dict does not depend on order)
    model.eval()
    # strict load of a reference-shaped checkpoint works; a wrong one fails
    sd = synthetic.random_state_dict(synthetic.stage3_args(), seed=1)
    model.load_state_dict(sd)
    sd.pop('transformer.out.bias')
    with pytest.raises(RuntimeError):
        model.load_state_dict(sd)


def test_forward_on_cpu_raises():
    from biom3_b200.Stage3_source import cond_diff_transformer_layer as mod
    model = mod.get_model(synthetic.stage3_args(transformer_depth=1), (32, 32), 29)
    with pytest.raises(RuntimeError, match='no CPU path'):
        model(torch.zeros(1, 1024), torch.zeros(1).long(), torch.zeros(1, 512))


def test_install_as_stage3_source():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k.startswith('Stage3_source')}
    try:
        biom3_b200.install_as_stage3_source()
        import Stage3_source.PL_wrapper  # noqa: F401  (the reference CLI imports it)
        import Stage3_source.cond_diff_transformer_layer as a
        import Stage3_source.sampling_analysis as b
        import Stage3_source.animation_tools as c
        assert a.get_model and b.batch_generate_denoised_sampled and b.predict_next_index
        assert c.convert_num_to_char(synthetic.TOKENS, np.array([1, 2, 3, 22])) == '<START>AC<END>'
    finally:
        for k in [k for k in sys.modules if k.startswith('Stage3_source')]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_lazy_lists_match_reference_contract():
    from biom3_b200.Stage3_source.sampling_analysis import _LazyStates, _LazyTimes
    traj = np.arange(4 * 2 * 8, dtype=np.uint8).reshape(4, 2, 8) % 29
    s = _LazyStates(traj)
    assert len(s) == 4 and s[-1].shape == (2, 1, 8) and s[-1].dtype == np.int64
    np.testing.assert_array_equal(s[1][:, 0], traj[1])
    assert [x.shape for x in s[1:3]] == [(2, 1, 8)] * 2
    for i, entry in enumerate(s[-1]):                      # how the reference CLI iterates (:121-122)
        np.testing.assert_array_equal(entry[0], traj[-1][i])
    t = _LazyTimes(5, 4, 2)
    assert len(t) == 4 and t[0].shape == (2, 1) and t[0].dtype == np.int64 and int(t[-1][0, 0]) == 8
    with pytest.raises(IndexError):
        t[4]


def test_cli_helpers(tmp_path):
    from biom3_b200 import run_ProteoScribe_sample as cli
    # the stage3_config.json schema: flat keys, training-only keys present and ignored
    d = vars(synthetic.stage3_args())
    d.update(device='cuda', precision='fp16', choose_optim='DeepSpeedCPUAdam', learning_rate=1e-4, epochs=1000,
             model_option='transformer', num_y_class_labels=6, task='proteins', nested=dict(a=1))
    jp = tmp_path / 'stage3_config.json'
    jp.write_text(json.dumps(d))
    cfg = cli.convert_to_namespace(cli.load_json_config(str(jp)))
    assert cfg.nested.a == 1
    assert cfg.diffusion_steps == 1024 and cfg.transformer_local_heads == 8 and cfg.num_replicas == 5
    eng_cfg = __import__('biom3_b200.engine', fromlist=['x']).config_from_args(cfg)
    assert (eng_cfg.seq_len, eng_cfg.dim, eng_cfg.heads, eng_cfg.depth, eng_cfg.num_classes) == (1024, 512, 16, 16, 29)
    ids = np.array([1, 2, 0, 3, 22, 23])
    assert cli.clean_sequence(synthetic.TOKENS, ids) == 'A-C'     # '-' (id 0) is kept, as in the reference
    a = cli.parse_arguments(['--json_path', 'a', '--model_path', 'b', '--input_path', 'c', '--output_path', 'd'])
    assert a.output_path == 'd'


def test_unit_plan_matches_reference_loop_nest():
    from biom3_b200 import distributed as bd
    assert bd.plan_units(2, 5, 32) == [(0, 0, 5), (1, 0, 5)]
    assert bd.plan_units(1, 70, 32) == [(0, 0, 32), (0, 32, 32), (0, 64, 6)]
    assert len(bd.plan_units(16, 32, 32)) == 16 and len(bd.plan_units(1, 512, 64)) == 8
    got = sorted(sum((bd.units_for_rank(16, r, 8) for r in range(8)), []))
    assert got == list(range(16))


def test_facilitator_mirror_keys():
    """Same state-dict keys/shapes as weight_norm(nn.Linear, dim=None) in the reference Facilitator."""
    from biom3_b200.Stage1_source.model import Facilitator
    m = Facilitator(512, 1024, 512, dropout=0.0)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {'main.0.bias': (1024,), 'main.0.weight_g': (), 'main.0.weight_v': (1024, 512),
                   'main.3.bias': (512,), 'main.3.weight_g': (), 'main.3.weight_v': (512, 1024)}
    m.load_state_dict(synthetic.facilitator_state_dict())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU path'):
            m(torch.zeros(2, 512))


def test_inpainting_helpers_match_real_reference_fixture():
    """corrupt_samples and the mask / token helpers against tests/golden/inpaint_b3.npz (the REAL reference run with
    the same global seed); host logic only, no device."""
    import ast
    import numpy as np
    from biom3_b200.Stage3_source import sampling_analysis as samp
    from biom3_b200.Stage3_source import transformer_training_helper as th
    z = np.load(os.path.join(GOLDEN, 'inpaint_b3.npz'))
    args = synthetic.stage3_args(**ast.literal_eval(str(z['overrides'])))
    args.task = 'proteins'
    real = torch.from_numpy(z['realization'])
    torch.manual_seed(41)
    masked, path, idx = samp.corrupt_samples(args, real, 0.625)
    assert np.array_equal(masked.numpy(), z['corrupt_masked'])
    assert np.array_equal(path.numpy(), z['corrupt_path'])
    assert np.array_equal(idx.numpy(), z['corrupt_idx']) and idx.shape == (1,)
    # the state a resumed decode starts from: tokens where path < idx, mask (0) elsewhere
    tokens, bs, L = th.create_token_labels(args, real)
    assert (bs, L) == (3, 256) and torch.equal(tokens, real.squeeze(1) + 1)
    assert torch.equal(masked == 0, path >= idx)
    assert torch.equal(masked[path < idx], tokens[path < idx])
    # per-sample idx [B, 1] as cond_autocomplete_real_samples passes it
    m2 = th.create_mask_at_random_path_index(torch.from_numpy(z['auto_path']), torch.from_numpy(z['auto_idx']), 3, L)
    assert np.array_equal(m2.numpy(), z['auto_mask'])
    assert np.array_equal(th.mask_realizations(tokens, m2).numpy(), z['auto_masked'])
    assert torch.equal(th.create_sampling_location_mask(path, idx, 3, L).sum(1), torch.ones(3, dtype=torch.long))
    assert torch.equal(th.create_mask_at_future_path_index(path, idx, 3, L).sum(1), torch.full((3,), L - 161))
    args.task = 'MNIST'
    t2, _, _ = th.create_token_labels(args, torch.tensor([[[0, 1, 1, 0]]]))
    assert t2.tolist() == [[1, 2, 2, 1]]
    lp = th.log_prob_of_unsampled_locations(torch.full((2, 4), -1.0), torch.tensor([[1, 0, 0, 1], [0, 0, 0, 0]]))
    assert lp.tolist() == [-2.0, -4.0]


def test_extract_samples_with_labels():
    from biom3_b200.Stage3_source import sampling_analysis as samp
    data = [(torch.arange(12).reshape(4, 3), torch.tensor([0, 1, 1, 0])), (torch.zeros(2, 3, dtype=torch.long), torch.tensor([1, 1]))]
    out = samp.extract_samples_with_labels(data, 1, 3)
    assert len(out['sample']) == 3 and out['sample'][0].tolist() == [4, 5, 6] and out['sample'][2].tolist() == [1, 1, 1]
    out = samp.extract_samples_with_labels([(torch.arange(6).reshape(2, 3), torch.tensor([1, 0]))], 1, 5, pad_included=True)
    assert len(out['sample']) == 1 and out['sample'][0].tolist() == [0, 1, 2]


def test_random_paths_has_no_cpu_path():
    from biom3_b200 import engine
    with pytest.raises(RuntimeError, match='no CPU path'):
        engine.random_paths(2, 128, 1, 'cpu')
