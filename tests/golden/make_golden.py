#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REAL in-tree reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What is real and what is stood in:
  * real, imported unmodified from /root/reference:
      Stage3_source/cond_diff_transformer_layer.py   (get_model, forward, conditioning layout)
      Stage3_source/transformer_training_helper.py   (cond_predict_conditional_prob)
      Stage3_source/sampling_analysis.py             (predict_next_index, batch_generate_denoised_sampled)
      Stage3_source/animation_tools.py               (convert_num_to_char)
  * stood in, because the wheels are not installed and there is no network:
      linear_attention_transformer / axial_positional_embedding -> oracle/upstream_blocks.py
      Bio, matplotlib, imageio -> empty stand-ins (only imported, never called on this path)

So the fixtures pin everything the reference keeps in its own tree (conditioning
layout, time embedding, output permute, softmax axis, OneHotCategorical draw, the
cross-sample unmask write, return lists); the transformer block itself stays
"parity unpinned" (oracle/upstream_blocks.py header).

The reference sampler draws from torch's global CPU generator.  Each case seeds it
with ``torch.manual_seed(noise_seed)`` right before the call; the consumer
(tests/test_oracle_vs_reference.py) rebuilds the same Exp(1) stream with
``torch.empty(B*L, C).exponential_(1)`` per step after the same seed.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from biom3_b200 import synthetic  # noqa: E402
from oracle import upstream_blocks  # noqa: E402


def _install_stand_ins():
    lat = types.ModuleType('linear_attention_transformer')
    lat.LinearAttentionTransformer = upstream_blocks.LinearAttentionTransformer
    sys.modules['linear_attention_transformer'] = lat
    ax = types.ModuleType('axial_positional_embedding')
    ax.AxialPositionalEmbedding = upstream_blocks.AxialPositionalEmbedding
    sys.modules['axial_positional_embedding'] = ax
    for name in ('Bio', 'Bio.Align', 'matplotlib', 'matplotlib.pyplot', 'imageio', 'esm'):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)


def import_reference():
    _install_stand_ins()
    sys.path.insert(0, REF)
    import Stage3_source.cond_diff_transformer_layer as mod
    import Stage3_source.sampling_analysis as samp
    import Stage3_source.transformer_training_helper as helper
    import Stage3_source.animation_tools as ani
    assert mod.__file__.startswith(REF) and samp.__file__.startswith(REF)
    return mod, samp, helper, ani


CASES = {
    # name: (arg overrides, batch, n_iters, start_step)
    'tiny_b3': (dict(diffusion_steps=256, transformer_dim=64, transformer_heads=4, transformer_depth=2,
                     transformer_local_heads=2, transformer_local_size=64, text_emb_dim=32), 3, 256, 0),
    'tiny_b1': (dict(diffusion_steps=256, transformer_dim=64, transformer_heads=4, transformer_depth=2,
                     transformer_local_heads=2, transformer_local_size=64, text_emb_dim=32), 1, 256, 0),
    'tiny_resume_b2': (dict(diffusion_steps=256, transformer_dim=64, transformer_heads=4, transformer_depth=2,
                            transformer_local_heads=2, transformer_local_size=64, text_emb_dim=32), 2, 56, 200),
    # shapes the sm_100a kernels accept (head dim 32, window 128, dim % 256 == 0): compared on the GPU
    'gpu_small_b3': (dict(diffusion_steps=256, transformer_dim=256, transformer_heads=8, transformer_depth=2,
                          transformer_local_heads=4, transformer_local_size=128, text_emb_dim=64), 3, 256, 0),
    'gpu_resume_b2': (dict(diffusion_steps=256, transformer_dim=256, transformer_heads=8, transformer_depth=2,
                           transformer_local_heads=4, transformer_local_size=128, text_emb_dim=64), 2, 96, 160),
    'mid_b2': (dict(diffusion_steps=512, transformer_dim=128, transformer_heads=4, transformer_depth=3,
                    transformer_local_heads=2, transformer_local_size=128, text_emb_dim=512), 2, 512, 0),
}


def run_case(mod, samp, name):
    over, B, n_iters, start = CASES[name]
    args = synthetic.stage3_args(**over)
    L, C = args.diffusion_steps, args.num_classes
    sd = synthetic.random_state_dict(args, seed=11, perturb_norm=True)
    model = mod.get_model(args, (args.image_size, args.image_size), C)
    model.load_state_dict(sd, strict=True)
    model.eval()
    z_c = synthetic.synthetic_z_c(1, args.text_emb_dim, seed=12).repeat(B, 1)
    path = synthetic.synthetic_paths(B, L, seed=13)
    g = torch.Generator().manual_seed(14)
    if start > 0:
        # a partially unmasked start state: positions whose path index < start hold a token
        toks = torch.randint(1, C, (B, L), generator=g)
        state0 = torch.where(path < start, toks, torch.zeros_like(toks)).float()
    else:
        state0 = torch.zeros(B, L)
    # forward fixture: random tokens, distinct steps per sample
    x = torch.randint(0, C, (B, L), generator=g)
    t = torch.randint(0, L, (B,), generator=g)
    with torch.no_grad():
        logits = model(x, t, z_c)
    # sampler fixture: the real loop, global generator seeded.
    # On a CPU device ``state.cpu().numpy()`` (sampling_analysis.py:259-260) ALIASES the live
    # state tensor, so every list entry the reference returns here is the same final array; on
    # its intended CUDA device each entry is a snapshot.  The fixture records the snapshots the
    # CUDA run would return: the model input of call i+1 is the state after step i.
    seen_x, seen_t = [], []

    class Recorder(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, t, y_c):
            seen_x.append(x.detach().clone())
            seen_t.append(t.detach().clone())
            return self.inner(x=x, t=t, y_c=y_c)

    noise_seed = 15
    torch.manual_seed(noise_seed)
    states, times = samp.batch_generate_denoised_sampled(
        args=args, model=Recorder(model).eval(), extract_digit_samples=state0.clone(),
        extract_time=torch.full((B,), start).long(), extract_digit_label=z_c, sampling_path=path)
    assert len(states) == n_iters == len(seen_x), (len(states), n_iters)
    assert all(s is not None and np.array_equal(s, states[-1]) for s in states)   # the aliasing
    snaps = [sx.numpy()[:, None, :] for sx in seen_x[1:]] + [states[-1]]
    traj = np.stack(snaps).astype(np.uint8)           # [T, B, 1, L]
    tt = np.stack([st.numpy()[:, None] for st in seen_t]).astype(np.int32)   # [T, B, 1]
    np.savez_compressed(
        os.path.join(HERE, f'{name}.npz'),
        x=x.numpy().astype(np.uint8), t=t.numpy().astype(np.int32), z_c=z_c.numpy(),
        logits=logits.numpy(), path=path.numpy().astype(np.int32), state0=state0.numpy().astype(np.uint8),
        start=np.int32(start), noise_seed=np.int32(noise_seed), traj=traj, times=tt,
        weight_seed=np.int32(11),
        overrides=np.array(repr(over)))
    print(name, 'logits', tuple(logits.shape), 'traj', traj.shape)


def full_config_forward(mod):
    """stage3_config.json shape, B=2 (distinct steps), forward only: logits fixture (fp32)."""
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=0)
    model = mod.get_model(args, (32, 32), 29)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(21)
    B = 2
    x = torch.randint(0, 29, (B, 1024), generator=g)
    x[1, 300:] = 0                                    # a partly masked row
    t = torch.tensor([17, 900])
    z_c = synthetic.synthetic_z_c(B, 512, seed=1)
    with torch.no_grad():
        logits = model(x, t, z_c)
    np.savez_compressed(os.path.join(HERE, 'full_forward_b2.npz'),
                        x=x.numpy().astype(np.uint8), t=t.numpy().astype(np.int32), z_c=z_c.numpy(),
                        logits=logits.numpy().astype(np.float32), weight_seed=np.int32(0))
    print('full_forward_b2', tuple(logits.shape))


def facilitator_case():
    """The REAL Facilitator class (Stage1_source/model.py:473-493; `esm` stood in, never called), trained-looking
    weights (weight_g != ||V||), z_t ~ N(0, 1.45^2) (README reports |z_t| ~ 33 at dim 512)."""
    import Stage1_source.model as s1
    assert s1.__file__.startswith(REF)
    model = s1.Facilitator(in_dim=512, hid_dim=1024, out_dim=512, dropout=0.0)
    sd = synthetic.facilitator_state_dict(512, 1024, seed=31)
    model.load_state_dict(sd)                       # strict: the real class accepts our key schema
    model.eval()
    z_t = torch.randn(16, 512, generator=torch.Generator().manual_seed(32)) * 1.45
    with torch.no_grad():
        z_c = model(z_t)
    np.savez_compressed(os.path.join(HERE, 'facilitator_p16.npz'), z_t=z_t.numpy(), z_c=z_c.numpy(), weight_seed=np.int32(31))
    print('facilitator_p16', tuple(z_c.shape), sorted(sd.keys()))


def inpaint_case(mod, samp, helper):
    """the partial-start / inpainting entry points of the REAL reference
    (sampling_analysis.py:21-61 cond_autocomplete_real_samples, :96-119 corrupt_samples, :152-201
    generate_denoised_sampled) and the mask helpers they call (transformer_training_helper.py:16-44, 187-232),
    at the GPU-testable small shape.  RNG: torch's global CPU generator, seeded right before each call."""
    over = CASES['gpu_small_b3'][0]
    args = synthetic.stage3_args(**over)
    args.task = 'proteins'
    L, C, B = args.diffusion_steps, args.num_classes, 3
    sd = synthetic.random_state_dict(args, seed=11, perturb_norm=True)
    model = mod.get_model(args, (args.image_size, args.image_size), C)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(40)
    realization = torch.randint(0, C - 1, (B, 1, L), generator=g)       # raw ids; tokens are ids + 1 (:203)
    z_c = synthetic.synthetic_z_c(B, args.text_emb_dim, seed=12)
    out = {}
    torch.manual_seed(41)
    masked, path, idx = samp.corrupt_samples(args, realization, 0.625)
    out.update(corrupt_masked=masked.numpy(), corrupt_path=path.numpy(), corrupt_idx=idx.numpy())
    torch.manual_seed(42)
    idx2 = torch.tensor([[40], [128], [200]])
    dist, probs, real_masked, real_tokens, log_prob, path2, mask2 = samp.cond_autocomplete_real_samples(
        model, args, realization, z_c, idx2)
    out.update(auto_idx=idx2.numpy(), auto_probs=probs.numpy(), auto_masked=real_masked.numpy(),
               auto_tokens=real_tokens.numpy(), auto_log_prob=log_prob.numpy(), auto_path=path2.numpy(),
               auto_mask=mask2.numpy())
    # single-sample resume loop from the corrupted state of sample 0 (start = corrupt idx = 160)
    seen_x = []

    class Recorder(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, t, y_c):
            seen_x.append(x.detach().clone())
            return self.inner(x=x, t=t, y_c=y_c)

    start = int(idx.item())
    torch.manual_seed(43)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints a shape every step (:177)
        states, times = samp.generate_denoised_sampled(
            args=args, model=Recorder(model).eval(), extract_digit_samples=masked[0:1].clone().float(),
            extract_time=torch.tensor([start]), extract_digit_label=z_c[0:1], sampling_path=path[0:1])
    snaps = [sx.numpy()[:, None, :] for sx in seen_x[1:]] + [states[-1]]
    out.update(single_start=np.int32(start), single_noise_seed=np.int32(43),
               single_traj=np.stack(snaps).astype(np.uint8), single_len=np.int32(len(states)),
               single_state_shape=np.array(states[0].shape), single_time_dtype=np.array(str(times[0].dtype)),
               single_time_shape=np.array(times[0].shape, dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, 'inpaint_b3.npz'), realization=realization.numpy(), z_c=z_c.numpy(),
                        weight_seed=np.int32(11), overrides=np.array(repr(over)), **out)
    print('inpaint_b3', {k: getattr(v, 'shape', v) for k, v in out.items()})


class _StopAfter(Exception):
    pass


def full_b64_case(mod, samp):
    """BASELINE.json configs[1] at its own size: stage3_config.json shape, ONE reference batch of 64 sequences
    (so the cross-sample unmask write couples 64 x 64 positions per step), LayerNorm affine parameters moved off
    1 / 0.  The real sampler loop, three steps from the all-mask state (the loop is stopped by an exception raised at
    the 4th forward call: everything recorded before it is the reference's own work) and its natural last three
    steps (start = 1021 with a state that already holds 1021 tokens per row).  Per step the fixture keeps the state,
    and for the 64 x 64 selected (sample, location) pairs the race margin of the reference's own draw; plus a
    slice of the step-1 logits.  ~3 min of CPU."""
    from oracle import sampler as osamp
    args = synthetic.stage3_args()
    L, C, B = args.diffusion_steps, args.num_classes, 64
    sd = synthetic.random_state_dict(args, seed=51, perturb_norm=True)
    model = mod.get_model(args, (32, 32), C)
    model.load_state_dict(sd, strict=True)
    model.eval()
    z_c = synthetic.synthetic_z_c(1, args.text_emb_dim, seed=52).repeat(B, 1)
    path = synthetic.synthetic_paths(B, L, seed=53)
    inv = torch.argsort(path, dim=1)                      # inv[b, t] = location sample b unmasks at step t
    out = dict(weight_seed=np.int32(51), z_seed=np.int32(52), path_seed=np.int32(53), B=np.int32(B))
    for tag, start, noise_seed in (('early', 0, 54), ('late', L - 3, 55)):
        g = torch.Generator().manual_seed(56)
        toks = torch.randint(1, C, (B, L), generator=g)
        state0 = torch.where(path < start, toks, torch.zeros_like(toks)).float()
        seen_x, seen_logits = [], []

        class Recorder(torch.nn.Module):
            def __init__(self, inner):
                super().__init__()
                self.inner = inner

            def forward(self, x, t, y_c):
                seen_x.append(x.detach().clone())
                if len(seen_x) == 4:
                    raise _StopAfter()
                lg = self.inner(x=x, t=t, y_c=y_c)
                seen_logits.append(lg.detach().clone())
                return lg

        torch.manual_seed(noise_seed)
        final = None
        try:
            states, _ = samp.batch_generate_denoised_sampled(
                args=args, model=Recorder(model).eval(), extract_digit_samples=state0.clone(),
                extract_time=torch.full((B,), start).long(), extract_digit_label=z_c, sampling_path=path)
            final = states[-1][:, 0]
        except _StopAfter:
            pass
        T = 3
        assert len(seen_logits) == T
        snaps = [sx.numpy() for sx in seen_x[1:]]
        if final is not None:
            snaps.append(final)
        traj = np.stack(snaps[:T]).astype(np.uint8)         # [3, B, L] state after each step
        noise = osamp.reference_noise_stream(noise_seed, T, B, L, C)
        margins = np.zeros((T, B, B), dtype=np.float32)
        for s in range(T):
            mg = osamp.race_margins(seen_logits[s], noise[s])
            loc = inv[:, start + s]                         # [B] current location of every sample
            margins[s] = mg[:, loc].numpy()                 # [b', b] = margin at (sample b', location of sample b)
            # the recorded trajectory is what these logits + this noise give (the noise stream is the real one)
            x_in = seen_x[s].clone()
            tok = osamp.sample_tokens(seen_logits[s], noise[s])
            x_in[:, loc] = tok[:, loc]
            assert np.array_equal(x_in.numpy(), traj[s]), (tag, s)
        out.update({f'{tag}_start': np.int32(start), f'{tag}_noise_seed': np.int32(noise_seed),
                    f'{tag}_state0': state0.numpy().astype(np.uint8), f'{tag}_traj': traj,
                    f'{tag}_margins': margins.astype(np.float16)})
        if tag == 'early':
            out['early_logits1'] = seen_logits[1][:4, :, ::4].numpy().astype(np.float32)    # step 1, samples 0-3, every 4th position
        print('full_b64', tag, 'traj', traj.shape, 'min margin', float(margins.min()))
    np.savez_compressed(os.path.join(HERE, 'full_b64_g64.npz'), **out)


def cli_units_case(mod):
    """a1: the REAL batch_stage3_generate_sequences (run_ProteoScribe_sample.py:60-126; its unused imports
    pytorch_lightning and Stage3_source.PL_wrapper are stood in) for 2 prompts x 3 replicas in batches of 2 = four
    units of 2, 1, 2, 1 sequences, at the GPU-testable small shape.  Everything random comes from torch's global CPU
    generator, seeded once: per unit first the paths (randperm), then 256 steps of Exp(1) draws."""
    import json
    for name in ('pytorch_lightning', 'Stage3_source.PL_wrapper'):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    import run_ProteoScribe_sample as ref_cli
    assert ref_cli.__file__.startswith(REF)
    over = CASES['gpu_small_b3'][0]
    args = synthetic.stage3_args(**over, num_replicas=3, batch_size_sample=2)
    args.device = 'cpu'
    C = args.num_classes
    sd = synthetic.random_state_dict(args, seed=11, perturb_norm=True)
    model = mod.get_model(args, (args.image_size, args.image_size), C)
    model.load_state_dict(sd, strict=True)
    model.eval()
    z_c = synthetic.synthetic_z_c(2, args.text_emb_dim, seed=61)
    import contextlib
    import io
    torch.manual_seed(62)
    with contextlib.redirect_stdout(io.StringIO()):
        d = ref_cli.batch_stage3_generate_sequences(args=args, model=model, z_t=z_c)
    np.savez_compressed(os.path.join(HERE, 'cli_units.npz'), overrides=np.array(repr(over)), weight_seed=np.int32(11),
                        z_seed=np.int32(61), global_seed=np.int32(62), num_replicas=np.int32(3), batch_size_sample=np.int32(2),
                        result=np.array(json.dumps(d)))
    print('cli_units', {k: [len(s) for s in v] for k, v in d.items()})


def main():
    torch.set_num_threads(os.cpu_count())
    mod, samp, helper, ani = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'inpaint':
        inpaint_case(mod, samp, helper)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'full_b64':
        full_b64_case(mod, samp)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'cli_units':
        cli_units_case(mod)
        return
    for name in CASES:
        run_case(mod, samp, name)
    inpaint_case(mod, samp, helper)
    cli_units_case(mod)
    full_config_forward(mod)
    full_b64_case(mod, samp)
    facilitator_case()
    # key schema of the real model at the stage3_config.json shape
    import json
    args = synthetic.stage3_args()
    real = mod.get_model(args, (32, 32), 29)
    with open(os.path.join(HERE, 'state_dict_keys.json'), 'w') as f:
        json.dump({k: list(v.shape) for k, v in real.state_dict().items()}, f, indent=0)
    toks = np.arange(29)
    s = ani.convert_num_to_char(synthetic.TOKENS, toks)
    with open(os.path.join(HERE, 'convert_num_to_char.txt'), 'w') as f:
        f.write(s)
    print('done')


if __name__ == '__main__':
    main()
