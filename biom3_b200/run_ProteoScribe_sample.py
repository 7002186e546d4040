"""ProteoScribe sampling CLI: same flags, config schema and result dictionary as the reference's
/root/reference/run_ProteoScribe_sample.py, on the CUDA engine.

    python -m biom3_b200.run_ProteoScribe_sample --json_path stage3_config.json \\
        --model_path pytorch_model.bin --input_path z_c.pt --output_path out.pt

Under ``torchrun --nproc-per-node N`` the (prompt, replica-batch) units are sharded over the N GPUs
and the token ids are all-gathered at the end (biom3_b200/distributed.py); rank 0 prints the same
dictionary.  Unlike the reference (which parses --output_path and never writes it, :140-141, :170)
the dictionary is also saved there with torch.save.
"""
from __future__ import annotations

import argparse
import json
from argparse import Namespace

import numpy as np
import torch

from . import distributed as bdist
from . import synthetic
from .Stage3_source import animation_tools as Stage3_ani_tools
from .Stage3_source import cond_diff_transformer_layer as Stage3_mod
from .Stage3_source import sampling_analysis as Stage3_sample_tools


def load_json_config(json_path):
    with open(json_path, 'r') as f:
        return json.load(f)


def convert_to_namespace(config_dict):
    for key, value in config_dict.items():
        if isinstance(value, dict):
            config_dict[key] = convert_to_namespace(value)
    return Namespace(**config_dict)


def prepare_model(args, config_args):
    """get_model + strict load_state_dict + eval (reference :38-55)."""
    model = Stage3_mod.get_model(args=config_args, data_shape=(config_args.image_size, config_args.image_size),
                                 num_classes=config_args.num_classes)
    model.load_state_dict(torch.load(args.model_path, map_location='cpu'))
    model.eval()
    print(f"Stage 3 model loaded from: {args.model_path} (loaded on {config_args.device})")
    return model


def clean_sequence(tokens, ids) -> str:
    s = Stage3_ani_tools.convert_num_to_char(tokens, ids)
    return s.replace('<START>', '').replace('<END>', '').replace('<PAD>', '')


@torch.no_grad()
def batch_stage3_generate_sequences(args, model, z_t, unit_paths=None, unit_noise=None, unit_seeds=None):
    """Reference semantics (:60-126): for every prompt, num_replicas sequences in batches of
    batch_size_sample, one fresh random permutation path per sample; returns
    {'replica_i': [sequence for each prompt]}.  Units run on this rank's GPU; with several ranks
    they are sharded and gathered.

    Everything random is drawn up front, for ALL units and in unit order, on every rank — first the sampling paths
    (torch's CPU generator, like the reference's ``torch.randperm``, :108), then one Philox seed per unit for the
    on-device Exp(1) noise — so a unit's tokens depend on its index only, not on the number of GPUs or on which rank
    runs it.  Parity hooks (tests): ``unit_paths[u]`` int64 [bs, L], ``unit_noise[u]`` fp32 [L, bs*L, C] explicit
    Exp(1) draws, ``unit_seeds[u]`` Philox seeds, each indexed by unit id."""
    if isinstance(z_t, list) and all(isinstance(item, torch.Tensor) for item in z_t):
        z_t = torch.stack(z_t)
    rank, world, _ = bdist.env_rank_world()
    model.to(args.device)
    tokens = synthetic.TOKENS
    L = args.diffusion_steps
    units = bdist.plan_units(len(z_t), args.num_replicas, args.batch_size_sample)
    mine = bdist.units_for_rank(len(units), rank, world)
    rows = min(args.batch_size_sample, args.num_replicas)
    local = torch.zeros(len(mine), rows, L, dtype=torch.uint8, device=args.device)
    device_paths = unit_paths is None and getattr(args, 'b200_device_paths', False)
    if unit_paths is not None:
        paths = list(unit_paths)
    elif device_paths:
        # permutations drawn on the GPU (biom3_random_paths), one Philox seed per unit taken from torch's global generator
        path_seeds = [int(torch.randint(0, 2 ** 62, (1,)).item()) for _ in units]
        paths = [None] * len(units)
    else:
        paths = [torch.stack([torch.randperm(L) for _ in range(bs)]) for (_, _, bs) in units]
    if unit_seeds is None and unit_noise is None:
        unit_seeds = [int(torch.randint(0, 2 ** 62, (1,)).item()) for _ in units]
    # Units are independent, so consecutive units of equal size are fused into one launch of up to
    # ``args.b200_rows_per_launch`` (default 64) rows with ``group = unit size``: the unmask write stays inside each unit and
    # every unit keeps its own noise seed, so its tokens do not depend on what it was fused with.
    max_rows = int(getattr(args, 'b200_rows_per_launch', 64) or 64)
    slot = 0
    while slot < len(mine):
        bs = units[mine[slot]][2]
        n = 1
        while slot + n < len(mine) and units[mine[slot + n]][2] == bs and (n + 1) * bs <= max_rows:
            n += 1
        batch = mine[slot:slot + n]
        if device_paths:
            from . import engine as _engine
            for uid in batch:
                paths[uid] = _engine.random_paths(bs, L, path_seeds[uid], args.device)
        z = torch.cat([z_t[units[uid][0]].unsqueeze(0).repeat(bs, 1) for uid in batch]).to(args.device)
        pth = torch.cat([paths[uid].to(args.device) for uid in batch])
        noise = None if unit_noise is None else torch.cat([unit_noise[uid] for uid in batch], dim=1)
        states, _ = Stage3_sample_tools.batch_generate_denoised_sampled(
            args=args, model=model, extract_digit_samples=torch.zeros(n * bs, L),
            extract_time=torch.zeros(n * bs).long(), extract_digit_label=z, sampling_path=pth,
            noise=noise, group=bs, group_seeds=None if unit_seeds is None else [unit_seeds[uid] for uid in batch],
            final_only=True)
        last = torch.from_numpy(states[-1][:, 0, :].astype(np.uint8)).to(args.device)
        for k in range(n):
            local[slot + k, :bs] = last[k * bs:(k + 1) * bs]
        slot += n
    allt = bdist.gather_unit_tokens(local, mine, len(units), rows).cpu().numpy()
    design_sequence_dict = {f'replica_{ii}': [] for ii in range(args.num_replicas)}
    for uid, (p, start, bs) in enumerate(units):
        for i in range(bs):
            design_sequence_dict[f'replica_{start + i}'].append(clean_sequence(tokens, allt[uid, i]))
    return design_sequence_dict


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(description='BioM3 ProteoScribe sampling (B200-native)')
    parser.add_argument('--json_path', type=str, required=True)
    parser.add_argument('--model_path', type=str, required=True)
    parser.add_argument('--input_path', type=str, required=True)
    parser.add_argument('--output_path', type=str, required=True)
    parser.add_argument('--seed', type=int, default=None,
                        help='torch.manual_seed before sampling (the reference never seeds; default: unseeded like it)')
    return parser.parse_args(argv)


def main(argv=None):
    cli = parse_arguments(argv)
    config_args = convert_to_namespace(load_json_config(cli.json_path))
    rank, world, local = bdist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError('biom3_b200 needs a CUDA device (sm_100a); there is no CPU path')
    config_args.device = f'cuda:{local}'
    embedding_dataset = torch.load(cli.input_path)
    model = prepare_model(args=cli, config_args=config_args)
    if cli.seed is not None:
        torch.manual_seed(cli.seed)            # every rank: paths and noise seeds are drawn for all units on every rank
    design_sequence_dict = batch_stage3_generate_sequences(args=config_args, model=model, z_t=embedding_dataset['z_c'])
    if rank == 0:
        torch.save(design_sequence_dict, cli.output_path)
        print(f'{design_sequence_dict=}')
    return design_sequence_dict


if __name__ == '__main__':
    main()
