"""Mirror of the inference part of /root/reference/run_Facilitator_sample.py (:76-121): load a dict with
``z_t`` [P, emb_dim], compute ``z_c = Facilitator(z_t)`` on the GPU, add key ``'z_c'`` and save — the
on-disk format run_ProteoScribe_sample.py consumes (``embedding_dataset['z_c']``, its :158-167).  The
MSE/MMD diagnostics the reference script prints are not reproduced."""
from __future__ import annotations

import argparse
import json
from argparse import Namespace

import torch

from .Stage1_source.model import Facilitator


def prepare_model(config_args, model_path):
    model = Facilitator(in_dim=config_args.emb_dim, hid_dim=config_args.hid_dim, out_dim=config_args.emb_dim,
                        dropout=config_args.dropout)
    model.load_state_dict(torch.load(model_path, map_location='cpu'))
    return model.eval()


def main(argv=None):
    ap = argparse.ArgumentParser(description='BioM3 Facilitator sampling (B200-native)')
    ap.add_argument('--json_path', required=True)
    ap.add_argument('--model_path', required=True)
    ap.add_argument('--input_data_path', required=True)
    ap.add_argument('--output_data_path', required=True)
    a = ap.parse_args(argv)
    with open(a.json_path) as f:
        cfg = Namespace(**json.load(f))
    model = prepare_model(cfg, a.model_path)
    data = torch.load(a.input_data_path)
    data['z_c'] = model(data['z_t']).cpu()
    torch.save(data, a.output_data_path)
    print(f"z_c {tuple(data['z_c'].shape)} written to {a.output_data_path}")
    return data


if __name__ == '__main__':
    main()
