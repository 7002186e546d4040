"""biom3_b200 — B200-native ProteoScribe sampling (the Stage-3 hot path of amelie-iska/BioM3).

Layout: csrc/ (sm_100a kernels + C ABI), engine.py (ctypes handle), Stage3_source/ (drop-in mirror
of the reference's module/function names), run_ProteoScribe_sample.py (the reference CLI)."""
from __future__ import annotations

import importlib
import sys

__all__ = ['install_as_stage3_source']

_MIRRORED = ['PL_wrapper', 'cond_diff_transformer_layer', 'sampling_analysis', 'animation_tools',
             'transformer_training_helper', 'transformer_sampling_helper']


def install_as_stage3_source() -> None:
    """Make ``import Stage3_source.<module>`` resolve to this package's mirror, so the reference's
    unmodified run_ProteoScribe_sample.py (imports at :9-13) runs on the CUDA engine."""
    pkg = importlib.import_module('biom3_b200.Stage3_source')
    sys.modules['Stage3_source'] = pkg
    for name in _MIRRORED:
        sys.modules[f'Stage3_source.{name}'] = importlib.import_module(f'biom3_b200.Stage3_source.{name}')
