"""Mirror of the on-path part of /root/reference/Stage3_source/transformer_training_helper.py:
``cond_predict_conditional_prob`` (:432-455).  The training loss / metrics in that file are out of
scope."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.distributions import OneHotCategorical


def cond_predict_conditional_prob(model, real_token_masked, y_c, idx, args):
    """logits = model(x, t, y_c) on the CUDA engine; softmax over the class axis (dim=1);
    OneHotCategorical over [B, L, C] (:443-449).  Returns (distribution, probs [B, C, L])."""
    logits = model(x=real_token_masked, t=idx.view(-1,), y_c=y_c)
    probs = F.softmax(logits, dim=1)
    conditional_prob = OneHotCategorical(probs=probs.permute(0, 2, 1))
    return conditional_prob, probs
