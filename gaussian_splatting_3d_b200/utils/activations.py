"""Activation / inverse-activation tables (reference utils/activations.py)."""
import numpy as np
import torch


def _dual(number_fn, tensor_fn):
    return lambda x: tensor_fn(x) if isinstance(x, torch.Tensor) else number_fn(x)


def _logit(p):
    return float(np.log(p) - np.log1p(-p))


activations = dict(
    abs=torch.abs,
    relu=torch.nn.functional.relu,
    sigmoid=torch.sigmoid,
    nothing=lambda x: x,
    exp=torch.exp,
)

inv_activations = dict(
    abs=_dual(np.abs, torch.abs),
    nothing=lambda x: x,
    sigmoid=_dual(_logit, torch.logit),
    relu=lambda x: x,
    exp=_dual(np.log, torch.log),
)
