"""Parameter activations and their inverses, looked up by the names the reference's configs use
(`svec_act`, `alpha_act`, `color_act`: abs | relu | sigmoid | nothing | exp; reference
utils/activations.py).  `activations[name]` maps a tensor; `inv_activations[name]` accepts a tensor or a
Python number (the configs' `svec_init` / `alpha_init` scalars) and returns the same kind."""
import math

import torch


def _identity(x):
    return x


class _Inverse:
    """Inverse activation that dispatches on the argument kind (tensor -> torch op, number -> math)."""

    def __init__(self, on_tensor, on_number):
        self._on_tensor, self._on_number = on_tensor, on_number

    def __call__(self, x):
        return self._on_tensor(x) if torch.is_tensor(x) else self._on_number(float(x))


def _logit_number(p):
    return math.log(p) - math.log1p(-p)


_TABLE = {
    # name: (forward on tensors, inverse)
    "abs": (torch.abs, _Inverse(torch.abs, abs)),
    "relu": (torch.nn.functional.relu, _identity),
    "sigmoid": (torch.sigmoid, _Inverse(torch.logit, _logit_number)),
    "nothing": (_identity, _identity),
    "exp": (torch.exp, _Inverse(torch.log, math.log)),
}

activations = {name: fwd for name, (fwd, _) in _TABLE.items()}
inv_activations = {name: inv for name, (_, inv) in _TABLE.items()}
