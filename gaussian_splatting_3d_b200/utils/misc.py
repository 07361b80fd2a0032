"""Small helpers the renderer modules import (reference utils/misc.py:13-50,131-136)."""
import time

import torch

_timing_ = False
_t0 = None


def step_check(step, step_size, run_at_zero=False) -> bool:
    if step_size == 0:
        return False
    return (run_at_zero or step != 0) and step % step_size == 0


def tic():
    global _t0
    if _timing_:
        _t0 = time.time()


def toc(name=""):
    if _timing_ and _t0 is not None:
        print(f"{name} Elapsed time is {time.time() - _t0} seconds.")


def print_info(x, name="tensor"):
    if isinstance(x, torch.Tensor) and x.numel():
        xf = x.float()
        print(f"{name}: shape {tuple(x.shape)} min {xf.min().item():.4g} max {xf.max().item():.4g} "
              f"mean {xf.mean().item():.4g}")
    else:
        print(f"{name}: {x}")


def lineprofiler(fn):
    return fn


class Config(dict):
    """Minimal attribute-style config with `.get`, standing in for the OmegaConf object the
    reference passes as `cfg` (hydra/omegaconf are not available offline)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
