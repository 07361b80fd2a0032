"""Learning-rate schedules used by SHRenderer.get_optimizer (reference utils/schedulers.py)."""
import numpy as np


def exp_decay(tot_steps, lr_start, lr_end, warmup_steps=0, warmup_type="linear"):
    def at(step):
        if step < warmup_steps:
            return lr_start * (step / warmup_steps) if warmup_type == "linear" else None
        t = np.clip((step - warmup_steps) / (tot_steps - warmup_steps), 0, 1)
        return np.exp(np.log(lr_start) * (1 - t) + np.log(lr_end) * t)

    return at


def cosine_decay(tot_steps, lr_start, lr_end, warmup_steps=0, warmup_type="linear"):
    def at(step):
        if step < warmup_steps:
            return lr_start * (step / warmup_steps)
        progress = (step - warmup_steps) / (tot_steps - warmup_steps)
        return lr_end + (lr_start - lr_end) * (1 + np.cos(np.pi * progress)) / 2

    return at


def no_decay(tot_steps, lr_start, lr_end, warmup_steps=0, warmup_type="linear"):
    return lambda step: lr_start


lr_schedulers = dict(nothing=no_decay, cosine=cosine_decay, exp=exp_decay)
