"""The three helpers of the reference's util.py that sit on the hot path
(Neural_network/VI_HMC/util.py:13-25 seed, :106-118 NaN guard, :121-136 flatten/unflatten)."""
from __future__ import annotations

import random
import time

import numpy as np
import torch


def set_random_seed(seed=None):
    """util.py:13-22.  Unlike the reference this is NOT called at import time: engine runs are keyed by an
    explicit Philox ``seed`` argument so they are reproducible."""
    if seed is None:
        seed = int((time.time() * 1e6) % 1e8)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def has_nan_or_inf(value) -> bool:
    if torch.is_tensor(value):
        v = torch.sum(value)
        return bool(torch.isnan(v)) or bool(torch.isinf(v))
    value = float(value)
    return value in (float("inf"), float("-inf")) or value != value


class LogProbError(Exception):
    pass


def flatten(model: torch.nn.Module) -> torch.Tensor:
    return torch.cat([p.flatten() for p in model.parameters()])


def unflatten(model: torch.nn.Module, flattened_params: torch.Tensor):
    if flattened_params.dim() != 1:
        raise ValueError('Expecting a 1d flattened_params')
    params_list, i = [], 0
    for val in list(model.parameters()):
        length = val.nelement()
        params_list.append(flattened_params[i:i + length].view_as(val))
        i += length
    return params_list
