"""The VI -> HMC split selector of the reference (Neural_network/VI/sensitivity.py), on the CUDA engine.

Reference surface mirrored here (names and meaning kept):
  eval_std_dydw(valid_data, model, mean_params, std_params) -> ndarray[D]    sensitivity.py:71-98
      scores = mean over the validation inputs of (d output / d parameter)^2, times std_params^2
      (eval_jac :101-126: jacrev of the functional model, squared, mean over the data and output dimensions)
  captured_var(imp, var_threshold) -> int                                     sensitivity.py:129-166 (without the plot)
  select_indices(imp, var_threshold) -> sorted int64 indices                  run() :226-231, what gradient_indices_<uid>.npy holds

``model`` may be the reference's ``nn.Sequential`` (Linear / activation / ... / Linear) or an :class:`vihmc.spec.MLPArch`.
Only the small-MLP family (the BNN configs) is implemented; the per-sample Jacobian runs in
``mlp_small_sensitivity_kernel`` (one CTA per chunk of validation points), there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, engine
from .spec import LogProbSpec, MLPArch


def arch_of(model) -> MLPArch:
    """MLPArch of the reference's nn.Sequential BNN (sensitivity.py:169-205 get_model)."""
    if isinstance(model, MLPArch):
        return model
    linears = [m for m in model if isinstance(m, torch.nn.Linear)]
    acts = [m for m in model if not isinstance(m, torch.nn.Linear)]
    name = type(acts[0]).__name__.lower() if acts else "tanh"
    act = {"tanh": "tanh", "relu": "relu", "sin": "sine"}.get(name)
    if act is None:
        raise ValueError("Activation should be relu, sine or tanh")
    return MLPArch(in_dim=linears[0].in_features, widths=tuple(l.out_features for l in linears[:-1]),
                   out_dim=linears[-1].out_features, act=act, last_bias=linears[-1].bias is not None)


def eval_std_dydw(valid_data, model, mean_params: torch.Tensor, std_params: torch.Tensor) -> np.ndarray:
    x, _ = valid_data
    arch = arch_of(model)
    x = x.detach().float().cpu().reshape(-1, arch.in_dim)
    # the problem struct carries the architecture and the validation inputs; targets, prior and likelihood are unused
    spec = LogProbSpec(arch=arch, x=x, y=torch.zeros(x.shape[0], 1), loss="NLL", tau_out=1.0, prior_sigma_scalar=1.0)
    prep = engine.prepare(spec)
    dev = prep.device
    w = mean_params.detach().to(device=dev, dtype=torch.float32).contiguous()
    sg = std_params.detach().to(device=dev, dtype=torch.float32).contiguous()
    if w.numel() != spec.D or sg.numel() != spec.D:
        raise ValueError(f"mean_params / std_params must have D = {spec.D} entries")
    lib = _lib.load()
    out = torch.empty(spec.D, dtype=torch.float32, device=dev)
    need = int(lib.vihmc_mlp_sensitivity_workspace_bytes(C.byref(prep.problem)))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.vihmc_mlp_sensitivity(C.byref(prep.problem), w.data_ptr(), sg.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                             ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return out.cpu().numpy()


def captured_var(imp: np.ndarray, var_threshold: float) -> int:
    """Number of parameters whose sorted cumulative share of the total score stays <= var_threshold."""
    per = np.cumsum(np.sort(imp)[::-1]) / np.sum(imp)
    return int(np.sum(per <= var_threshold))


def select_indices(imp: np.ndarray, var_threshold: float) -> np.ndarray:
    """Sorted indices of the most sensitive parameters: the content of gradient_indices_<uid>.npy."""
    num = captured_var(imp, var_threshold)
    return np.sort(np.argsort(-imp)[:num]).astype(np.int64)
