"""The VI -> HMC split selector of the reference (Neural_network/VI/sensitivity.py), on the CUDA engine.

Reference surface mirrored here (names and meaning kept):
  eval_std_dydw(valid_data, model, mean_params, std_params) -> ndarray[D]    sensitivity.py:71-98
      scores = mean over the validation inputs of (d output / d parameter)^2, times std_params^2
      (eval_jac :101-126: jacrev of the functional model, squared, mean over the data and output dimensions)
  captured_var(imp, var_threshold) -> int                                     sensitivity.py:129-166 (without the plot)
  select_indices(imp, var_threshold) -> sorted int64 indices                  run() :226-231, what gradient_indices_<uid>.npy holds

``model`` may be the reference's ``nn.Sequential`` (Linear / activation / ... / Linear) or an :class:`vihmc.spec.MLPArch`:
the per-sample Jacobian runs in ``mlp_small_sensitivity_kernel`` (one CTA per chunk of validation points).  A DeepONet
(the reference's module or a :class:`vihmc.spec.DeepONetArch`; Operator_network/VI/sensitivity.py:61-126) goes to
``vihmc_deeponet_sensitivity`` (csrc/don_sensitivity.cu), which gets the same mean of squared Jacobian entries from K seeded
back-propagations per row instead of the N x P x D Jacobian.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, engine
from .spec import DeepONetArch, LogProbSpec, MLPArch


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


def deeponet_arch_of(model) -> DeepONetArch:
    """DeepONetArch of the reference's DeepONet modules: Operator_network/VI/model.py:14-33 (one ``neurons`` width) or
    Operator_network/{HMC,VI_HMC}/model.py (``width_branch`` / ``width_trunk``)."""
    if isinstance(model, DeepONetArch):
        return model
    act = {"tanh": "tanh", "relu": "relu"}.get(type(model.act).__name__.lower())
    if act is None:
        raise ValueError("activation should be relu or tanh")
    wb = getattr(model, "width_branch", getattr(model, "neurons", None))
    wt = getattr(model, "width_trunk", getattr(model, "neurons", None))
    return DeepONetArch(width_branch=wb, width_trunk=wt, in_branch=model.in_branch, in_trunk=model.in_trunk,
                        depth_branch=model.depth_branch, depth_trunk=model.depth_trunk, output_neurons=model.output_neurons,
                        act=act, impose_bc=bool(getattr(model, "impose_bc", True)))


def _is_deeponet(model) -> bool:
    return isinstance(model, DeepONetArch) or hasattr(model, "depth_branch")


def _deeponet_batch_scores(arch: DeepONetArch, x_branch, x_trunk, w, sg) -> torch.Tensor:
    x1 = x_branch.detach().float().cpu().reshape(-1, arch.in_branch)
    x2 = x_trunk.detach().float().cpu().reshape(-1, x_trunk.shape[-1])
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=torch.zeros(x1.shape[0], x2.shape[0]), loss="NLL", tau_out=1.0,
                       prior_sigma_scalar=1.0)
    prep = engine.prepare(spec)
    dev = prep.device
    lib = _lib.load()
    out = torch.empty(spec.D, dtype=torch.float32, device=dev)
    need = int(lib.vihmc_deeponet_sensitivity_workspace_bytes(C.byref(prep.problem)))
    if need == 0:
        _lib.check(lib.vihmc_deeponet_sensitivity(C.byref(prep.problem), None, None, None, None, 0, None))   # raises with the reason
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.vihmc_deeponet_sensitivity(C.byref(prep.problem), w.data_ptr(), sg.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return out


def eval_std_dydw_deeponet(valid_data, model, mean_params: torch.Tensor, std_params: torch.Tensor) -> np.ndarray:
    """Operator_network/VI/sensitivity.py:61-98.  ``valid_data`` is one batch ``(x_branch [n,1,in_branch], x_trunk [1,p,2], ...)``
    -- the mean runs over all n functions and the p trunk points they share -- or an iterable of such batches (the reference's
    DataLoader, batch size 1 with its own random trunk subset per function, utils.py:38-40): per-batch scores are averaged with
    weight 1 / num_batches exactly as :90-94 does."""
    arch = deeponet_arch_of(model)
    dev = engine._require_cuda(None)
    w = mean_params.detach().to(device=dev, dtype=torch.float32).contiguous()
    sg = std_params.detach().to(device=dev, dtype=torch.float32).contiguous()
    if w.numel() != arch.num_params or sg.numel() != arch.num_params:
        raise ValueError(f"mean_params / std_params must have D = {arch.num_params} entries")
    if isinstance(valid_data, (tuple, list)) and len(valid_data) >= 2 and torch.is_tensor(valid_data[0]):
        batches = [valid_data]
    else:
        batches = list(valid_data)
    total = None
    for b in batches:
        s = _deeponet_batch_scores(arch, b[0], b[1], w, sg)
        total = s if total is None else total + s
    return (total / len(batches)).cpu().numpy()


def eval_std_dydw(valid_data, model, mean_params: torch.Tensor, std_params: torch.Tensor) -> np.ndarray:
    if _is_deeponet(model):
        return eval_std_dydw_deeponet(valid_data, model, mean_params, std_params)
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
