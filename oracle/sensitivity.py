"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sensitivity scores.

Follows Neural_network/VI/sensitivity.py:71-126 (eval_std_dydw / eval_jac): the per-sample Jacobian of the scalar network
output with respect to every parameter, squared, averaged over the validation inputs (and the single output dimension),
times the squared VI standard deviation.  Pinned by tests/golden/bnn_sensitivity.npz, produced by the REAL reference
functions (oracle/make_golden.py :: bnn_sensitivity_cases)."""
from __future__ import annotations

import numpy as np
import torch


def _act(name):
    return {"tanh": torch.tanh, "relu": torch.relu, "sine": torch.sin}[name]


def scores(x: torch.Tensor, widths, act: str, mean_params: torch.Tensor, std_params: torch.Tensor, last_bias: bool = True,
           dtype=torch.float32) -> np.ndarray:
    """scores[i] = std_i^2 * mean_n (d o(x_n) / d w_i)^2 for Linear(1,w0)-act-...-Linear(w_last,1)."""
    f = _act(act)
    dims, prev = [], x.shape[1]
    for w in list(widths) + [1]:
        dims.append((w, prev))
        prev = w
    w = mean_params.to(dtype)
    out = torch.zeros_like(w)
    for n in range(x.shape[0]):
        p = w.clone().requires_grad_()
        h, off = x[n].to(dtype), 0
        for li, (o, i) in enumerate(dims):
            W = p[off:off + o * i].reshape(o, i)
            off += o * i
            h = W @ h
            if li < len(dims) - 1 or last_bias:
                h = h + p[off:off + o]
                off += o
            if li < len(dims) - 1:
                h = f(h)
        (g,) = torch.autograd.grad(h.sum(), p)
        out += g * g
    return (out / x.shape[0] * std_params.to(dtype) ** 2).detach().numpy()


def deeponet_scores(x_branch: torch.Tensor, x_trunk: torch.Tensor, arch_kwargs: dict, mean_params: torch.Tensor,
                    std_params: torch.Tensor, dtype=torch.float64) -> np.ndarray:
    """Operator_network/VI/sensitivity.py:61-126 for one batch: scores[i] = std_i^2 * mean_{n,p} (d out[n,p] / d w_i)^2 with
    out = the functional DeepONet (my_make_func.py:46-85, restated in oracle/closures.py::deeponet_forward).  The Jacobian is
    taken one output at a time with autograd (small cases only).  x_branch [N,1,in_branch], x_trunk [1,P,2].
    Pinned by tests/golden/deeponet_sensitivity.npz (the reference's own eval_std_dydw, oracle/make_golden.py)."""
    from . import closures as oc

    slots = oc.deeponet_layout(arch_kwargs["width_branch"], arch_kwargs["width_trunk"], arch_kwargs["in_branch"],
                               arch_kwargs["in_trunk"], arch_kwargs["depth_branch"], arch_kwargs["depth_trunk"],
                               arch_kwargs["output_neurons"])
    w = mean_params.to(dtype).clone().requires_grad_()
    out = oc.deeponet_forward(x_branch.to(dtype), x_trunk.to(dtype), oc.unflatten(slots, w), arch_kwargs["depth_branch"],
                              arch_kwargs["depth_trunk"], arch_kwargs["act"], arch_kwargs.get("impose_bc", True)).reshape(-1)
    acc = torch.zeros_like(w)
    for k in range(out.numel()):
        (g,) = torch.autograd.grad(out[k], w, retain_graph=True)
        acc += g * g
    return (acc / out.numel() * std_params.to(dtype) ** 2).detach().numpy()
