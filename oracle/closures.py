"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's log-posterior closures.

Each closure is ``log_prob(q) -> scalar`` built from the same torch ops, in the same order, as
the reference closure it restates, so that (a) values/gradients agree with the reference to
rounding and (b) timing it is a fair stand-in for the reference's PyTorch-eager CPU path
(``bench.py`` cpu_baseline, kind "port").  Gradients are taken with ``torch.autograd.grad`` exactly
as hamiltorch does.  ``dtype=torch.float64`` gives the fp64 twin used to attribute error.

Pinned by tests/golden/{bnn_vi_hmc,deeponet}_*.npz, which were generated from the real reference
closures by oracle/make_golden.py.

Reference lines restated
------------------------
BNN VI-HMC      Neural_network/VI_HMC/main_VI_HMC.py:96-151 (prior :101-112, likelihood :132-136)
                Neural_network/VI_HMC/my_make_func.py:52-73 (scatter :56-57, MLP :61-71)
                Neural_network/VI_HMC/util.py:121-136 (flatten / unflatten)
BNN full HMC    hamiltorch.sample_model's closure (third-party, absent): Gaussian prior with
                per-tensor precision tau, 'regression' likelihood -0.5*tau_out*sum((o-y)^2);
                documented in the docstring copied at main_VI_HMC.py:30-75; call site
                Neural_network/HMC/main_regression_hmc.py:124-127
DeepONet        Operator_network/VI_HMC/main_VI_HMC_burgers.py:86-178
                Operator_network/VI_HMC/my_make_func.py:44-83 (lambda layer :33-36)
                Operator_network/VI_HMC/model.py:26,33-34 (parameter order: b, branch, trunk)
DeepONet split  Operator_network/HMC/main_HMC_splitting.py:134-204, 209-258, 28-54
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# parameter layouts (flat vector == torch.cat([p.flatten() for p in model.parameters()]))
# --------------------------------------------------------------------------------------------


@dataclass
class TensorSlot:
    name: str
    offset: int
    shape: Tuple[int, ...]

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def mlp_layout(in_dim: int, widths: Sequence[int], out_dim: int = 1, last_bias: bool = True) -> List[TensorSlot]:
    """nn.Sequential(Linear(in,w0),act,...,Linear(w_last,out,bias)) -- main_VI_HMC.py:323-333."""
    slots, off, prev = [], 0, in_dim
    dims = list(widths) + [out_dim]
    for li, w in enumerate(dims):
        slots.append(TensorSlot(f"W{li}", off, (w, prev)))
        off += w * prev
        if li < len(dims) - 1 or last_bias:
            slots.append(TensorSlot(f"b{li}", off, (w,)))
            off += w
        prev = w
    return slots


def deeponet_layout(width_branch=100, width_trunk=100, in_branch=101, in_trunk=5, depth_branch=9, depth_trunk=9,
                    output_neurons=100) -> List[TensorSlot]:
    """DeepONet parameter order: scalar b FIRST, then branch, then trunk (model.py:26,33-34,42-62)."""
    slots, off = [TensorSlot("b", 0, ())], 1

    def stack(prefix, in_dim, width, depth):
        nonlocal off
        dims = [width] * (depth - 1) + [output_neurons]
        prev = in_dim
        for li, w in enumerate(dims):
            slots.append(TensorSlot(f"{prefix}W{li}", off, (w, prev)))
            off += w * prev
            slots.append(TensorSlot(f"{prefix}b{li}", off, (w,)))
            off += w
            prev = w

    stack("br", in_branch, width_branch, depth_branch)
    stack("tr", in_trunk, width_trunk, depth_trunk)
    return slots


def layout_numel(slots: Sequence[TensorSlot]) -> int:
    return slots[-1].offset + slots[-1].numel


def unflatten(slots: Sequence[TensorSlot], flat: torch.Tensor) -> List[torch.Tensor]:
    """util.py:125-136 -- consecutive row-major views of the flat vector."""
    if flat.dim() != 1:
        raise ValueError("Expecting a 1d flattened_params")
    return [flat[s.offset:s.offset + s.numel].view(s.shape) for s in slots]


_ACTS = {"tanh": torch.tanh, "relu": F.relu, "sine": torch.sin}


# --------------------------------------------------------------------------------------------
# functional models
# --------------------------------------------------------------------------------------------


def scatter_vi(frozen: Optional[torch.Tensor], sens_ind: Optional[np.ndarray], q: torch.Tensor) -> torch.Tensor:
    """my_make_func.py:56-57 -- W = clone(mu_VI); W[ind] = q  (autograd flows to q only)."""
    if frozen is None:
        return q
    full = frozen.clone()
    full[sens_ind] = q
    return full


def mlp_forward(x: torch.Tensor, weights: List[torch.Tensor], n_hidden: int, act: str, last_bias: bool) -> torch.Tensor:
    """my_make_func.py:60-73 -- linear/act x n_hidden, then the output linear."""
    a = _ACTS[act]
    c = 0
    h = x
    for _ in range(n_hidden):
        h = a(F.linear(h, weights[c], weights[c + 1]))
        c += 2
    return F.linear(h, weights[c], weights[c + 1]) if last_bias else F.linear(h, weights[c])


def trunk_features(x2: torch.Tensor) -> torch.Tensor:
    """my_make_func.py:33-36,63-65 -- [t, sin2pi x, sin4pi x, cos2pi x, cos4pi x]; x2[...,0]=t, x2[...,1]=x."""
    x = x2[:, :, 1]
    lam = torch.stack([torch.sin(2 * np.pi * x), torch.sin(4 * np.pi * x),
                       torch.cos(2 * np.pi * x), torch.cos(4 * np.pi * x)], dim=2)
    return torch.cat([x2[:, :, 0].unsqueeze(dim=2), lam], dim=2)


def deeponet_forward(x1: torch.Tensor, x2: torch.Tensor, weights: List[torch.Tensor], depth_branch: int,
                     depth_trunk: int, act: str, impose_bc: bool = True) -> torch.Tensor:
    """my_make_func.py:52-82 -- x1 (N,1,in_branch), x2 (1,P,2) -> (N,1,P)."""
    a = _ACTS[act]
    c = 1  # weights[0] is the scalar output bias
    xb = a(F.linear(x1, weights[c], weights[c + 1]))
    c += 2
    for _ in range(depth_branch - 2):
        xb = a(F.linear(xb, weights[c], weights[c + 1]))
        c += 2
    xb = F.linear(xb, weights[c], weights[c + 1])
    c += 2
    xt = trunk_features(x2) if impose_bc else x2
    xt = a(F.linear(xt, weights[c], weights[c + 1]))
    c += 2
    for _ in range(depth_trunk - 2):
        xt = a(F.linear(xt, weights[c], weights[c + 1]))
        c += 2
    xt = F.linear(xt, weights[c], weights[c + 1])
    out = torch.einsum("...i,...i->...", xb, xt)
    return torch.unsqueeze(out, 1) + weights[0]


# --------------------------------------------------------------------------------------------
# likelihood + prior pieces
# --------------------------------------------------------------------------------------------


def gaussian_log_prob_sum(q: torch.Tensor, loc, scale) -> torch.Tensor:
    """torch.distributions.Normal(loc, scale).log_prob(q).sum()."""
    return torch.distributions.Normal(loc, scale).log_prob(q).sum()


def log_likelihood(output: torch.Tensor, y: torch.Tensor, loss: str, tau_out: float) -> torch.Tensor:
    """main_VI_HMC.py:132-136 -- 'regression': tau_out is a precision; 'NLL': tau_out is a variance."""
    if loss == "regression":
        return -0.5 * tau_out * ((output - y) ** 2).sum(0)
    if loss == "NLL":
        return -torch.nn.GaussianNLLLoss(reduction="sum")(output, y, tau_out * torch.ones_like(output))
    raise NotImplementedError(loss)


def sliced_prior_sigma(d: int, tensor_numels: Sequence[int], prior_vars: Sequence[float]) -> np.ndarray:
    """Per-coordinate prior std implied by the slice loop at main_VI_HMC.py:107-112.

    The loop walks the REDUCED vector with the FULL tensors' lengths, so coordinate i of q gets the
    variance of whichever full-tensor slice [i_prev, i_prev+numel) contains i; coordinates beyond the
    last slice get no prior at all (sigma = inf).  With the shipped equal variances this is isotropic.
    """
    sig = np.full(d, np.inf, dtype=np.float64)
    i_prev = 0
    for n, v in zip(tensor_numels, prior_vars):
        hi = min(d, i_prev + n)
        if i_prev < hi:
            sig[i_prev:hi] = float(v) ** 0.5
        i_prev += n
    return sig


# --------------------------------------------------------------------------------------------
# closures
# --------------------------------------------------------------------------------------------


@dataclass
class BnnLogProb:
    """BNN log-posterior: VI-HMC form (main_VI_HMC.py) or hamiltorch.sample_model form.

    prior: either ("sliced", [var per parameter tensor]) as main_VI_HMC.py:90-91,107-112,
           ("tau", [precision per parameter tensor]) as hamiltorch.sample_model (std = tau**-0.5,
           slices over the FULL vector), or ("loc_scale", mu[d], sigma[d]) as cfg.load_prior :88,105.
    """
    x: torch.Tensor
    y: torch.Tensor
    widths: Sequence[int]
    act: str = "tanh"
    last_bias: bool = True
    loss: str = "NLL"
    tau_out: float = 0.0025
    prior: tuple = ("sliced", None)
    prior_scale: float = 1.0
    frozen: Optional[torch.Tensor] = None      # VI means, [D]
    sens_ind: Optional[np.ndarray] = None      # sorted int64 [d]
    dtype: torch.dtype = torch.float32
    slots: List[TensorSlot] = field(init=False)

    def __post_init__(self):
        self.slots = mlp_layout(self.x.shape[1], self.widths, self.y.shape[1], self.last_bias)
        self.x = self.x.to(self.dtype)
        self.y = self.y.to(self.dtype)
        if self.frozen is not None:
            self.frozen = self.frozen.to(self.dtype)
        kind = self.prior[0]
        numels = [s.numel for s in self.slots]
        if kind == "sliced":
            self._dists = [torch.distributions.Normal(torch.zeros((), dtype=self.dtype),
                                                      torch.tensor(v, dtype=self.dtype) ** 0.5) for v in self.prior[1]]
        elif kind == "tau":
            self._dists = [torch.distributions.Normal(torch.zeros((), dtype=self.dtype),
                                                      torch.tensor(t, dtype=self.dtype) ** -0.5) for t in self.prior[1]]
        elif kind == "loc_scale":
            self._dists = [torch.distributions.Normal(self.prior[1].to(self.dtype), self.prior[2].to(self.dtype))]
        else:
            raise ValueError(kind)
        self._numels = numels

    @property
    def D(self) -> int:
        return layout_numel(self.slots)

    def log_prior(self, q: torch.Tensor) -> torch.Tensor:
        l_prior = torch.zeros_like(q[0], requires_grad=True)
        if self.prior[0] == "loc_scale":
            return self._dists[0].log_prob(q).sum() + l_prior
        i_prev = 0
        for n, dist in zip(self._numels, self._dists):
            l_prior = dist.log_prob(q[i_prev:n + i_prev]).sum() + l_prior
            i_prev += n
        return l_prior

    def forward(self, q: torch.Tensor, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        full = scatter_vi(self.frozen, self.sens_ind, q)
        return mlp_forward(self.x if x is None else x.to(self.dtype), unflatten(self.slots, full),
                           len(self.widths), self.act, self.last_bias)

    def __call__(self, q: torch.Tensor) -> torch.Tensor:
        out = self.forward(q)
        ll = log_likelihood(out, self.y, self.loss, self.tau_out)
        return ll + self.log_prior(q) / self.prior_scale


@dataclass
class DeepONetLogProb:
    """DeepONet log-posterior (main_VI_HMC_burgers.py:86-178 / main_HMC_splitting.py:134-204).

    prior: N(0, sqrt(prior_var)) over the whole reduced vector (:96-102) or (mu, sigma) under load_prior.
    """
    x1: torch.Tensor                 # (N,1,in_branch)
    x2: torch.Tensor                 # (1,P,2)
    y: torch.Tensor                  # (N,P)
    width_branch: int = 100
    width_trunk: int = 100
    in_branch: int = 101
    in_trunk: int = 5
    depth_branch: int = 9
    depth_trunk: int = 9
    output_neurons: int = 100
    act: str = "tanh"
    impose_bc: bool = True
    loss: str = "NLL"
    tau_out: float = 1.0
    prior_var: float = 0.01
    prior_loc_scale: Optional[tuple] = None
    prior_scale: float = 1.0
    frozen: Optional[torch.Tensor] = None
    sens_ind: Optional[np.ndarray] = None
    dtype: torch.dtype = torch.float32
    sample_p: Optional[int] = None   # cfg.sample_data: a fresh random.sample(range(P), cfg.p) of trunk points per call (:127-137)
    slots: List[TensorSlot] = field(init=False)

    def __post_init__(self):
        self.slots = deeponet_layout(self.width_branch, self.width_trunk, self.in_branch, self.in_trunk,
                                     self.depth_branch, self.depth_trunk, self.output_neurons)
        self.x1, self.x2, self.y = (t.to(self.dtype) for t in (self.x1, self.x2, self.y))
        if self.frozen is not None:
            self.frozen = self.frozen.to(self.dtype)
        if self.prior_loc_scale is not None:
            self._dist = torch.distributions.Normal(self.prior_loc_scale[0].to(self.dtype),
                                                    self.prior_loc_scale[1].to(self.dtype))
        else:
            self._dist = torch.distributions.Normal(torch.zeros((), dtype=self.dtype),
                                                    torch.tensor(self.prior_var, dtype=self.dtype) ** 0.5)

    @property
    def D(self) -> int:
        return layout_numel(self.slots)

    def forward(self, q: torch.Tensor, data=None) -> torch.Tensor:
        x1, x2 = (self.x1, self.x2) if data is None else (data[0].to(self.dtype), data[1].to(self.dtype))
        full = scatter_vi(self.frozen, self.sens_ind, q)
        out = deeponet_forward(x1, x2, unflatten(self.slots, full), self.depth_branch, self.depth_trunk,
                               self.act, self.impose_bc)
        return out.squeeze(1)

    def __call__(self, q: torch.Tensor) -> torch.Tensor:
        l_prior = self._dist.log_prob(q).sum() + torch.zeros_like(q[0], requires_grad=True)
        if self.sample_p is not None:
            import random
            ind = random.sample(range(self.x2.shape[1]), self.sample_p)
            out = self.forward(q, data=(self.x1, self.x2[:, ind]))
            y = self.y[:, ind]
        else:
            out = self.forward(q)
            y = self.y
        assert out.shape == y.shape
        ll = log_likelihood(out, y, self.loss, self.tau_out)
        return ll + l_prior / self.prior_scale


def split_deeponet(base_kwargs: dict, x1: torch.Tensor, x2: torch.Tensor, y: torch.Tensor, num_splits: int
                   ) -> List[DeepONetLogProb]:
    """main_HMC_splitting.py:28-54 + :209-258 -- equal row blocks, shared trunk, prior/num_splits each."""
    if x1.shape[0] % num_splits != 0:
        raise ValueError("Number of splits does not split the data equally")
    n = x1.shape[0] // num_splits
    return [DeepONetLogProb(x1=x1[i * n:(i + 1) * n], x2=x2, y=y[i * n:(i + 1) * n],
                            prior_scale=float(num_splits), **base_kwargs) for i in range(num_splits)]


def value_and_grad(log_prob: Callable[[torch.Tensor], torch.Tensor], q: torch.Tensor):
    """hamiltorch's params_grad: detach, requires_grad_, autograd.grad of the scalar closure."""
    p = q.detach().requires_grad_()
    lp = log_prob(p)
    (g,) = torch.autograd.grad(lp.sum(), p)
    return lp.detach().reshape(()), g
