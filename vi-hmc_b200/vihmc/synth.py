"""Synthetic workloads for the configurations BASELINE.json names.

The reference bundles only the 20-point BNN regression set; its VI artefacts
(``means_flattened_<uid>``, ``stds_flattened_<uid>``, ``gradient_indices_<uid>.npy``) and the Burgers
``.mat`` are not shipped.  Everything here is generated with ``numpy.random.RandomState`` so the same
arrays come out on any machine and any torch version.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from .spec import DeepONetArch, MLPArch

_DATA = os.path.join(os.path.dirname(__file__), "data", "bnn_regression.npz")


def bnn_data() -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """The reference's bundled regression set (Neural_network/Data/{x,y}_{train,val}): x_train (20,1),
    y_train (20,1), x_val (300,1), y_val (300,1), float32; y = 4 sin 4x + 5 cos 12x + noise
    (main_regression_hmc.py:40-48)."""
    z = np.load(_DATA)
    return tuple(torch.from_numpy(z[k].copy()) for k in ("x_train", "y_train", "x_val", "y_val"))


def bnn_arch() -> MLPArch:
    """1 -> 10 -> 10 -> 1 tanh, bias on (Neural_network/VI_HMC/config.py:12-17)."""
    return MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)


def bnn_vi_artifacts(D: int = 141, d: int = 40, seed: int = 1):
    """cfg2 artefacts: mu = 0.5 randn(D), sigma = 0.01 + 0.1 rand(D) (seed), ind = sorted choice(D,d) (seed 0)."""
    rs = np.random.RandomState(seed)
    mu = (0.5 * rs.randn(D)).astype(np.float32)
    sigma = (0.01 + 0.1 * rs.rand(D)).astype(np.float32)
    ind = np.sort(np.random.RandomState(0).choice(D, d, replace=False)).astype(np.int64)
    return torch.from_numpy(mu), torch.from_numpy(sigma), ind


def wide_bnn_data(n: int = 100_000, seed: int = 0):
    """cfg5 data: x ~ U(-1,1), y = 4 sin 4x + 5 cos 12x + N(0, 0.05^2)."""
    rs = np.random.RandomState(seed)
    x = rs.uniform(-1.0, 1.0, size=(n, 1))
    y = 4 * np.sin(4 * x) + 5 * np.cos(12 * x) + 0.05 * rs.randn(n, 1)
    return torch.from_numpy(x.astype(np.float32)), torch.from_numpy(y.astype(np.float32))


def default_linear_init(arch, seed: int = 0) -> torch.Tensor:
    """Flat parameter vector with nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) scale, numpy RNG."""
    rs = np.random.RandomState(seed)
    parts = []
    if isinstance(arch, MLPArch):
        dims = arch.layer_dims
        for li, (o, i) in enumerate(dims):
            b = 1.0 / np.sqrt(i)
            parts.append(rs.uniform(-b, b, size=o * i))
            if li < len(dims) - 1 or arch.last_bias:
                parts.append(rs.uniform(-b, b, size=o))
    else:
        parts.append(np.zeros(1))
        for which in ("branch", "trunk"):
            for o, i in arch.stack_dims(which):
                b = 1.0 / np.sqrt(i)
                parts.append(rs.uniform(-b, b, size=o * i))
                parts.append(rs.uniform(-b, b, size=o))
    return torch.from_numpy(np.concatenate(parts).astype(np.float32))


def trunk_grid(n_t: int = 101, n_x: int = 101) -> np.ndarray:
    """Time-major (t_j, x_i) grid on [0,1]^2, shape (n_t*n_x, 2); col0 = t, col1 = x."""
    t = np.linspace(0.0, 1.0, n_t)
    x = np.linspace(0.0, 1.0, n_x)
    tt, xx = np.meshgrid(t, x, indexing="ij")
    return np.stack([tt.ravel(), xx.ravel()], axis=1)


def _features(x2: np.ndarray) -> np.ndarray:
    t, x = x2[:, 0], x2[:, 1]
    return np.stack([t, np.sin(2 * np.pi * x), np.sin(4 * np.pi * x), np.cos(2 * np.pi * x), np.cos(4 * np.pi * x)], 1)


def deeponet_numpy_forward(arch: DeepONetArch, theta: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """float64 DeepONet forward used only to manufacture the teacher targets below."""
    act = {"tanh": np.tanh, "relu": lambda z: np.maximum(z, 0.0)}[arch.act]
    off = 1

    def stack(h, dims):
        nonlocal off
        for li, (o, i) in enumerate(dims):
            w = theta[off:off + o * i].reshape(o, i)
            off += o * i
            b = theta[off:off + o]
            off += o
            h = h @ w.T + b
            if li < len(dims) - 1:
                h = act(h)
        return h

    xb = stack(x1.astype(np.float64), arch.stack_dims("branch"))
    feats = _features(x2.astype(np.float64)) if arch.impose_bc else x2.astype(np.float64)
    xt = stack(feats, arch.stack_dims("trunk"))
    return xb @ xt.T + theta[0]


def burgers_like(arch: DeepONetArch = DeepONetArch(), n_train: int = 1000, n_t: int = 101, n_x: int = 101, seed: int = 0,
                 out_scale: float = 1.0, trunk_scale: float = 1.0):
    """Burgers-SHAPED synthetic operator data (the real .mat is not bundled; Operator_network/Data/data.txt).

    branch inputs: n_train draws of a periodic Gaussian random field on ``arch.in_branch`` sensors
    (RBF in sin(pi dx), length-scale 0.2, amplitude 0.1); trunk grid n_t x n_x time-major;
    y = teacher-DeepONet(theta*) + 0.01 N(0,1) with theta* ~ N(0, 0.1^2).
    out_scale != 1 multiplies the teacher's outputs: theta*'s scalar output bias and the last branch layer (weights and bias)
    are scaled, so the scaled teacher is still a member of the architecture.  With theta* ~ N(0, 0.1^2) the raw outputs have rms
    1.3; out_scale = 0.15 puts the targets in the +-0.2 range SURVEY.md 8(d) specifies for the Burgers-shaped data (the bench
    legs use it).  trunk_scale != 1 scales the last trunk layer the same way (the curvature of the log-posterior along the branch
    weights grows with the square of the trunk features, so the feature scale decides which leapfrog step sizes are stable).
    The defaults 1.0 are what the golden vectors were generated with.
    Returns x1 (N,in_branch) f32, x2 (P,2) f32, y (N,P) f32, theta_star (D,) f32.
    """
    rs = np.random.RandomState(seed)
    m = arch.in_branch
    s = np.linspace(0.0, 1.0, m)
    dist = np.sin(np.pi * np.abs(s[:, None] - s[None, :]))
    cov = 0.1 ** 2 * np.exp(-2.0 * dist ** 2 / 0.2 ** 2) + 1e-10 * np.eye(m)
    chol = np.linalg.cholesky(cov)
    x1 = (chol @ rs.randn(m, n_train)).T
    x2 = trunk_grid(n_t, n_x)
    theta = 0.1 * rs.randn(arch.num_params)
    if out_scale != 1.0:
        theta[0] *= out_scale
        off = 1
        dims = arch.stack_dims("branch")
        for o, i in dims[:-1]:
            off += o * i + o
        o, i = dims[-1]
        theta[off:off + o * i + o] *= out_scale
    if trunk_scale != 1.0:
        off = 1 + sum(o * i + o for o, i in arch.stack_dims("branch"))
        dims = arch.stack_dims("trunk")
        for o, i in dims[:-1]:
            off += o * i + o
        o, i = dims[-1]
        theta[off:off + o * i + o] *= trunk_scale
    y = deeponet_numpy_forward(arch, theta, x1, x2) + 0.01 * rs.randn(n_train, x2.shape[0])
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return f32(x1), f32(x2), f32(y), f32(theta)


def deeponet_vi_artifacts(theta_star: torch.Tensor, frac: float = 0.10, seed: int = 1):
    """cfg4 artefacts: mu = theta* + 0.01 randn, sigma = 0.001 + 0.01 rand, ind = sorted random 10 %."""
    rs = np.random.RandomState(seed)
    D = int(theta_star.numel())
    mu = (theta_star.numpy().astype(np.float64) + 0.01 * rs.randn(D)).astype(np.float32)
    sigma = (0.001 + 0.01 * rs.rand(D)).astype(np.float32)
    d = max(1, int(round(frac * D)))
    ind = np.sort(rs.choice(D, d, replace=False)).astype(np.int64)
    return torch.from_numpy(mu), torch.from_numpy(sigma), ind
