"""Bayes-by-Backprop trainer on the CUDA engine -- produces the VI artefacts the VI-HMC path loads.

Reference surface mirrored (names and meaning of the config keys kept):
  Neural_network/VI/main_regression_VI.py   run() :279-346, train_model :75-124, validate_model :127-170
  Neural_network/VI/config.py               priors {prior_mu, prior_sigma, posterior_mu_initial, posterior_rho_initial}, lr_start,
                                            lr_patience, epochs, num_ens, beta_type (float), noise (std; the NLL variance is noise**2)
  Operator_network/VI/main_VI_deeponet.py   the same loop over mini-batches, loss = NLL(mean) * train_size + beta * KL
  sensitivity.py:211-222                    means_flattened_<uid> = the mu's, stds_flattened_<uid> = softplus(rho)

One optimiser step of the reference is ``num_ens`` forward/backward passes of a Bayesian net whose weights are redrawn as
``mu + softplus(rho) * eps``; here it is one ``vihmc_logp_grad`` call with ``num_ens`` parameter vectors on a prior-free
:class:`LogProbSpec` (the hot-path kernel: d loglik / d W for every draw) between ``vihmc_vi_draw`` and ``vihmc_vi_step``
(csrc/vi_bbb.cu).  Learning rate, step count, plateau scheduler and the best-validation snapshot live on the device, so an
epoch is a fixed launch sequence; it is captured once in a CUDA graph and replayed ``epochs`` times without host synchronisation.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, engine
from .spec import LogProbSpec

DEFAULT_PRIORS = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_initial": (0.0, 0.1), "posterior_rho_initial": (-3.0, 0.1)}


@dataclass
class VIResult:
    mu: torch.Tensor            # [D] final variational means (flat, model.parameters() order of the deterministic net)
    rho: torch.Tensor           # [D]
    best_mu: torch.Tensor       # [D] parameters at the lowest validation loss (the reference's max_model checkpoint)
    best_rho: torch.Tensor
    history: torch.Tensor       # [epochs, 3]: train loss, validation loss, learning rate used in that epoch
    steps: int

    @property
    def sigma(self) -> torch.Tensor:
        return torch.log1p(torch.exp(self.rho))

    @property
    def best_sigma(self) -> torch.Tensor:
        return torch.log1p(torch.exp(self.best_rho))


def likelihood_only(spec: LogProbSpec) -> LogProbSpec:
    """The spec with its prior removed (sigma = inf on every coordinate): vihmc_logp_grad then returns the log-likelihood and its
    gradient, which is what the ELBO's data term needs; the KL term is analytic."""
    if spec.sens_ind is not None:
        raise ValueError("VI training runs over the full weight vector (no VI-HMC split)")
    return dataclasses.replace(spec, prior_mu=None, prior_sigma=torch.full((spec.D,), float("inf")), prior_scale=1.0)


def init_posterior(D: int, priors: dict, seed: int = 0):
    """mu ~ N(*posterior_mu_initial), rho ~ N(*posterior_rho_initial)  (BBBLinear.reset_parameters, BBBLinear.py:45-51)."""
    g = torch.Generator().manual_seed(seed)
    m0, s0 = priors["posterior_mu_initial"]
    m1, s1 = priors["posterior_rho_initial"]
    return m0 + s0 * torch.randn(D, generator=g), m1 + s1 * torch.randn(D, generator=g)


def train_bbb(train_specs: Union[LogProbSpec, Sequence[LogProbSpec]], valid_specs: Union[LogProbSpec, Sequence[LogProbSpec], None] = None,
              priors: Optional[dict] = None, lr_start: float = 1e-2, lr_patience: int = 5000, min_lr: float = 1e-5, epochs: int = 1000,
              num_ens: int = 10, beta: float = 1.0, nll_scale: Union[float, Sequence[float]] = 1.0,
              valid_nll_scale: Union[float, Sequence[float]] = 1.0, seed: int = 0, mu0: Optional[torch.Tensor] = None,
              rho0: Optional[torch.Tensor] = None, inject_eps: Optional[torch.Tensor] = None, kl_form: int = 0,
              use_graph: bool = True) -> VIResult:
    """Train a mean-field Gaussian posterior over the full weight vector.

    train_specs: one :class:`LogProbSpec` per mini-batch (the BNN reference is full batch: one spec); priors inside the specs are
    ignored.  nll_scale multiplies the summed Gaussian NLL of a batch (1 for the BNN trainer; train_size / batch elements for the
    DeepONet trainer, whose loss is the mean NLL times train_size).  inject_eps [steps, num_ens, D] replaces the Philox draws."""
    pri = dict(DEFAULT_PRIORS if priors is None else priors)
    tspecs = [train_specs] if isinstance(train_specs, LogProbSpec) else list(train_specs)
    vspecs = [] if valid_specs is None else ([valid_specs] if isinstance(valid_specs, LogProbSpec) else list(valid_specs))
    tpre = [engine.prepare(likelihood_only(s)) for s in tspecs]
    vpre = [engine.prepare(likelihood_only(s)) for s in vspecs]
    dev = tpre[0].device
    D, E, nb = tpre[0].spec.D, int(num_ens), len(tpre)
    scales = [float(nll_scale)] * nb if np.isscalar(nll_scale) else [float(v) for v in nll_scale]
    vscale = float(valid_nll_scale) if np.isscalar(valid_nll_scale) else float(valid_nll_scale[0])
    if mu0 is None or rho0 is None:
        mu0, rho0 = init_posterior(D, pri, seed)
    mu = mu0.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
    rho = rho0.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
    if mu.numel() != D or rho.numel() != D:
        raise ValueError(f"mu0 / rho0 must have D = {D} entries")
    lib = _lib.load()
    cfg = _lib.ViCfg(num_ens=E, patience=int(lr_patience), kl_form=int(kl_form), lr_start=lr_start, min_lr=min_lr, lr_factor=0.1,
                     plateau_threshold=1e-4, beta=float(beta), prior_mu=float(pri["prior_mu"]), prior_sigma=float(pri["prior_sigma"]),
                     adam_b1=0.9, adam_b2=0.999, adam_eps=1e-8, seed=seed)
    nbytes = int(lib.vihmc_vi_workspace_bytes(D, E))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    wsp = ws.data_ptr()
    W_ptr, g_ptr, lp_ptr = (int(lib.vihmc_vi_buffer(wsp, D, E, k)) for k in (0, 1, 2))
    bm_ptr, br_ptr = (int(lib.vihmc_vi_buffer(wsp, D, E, k)) for k in (4, 5))
    history = torch.zeros((epochs, 3), dtype=torch.float32, device=dev)
    valid_logp = torch.zeros(max(len(vpre), 1), dtype=torch.float32, device=dev)
    eps_dev = None
    if inject_eps is not None:
        eps_dev = inject_eps.detach().to(device=dev, dtype=torch.float32).contiguous()
        if tuple(eps_dev.shape) != (epochs * nb, E, D):
            raise ValueError(f"inject_eps must be [epochs * batches, num_ens, D] = {(epochs * nb, E, D)}")
    wss_t = [p.workspace(E) for p in tpre]
    wss_v = [p.workspace(1) for p in vpre]

    def one_epoch():
        st = torch.cuda.current_stream(dev).cuda_stream
        for b, p in enumerate(tpre):
            _lib.check(lib.vihmc_vi_draw(C.byref(cfg), D, mu.data_ptr(), rho.data_ptr(), None if eps_dev is None else eps_dev.data_ptr(),
                                         wsp, nbytes, st))
            _lib.check(lib.vihmc_logp_grad(C.byref(p.problem), E, W_ptr, lp_ptr, g_ptr, wss_t[b].data_ptr(), wss_t[b].numel(), st))
            _lib.check(lib.vihmc_vi_step(C.byref(cfg), D, scales[b], mu.data_ptr(), rho.data_ptr(), wsp, nbytes, st))
        for b, p in enumerate(vpre):   # model.eval(): weights = mu
            _lib.check(lib.vihmc_logp_grad(C.byref(p.problem), 1, mu.data_ptr(), valid_logp.data_ptr() + 4 * b, None,
                                           wss_v[b].data_ptr(), wss_v[b].numel(), st))
        _lib.check(lib.vihmc_vi_epoch_end(C.byref(cfg), D, valid_logp.data_ptr(), len(vpre), vscale, mu.data_ptr(), rho.data_ptr(),
                                          history.data_ptr(), wsp, nbytes, st))

    with torch.cuda.device(dev):
        _lib.check(lib.vihmc_vi_init(C.byref(cfg), D, wsp, nbytes, torch.cuda.current_stream(dev).cuda_stream))
        if use_graph and epochs > 2:
            # capture one epoch on a side stream (kernel attributes were set by an eager epoch first) and replay it
            one_epoch()
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    one_epoch()
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(epochs - 1):
                graph.replay()
        else:
            for _ in range(epochs):
                one_epoch()
        best_mu = torch.empty(D, dtype=torch.float32, device=dev)
        best_rho = torch.empty(D, dtype=torch.float32, device=dev)
        best_mu.copy_(_view(bm_ptr, D, dev))
        best_rho.copy_(_view(br_ptr, D, dev))
        torch.cuda.current_stream(dev).synchronize()
    if not vpre:   # no validation set: the "best" snapshot is the final state
        best_mu, best_rho = mu.clone(), rho.clone()
    return VIResult(mu.cpu(), rho.cpu(), best_mu.cpu(), best_rho.cpu(), history.cpu(), steps=epochs * nb)


def _view(ptr: int, n: int, dev) -> torch.Tensor:
    """A torch view of n floats of caller-owned device memory (the trainer's workspace, alive for the duration of the call)."""
    class _Mem:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Mem(), device=dev)


def save_artifacts(directory: str, uid: str, result: VIResult, best: bool = True) -> None:
    """means_flattened_<uid> / stds_flattened_<uid> as sensitivity.py:211-222 derives them from the best checkpoint."""
    from . import artifacts

    artifacts.save_vi_artifacts(directory, uid, result.best_mu if best else result.mu, result.best_sigma if best else result.sigma)
