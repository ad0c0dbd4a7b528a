"""The validation step the reference runs right after sampling, on the CUDA engine.

  Neural_network/VI_HMC/main_VI_HMC.py:384-429 (validate)            expected validation log probability, expected / final MSE
  Operator_network/VI_HMC/main_VI_HMC_burgers.py:290-301             the same + min MSE, and sample_mse_<dt>.npy (one MSE per draw)

The reference materialises every prediction (S x N x P floats: 37 GB for 900 DeepONet draws on the full grid) and averages on the host.
The per-draw MSE does not need the predictions: for the Gaussian likelihood the value-only log-posterior kernel already reduces
sum (o - y)^2 over the validation set,
    NLL        loglik = -0.5 (n log v + SS / v)   =>   SS = -2 v loglik - n v log v
    regression loglik = -0.5 tau SS               =>   SS = -2 loglik / tau
so one vihmc_logp_grad (value only) call per chain batch gives the S log-probabilities and the S MSEs.  No CPU fallback.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import engine
from .spec import LogProbSpec


def _validation_spec(spec: LogProbSpec, x=None, y=None, data=None) -> LogProbSpec:
    if data is not None:
        x1, x2, yv = data
        return dataclasses.replace(spec, x=x1.reshape(x1.shape[0], -1).float().cpu(), x2=x2.reshape(-1, x2.shape[-1]).float().cpu(),
                                   y=yv.float().cpu())
    if x is not None and y is not None:
        return dataclasses.replace(spec, x=x.float().cpu(), y=y.float().cpu())
    raise RuntimeError('Val data not defined (i.e. arguments x, y, val_loader are all not defined)')


def sample_log_prob_and_mse(spec: LogProbSpec, samples, x=None, y=None, data=None, chunk: int = 256):
    """(log_prob [S], mse [S]) of every draw on the validation data: log_prob = log-likelihood + log-prior / prior_scale as
    predict_model's closure returns it, mse = mean (o - y)^2 over all validation outputs."""
    vspec = _validation_spec(spec, x, y, data)
    q = torch.stack(list(samples)) if isinstance(samples, (list, tuple)) else samples
    q = q.reshape(-1, vspec.d).float()
    full = engine.prepare(vspec)
    lik = engine.prepare(dataclasses.replace(vspec, prior_mu=None, prior_sigma=torch.full((vspec.d,), float("inf")), prior_scale=1.0))
    n = vspec.y.numel()
    lp, ll = [], []
    for i in range(0, q.shape[0], chunk):
        lp.append(engine.logp_grad(full, q[i:i + chunk], need_grad=False)[0])
        ll.append(engine.logp_grad(lik, q[i:i + chunk], need_grad=False)[0])
    lp, ll = torch.cat(lp).double(), torch.cat(ll).double()
    if vspec.loss == "NLL":
        v = max(float(vspec.tau_out), 1e-6)
        ss = -2.0 * v * ll - n * v * math.log(v)
    else:
        ss = -2.0 * ll / float(vspec.tau_out)
    return lp.float().cpu(), (ss / n).float().cpu()


def validate(spec: LogProbSpec, samples, burn: int = 0, x=None, y=None, data=None, out_dir: Optional[str] = None,
             uid: Optional[str] = None) -> Dict[str, float]:
    """The numbers the reference prints after sampling (main_VI_HMC_burgers.py:293-300); with out_dir / uid also writes
    sample_mse_<uid>.npy exactly as :301 does (one float32 per post-burn draw)."""
    q = torch.stack(list(samples)) if isinstance(samples, (list, tuple)) else samples
    q = q.reshape(-1, q.shape[-1])[burn:]
    lp, mse = sample_log_prob_and_mse(spec, q, x=x, y=y, data=data)
    out = {"expected_log_prob": float(lp.mean()), "expected_mse": float(mse.mean()), "final_mse": float(mse[-1]),
           "min_mse": float(mse.min()), "draws": int(mse.numel())}
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, f"sample_mse_{uid}.npy"), mse.numpy().astype(np.float32))
    return out
