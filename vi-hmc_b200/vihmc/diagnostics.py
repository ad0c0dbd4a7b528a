"""Convergence diagnostics for batched chains.  The reference computes none (grep for ess|rhat|autocorr in
/root/reference finds nothing), so the definitions are fixed here:

  * rank-normalised split-R-hat and bulk-ESS, Vehtari, Gelman, Simpson, Carpenter, Buerkner (2021);
  * autocorrelations by FFT, truncated with Geyer's initial positive / monotone sequence;
  * computed per coordinate over post-burn draws, chains pooled; report min and median over coordinates.

Draws are tensors [S, C, ...] (draw, chain, coordinate...), on any device (torch ops only).
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def split_chains(x: torch.Tensor) -> torch.Tensor:
    """[S, C, ...] -> [S//2, 2C, ...]: each chain is cut into two halves (odd S drops the middle draw)."""
    S = x.shape[0]
    h = S // 2
    return torch.cat([x[:h], x[S - h:]], dim=1)


def rank_normalize(x: torch.Tensor) -> torch.Tensor:
    """z-scores of the pooled ranks per coordinate: z = Phi^-1((r - 3/8) / (n + 1/4)), average ranks for ties."""
    S, C = x.shape[:2]
    flat = x.reshape(S * C, -1).double()
    n = flat.shape[0]
    order = flat.argsort(dim=0, stable=True)
    ranks = torch.empty_like(flat)
    ar = torch.arange(1, n + 1, dtype=torch.float64, device=x.device).unsqueeze(1).expand_as(flat)
    ranks.scatter_(0, order, ar)
    # average the ranks of tied values (rejected proposals repeat draws exactly)
    srt = flat.gather(0, order)
    new_run = torch.ones_like(srt, dtype=torch.bool)
    new_run[1:] = srt[1:] != srt[:-1]
    run_id = new_run.long().cumsum(0) - 1
    sums = torch.zeros_like(srt).scatter_add_(0, run_id, ar)
    cnts = torch.zeros_like(srt).scatter_add_(0, run_id, torch.ones_like(srt))
    avg_sorted = (sums / cnts.clamp_min(1)).gather(0, run_id)
    ranks.scatter_(0, order, avg_sorted)
    z = torch.special.ndtri((ranks - 0.375) / (n + 0.25))
    return z.reshape(x.shape)


def _rhat_plain(x: torch.Tensor) -> torch.Tensor:
    """Potential scale reduction of [S, C, ...] without splitting or rank normalisation."""
    x = x.double()
    S = x.shape[0]
    chain_mean = x.mean(0)
    chain_var = x.var(0, unbiased=True)
    W = chain_var.mean(0)
    B = S * chain_mean.var(0, unbiased=True)
    var_plus = (S - 1) / S * W + B / S
    return torch.sqrt(var_plus / W)


def rhat_from_moments(chain_mean: torch.Tensor, chain_var: torch.Tensor, S: int) -> torch.Tensor:
    """Split-free R-hat from per-chain sufficient statistics [C, ...] (what ranks exchange across GPUs)."""
    W = chain_var.double().mean(0)
    B = S * chain_mean.double().var(0, unbiased=True)
    return torch.sqrt(((S - 1) / S * W + B / S) / W)


def half_chain_moments(x: torch.Tensor):
    """Per half-chain mean / unbiased variance of [S, C, ...] -> two tensors [2C, ...] and the half length."""
    xs = split_chains(x).double()
    return xs.mean(0), xs.var(0, unbiased=True), xs.shape[0]


def split_rhat(x: torch.Tensor) -> torch.Tensor:
    return _rhat_plain(split_chains(x))


def rank_split_rhat(x: torch.Tensor) -> torch.Tensor:
    """max of the bulk and the folded (tail) rank-normalised split-R-hat."""
    xs = split_chains(x)
    bulk = _rhat_plain(rank_normalize(xs))
    med = xs.reshape(-1, *xs.shape[2:]).double().median(0).values
    folded = _rhat_plain(rank_normalize((xs.double() - med).abs()))
    return torch.maximum(bulk, folded)


def _autocov(x: torch.Tensor) -> torch.Tensor:
    """FFT autocovariance along dim 0 of [S, ...] (biased, divides by S)."""
    S = x.shape[0]
    n = 1 << (2 * S - 1).bit_length()
    xc = x - x.mean(0, keepdim=True)
    f = torch.fft.rfft(xc, n=n, dim=0)
    ac = torch.fft.irfft(f * f.conj(), n=n, dim=0)[:S]
    return ac / S


def ess(x: torch.Tensor) -> torch.Tensor:
    """Effective sample size of [S, C, ...] (chains pooled), Geyer initial monotone sequence (Stan's estimator)."""
    x = x.double()
    S, C = x.shape[:2]
    if S < 4:
        return torch.full(x.shape[2:], float("nan"), dtype=torch.float64, device=x.device)
    acov = _autocov(x)                                 # [S, C, ...]
    chain_mean = x.mean(0)
    mean_var = acov[0].mean(0) * S / (S - 1)
    var_plus = mean_var * (S - 1) / S
    if C > 1:
        var_plus = var_plus + chain_mean.var(0, unbiased=True)
    rho = 1.0 - (mean_var - acov.mean(1)) / var_plus   # [S, ...]
    rho[0] = 1.0
    # pair sums P_t = rho_{2t} + rho_{2t+1}; keep while positive, enforce monotone decrease
    T = (S // 2) * 2
    pairs = rho[:T].reshape(T // 2, 2, *rho.shape[1:]).sum(1)
    positive = (pairs > 0).long().cumprod(0).bool()
    pairs = torch.where(positive, pairs, torch.zeros_like(pairs))
    pairs = torch.cummin(pairs, dim=0).values
    tau = -1.0 + 2.0 * pairs.sum(0)
    tau = torch.maximum(tau, torch.full_like(tau, 1.0 / math.log10(max(S * C, 11))))
    return S * C / tau


def bulk_ess(x: torch.Tensor) -> torch.Tensor:
    return ess(rank_normalize(split_chains(x)))


def summarize(draws: torch.Tensor, logp: torch.Tensor = None) -> Dict[str, float]:
    """min / median over coordinates of rank-normalised split-R-hat and bulk-ESS (+ the scalar log-posterior)."""
    d = draws.reshape(draws.shape[0], draws.shape[1], -1)
    r, e = rank_split_rhat(d), bulk_ess(d)
    out = {"rhat_max": float(r.max()), "rhat_median": float(r.median()), "ess_bulk_min": float(e.min()),
           "ess_bulk_median": float(e.median()), "draws": int(d.shape[0]), "chains": int(d.shape[1])}
    if logp is not None:
        lp = logp.reshape(logp.shape[0], logp.shape[1], 1)
        out["rhat_logp"] = float(rank_split_rhat(lp).max())
        out["ess_bulk_logp"] = float(bulk_ess(lp).min())
    return out
