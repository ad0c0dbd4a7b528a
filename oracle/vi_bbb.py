"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's Bayes-by-Backprop training loop.

Follows Neural_network/VI/main_regression_VI.py:75-124 (train_model: loss = mean over num_ens draws of ELBO, one Adam step),
:127-170 (validate_model: model.eval(), weights = mu), :300-335 (Adam(lr_start), ReduceLROnPlateau(patience, min_lr=1e-5),
checkpoint on valid_loss <= valid_loss_min), layers/BBB/BBBLinear.py:53-79 (W = mu + eps * log1p(exp(rho)); kl_loss with the
(prior, posterior) arguments swapped into calculate_kl, metrics.py:47-49) and metrics.py:12-20 (ELBO = gaussian_nll_loss(sum) +
beta * kl).  The eps stream is an input so that the CUDA trainer can be compared step for step.
Pinned by tests/golden/bnn_vi_training.npz: the REAL reference classes run on the same eps stream (oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import closures as oc


def kl_reference(mu, sigma, prior_mu, prior_sigma):
    """calculate_kl(mu_q=prior_mu, sig_q=prior_sigma, mu_p=mu, sig_p=sigma) -- the call as BBBLinear.kl_loss makes it."""
    pm, ps = torch.as_tensor(prior_mu, dtype=mu.dtype), torch.as_tensor(prior_sigma, dtype=mu.dtype)
    return 0.5 * (2 * torch.log(sigma / ps) - 1 + (ps / sigma).pow(2) + ((mu - pm) / sigma).pow(2)).sum()


def train(x, y, x_val, y_val, widths, act, mu0, rho0, eps, noise_var, prior_mu, prior_sigma, lr_start, lr_patience, beta=1.0,
          min_lr=1e-5, dtype=torch.float32):
    """eps [epochs, num_ens, D].  Returns (mu, rho, best_mu, best_rho, history [epochs, 3] = train loss, valid loss, lr)."""
    slots = oc.mlp_layout(x.shape[1], widths, y.shape[1], True)
    mu = mu0.to(dtype).clone().requires_grad_()
    rho = rho0.to(dtype).clone().requires_grad_()
    x, y, x_val, y_val = (t.to(dtype) for t in (x, y, x_val, y_val))
    opt = torch.optim.Adam([mu, rho], lr=lr_start)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=lr_patience, min_lr=min_lr)
    hist, best, best_state = [], float("inf"), None

    def forward(w):
        return oc.mlp_forward(x_in, oc.unflatten(slots, w), len(widths), act, True)

    for ep in range(eps.shape[0]):
        lr_used = opt.param_groups[0]["lr"]
        opt.zero_grad()
        loss = 0.0
        for j in range(eps.shape[1]):
            sigma = torch.log1p(torch.exp(rho))
            w = mu + eps[ep, j].to(dtype) * sigma
            x_in = x
            pred = forward(w)
            loss = loss + F.gaussian_nll_loss(pred, y, noise_var * torch.ones_like(y), reduction="sum") + beta * kl_reference(
                mu, sigma, prior_mu, prior_sigma)
        loss = loss / eps.shape[1]
        loss.backward()
        opt.step()
        with torch.no_grad():
            sigma = torch.log1p(torch.exp(rho))
            x_in = x_val
            vloss = F.gaussian_nll_loss(forward(mu), y_val, noise_var * torch.ones_like(y_val), reduction="sum") + beta * kl_reference(
                mu, sigma, prior_mu, prior_sigma)
        sched.step(float(vloss))
        hist.append([float(loss.detach()), float(vloss), lr_used])
        if float(vloss) <= best:
            best, best_state = float(vloss), (mu.detach().clone(), rho.detach().clone())
    return mu.detach(), rho.detach(), best_state[0], best_state[1], np.asarray(hist, dtype=np.float64)


def train_batches(batches, valid_batches, forward, mu0, rho0, eps, noise_var, prior_mu, prior_sigma, lr_start, lr_patience, train_size,
                  valid_size, beta=1.0, min_lr=1e-5, dtype=torch.float32):
    """The mini-batch form of the loop, Operator_network/VI/main_VI_deeponet.py:56-80 (train) and :105-118 (validate) with
    metrics.py:29-31: loss = gaussian_nll_loss(mean) * train_size + beta * kl per batch, one Adam step per batch, epoch losses =
    means over the batches.  batches: list of (inputs, y); forward(w, inputs) -> prediction shaped like y.  eps [epochs * n_batches,
    num_ens, D].  Pinned by tests/golden/deeponet_vi_training.npz: the reference's own Bayesian_DeepONet training loop on the same eps stream."""
    mu = mu0.to(dtype).clone().requires_grad_()
    rho = rho0.to(dtype).clone().requires_grad_()
    opt = torch.optim.Adam([mu, rho], lr=lr_start)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=lr_patience, min_lr=min_lr)
    nb = len(batches)
    hist = []
    for ep in range(eps.shape[0] // nb):
        lr_used = opt.param_groups[0]["lr"]
        tl = 0.0
        for b, (inp, y) in enumerate(batches):
            opt.zero_grad()
            loss = 0.0
            for j in range(eps.shape[1]):
                sigma = torch.log1p(torch.exp(rho))
                w = mu + eps[ep * nb + b, j].to(dtype) * sigma
                pred = forward(w, inp)
                loss = loss + F.gaussian_nll_loss(pred, y.to(dtype), noise_var * torch.ones_like(pred), reduction="mean") * train_size \
                    + beta * kl_reference(mu, sigma, prior_mu, prior_sigma)
            loss = loss / eps.shape[1]
            loss.backward()
            opt.step()
            tl += float(loss.detach())
        vl = 0.0
        with torch.no_grad():
            sigma = torch.log1p(torch.exp(rho))
            for inp, y in valid_batches:
                pred = forward(mu, inp)
                vl += float(F.gaussian_nll_loss(pred, y.to(dtype), noise_var * torch.ones_like(pred), reduction="mean") * valid_size
                            + beta * kl_reference(mu, sigma, prior_mu, prior_sigma))
        vl /= max(len(valid_batches), 1)
        sched.step(vl)
        hist.append([tl / nb, vl, lr_used])
    return mu.detach(), rho.detach(), np.asarray(hist, dtype=np.float64)
