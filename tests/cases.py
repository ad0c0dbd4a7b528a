"""Shared builders: the same synthetic inputs as oracle/make_golden.py, turned into oracle closures
(CPU checker) and into product LogProbSpecs (thing under test)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import closures as oc
from vihmc import synth
from vihmc.spec import DeepONetArch, LogProbSpec, MLPArch, sliced_prior_sigma

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BNN_CASES = ["d40_nll", "d141_nll", "d14_nll", "d70_regression", "d40_relu", "d40_sine", "d40_loadprior"]

DON_ARCHS = {
    "small": (DeepONetArch(width_branch=16, width_trunk=16, in_branch=12, depth_branch=3, depth_trunk=4,
                           output_neurons=8), 6, 5, 7, 0.25),
    "full": (DeepONetArch(), 8, 3, 11, 0.10),
}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bnn_case(golden, name):
    g = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}
    d, loss, act = int(g["d"]), str(g["loss"]), str(g["act"])
    tau_out, load_prior, prior_var = float(g["tau_out"]), bool(g["load_prior"]), float(g["prior_var"])
    mu, sigma, ind = synth.bnn_vi_artifacts(141, d, seed=1)
    return dict(d=d, loss=loss, act=act, tau_out=tau_out, load_prior=load_prior, prior_var=prior_var,
                mu=mu, sigma=sigma, ind=ind, q=g["q"], logp=g["logp"], grad=g["grad"])


def bnn_oracle(case, dtype=torch.float32):
    x, y, _, _ = synth.bnn_data()
    if case["load_prior"]:
        prior = ("loc_scale", case["mu"][case["ind"]], case["sigma"][case["ind"]])
    else:
        prior = ("sliced", [case["prior_var"]] * 6)
    return oc.BnnLogProb(x=x, y=y, widths=(10, 10), act=case["act"], last_bias=True, loss=case["loss"],
                         tau_out=case["tau_out"], prior=prior, frozen=case["mu"], sens_ind=case["ind"], dtype=dtype)


def bnn_spec(case) -> LogProbSpec:
    x, y, _, _ = synth.bnn_data()
    arch = MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act=case["act"], last_bias=True)
    if case["load_prior"]:
        pm, ps, pss = case["mu"][case["ind"]].clone(), case["sigma"][case["ind"]].clone(), 1.0
    else:
        sig = sliced_prior_sigma(case["d"], arch.tensor_numels(), [case["prior_var"]] * 6)
        pm, ps, pss = None, torch.from_numpy(sig.astype(np.float32)), 1.0
    return LogProbSpec(arch=arch, x=x, y=y, loss=case["loss"], tau_out=case["tau_out"], prior_mu=pm, prior_sigma=ps,
                       prior_sigma_scalar=pss, frozen=case["mu"], sens_ind=case["ind"], vi_sigma=case["sigma"])


def don_inputs(name):
    arch, n_train, n_t, n_x, frac = DON_ARCHS[name]
    x1, x2, y, theta = synth.burgers_like(arch, n_train=n_train, n_t=n_t, n_x=n_x, seed=0)
    mu, sigma, ind = synth.deeponet_vi_artifacts(theta, frac=frac, seed=1)
    return dict(arch=arch, x1=x1, x2=x2, y=y, theta=theta, mu=mu, sigma=sigma, ind=ind)


def _don_kwargs(arch: DeepONetArch, dtype):
    return dict(width_branch=arch.width_branch, width_trunk=arch.width_trunk, in_branch=arch.in_branch,
                in_trunk=arch.in_trunk, depth_branch=arch.depth_branch, depth_trunk=arch.depth_trunk,
                output_neurons=arch.output_neurons, act=arch.act, impose_bc=arch.impose_bc, loss="NLL", tau_out=1.0,
                prior_var=0.1 ** 2, dtype=dtype)


def don_oracle(inp, mode, dtype=torch.float32):
    """mode: 'vi' (reduced vector), 'full', or 'split' (list of 2 closures)."""
    kw = _don_kwargs(inp["arch"], dtype)
    x1, x2, y = inp["x1"].unsqueeze(1), inp["x2"].unsqueeze(0), inp["y"]
    if mode == "vi":
        return oc.DeepONetLogProb(x1=x1, x2=x2, y=y, frozen=inp["mu"], sens_ind=inp["ind"], **kw)
    if mode == "full":
        return oc.DeepONetLogProb(x1=x1, x2=x2, y=y, **kw)
    return oc.split_deeponet(kw, x1, x2, y, 2)


def don_spec(inp, mode):
    arch = inp["arch"]
    common = dict(arch=arch, x2=inp["x2"], loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    if mode == "vi":
        return LogProbSpec(x=inp["x1"], y=inp["y"], frozen=inp["mu"], sens_ind=inp["ind"], vi_sigma=inp["sigma"], **common)
    if mode == "full":
        return LogProbSpec(x=inp["x1"], y=inp["y"], **common)
    n = inp["x1"].shape[0] // 2
    return [LogProbSpec(x=inp["x1"][i * n:(i + 1) * n], y=inp["y"][i * n:(i + 1) * n], prior_scale=2.0, **common)
            for i in range(2)]
