"""Recover a :class:`LogProbSpec` from one of the reference's OWN ``log_prob_func`` closures.

The reference's drivers build a Python closure and hand it to ``hamiltorch.samplers.sample``
(Neural_network/VI_HMC/main_VI_HMC.py:373-380, Operator_network/VI_HMC/main_VI_HMC_burgers.py:279-287,
Operator_network/HMC/main_HMC_splitting.py:355-369, Operator_network/HMC/NUTS_DeepOnets.py:283-290).  A CUDA
engine cannot call Python per leapfrog step, but everything such a closure computes is determined by what it
*captured*, and Python exposes that: ``fn.__code__.co_freevars`` / ``fn.__closure__`` hold the data tensors, the
prior ``torch.distributions.Normal`` objects, the likelihood name and scale, and ``fmodel`` -- a bound method whose
``__self__`` is the reference's ``Functional_Net`` / ``Functional_DeepONet`` with the VI means, the sampled index
set and the ``nn.Module`` that fixes the layer table.  ``spec_from_closure`` reads those and nothing else, so the
reference's ``main_*.py`` run UNMODIFIED on top of the ``hamiltorch`` drop-in package next to this one.

Because introspection could silently mis-read a closure that was edited upstream, ``verify_closure`` evaluates the
closure itself ONCE (value and ``autograd.grad``, the reference's own code on its own device) at the initial
state and compares with the engine's result for the recovered spec; a mismatch raises.  This is a set-up check,
not a fallback: the sampling arithmetic never runs through the closure.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .spec import DeepONetArch, LogProbSpec, MLPArch

__all__ = ["ClosureError", "closure_freevars", "spec_from_closure", "verify_closure"]


class ClosureError(TypeError):
    """The callable is not one of the reference's log-posterior closures (or captured something unsupported)."""


def closure_freevars(fn) -> Dict[str, object]:
    """{free variable name: captured object}; cells that were never filled (``nll_loss`` when the loss is not NLL) are skipped."""
    code, cells = getattr(fn, "__code__", None), getattr(fn, "__closure__", None)
    if code is None or cells is None:
        raise ClosureError("log_prob_func is neither a LogProbSpec nor a Python closure with captured variables")
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def _act_name(f) -> str:
    """F.tanh / F.relu / the reference's Sin module (my_make_func.py:19-21,36-41; Operator...:24-28) -> spec name."""
    name = getattr(f, "__name__", type(f).__name__).lower()
    if name in ("tanh", "relu"):
        return name
    if name in ("sin", "sine"):
        return "sine"
    raise ClosureError(f"unrecognised activation {f!r} in the functional model")


def _cpu(t) -> torch.Tensor:
    return torch.as_tensor(t).detach().to(torch.float32).cpu()


def _normal(dist):
    if not isinstance(dist, torch.distributions.Normal):
        raise ClosureError(f"prior {dist!r} is not a torch.distributions.Normal")
    return dist.loc.detach().float().cpu(), dist.scale.detach().float().cpu()


def _sliced_prior(d: int, numels, dists):
    """main_VI_HMC.py:107-112: the loop walks the REDUCED vector with the FULL tensors' lengths; coordinate i gets the
    Normal of whichever slice contains it, coordinates past the last slice get no prior (sigma = inf)."""
    mu = np.zeros(d, dtype=np.float32)
    sig = np.full(d, np.inf, dtype=np.float32)
    i_prev = 0
    for n, dist in zip(numels, dists):
        loc, scale = _normal(dist)
        if loc.numel() != 1:
            raise ClosureError("per-tensor priors must be scalar Normals")
        hi = min(d, i_prev + int(n))
        if i_prev < hi:
            mu[i_prev:hi] = float(loc)
            sig[i_prev:hi] = float(scale)
        i_prev += int(n)
    return mu, sig


def spec_from_closure(fn) -> LogProbSpec:
    """Structured specification of a reference ``log_prob_func`` closure (see module docstring)."""
    fv = closure_freevars(fn)
    fmodel = fv.get("fmodel")
    owner = getattr(fmodel, "__self__", None)
    if owner is None or "dist_list" not in fv or "model_loss" not in fv or "tau_out" not in fv:
        raise ClosureError("not a reference define_model_log_prob closure: expected captured fmodel (bound method of the "
                           f"functional model), dist_list, model_loss and tau_out; found {sorted(fv)}")
    cfg = getattr(fn, "__globals__", {}).get("cfg")
    load_prior = bool(getattr(cfg, "load_prior", False))
    predict = bool(fv.get("predict", False))
    model = getattr(owner, "model", None) or fv.get("model")
    if not isinstance(model, torch.nn.Module):
        raise ClosureError("the functional model carries no nn.Module to read the layer table from")

    mus = getattr(owner, "learned_mus", None)
    frozen = sens_ind = vi_sigma = None
    if mus is not None:
        # my_make_func.py:56-57 scatters into sampled_weights (== learned_mus unless the redraw hook has fired)
        frozen = _cpu(getattr(owner, "sampled_weights", mus))
        sens_ind = np.asarray(owner.sensitive_ind, dtype=np.int64).reshape(-1)
        sig = getattr(owner, "learned_sigmas", None)
        vi_sigma = None if sig is None else _cpu(sig)

    is_deeponet = hasattr(owner, "depth_branch")
    trunk_subsample = None
    if is_deeponet:
        base = DeepONetArch.from_module(model)
        arch = DeepONetArch(width_branch=base.width_branch, width_trunk=base.width_trunk, in_branch=base.in_branch,
                            in_trunk=base.in_trunk, depth_branch=int(owner.depth_branch), depth_trunk=int(owner.depth_trunk),
                            output_neurons=base.output_neurons, act=_act_name(owner.act),
                            impose_bc=bool(getattr(owner, "impose_bc", True)))
        if "tr_data" not in fv:
            raise ClosureError("DeepONet closure without captured tr_data")
        # cfg.sample_data: a fresh random.sample(range(P), cfg.p) of the trunk points per closure call (main_VI_HMC_burgers.py:127-137);
        # only closures whose body reads cfg.sample_data have it (the full-HMC drivers do not)
        if not predict and "sample_data" in fn.__code__.co_names and bool(getattr(cfg, "sample_data", False)):
            trunk_subsample = int(cfg.p)
        x1, x2, y = fv["tr_data"]
        x, x2, y = _cpu(x1).reshape(x1.shape[0], -1), _cpu(x2).reshape(-1, x2.shape[-1]), _cpu(y)
    else:
        base = MLPArch.from_module(model)
        arch = MLPArch(in_dim=base.in_dim, widths=base.widths, out_dim=base.out_dim, act=_act_name(owner.activation),
                       last_bias=bool(getattr(owner, "bias", base.last_bias)))
        if len(arch.widths) != int(owner.depth) + 1:
            raise ClosureError(f"functional depth {owner.depth} does not match the module ({len(arch.widths)} hidden layers)")
        if fv.get("x") is None:
            raise NotImplementedError("prior-only sampling (x is None, main_VI_HMC.py:114-116) is not on the accelerated path")
        x, x2, y = _cpu(fv["x"]), None, _cpu(fv["y"])
    D = arch.num_params
    d = D if sens_ind is None else int(sens_ind.size)

    dists = list(fv["dist_list"])
    prior_mu = prior_sigma = None
    scal = 1.0
    if load_prior or "params_flattened_list" not in fv:
        # one Normal over the whole sampled vector (main_VI_HMC.py:104-105, main_VI_HMC_burgers.py:96-102); a closure whose
        # body runs the per-tensor slice loop captures params_flattened_list (main_VI_HMC.py:107-112, NUTS_DeepOnets.py:144-150)
        loc, scale = _normal(dists[0])
        if scale.numel() == 1:
            scal = float(scale)
            prior_mu = None if float(loc) == 0.0 else torch.full((d,), float(loc))
        else:
            prior_mu, prior_sigma = loc.reshape(-1), scale.reshape(-1)
    else:
        mu_np, sig_np = _sliced_prior(d, fv["params_flattened_list"], dists)
        prior_mu = None if not mu_np.any() else torch.from_numpy(mu_np)
        prior_sigma = torch.from_numpy(sig_np)

    tau_out = fv["tau_out"]
    spec = LogProbSpec(arch=arch, x=x, x2=x2, y=y, loss=fv["model_loss"], tau_out=float(tau_out), prior_mu=prior_mu,
                       prior_sigma=prior_sigma, prior_sigma_scalar=scal, prior_scale=float(fv.get("prior_scale", 1.0)),
                       frozen=frozen, sens_ind=sens_ind, vi_sigma=vi_sigma, predict=predict, trunk_subsample=trunk_subsample)
    spec.validate()
    return spec


def specs_from(log_prob_func) -> List[LogProbSpec]:
    """A spec, a closure, or a list of either (``Integrator.SPLITTING``) -> list of specs."""
    items = list(log_prob_func) if isinstance(log_prob_func, (list, tuple)) else [log_prob_func]
    return [it if isinstance(it, LogProbSpec) or not callable(it) else spec_from_closure(it) for it in items]


def verify_closure(fn, spec: LogProbSpec, q: torch.Tensor, rtol: float = 1e-4, prepared=None) -> Dict[str, float]:
    """Evaluate the closure once (its own code, its own device) at ``q`` and compare value and gradient with the engine's
    for the recovered spec.  Tolerances: |dlogp| <= rtol*|logp|, max|dgrad| <= rtol*max|grad| (both sides are fp32
    sums over the data set in different orders, so this is 10x the parity bar of the kernel tests)."""
    from . import engine

    q = q.detach().reshape(-1).to(torch.float32)
    dev = spec.x.device
    for v in closure_freevars(fn).values():
        if torch.is_tensor(v):
            dev = v.device
            break
    import random

    rstate = random.getstate()   # a subsampling closure draws its trunk subset from Python's global generator: replay it for the engine
    p = q.to(dev).clone().requires_grad_()
    out = fn(p)
    lp = (out[0] if isinstance(out, tuple) else out).sum()
    (g,) = torch.autograd.grad(lp, p)
    lp_ref, g_ref = float(lp.detach()), g.detach().cpu()
    if spec.trunk_subsample is not None:
        random.setstate(rstate)
        prep = engine.prepare(prepared if prepared is not None else spec)
        sub = engine._TrunkSubset(prep, random.sample(range(spec.P), int(spec.trunk_subsample)))
        lp_eng, g_eng = sub.logp_grad(engine._to_dev(q.reshape(1, -1), prep.device))
        random.setstate(rstate)   # the check leaves the generator where the caller had it
    else:
        lp_eng, g_eng = engine.logp_grad(prepared if prepared is not None else spec, q.reshape(1, -1).cpu())
    lp_eng, g_eng = float(lp_eng.reshape(-1)[0]), g_eng.reshape(-1).cpu()
    err_lp = abs(lp_eng - lp_ref) / max(abs(lp_ref), 1e-30)
    scale = max(float(g_ref.abs().max()), 1e-30)
    err_g = float((g_eng - g_ref).abs().max()) / scale
    if not (err_lp <= rtol and err_g <= rtol):
        raise ClosureError(f"the recovered specification does not reproduce the closure at params_init: "
                           f"log-posterior {lp_eng!r} vs {lp_ref!r} (rel {err_lp:.2e}), gradient rel {err_g:.2e} > {rtol:g}")
    return {"logp_rel_err": err_lp, "grad_rel_err": err_g}
