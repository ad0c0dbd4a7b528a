"""TEST INFRASTRUCTURE ONLY -- restatement of the third-party sampler the reference calls.

PARITY UNPINNED.  The reference does not own its sampler: every ``main_*.py`` calls
``hamiltorch.samplers.sample`` / ``hamiltorch.sample_model`` and ``requirements.txt:1`` installs
hamiltorch from an UNPINNED git HEAD (github.com/AdamCobb/hamiltorch).  The package is not in this
container, not vendored by the reference, and there is no network, so this file restates its
published algorithm (Neal 2011 leapfrog HMC, identity mass; Cobb & Jalaian 2021 symmetric split
integrator; Hoffman & Gelman 2014 dual averaging) together with hamiltorch's bookkeeping as the
reference's call sites rely on it.  This restatement is the specification the CUDA engine is tested
against; the log-posterior closures it drives ARE pinned against reference code (closures.py).

Reference call sites that fix the surface
-----------------------------------------
  Neural_network/VI_HMC/main_VI_HMC.py:379-381,420          sample(...); np.save; params[burn:]
  Neural_network/HMC/main_regression_hmc.py:124-127         sample_model(..., 'regression', tau_list)
  Operator_network/VI_HMC/main_VI_HMC_burgers.py:286-290    sample(..., sampler=Sampler.HMC)
  Operator_network/HMC/main_HMC_splitting.py:362-369        integrator=Integrator.SPLITTING (+NUTS, burn)
  Operator_network/HMC/NUTS_DeepOnets.py:289-290            sampler=Sampler.HMC_NUTS, burn=cfg.burn

Behaviour restated (and mirrored bit-for-bit in spirit by the CUDA engine)
--------------------------------------------------------------------------
* per sample n = 0..S-1: p ~ N(0,I); H0 = -logp(q) + 0.5 p.p; leapfrog; H1; rho = min(0, H0-H1);
  accept iff rho >= log(u), u ~ U(0,1) drawn AFTER the trajectory.
* leapfrog: p += eps/2 g(q); repeat L times { q += eps p; g = grad logp(q); p += eps g }; p -= eps/2 g.
* non-finite log-probability => LogProbError => the proposal is rejected (util.py:106-118).
* storage: the returned list starts with params_init; iteration n appends only when n > burn
  (accepted proposal, or a repeat of the last stored entry on reject); while n <= burn rejects fall
  back to the last accepted burn state.  Returned length = num_samples - burn.  Consequence kept on
  purpose: the first post-burn rejection falls back to the last STORED entry, i.e. params_init.
* split integrator over M closures (Hamiltonian = sum of the M closures): per step a forward sweep
  m = 0..M-1 { p += eps/2 g_m(q); if m < M-1: q += eps/(2(M-1)) p } then the mirrored reverse sweep.
* HMC_NUTS here means fixed-L HMC with dual-averaging step-size adaptation while n < burn, then
  eps = eps_bar (desired accept 0.8, gamma 0.05, t0 10, kappa 0.75, mu = log(10 eps0)).

``momenta`` / ``uniforms`` inject the random streams so the CUDA path can be fed identical draws.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Union

import torch

LogProb = Callable[[torch.Tensor], torch.Tensor]


class Sampler:
    HMC = 1
    HMC_NUTS = 3


class Integrator:
    IMPLICIT = 1   # hamiltorch's default value; for Sampler.HMC it selects plain leapfrog
    EXPLICIT = 2
    SPLITTING = 3


class LogProbError(Exception):
    pass


def has_nan_or_inf(value: torch.Tensor) -> bool:
    v = torch.sum(value)
    return bool(torch.isnan(v)) or bool(torch.isinf(v))


def params_grad(log_prob_func: LogProb, q: torch.Tensor) -> torch.Tensor:
    p = q.detach().requires_grad_()
    lp = log_prob_func(p)
    return torch.autograd.grad(lp.sum(), p)[0]


def gibbs(q: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    if generator is None:
        return torch.distributions.Normal(torch.zeros_like(q), torch.ones_like(q)).sample()
    return torch.randn(q.shape, dtype=q.dtype, generator=generator)


def hamiltonian(q: torch.Tensor, p: torch.Tensor, log_prob_func: Union[LogProb, Sequence[LogProb]]) -> torch.Tensor:
    with torch.no_grad():
        if isinstance(log_prob_func, (list, tuple)):
            log_prob = 0
            for f in log_prob_func:
                log_prob = log_prob + f(q).sum()
                if has_nan_or_inf(log_prob):
                    raise LogProbError()
        else:
            log_prob = log_prob_func(q).sum()
            if has_nan_or_inf(log_prob):
                raise LogProbError()
        return -log_prob + 0.5 * torch.dot(p, p)


def leapfrog(q: torch.Tensor, p: torch.Tensor, log_prob_func, steps: int, step_size: float,
             integrator: int = Integrator.IMPLICIT):
    q = q.detach().clone()
    p = p.clone()
    if integrator != Integrator.SPLITTING:
        g = params_grad(log_prob_func, q)
        p += 0.5 * step_size * g
        for _ in range(steps):
            q = q + step_size * p
            g = params_grad(log_prob_func, q)
            p += step_size * g
        p = p - 0.5 * step_size * g
        return q, p
    M = len(log_prob_func)
    if M == 1:
        raise NotImplementedError("splitting needs at least two closures")
    k_div = (M - 1) * 2
    for _ in range(steps):
        for m in range(M):
            g = params_grad(log_prob_func[m], q)
            p += 0.5 * step_size * g
            if m < M - 1:
                q += (step_size / k_div) * p
        for m in reversed(range(M)):
            g = params_grad(log_prob_func[m], q)
            p += 0.5 * step_size * g
            if m > 0:
                q += (step_size / k_div) * p
    return q, p


def adaptation(rho: float, t: int, step_size_init: float, H_t: float, eps_bar: float, desired_accept_rate: float = 0.8):
    """Dual averaging (Hoffman & Gelman 2014, alg. 5) with hamiltorch's constants, in fp32 like hamiltorch."""
    t = t + 1
    if math.isnan(rho) or math.isinf(rho):
        alpha = 0.0
    else:
        alpha = min(1.0, float(torch.exp(torch.FloatTensor([rho]))))
    mu = float(torch.log(10 * torch.FloatTensor([step_size_init])))
    gamma, t0, kappa = 0.05, 10, 0.75
    H_t = (1 - (1 / (t + t0))) * H_t + (1 / (t + t0)) * (desired_accept_rate - alpha)
    x_new = mu - (t ** 0.5) / gamma * H_t
    step_size = float(torch.exp(torch.FloatTensor([x_new])))
    x_new_bar = t ** -kappa * x_new + (1 - t ** -kappa) * torch.log(torch.FloatTensor([eps_bar]))
    eps_bar = float(torch.exp(x_new_bar))
    return step_size, eps_bar, H_t


def sample(log_prob_func, params_init: torch.Tensor, num_samples: int = 10, num_steps_per_sample: int = 10,
           step_size: float = 0.1, burn: int = 0, sampler: int = Sampler.HMC, integrator: int = Integrator.IMPLICIT,
           desired_accept_rate: float = 0.8, debug=False, momenta: Optional[torch.Tensor] = None,
           uniforms: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
           trace: Optional[dict] = None) -> List[torch.Tensor]:
    """Restated hamiltorch.samplers.sample (see module docstring).  ``trace`` (optional dict) receives
    per-iteration lists 'H0','H1','accept','step_size' for the parity tests."""
    if params_init.dim() != 1:
        raise RuntimeError("params_init must be a 1d tensor.")
    if burn >= num_samples:
        raise RuntimeError("burn must be less than num_samples.")
    nuts = False
    if sampler == Sampler.HMC_NUTS:
        if burn == 0:
            raise RuntimeError("burn must be greater than 0 for NUTS.")
        nuts, step_size_init, H_t, eps_bar = True, step_size, 0.0, 1.0

    params = params_init.clone()
    param_burn_prev = params_init.clone()
    ret_params = [params.clone()]
    num_rejected = 0
    if trace is not None:
        for k in ("H0", "H1", "accept", "step_size"):
            trace.setdefault(k, [])

    for n in range(num_samples):
        rho = float("nan")
        h0 = h1 = float("nan")
        accepted = False
        if trace is not None:
            trace["step_size"].append(step_size)
        try:
            momentum = momenta[n].clone() if momenta is not None else gibbs(params, generator)
            ham = hamiltonian(params, momentum, log_prob_func)
            h0 = float(ham)
            q_new, p_new = leapfrog(params, momentum, log_prob_func, num_steps_per_sample, step_size, integrator)
            new_ham = hamiltonian(q_new, p_new, log_prob_func)
            h1 = float(new_ham)
            rho = min(0.0, float(-new_ham + ham))
            if uniforms is not None:
                log_u = torch.log(uniforms[n].reshape(1).to(torch.float32))
            elif generator is not None:
                log_u = torch.log(torch.rand(1, generator=generator))
            else:
                log_u = torch.log(torch.rand(1))
            if bool(rho >= log_u):
                accepted = True
                params = q_new
                if n > burn:
                    ret_params.append(q_new.clone())
                else:
                    param_burn_prev = q_new.clone()
            else:
                num_rejected += 1
                if n > burn:
                    params = ret_params[-1].clone()
                    ret_params.append(ret_params[-1].clone())
                else:
                    params = param_burn_prev.clone()
            if nuts and n <= burn:
                if n < burn:
                    step_size, eps_bar, H_t = adaptation(rho, n, step_size_init, H_t, eps_bar, desired_accept_rate)
                if n == burn:
                    step_size = eps_bar
        except LogProbError:
            num_rejected += 1
            if n > burn:
                params = ret_params[-1].clone()
                ret_params.append(ret_params[-1].clone())
            else:
                params = param_burn_prev.clone()
            if nuts and n <= burn:
                step_size, eps_bar, H_t = adaptation(float("nan"), n, step_size_init, H_t, eps_bar, desired_accept_rate)
            if nuts and n == burn:
                step_size = eps_bar
        if trace is not None:
            trace["H0"].append(h0)
            trace["H1"].append(h1)
            trace["accept"].append(accepted)
    if trace is not None:
        trace["acceptance_rate"] = 1 - num_rejected / num_samples
        trace["final_step_size"] = step_size
    return [t.detach() for t in ret_params]
