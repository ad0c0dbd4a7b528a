"""``hamiltorch.samplers`` as the reference calls it (see the package docstring for the call sites)."""
from __future__ import annotations

import os

import torch

from vihmc import samplers as _impl
from vihmc.samplers import Integrator, Metric, Sampler  # noqa: F401
from vihmc.samplers import define_model_log_prob_hamiltorch as define_model_log_prob  # noqa: F401


def _fresh_seed() -> int:
    """Upstream draws momenta from torch's global generator, so a script that seeds torch is reproducible and successive
    chains (the reference's ``for run_num in range(cfg.num_chains)`` loop, main_VI_HMC.py:458-460) differ.  The engine's
    Philox streams are keyed by an explicit seed: take it from the same global generator to keep both properties."""
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


def _num_chains(kw):
    n = kw.pop("num_chains", None)
    if n is None and os.environ.get("VIHMC_NUM_CHAINS"):
        n = int(os.environ["VIHMC_NUM_CHAINS"])
    return n


def sample(log_prob_func, params_init, num_samples=10, num_steps_per_sample=10, step_size=0.1, burn=0, jitter=None,
           inv_mass=None, normalizing_const=1., softabs_const=None, explicit_binding_const=100,
           fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, jitter_max_tries=10, sampler=Sampler.HMC,
           integrator=Integrator.IMPLICIT, metric=Metric.HESSIAN, debug=False, desired_accept_rate=0.8,
           store_on_GPU=True, pass_grad=None, verbose=False, **engine_kw):
    """hamiltorch.samplers.sample(log_prob_func, params_init, ...) -> list of ``num_samples - burn`` 1-D tensors.

    ``log_prob_func`` is the reference's closure (or a list of them with ``Integrator.SPLITTING``), or a
    ``vihmc.LogProbSpec``.  Samples come back on the host, so the reference's ``np.save(path, params_hmc)`` writes the
    same ``(S, d)`` float32 array."""
    if pass_grad is not None:
        raise NotImplementedError("pass_grad (a user-supplied gradient function) is not used by the reference")
    seed = engine_kw.pop("seed", None)
    return _impl.sample(log_prob_func, params_init, num_samples=num_samples, num_steps_per_sample=num_steps_per_sample,
                        step_size=step_size, burn=burn, inv_mass=inv_mass, sampler=sampler, integrator=integrator,
                        debug=debug, desired_accept_rate=desired_accept_rate, verbose=verbose,
                        num_chains=_num_chains(engine_kw), seed=_fresh_seed() if seed is None else seed, **engine_kw)


def sample_model(model, x, y, params_init, model_loss='multi_class_linear_output', num_samples=10,
                 num_steps_per_sample=10, step_size=0.1, burn=0, inv_mass=None, jitter=None, normalizing_const=1.,
                 softabs_const=None, explicit_binding_const=100, fixed_point_threshold=1e-5,
                 fixed_point_max_iterations=1000, jitter_max_tries=10, sampler=Sampler.HMC, integrator=Integrator.IMPLICIT,
                 metric=Metric.HESSIAN, debug=False, tau_out=1., tau_list=None, store_on_GPU=True,
                 desired_accept_rate=0.8, verbose=False, **engine_kw):
    """hamiltorch.sample_model (main_regression_hmc.py:124-127): Gaussian prior of precision tau per parameter tensor."""
    seed = engine_kw.pop("seed", None)
    return _impl.sample_model(model, x, y, params_init, model_loss=model_loss, num_samples=num_samples,
                              num_steps_per_sample=num_steps_per_sample, step_size=step_size, burn=burn, inv_mass=inv_mass,
                              normalizing_const=normalizing_const, sampler=sampler, integrator=integrator, debug=debug,
                              tau_out=tau_out, tau_list=tau_list, desired_accept_rate=desired_accept_rate, verbose=verbose,
                              num_chains=_num_chains(engine_kw), seed=_fresh_seed() if seed is None else seed, **engine_kw)


def predict_model(model, samples, x=None, y=None, test_loader=None, model_loss='multi_class_linear_output', tau_out=1.,
                  tau_list=None, verbose=False):
    """hamiltorch.predict_model (main_regression_hmc.py:153-155): outputs [S,N,O] and the log-probability of every sample
    on validation data, forward-only kernels."""
    if test_loader is not None:
        raise NotImplementedError("predict_model(test_loader=...) is not used by the reference; pass x and y")
    numels = [w.nelement() for w in model.parameters()]
    shapes = [w.shape for w in model.parameters()]
    if tau_list is None:
        tau_list = [torch.tensor(1.) for _ in numels]
    if x is None or y is None:
        raise RuntimeError('Val data not defined (i.e. arguments x, y, val_loader are all not defined)')
    spec = define_model_log_prob(model, model_loss, x, y, numels, shapes, tau_list, tau_out, predict=True)
    pred, logp = _impl.predict_model(spec, samples, x=x, y=y)
    if verbose:
        print('\nExpected validation log probability: {:.2f}'.format(torch.stack(logp).mean()))
    return pred, logp
