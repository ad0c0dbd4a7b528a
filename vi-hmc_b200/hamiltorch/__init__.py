"""Drop-in for the ``hamiltorch`` package AS THE REFERENCE USES IT, backed by the vihmc CUDA engine.

The reference (ponkrshnan/VI-HMC) installs hamiltorch from git HEAD (requirements.txt:1) and touches exactly this
surface of it:

    import hamiltorch                                Neural_network/HMC/main_regression_hmc.py:12
    hamiltorch.util.flatten(net)                     main_regression_hmc.py:115
    hamiltorch.sample_model(net, x, y, ...)          main_regression_hmc.py:124-127
    hamiltorch.predict_model(net, x=, y=, ...)       main_regression_hmc.py:153-155
    import hamiltorch.samplers as samplers           Neural_network/VI_HMC/main_VI_HMC.py:21
    from hamiltorch import samplers                  Operator_network/VI_HMC/main_VI_HMC_burgers.py:16,
                                                     Operator_network/HMC/main_HMC_splitting.py:17, NUTS_DeepOnets.py:15
    samplers.sample(log_prob_func, params_init, ...) main_VI_HMC.py:379-380, main_VI_HMC_burgers.py:286-287,
                                                     main_HMC_splitting.py:362-369, NUTS_DeepOnets.py:289-290
    samplers.Sampler.HMC / HMC_NUTS, samplers.Integrator.SPLITTING

With ``vi-hmc_b200/`` on ``PYTHONPATH`` ahead of (or instead of) the upstream package, the reference's ``main_*.py`` and
``config.py`` run unmodified: ``samplers.sample`` receives the reference's own closure, recovers what it captured
(``vihmc.closure``), checks the recovered log-posterior against the closure once at ``params_init`` and samples on the
GPU.  One call = one chain, as upstream; ``num_chains=`` (or ``VIHMC_NUM_CHAINS``) runs many chains in the same call.
There is no CPU fallback: without the CUDA library or a GPU every sampling call raises ``vihmc._lib.VihmcError``.
"""
from vihmc import util  # noqa: F401  (flatten / unflatten / set_random_seed, as hamiltorch.util)
from vihmc.util import set_random_seed  # noqa: F401

from . import samplers  # noqa: F401
from .samplers import Integrator, Metric, Sampler, predict_model, sample, sample_model  # noqa: F401

__version__ = "0.4.0.dev1+vihmc"
