"""Bayes-by-Backprop trainer: oracle vs golden vectors produced by the reference's own Bayesian_Net /
train_model / validate_model / Adam / ReduceLROnPlateau loop on a recorded eps stream, and the CUDA trainer vs the same vectors."""
import numpy as np
import pytest
import torch

import cases
from oracle import vi_bbb as ovi
from vihmc import synth

VI_TRAIN_CASES = ["adam_6x3", "plateau_8x2"]


def _case(name):
    g = cases.load_golden("bnn_vi_training.npz")
    c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    noise_var, prior_mu, prior_sigma, beta, lr, patience = c["cfg"]
    return c, dict(noise_var=float(noise_var), prior_mu=float(prior_mu), prior_sigma=float(prior_sigma), beta=float(beta), lr_start=float(lr),
                   lr_patience=int(patience))


@pytest.mark.parametrize("name", VI_TRAIN_CASES)
def test_oracle_matches_reference_training_run(name):
    c, kw = _case(name)
    x, y, xv, yv = synth.bnn_data()
    mu, rho, bmu, brho, hist = ovi.train(x, y, xv, yv, (10, 10), "tanh", torch.from_numpy(c["mu0"]), torch.from_numpy(c["rho0"]),
                                         torch.from_numpy(c["eps"]), **kw)
    np.testing.assert_allclose(hist[:, 2], c["history"][:, 2], rtol=1e-6)          # the learning-rate schedule
    np.testing.assert_allclose(hist[:, :2], c["history"][:, :2], rtol=1e-4)
    np.testing.assert_allclose(mu.numpy(), c["mu"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(rho.numpy(), c["rho"], rtol=1e-3, atol=1e-4)


def test_kl_term_is_the_swapped_argument_form():
    """BBBLinear.kl_loss passes (prior, posterior) into calculate_kl(mu_q, sig_q, mu_p, sig_p): KL(prior || q), not KL(q || prior)."""
    mu, sg = torch.tensor([0.3, -1.0]), torch.tensor([0.05, 0.7])
    got = float(ovi.kl_reference(mu, sg, 0.0, 1.0))
    p, q = torch.distributions.Normal(0.0, 1.0), torch.distributions.Normal(mu, sg)
    assert got == pytest.approx(float(torch.distributions.kl_divergence(p, q).sum()), rel=1e-6)
    assert got != pytest.approx(float(torch.distributions.kl_divergence(q, p).sum()), rel=1e-2)


def _specs():
    x, y, xv, yv = synth.bnn_data()
    arch = cases.MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
    mk = lambda a, b, v: cases.LogProbSpec(arch=arch, x=a, y=b, loss="NLL", tau_out=v, prior_sigma_scalar=1.0)
    return x, y, xv, yv, mk


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("name", VI_TRAIN_CASES)
def test_cuda_trainer_matches_reference_training_run(name, use_graph):
    from vihmc import vi

    c, kw = _case(name)
    x, y, xv, yv, mk = _specs()
    eps = torch.from_numpy(c["eps"])
    res = vi.train_bbb(mk(x, y, kw["noise_var"]), mk(xv, yv, kw["noise_var"]),
                       priors=dict(prior_mu=kw["prior_mu"], prior_sigma=kw["prior_sigma"]), lr_start=kw["lr_start"],
                       lr_patience=kw["lr_patience"], epochs=eps.shape[0], num_ens=eps.shape[1], beta=kw["beta"],
                       mu0=torch.from_numpy(c["mu0"]), rho0=torch.from_numpy(c["rho0"]), inject_eps=eps, use_graph=use_graph)
    hist = res.history.numpy()
    np.testing.assert_allclose(hist[:, 2], c["history"][:, 2], rtol=1e-6)          # the learning-rate schedule, decision by decision
    # losses are ~1e5..1e6 sums in fp32; Adam divides by sqrt(v), so coordinates whose gradient passes near zero amplify rounding
    np.testing.assert_allclose(hist[:, :2], c["history"][:, :2], rtol=2e-4)
    tol = 2e-4 if name == "adam_6x3" else 2e-2
    np.testing.assert_allclose(res.mu.numpy(), c["mu"], rtol=1e-3, atol=tol)
    np.testing.assert_allclose(res.rho.numpy(), c["rho"], rtol=1e-3, atol=tol)
    # the best-validation snapshot is the state after the epoch with the lowest validation loss
    best_ep = int(np.flatnonzero(hist[:, 1] <= np.minimum.accumulate(hist[:, 1]))[-1])
    if best_ep == hist.shape[0] - 1:
        assert torch.equal(res.best_mu, res.mu)
    else:
        assert not torch.equal(res.best_mu, res.mu)


@pytest.mark.gpu
def test_cuda_trainer_philox_run_learns_and_writes_reference_artifacts(tmp_path):
    """Philox eps, CUDA-graph replay, 400 epochs on the bundled data: the loss falls, runs are reproducible by seed, and the
    artefacts load back exactly as main_VI_HMC.py:76-77 reads them."""
    from vihmc import artifacts, vi

    x, y, xv, yv, mk = _specs()
    kw = dict(priors=None, lr_start=1e-2, lr_patience=5000, epochs=400, num_ens=10, beta=1.0, seed=3)
    a = vi.train_bbb(mk(x, y, 0.0025), mk(xv, yv, 0.0025), **kw)
    b = vi.train_bbb(mk(x, y, 0.0025), mk(xv, yv, 0.0025), **kw)
    assert torch.equal(a.mu, b.mu) and torch.equal(a.history, b.history)
    h = a.history.numpy()
    assert np.isfinite(h).all() and h[-50:, 0].mean() < 0.5 * h[:50, 0].mean()
    assert a.steps == 400 and (a.sigma > 0).all()
    vi.save_artifacts(str(tmp_path), "t", a)
    mu, sg = torch.load(tmp_path / "means_flattened_t"), torch.load(tmp_path / "stds_flattened_t")
    assert torch.equal(mu, a.best_mu) and torch.allclose(sg, a.best_sigma)
    assert mu.dtype == torch.float32 and mu.shape == (141,)
    c = vi.train_bbb(mk(x, y, 0.0025), mk(xv, yv, 0.0025), **{**kw, "seed": 4})
    assert not torch.equal(a.mu, c.mu)


@pytest.mark.gpu
def test_cuda_trainer_deeponet_minibatches_vs_restated_loop():
    """The DeepONet form of the loop (Operator_network/VI/main_VI_deeponet.py:56-118): two mini-batches of functions on the shared
    trunk grid, loss = NLL(mean) * train_size + beta KL, one Adam step per batch -- against the torch restatement on the same eps."""
    from vihmc import vi

    inp = cases.don_inputs("small")
    arch = inp["arch"]
    x1, x2, y = inp["x1"], inp["x2"], inp["y"]
    n, P, D = x1.shape[0], x2.shape[0], arch.num_params
    half = n // 2
    noise_var, epochs, E = 1.0, 4, 2
    mk = lambda a, b: cases.LogProbSpec(arch=arch, x=a, x2=x2, y=b, loss="NLL", tau_out=noise_var, prior_sigma_scalar=1.0)
    tspecs = [mk(x1[:half], y[:half]), mk(x1[half:], y[half:])]
    vspec = mk(x1, y)
    train_size = valid_size = float(n * P)
    g = torch.Generator().manual_seed(0)
    mu0 = inp["theta"] + 0.01 * torch.randn(D, generator=g)
    rho0 = -5.0 + 0.1 * torch.randn(D, generator=g)
    eps = torch.randn(epochs * 2, E, D, generator=g)
    pri = dict(prior_mu=0.0, prior_sigma=0.1)
    res = vi.train_bbb(tspecs, vspec, priors=pri, lr_start=1e-3, lr_patience=500, epochs=epochs, num_ens=E, beta=1.0,
                       nll_scale=[train_size / (half * P), train_size / ((n - half) * P)], valid_nll_scale=valid_size / (n * P),
                       mu0=mu0, rho0=rho0, inject_eps=eps)
    kw = cases._don_kwargs(arch, torch.float32)
    slots = cases.oc.deeponet_layout(kw["width_branch"], kw["width_trunk"], kw["in_branch"], kw["in_trunk"], kw["depth_branch"],
                                     kw["depth_trunk"], kw["output_neurons"])

    def forward(w, xb):
        return cases.oc.deeponet_forward(xb.unsqueeze(1), x2.unsqueeze(0), cases.oc.unflatten(slots, w), kw["depth_branch"],
                                         kw["depth_trunk"], kw["act"], kw["impose_bc"]).squeeze(1)
    mu, rho, hist = ovi.train_batches([(x1[:half], y[:half]), (x1[half:], y[half:])], [(x1, y)], forward, mu0, rho0, eps, noise_var,
                                      0.0, 0.1, 1e-3, 500, train_size, valid_size)
    np.testing.assert_allclose(res.history.numpy()[:, :2], hist[:, :2], rtol=2e-4)
    np.testing.assert_allclose(res.mu.numpy(), mu.numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(res.rho.numpy(), rho.numpy(), rtol=1e-3, atol=2e-4)


# ------------------------------------------------------------------------------------------------
# DeepONet form pinned by the reference's own Bayesian_DeepONet training loop (tests/golden/deeponet_vi_training.npz)
# ------------------------------------------------------------------------------------------------
def _don_vi_case():
    g = cases.load_golden("deeponet_vi_training.npz")
    inp = cases.don_inputs("small")
    noise_var, prior_mu, prior_sigma, beta, lr, patience = g["cfg"]
    return g, inp, dict(noise_var=float(noise_var), prior_mu=float(prior_mu), prior_sigma=float(prior_sigma), beta=float(beta),
                        lr=float(lr), patience=int(patience))


def _don_forward(inp):
    arch = inp["arch"]
    kw = cases._don_kwargs(arch, torch.float32)
    slots = cases.oc.deeponet_layout(kw["width_branch"], kw["width_trunk"], kw["in_branch"], kw["in_trunk"], kw["depth_branch"],
                                     kw["depth_trunk"], kw["output_neurons"])
    x2 = inp["x2"]

    def forward(w, xb):
        return cases.oc.deeponet_forward(xb.unsqueeze(1), x2.unsqueeze(0), cases.oc.unflatten(slots, w), kw["depth_branch"],
                                         kw["depth_trunk"], kw["act"], kw["impose_bc"]).squeeze(1)
    return forward


def test_deeponet_oracle_matches_reference_training_run():
    g, inp, kw = _don_vi_case()
    x1, y = inp["x1"], inp["y"]
    n, P = x1.shape[0], inp["x2"].shape[0]
    half, size = n // 2, float(n * P)
    mu, rho, hist = ovi.train_batches([(x1[:half], y[:half]), (x1[half:], y[half:])], [(x1, y)], _don_forward(inp), torch.from_numpy(g["mu0"]),
                                      torch.from_numpy(g["rho0"]), torch.from_numpy(g["eps"]), kw["noise_var"], kw["prior_mu"],
                                      kw["prior_sigma"], kw["lr"], kw["patience"], size, size, beta=kw["beta"])
    np.testing.assert_allclose(hist[:, :2], g["history"][:, :2], rtol=1e-5)
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rho.numpy(), g["rho"], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_cuda_trainer_deeponet_matches_reference_training_run(use_graph):
    from vihmc import vi

    g, inp, kw = _don_vi_case()
    arch, x1, x2, y = inp["arch"], inp["x1"], inp["x2"], inp["y"]
    n, P = x1.shape[0], x2.shape[0]
    half, size = n // 2, float(n * P)
    mk = lambda a, b: cases.LogProbSpec(arch=arch, x=a, x2=x2, y=b, loss="NLL", tau_out=kw["noise_var"], prior_sigma_scalar=1.0)
    eps = torch.from_numpy(g["eps"])
    res = vi.train_bbb([mk(x1[:half], y[:half]), mk(x1[half:], y[half:])], mk(x1, y),
                       priors=dict(prior_mu=kw["prior_mu"], prior_sigma=kw["prior_sigma"]), lr_start=kw["lr"], lr_patience=kw["patience"],
                       epochs=eps.shape[0] // 2, num_ens=eps.shape[1], beta=kw["beta"],
                       nll_scale=[size / (half * P), size / ((n - half) * P)], valid_nll_scale=size / (n * P),
                       mu0=torch.from_numpy(g["mu0"]), rho0=torch.from_numpy(g["rho0"]), inject_eps=eps, use_graph=use_graph)
    np.testing.assert_allclose(res.history.numpy()[:, :2], g["history"][:, :2], rtol=2e-4)
    np.testing.assert_allclose(res.history.numpy()[:, 2], g["history"][:, 2], rtol=1e-6)
    np.testing.assert_allclose(res.mu.numpy(), g["mu"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(res.rho.numpy(), g["rho"], rtol=1e-3, atol=2e-4)
