"""Sensitivity scores of the VI -> HMC split selector: oracle vs the golden vectors produced by the
reference's own eval_std_dydw (Neural_network/VI/sensitivity.py:71-126), and the CUDA kernel vs the same vectors."""
import numpy as np
import pytest
import torch

import cases
from oracle import sensitivity as osens

SENS_CASES = [("tanh_10x10", [10, 10], "tanh"), ("relu_10x10", [10, 10], "relu"), ("sine_16x16", [16, 16], "sine"), ("tanh_32", [32], "tanh")]
RTOL = 1e-5


def _case(name):
    g = cases.load_golden("bnn_sensitivity.npz")
    return tuple(torch.from_numpy(g[f"{name}/{k}"]) for k in ("x", "mu", "sigma")) + (g[f"{name}/scores"],)


@pytest.mark.parametrize("name,widths,act", SENS_CASES)
def test_oracle_matches_reference_golden(name, widths, act):
    x, mu, sigma, ref = _case(name)
    got = osens.scores(x, widths, act, mu, sigma)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


def test_selection_rule_matches_reference_run():
    """captured_var / select_indices (sensitivity.py:129-166, :226-231) on a hand-checkable score vector."""
    from vihmc import sensitivity as vs

    imp = np.array([0.05, 0.4, 0.1, 0.3, 0.15])
    assert vs.captured_var(imp, 0.90) == 3            # cumulative shares 0.4, 0.7, 0.85, 0.95, 1.0
    np.testing.assert_array_equal(vs.select_indices(imp, 0.90), np.array([1, 3, 4]))


@pytest.mark.gpu
@pytest.mark.parametrize("name,widths,act", SENS_CASES)
def test_kernel_matches_reference_golden(name, widths, act):
    from vihmc import sensitivity as vs
    from vihmc.spec import MLPArch

    x, mu, sigma, ref = _case(name)
    got = vs.eval_std_dydw((x, None), MLPArch(in_dim=1, widths=tuple(widths), act=act), mu, sigma)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
    # the reference's nn.Sequential is accepted as well
    act_mod = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}.get(act)
    if act_mod is not None:
        mods, prev = [], 1
        for w in widths:
            mods += [torch.nn.Linear(prev, w), act_mod()]
            prev = w
        mods.append(torch.nn.Linear(prev, 1))
        got2 = vs.eval_std_dydw((x, None), torch.nn.Sequential(*mods), mu, sigma)
        np.testing.assert_array_equal(got, got2)
    # and the selected subset is what the reference's run() would write to gradient_indices_<uid>.npy
    np.testing.assert_array_equal(vs.select_indices(got, 0.9), vs.select_indices(ref, 0.9))


# ------------------------------------------------------------------------------------------------
# DeepONet scores (Operator_network/VI/sensitivity.py:61-126)
# ------------------------------------------------------------------------------------------------
DON_SENS_CASES = [("tanh_small", 16, 12, 3, 4, 8, "tanh"), ("relu_small", 24, 10, 3, 3, 16, "relu"),
                  ("tanh_loader", 16, 12, 3, 4, 8, "tanh"), ("tanh_shipped", 100, 101, 9, 9, 100, "tanh")]
# the reference's scores are fp32 jacrev products; the checkers differ from them by fp32 rounding of sums of squares
RTOL_DON = 1e-4


def _don_case(name):
    g = cases.load_golden("deeponet_sensitivity.npz")
    batches, bi = [], 0
    while f"{name}/xb{bi}" in g.files:
        batches.append((torch.from_numpy(g[f"{name}/xb{bi}"]), torch.from_numpy(g[f"{name}/xt{bi}"])))
        bi += 1
    return batches, torch.from_numpy(g[f"{name}/mu"]), torch.from_numpy(g[f"{name}/sigma"]), g[f"{name}/scores"]


def _don_arch(width, in_branch, db, dt, K, act):
    from vihmc.spec import DeepONetArch

    return DeepONetArch(width_branch=width, width_trunk=width, in_branch=in_branch, in_trunk=5, depth_branch=db, depth_trunk=dt,
                        output_neurons=K, act=act, impose_bc=True)


@pytest.mark.parametrize("name,width,in_branch,db,dt,K,act", DON_SENS_CASES)
def test_deeponet_oracle_matches_reference_golden(name, width, in_branch, db, dt, K, act):
    batches, mu, sigma, ref = _don_case(name)
    kw = dict(width_branch=width, width_trunk=width, in_branch=in_branch, in_trunk=5, depth_branch=db, depth_trunk=dt,
              output_neurons=K, act=act, impose_bc=True)
    got = sum(osens.deeponet_scores(xb, xt, kw, mu, sigma) for xb, xt in batches) / len(batches)
    np.testing.assert_allclose(got, ref, rtol=RTOL_DON, atol=RTOL_DON * np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name,width,in_branch,db,dt,K,act", DON_SENS_CASES)
def test_deeponet_kernel_matches_reference_golden(name, width, in_branch, db, dt, K, act):
    from vihmc import sensitivity as vs

    batches, mu, sigma, ref = _don_case(name)
    arch = _don_arch(width, in_branch, db, dt, K, act)
    data = batches[0] if len(batches) == 1 else batches
    got = vs.eval_std_dydw(data, arch, mu, sigma)
    np.testing.assert_allclose(got, ref, rtol=RTOL_DON, atol=RTOL_DON * np.abs(ref).max())
    assert got[0] == pytest.approx(float(sigma[0]) ** 2)          # d out / d b = 1
    np.testing.assert_array_equal(vs.select_indices(got, 0.9), vs.select_indices(ref, 0.9))


@pytest.mark.gpu
def test_deeponet_kernel_at_the_shipped_size_against_the_fp64_oracle_on_a_subset():
    """N = 64 functions x the full 101 x 101 trunk grid of the shipped architecture (a 64 x 10201 x 172 401 Jacobian, 450 GB in fp32,
    never formed): the sum over trunk points is linear in the per-point contributions, so the kernel's scores on the full grid equal
    the average of its scores on a partition of the grid; one part is small enough for the fp64 oracle."""
    from vihmc import sensitivity as vs, synth

    arch = _don_arch(100, 101, 9, 9, 100, "tanh")
    x1, x2, y, theta = synth.burgers_like(arch, n_train=64, n_t=101, n_x=101, seed=0)
    g = torch.Generator().manual_seed(5)
    sigma = 0.001 + 0.01 * torch.rand(arch.num_params, generator=g)
    full = vs.eval_std_dydw((x1.unsqueeze(1), x2.unsqueeze(0)), arch, theta, sigma)
    assert np.isfinite(full).all() and (full >= 0).all()
    parts = [x2[i::3] for i in range(3)]
    sizes = np.array([p.shape[0] for p in parts], dtype=np.float64)
    avg = sum(vs.eval_std_dydw((x1.unsqueeze(1), p.unsqueeze(0)), arch, theta, sigma).astype(np.float64) * n for p, n in zip(parts, sizes)) / sizes.sum()
    np.testing.assert_allclose(full, avg, rtol=2e-4, atol=2e-4 * np.abs(full).max())
    # fp64 oracle on 2 functions x 6 trunk points
    kw = dict(width_branch=100, width_trunk=100, in_branch=101, in_trunk=5, depth_branch=9, depth_trunk=9, output_neurons=100,
              act="tanh", impose_bc=True)
    xb, xt = x1[:2].unsqueeze(1), x2[::1700].unsqueeze(0)
    ref = osens.deeponet_scores(xb, xt, kw, theta, sigma)
    got = vs.eval_std_dydw((xb, xt), arch, theta, sigma)
    np.testing.assert_allclose(got, ref, rtol=RTOL_DON, atol=RTOL_DON * np.abs(ref).max())
