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
