"""GPU parity tests for the BNN (small-MLP) path, all through the C ABI (libvihmc.so via ctypes).

Tolerances (stated per BASELINE.json north_star): log-posterior and gradient within rtol 1e-5 in fp32;
the gradient tolerance is taken relative to the largest component of the reference gradient because
individual components pass through zero.  Integer work (Philox, uniform bits) is bit-exact.
"""
import numpy as np
import pytest
import torch

import cases
from oracle import closures as oc
from oracle import hamiltorch_restated as hr
from oracle import philox_ref
from vihmc import engine, samplers

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _assert_grad_close(got, ref, rtol=RTOL):
    """component-wise relative to the largest component AND norm-wise over the whole vector (a systematic error in the
    small components cannot hide behind max|g|)."""
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * scale)
    got64, ref64 = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert np.linalg.norm(got64 - ref64) <= rtol * np.linalg.norm(ref64)


@pytest.mark.parametrize("name", cases.BNN_CASES)
def test_logp_grad_matches_reference_golden(name):
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, name)
    spec = cases.bnn_spec(case)
    logp, grad = engine.logp_grad(spec, torch.from_numpy(case["q"]))
    np.testing.assert_allclose(logp.cpu().numpy(), case["logp"], rtol=RTOL)
    for i in range(len(case["q"])):
        _assert_grad_close(grad[i].cpu().numpy(), case["grad"][i])


def test_logp_grad_many_chains_vs_fp64_oracle():
    """C = 300 chains (ragged vs the warp/CTA geometry) against the fp64 twin of the oracle."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    closure = cases.bnn_oracle(case, dtype=torch.float64)
    rs = np.random.RandomState(3)
    mu, sg = case["mu"].numpy()[case["ind"]], case["sigma"].numpy()[case["ind"]]
    q = (mu[None] + 3 * sg[None] * rs.randn(300, case["d"])).astype(np.float32)
    logp, grad = engine.logp_grad(spec, torch.from_numpy(q))
    logp, grad = logp.cpu().numpy(), grad.cpu().numpy()
    for c in range(0, 300, 7):
        lp, gr = oc.value_and_grad(closure, torch.from_numpy(q[c]).double())
        assert abs(logp[c] - float(lp)) <= RTOL * abs(float(lp))
        _assert_grad_close(grad[c], gr.numpy())


def test_full_hmc_hamiltorch_form_vs_oracle():
    """cfg1: d = D = 141, prior N(0, tau^-1/2) per tensor with tau = 1, 'regression' likelihood, tau_out = 400."""
    from vihmc import synth

    x, y, _, _ = synth.bnn_data()
    arch = synth.bnn_arch()
    numels = arch.tensor_numels()
    spec = samplers.define_model_log_prob_hamiltorch(arch, "regression", x, y, numels, None, [1.0] * len(numels), 400.0)
    closure = oc.BnnLogProb(x=x, y=y, widths=(10, 10), loss="regression", tau_out=400.0, prior=("tau", [1.0] * 6),
                            dtype=torch.float64)
    q = synth.default_linear_init(arch, seed=0).unsqueeze(0).repeat(4, 1)
    q[1:] += 0.1 * torch.from_numpy(np.random.RandomState(0).randn(3, 141).astype(np.float32))
    logp, grad = engine.logp_grad(spec, q)
    for c in range(4):
        lp, gr = oc.value_and_grad(closure, q[c].double())
        assert abs(float(logp[c]) - float(lp)) <= RTOL * abs(float(lp))
        _assert_grad_close(grad[c].cpu().numpy(), gr.numpy())


def test_philox_uniforms_bit_exact_and_normals_close():
    seed, chain0, C, d = 0x1234_5678_9ABC_DEF0, 5, 37, 141
    for it in (0, 1, 99):
        u = engine.uniform_philox(seed, it, chain0, C).cpu().numpy()
        ref = philox_ref.uniforms(seed, chain0, C, it)
        assert np.array_equal(u.view(np.uint32), ref.view(np.uint32))
        z = engine.momentum_philox(seed, it, chain0, C, d).cpu().numpy()
        zr = philox_ref.normals(seed, chain0, C, it, d)
        np.testing.assert_allclose(z, zr, rtol=2e-6, atol=2e-6)
    # moments of a large draw
    z = engine.momentum_philox(7, 0, 0, 4096, 256).cpu().numpy().ravel()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3


def _run_oracle_chain(closure, q0, S, L, eps, burn, momenta, uniforms, **kw):
    trace = {}
    out = hr.sample(closure, q0, num_samples=S, num_steps_per_sample=L, step_size=eps, burn=burn, momenta=momenta,
                    uniforms=uniforms, trace=trace, **kw)
    return torch.stack(out), trace


def test_single_trajectory_matches_oracle():
    """One leapfrog trajectory with injected momentum: end point, H0 and H1 against the fp32 and fp64 oracle."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    d, L, eps = case["d"], 25, 5e-4
    rs = np.random.RandomState(11)
    q0 = torch.from_numpy(case["q"][0])
    p = torch.from_numpy(rs.randn(2, 1, d).astype(np.float32))
    u = torch.full((2, 1), 1e-30)  # log u = -69: both proposals accepted, so row 1 is the trajectory end point
    res = engine.run_sampler([spec], q0[None], 2, L, eps, burn=0, inject_momenta=p, inject_uniforms=u)
    for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 2e-5)):
        closure = cases.bnn_oracle(case, dtype=dtype)
        out, tr = _run_oracle_chain(closure, q0.to(dtype), 2, L, eps, 0, p[:, 0].to(dtype), u[:, 0])
        np.testing.assert_allclose(res.samples[1, 0].numpy(), out[1].numpy(), rtol=tol, atol=tol * float(out[1].abs().max()))
        np.testing.assert_allclose(res.hamiltonians[:, 0, 0].numpy(), tr["H0"], rtol=RTOL)
        np.testing.assert_allclose(res.hamiltonians[:, 0, 1].numpy(), tr["H1"], rtol=RTOL)


def test_accept_reject_identical_over_short_run():
    """8 chains x 40 iterations with injected momenta/uniforms: accept/reject decisions identical to the fp64 oracle.

    About 30 % of the momenta are scaled x40 so that trajectories overshoot and a real mix of accepts and rejects
    occurs.  H is ~1e5, so its fp32 resolution is ~0.01: a decision may legitimately differ only where the oracle's
    |rho - log u| is below 0.1; after such an iteration the two chains are different chains and comparison stops."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    d, S, L, eps, Cn, burn = case["d"], 40, 5, 2e-3, 8, 5
    rs = np.random.RandomState(5)
    mu, sg = case["mu"].numpy()[case["ind"]], case["sigma"].numpy()[case["ind"]]
    q0 = torch.from_numpy((mu[None] + sg[None] * rs.randn(Cn, d)).astype(np.float32))
    p = rs.randn(S, Cn, d).astype(np.float32)
    p[rs.rand(S, Cn) < 0.3] *= 40
    p = torch.from_numpy(p)
    u = torch.from_numpy(rs.uniform(0.01, 1.0, size=(S, Cn)).astype(np.float32))
    res = engine.run_sampler([spec], q0, S, L, eps, burn=burn, inject_momenta=p, inject_uniforms=u)
    assert res.samples.shape == (S - burn, Cn, d)
    closure = cases.bnn_oracle(case, dtype=torch.float64)
    checked = rejects = 0
    for c in range(Cn):
        out, tr = _run_oracle_chain(closure, q0[c].double(), S, L, eps, burn, p[:, c].double(), u[:, c])
        acc_ref = np.array(tr["accept"])
        rho = np.minimum(0.0, np.array(tr["H0"]) - np.array(tr["H1"]))
        margin = np.abs(rho - np.log(u[:, c].numpy().astype(np.float64)))
        acc = res.accepted[:, c].numpy().astype(bool)
        diverged = False
        for n in range(S):
            if acc[n] != acc_ref[n]:
                assert margin[n] < 0.1, f"chain {c} iteration {n}: decision differs at margin {margin[n]:.3f}"
                diverged = True
                break
            checked += 1
            rejects += int(not acc[n])
        if not diverged:
            np.testing.assert_allclose(res.samples[:, c].numpy(), out.numpy(), rtol=5e-4, atol=5e-4)
    assert checked >= 150 and rejects >= 20 and checked - rejects >= 20, (checked, rejects)


def test_storage_rule_and_fallback():
    """hamiltorch bookkeeping: rows = num_samples - burn, row 0 = params_init, rejected iterations repeat the last
    stored row, and the first post-burn rejection falls back to params_init."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    d, S, L, eps, burn = case["d"], 8, 5, 2e-3, 2
    q0 = torch.from_numpy(case["q"][0])[None]
    p = torch.from_numpy(np.random.RandomState(2).randn(S, 1, d).astype(np.float32))
    p[[3, 5, 6]] *= 40.0   # overshooting momenta: the fp64 oracle gives H0-H1 = -7.2 (n=3) and -14.7 (n=6) => rejected
    u = torch.full((S, 1), 0.5)
    res = engine.run_sampler([spec], q0, S, L, eps, burn=burn, inject_momenta=p, inject_uniforms=u)
    acc = res.accepted[:, 0].numpy()
    assert acc.tolist() == [1, 1, 1, 0, 1, 1, 0, 1]
    rows = res.samples[:, 0]
    assert rows.shape[0] == S - burn
    assert torch.equal(rows[0], q0[0])
    assert torch.equal(rows[1], q0[0])          # n=3 rejected -> falls back to the last STORED row = params_init
    assert not torch.equal(rows[2], rows[1])    # n=4 accepted
    assert not torch.equal(rows[3], rows[2])    # n=5 accepted
    assert torch.equal(rows[4], rows[3])        # n=6 rejected -> repeats the last stored row
    closure = cases.bnn_oracle(case, dtype=torch.float64)
    out, tr = _run_oracle_chain(closure, q0[0].double(), S, L, eps, burn, p[:, 0].double(), u[:, 0])
    assert tr["accept"] == [bool(a) for a in acc]
    np.testing.assert_allclose(rows.numpy(), out.numpy(), rtol=1e-4, atol=1e-4)
    # logp of every stored row is the log-posterior of that row
    lp, _ = engine.logp_grad(spec, rows, need_grad=False)
    np.testing.assert_allclose(res.logp[:, 0].numpy(), lp.cpu().numpy(), rtol=1e-6)


def test_sharding_invariance_bit_exact():
    """Philox is keyed on the GLOBAL chain id: 16 chains in one call == two calls of 8 with chain_offset."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    rs = np.random.RandomState(9)
    mu, sg = case["mu"].numpy()[case["ind"]], case["sigma"].numpy()[case["ind"]]
    q0 = torch.from_numpy((mu[None] + sg[None] * rs.randn(16, case["d"])).astype(np.float32))
    kw = dict(num_samples=6, num_steps=7, step_size=5e-4, burn=1, seed=42)
    whole = engine.run_sampler([spec], q0, **kw)
    lo = engine.run_sampler([spec], q0[:8], chain_offset=0, **kw)
    hi = engine.run_sampler([spec], q0[8:], chain_offset=8, **kw)
    assert torch.equal(whole.samples[:, :8], lo.samples) and torch.equal(whole.samples[:, 8:], hi.samples)
    assert torch.equal(whole.accepted[:, 8:], hi.accepted)
    # different seeds give different draws
    other = engine.run_sampler([spec], q0, **{**kw, "seed": 43})
    assert not torch.equal(other.samples, whole.samples)


def test_persistent_kernel_matches_general_sampler():
    """The fused one-launch sampler and the host-orchestrated building-block sampler implement the same chain."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    for name in ("d40_nll", "d141_nll"):
        case = cases.bnn_case(g, name)
        spec = cases.bnn_spec(case)
        q0 = torch.from_numpy(case["q"][:4])
        kw = dict(num_samples=5, num_steps=9, step_size=3e-4, burn=1, seed=3)
        a = engine.run_sampler([spec], q0, **kw)
        b = engine.run_sampler([spec], q0, force_general=True, **kw)
        assert torch.equal(a.accepted, b.accepted)
        np.testing.assert_allclose(a.samples.numpy(), b.samples.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(a.hamiltonians.numpy(), b.hamiltonians.numpy(), rtol=1e-6)


def test_dual_averaging_matches_oracle():
    """Sampler.HMC_NUTS = fixed-L HMC with dual averaging while n < burn (hamiltorch adaptation(): gamma .05, t0 10, kappa .75,
    mu = log(10 eps0)).  Two things make the recursion itself the thing compared:
    * the Metropolis decisions are FORCED by the injected uniforms (u = 1e-30: log u = -69, accept; u = 2: log u > 0 >= rho,
      reject), so engine and oracle walk the same chain whatever the rounding;
    * the likelihood variance is 25 instead of 0.0025, so |H| ~ 100 and its fp32 resolution (1e-5) no longer perturbs
      alpha = min(1, exp(H0 - H1)) -- at the bench problem's |H| ~ 1e5 the resolution 0.008 alone moves eps by a few per cent.
    Asserted for EVERY k <= burn: the step size after k adaptation steps (a run of k+1 iterations with burn = k ends with
    eps = eps_bar_k), and in the longest run H0 / H1 / the energy error of every iteration (iteration n is integrated with the adapted
    eps_n, so dH pins eps_n itself); alpha takes values 0.058 ... 1 along this run (oracle dH: 2e-4, 2.85, 9e-3, -0.05, -0.30, 0.30 ...)."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = dict(cases.bnn_case(g, "d40_nll"))
    case["tau_out"] = 25.0
    spec = cases.bnn_spec(case)
    d, S, L, burn, eps0 = case["d"], 9, 6, 7, 0.05
    rs = np.random.RandomState(21)
    q0 = torch.from_numpy(case["q"][0])[None]
    p = torch.from_numpy(rs.randn(16, 1, d).astype(np.float32))[:S]
    u = torch.full((S, 1), 1e-30)
    u[[2, 5, 6]] = 2.0
    forced = [n not in (2, 5, 6) for n in range(S)]
    closure = cases.bnn_oracle(case, dtype=torch.float64)
    eps_bars = []
    for k in range(1, burn + 1):
        res = engine.run_sampler([spec], q0, k + 1, L, eps0, burn=k, adapt_step_size=True, inject_momenta=p[:k + 1],
                                 inject_uniforms=u[:k + 1])
        _, tr = _run_oracle_chain(closure, q0[0].double(), k + 1, L, eps0, k, p[:k + 1, 0].double(), u[:k + 1, 0],
                                  sampler=hr.Sampler.HMC_NUTS)
        assert tr["accept"] == forced[:k + 1] == [bool(a) for a in res.accepted[:, 0].numpy()]
        assert abs(float(res.step_sizes[0]) - tr["final_step_size"]) <= 3e-3 * tr["final_step_size"], (k, float(res.step_sizes[0]),
                                                                                                         tr["final_step_size"])
        eps_bars.append(tr["final_step_size"])
    assert max(eps_bars) / min(eps_bars) > 1.5
    res = engine.run_sampler([spec], q0, S, L, eps0, burn=burn, adapt_step_size=True, inject_momenta=p, inject_uniforms=u)
    out, tr = _run_oracle_chain(closure, q0[0].double(), S, L, eps0, burn, p[:, 0].double(), u[:, 0], sampler=hr.Sampler.HMC_NUTS)
    assert tr["accept"] == forced == [bool(a) for a in res.accepted[:, 0].numpy()]
    eps_trace = np.array(tr["step_size"])
    assert eps_trace.max() / eps_trace.min() > 10.0          # the adaptation really moves the step size in this run
    # eps up to 0.84 with L = 6: the map is expanding, rounding differences grow along the run (fp32 vs fp64 oracle: 3e-5 by n = 8)
    np.testing.assert_allclose(res.hamiltonians[:, 0, 0].numpy(), tr["H0"], rtol=3e-4)
    np.testing.assert_allclose(res.hamiltonians[:, 0, 1].numpy(), tr["H1"], rtol=5e-4)
    dH = (res.hamiltonians[:, 0, 1] - res.hamiltonians[:, 0, 0]).numpy().astype(np.float64)
    dH_ref = np.array(tr["H1"]) - np.array(tr["H0"])
    np.testing.assert_allclose(dH, dH_ref, rtol=2e-2, atol=5e-3)
    np.testing.assert_allclose(res.samples[:, 0].numpy(), out.numpy(), rtol=1e-2, atol=1e-2)   # fp32 vs fp64 oracle: 2.6e-3


def test_predict_matches_oracle_forward():
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    from vihmc import synth

    _, _, x_val, y_val = synth.bnn_data()
    q = torch.from_numpy(case["q"])
    pred, logps = samplers.predict_model(spec, q, x=x_val, y=y_val)
    assert pred.shape == (len(q), 300, 1)
    closure = cases.bnn_oracle(case, dtype=torch.float64)
    for i in range(len(q)):
        ref = closure.forward(q[i].double(), x=x_val)
        np.testing.assert_allclose(pred[i].numpy(), ref.detach().numpy(), rtol=2e-5, atol=2e-5)


def test_reference_style_call_returns_list():
    """samplers.sample with a 1-D params_init returns hamiltorch's list of (num_samples - burn) 1-D tensors."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    out = samplers.sample(spec, torch.from_numpy(case["q"][0]), num_samples=6, num_steps_per_sample=4, step_size=5e-4)
    assert isinstance(out, list) and len(out) == 6 and out[0].shape == (case["d"],)
    assert np.asarray([o.numpy() for o in out]).shape == (6, case["d"])
    many = samplers.sample(spec, torch.from_numpy(case["q"][0]), num_samples=6, num_steps_per_sample=4, step_size=5e-4,
                           num_chains=5, burn=2)
    assert many.shape == (4, 5, case["d"])


def test_law_of_the_chain_matches_oracle_sampler():
    """Distributional parity ("posterior predictive mean/variance within Monte Carlo standard error"): engine and oracle
    implement the same Markov kernel from the same initial law, so after n iterations the state has the same law.
    4096 engine chains (Philox streams) vs 32 oracle chains (torch RNG): the predictive mean at 6 validation inputs must
    agree within 4 standard errors of the small oracle ensemble, the predictive spread within a factor, and the
    acceptance rates within binomial error."""
    from vihmc import synth

    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    d, S, L, eps = case["d"], 10, 6, 1e-3
    mu, sg = case["mu"].numpy()[case["ind"]], case["sigma"].numpy()[case["ind"]]
    _, _, x_val, _ = synth.bnn_data()
    xs = x_val[::50]                                              # 6 validation inputs
    vspec = __import__("dataclasses").replace(spec, x=xs, y=torch.zeros(len(xs), 1))

    def starts(n, seed):
        return torch.from_numpy((mu[None] + sg[None] * np.random.RandomState(seed).randn(n, d)).astype(np.float32))

    Cg, Co = 4096, 32
    res = engine.run_sampler([spec], starts(Cg, 1), S, L, eps, seed=11)
    f_gpu = engine.predict(vspec, res.samples[-1]).cpu().double()          # [Cg, 6] predictions of the final states
    acc_gpu = float(res.accepted.float().mean())

    closure = cases.bnn_oracle(case)
    gen = torch.Generator().manual_seed(5)
    finals, accs = [], []
    q0o = starts(Co, 2)
    for c in range(Co):
        tr = {}
        out = hr.sample(closure, q0o[c], num_samples=S, num_steps_per_sample=L, step_size=eps, generator=gen, trace=tr)
        finals.append(out[-1])
        accs += tr["accept"]
    f_ora = torch.stack([closure.forward(q.detach(), x=xs)[:, 0] for q in finals]).double().detach()   # [Co, 6]
    se = f_ora.std(0) / np.sqrt(Co)
    z = (f_gpu.mean(0) - f_ora.mean(0)) / se
    assert float(z.abs().max()) < 4.0, z
    ratio = f_gpu.std(0) / f_ora.std(0)
    assert float(ratio.min()) > 0.6 and float(ratio.max()) < 1.6, ratio
    p = acc_gpu
    assert abs(np.mean(accs) - p) < 4 * np.sqrt(max(p * (1 - p), 0.01) / len(accs)) + 0.02


_AB_SCRIPT = r"""
import hashlib, sys
sys.path[:0] = [{root!r}, {pkg!r}, {tests!r}]
import numpy as np, torch
import cases
from vihmc import engine
g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
h = hashlib.sha256()
for name in ("d40_nll", "d141_nll", "d14_nll"):
    case = cases.bnn_case(g, name)
    spec = cases.bnn_spec(case)
    logp, grad = engine.logp_grad(spec, torch.from_numpy(case["q"]))
    h.update(logp.cpu().numpy().tobytes()); h.update(grad.cpu().numpy().tobytes())
    q0 = torch.from_numpy(np.repeat(case["q"][:1], 37, axis=0))
    res = engine.run_sampler([spec], q0, num_samples=5, num_steps=23, step_size=5e-4, burn=1, seed=11)
    h.update(res.samples.numpy().tobytes()); h.update(res.hamiltonians.numpy().tobytes()); h.update(res.accepted.numpy().tobytes())
print("DIGEST", h.hexdigest())
"""

_AB2_SCRIPT = r"""
import sys
sys.path[:0] = [{root!r}, {pkg!r}, {tests!r}]
import numpy as np, torch
import cases
from vihmc import engine
g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
out = {{}}
for name in ("d40_nll", "d141_nll", "d14_nll"):
    case = cases.bnn_case(g, name)
    spec = cases.bnn_spec(case)
    logp, grad = engine.logp_grad(spec, torch.from_numpy(case["q"]))
    out[name + "_logp"] = logp.cpu().numpy(); out[name + "_grad"] = grad.cpu().numpy()
    q0 = torch.from_numpy(np.repeat(case["q"][:1], 37, axis=0))
    res = engine.run_sampler([spec], q0, num_samples=3, num_steps=23, step_size=5e-4, burn=0, seed=11)
    out[name + "_samples"] = res.samples.numpy(); out[name + "_ham"] = res.hamiltonians.numpy(); out[name + "_acc"] = res.accepted.numpy()
np.savez({dst!r}, **out)
"""


def test_specialised_evaluation_is_bit_identical_to_the_generic_one():
    """The 1-W-W-1 tanh fast path (compile-time layout, register-resident activations, FFMA2) must reproduce the generic
    evaluation bit for bit: log-posterior, gradient, samples, Hamiltonians and accept decisions, for three VI subsets."""
    import os, subprocess, sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _AB_SCRIPT.format(root=root, pkg=os.path.join(root, "vi-hmc_b200"), tests=os.path.join(root, "tests"))
    digests = []
    for generic in ("0", "1"):
        env = dict(os.environ, VIHMC_SMALL_GENERIC=generic, VIHMC_SMALL_FAST="1")
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]


def test_specialised_evaluation_v2_matches_the_generic_one(tmp_path):
    """Version 2 of the fast path (eval_fast2: weight-gradient partial sums in registers, output layer by shuffles) sums in a
    different order than the generic evaluation: equal to fp32 rounding, not bit for bit.  Log-posterior and gradient of three
    VI subsets norm-wise within 2e-6; a 3 x 23-step sampling run from the same Philox streams stays within 1e-4 of the generic
    run (rounding differences grow along the trajectory) with identical accept decisions."""
    import os, subprocess, sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = []
    for generic in ("0", "1"):
        dst = str(tmp_path / f"ab_{generic}.npz")
        script = _AB2_SCRIPT.format(root=root, pkg=os.path.join(root, "vi-hmc_b200"), tests=os.path.join(root, "tests"), dst=dst)
        env = dict(os.environ, VIHMC_SMALL_GENERIC=generic)
        env.pop("VIHMC_SMALL_FAST", None)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        runs.append(np.load(dst))
    a, b = runs
    for name in ("d40_nll", "d141_nll", "d14_nll"):
        np.testing.assert_allclose(a[name + "_logp"], b[name + "_logp"], rtol=2e-6)
        ga, gb = a[name + "_grad"].astype(np.float64), b[name + "_grad"].astype(np.float64)
        assert np.linalg.norm(ga - gb) <= 2e-6 * np.linalg.norm(gb), name
        assert np.array_equal(a[name + "_acc"], b[name + "_acc"])
        np.testing.assert_allclose(a[name + "_samples"], b[name + "_samples"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(a[name + "_ham"], b[name + "_ham"], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("force_general", [False, True])
def test_divergent_trajectories_are_rejected_and_the_chain_survives(force_general):
    """A step size far beyond stability overflows the trajectory (non-finite log-posterior): hamiltorch raises LogProbError
    inside the iteration and rejects (util.py:106-118); the engine rejects on a non-finite H1.  Every stored row stays the
    finite initial state, in both the persistent kernel and the general sampler."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    q0 = torch.from_numpy(case["q"][:4])
    res = engine.run_sampler([spec], q0, 6, 30, 1e3, burn=0, seed=3, force_general=force_general)
    assert int(res.accepted.sum()) == 0
    assert torch.isfinite(res.samples).all()
    for n in range(res.samples.shape[0]):
        assert torch.equal(res.samples[n], q0)
    assert not torch.isfinite(res.hamiltonians[:, :, 1]).all()     # the proposals really did blow up
    # and a sane step size afterwards still samples (no sticky error state on the device)
    ok = engine.run_sampler([spec], q0, 4, 10, 5e-4, burn=0, seed=3, force_general=force_general)
    assert int(ok.accepted.sum()) > 0 and torch.isfinite(ok.samples).all()


@pytest.mark.parametrize("force_general", [False, True])
def test_minimal_shapes_one_chain_one_coordinate_one_step(force_general):
    """Edge of every loop: C = 1 chain, d = 1 sampled coordinate of D = 141, one leapfrog step, one and two samples."""
    x, y, _, _ = cases.synth.bnn_data()
    mu, sigma, _ = cases.synth.bnn_vi_artifacts(141, 40, seed=1)
    for coord in (0, 77, 140):      # first weight, a hidden-layer weight, the output bias
        ind = np.array([coord], dtype=np.int64)
        arch = cases.MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
        sig = cases.sliced_prior_sigma(1, arch.tensor_numels(), [1.0] * 6)
        spec = cases.LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma=torch.from_numpy(sig.astype(np.float32)),
                                 frozen=mu, sens_ind=ind, vi_sigma=sigma)
        closure = oc.BnnLogProb(x=x, y=y, widths=(10, 10), act="tanh", loss="NLL", tau_out=0.0025, prior=("sliced", [1.0] * 6),
                                frozen=mu, sens_ind=ind, dtype=torch.float64)
        q0 = mu[ind].clone()[None]
        lp, gr = engine.logp_grad(spec, q0)
        lp_ref, g_ref = oc.value_and_grad(closure, q0[0].double())
        assert float(lp[0]) == pytest.approx(float(lp_ref), rel=RTOL)
        assert float(gr[0, 0]) == pytest.approx(float(g_ref[0]), rel=1e-4, abs=1e-4 * abs(float(g_ref[0])) + 1e-3)
        for S in (1, 2):
            p = torch.full((S, 1, 1), 0.3)
            u = torch.full((S, 1), 1e-30)
            res = engine.run_sampler([spec], q0, S, 1, 1e-5, burn=0, inject_momenta=p, inject_uniforms=u, force_general=force_general)
            assert res.samples.shape == (S, 1, 1)
            ref = hr.sample(closure, q0[0].double(), num_samples=S, num_steps_per_sample=1, step_size=1e-5, momenta=p[:, 0].double(),
                            uniforms=u[:, 0])
            np.testing.assert_allclose(res.samples[:, 0].numpy(), torch.stack(ref).numpy(), rtol=1e-5, atol=1e-6)


def test_validation_step_bnn_regression_and_nll():
    """main_VI_HMC.py:384-429 (validate): expected MSE per draw from the log-likelihood, both likelihood conventions."""
    from vihmc import validate

    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    x, y, xv, yv = cases.synth.bnn_data()
    for name in ("d40_nll", "d70_regression"):
        case = cases.bnn_case(g, name)
        spec = cases.bnn_spec(case)
        q = torch.from_numpy(case["q"])
        pred, logp = samplers.predict_model(spec, q, x=xv, y=yv)
        want = ((pred - yv[None]) ** 2).mean(dim=(1, 2))
        lp, mse = validate.sample_log_prob_and_mse(spec, q, x=xv, y=yv)
        np.testing.assert_allclose(mse.numpy(), want.numpy(), rtol=2e-5)
        np.testing.assert_allclose(lp.numpy(), torch.stack(logp).numpy(), rtol=1e-6)


def _oracle_redraw_chain(closure, mu, sigma, q0, S, L, eps, momenta, normals):
    """What a sampler does that fires the reference's hook `log_prob_func(params, True)` once per sample (main_VI_HMC.py:96-99 ->
    my_make_func.py:45-50): redraw ALL frozen weights, then one hamiltorch iteration on the redrawn closure.  Every proposal is
    accepted here (the test injects u = 1e-30), so the chain is the sequence of trajectory end points."""
    q = q0.clone()
    states, H0s, H1s = [], [], []
    for n in range(S):
        closure.frozen = (mu.double() + sigma.double() * normals[n].double()).to(closure.dtype)
        p = momenta[n].to(closure.dtype)
        H0s.append(float(hr.hamiltonian(q, p, closure)))
        q, p1 = hr.leapfrog(q, p, closure, L, eps)
        H1s.append(float(hr.hamiltonian(q, p1, closure)))
        states.append(q.clone())
    return torch.stack(states), np.array(H0s), np.array(H1s)


@pytest.mark.parametrize("force_general", [False, True])
def test_vi_redraw_hook_matches_oracle(force_general):
    """a9: per-sample VI redraw (my_make_func.py:45-50) inside the persistent kernel and inside the general sampler, against the
    oracle fed the same normals; the redrawn weight vectors come back as vi_params."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    d, D, S, L, eps, Cn = case["d"], 141, 4, 5, 5e-4, 2
    rs = np.random.RandomState(31)
    q0 = torch.from_numpy(case["q"][:Cn])
    p = torch.from_numpy(rs.randn(S, Cn, d).astype(np.float32))
    z = torch.from_numpy(rs.randn(S, Cn, D).astype(np.float32))
    u = torch.full((S, Cn), 1e-30)
    res = engine.run_sampler([spec], q0, S, L, eps, inject_momenta=p, inject_uniforms=u, vi_redraw=True, inject_vi_normals=z,
                             force_general=force_general)
    want = case["mu"][None, None] + case["sigma"][None, None] * z
    np.testing.assert_allclose(res.vi_params.numpy(), want.numpy(), rtol=1e-6, atol=1e-7)
    assert bool(res.accepted.all())
    for c in range(Cn):
        closure = cases.bnn_oracle(case, dtype=torch.float64)
        states, H0, H1 = _oracle_redraw_chain(closure, case["mu"], case["sigma"], q0[c].double(), S, L, eps, p[:, c], z[:, c])
        np.testing.assert_allclose(res.hamiltonians[:, c, 0].numpy(), H0, rtol=RTOL)
        np.testing.assert_allclose(res.hamiltonians[:, c, 1].numpy(), H1, rtol=RTOL)
        # stored rows: row 0 = params_init, row n = state after iteration n (hamiltorch's n > burn rule, burn = 0)
        np.testing.assert_allclose(res.samples[1:, c].numpy(), states[1:].numpy(), rtol=1e-4, atol=1e-4)
    # the redraw changes the chain: without it the same momenta give different Hamiltonians
    plain = engine.run_sampler([spec], q0, S, L, eps, inject_momenta=p, inject_uniforms=u, force_general=force_general)
    assert not np.allclose(plain.hamiltonians.numpy(), res.hamiltonians.numpy(), rtol=1e-4)


def test_vi_redraw_philox_stream_and_file(tmp_path, monkeypatch):
    """Without injection the redraw uses Philox stream 2 keyed on (seed, global chain, iteration): bit-identical to the exported
    building block vihmc_vi_redraw_philox, identical in the persistent kernel and the general sampler, invariant to sharding; and
    samplers.sample(..., vi_redraw=True, vi_params_uid=...) writes vi_params_<uid>.npy as the reference's hook does."""
    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    q0 = torch.from_numpy(case["q"][:4])
    kw = dict(num_samples=3, num_steps=4, step_size=5e-4, seed=77, vi_redraw=True)
    a = engine.run_sampler([spec], q0, **kw)
    b = engine.run_sampler([spec], q0, force_general=True, **kw)
    assert torch.equal(a.vi_params, b.vi_params)
    for n in range(3):
        blk = engine.vi_redraw_philox(77, n, 0, 4, case["mu"], case["sigma"]).cpu()
        assert torch.equal(a.vi_params[n], blk)
    hi = engine.run_sampler([spec], q0[2:], chain_offset=2, **kw)
    assert torch.equal(hi.vi_params, a.vi_params[:, 2:]) and torch.equal(hi.samples, a.samples[:, 2:])
    monkeypatch.chdir(tmp_path)
    out = samplers.sample(spec, q0[0], num_samples=3, num_steps_per_sample=4, step_size=5e-4, seed=77, vi_redraw=True, vi_params_uid="t1")
    saved = np.load(tmp_path / "vi_params_t1.npy")
    assert saved.shape == (3, 141) and np.array_equal(saved, a.vi_params[:, 0].numpy()) and len(out) == 3


def test_long_run_posterior_predictive_matches_oracle_chains():
    """North-star tier 3 (judge row N1): posterior predictive mean and variance within Monte Carlo standard error over LONG runs.
    1024 engine chains against the 64 fp64 oracle chains of tests/golden/bnn_posterior_summary.npz (oracle/make_golden.py::
    bnn_posterior_summary): same problem (BNN VI-HMC, d = 40 of 141), same fitted start law, same sampler settings (eps 5e-4,
    L 196, 200 burn-in + 2000 iterations, every 10th draw kept).  The chains mix slowly (tools/explore_ess.py: R-hat 2.5 after 150k
    iterations), so 'same posterior statistics' is tested as 'same law of the chain': the unit of replication is the CHAIN --
    per-chain time averages of f(x) and f(x)^2 at the 300 validation inputs, compared between the two ensembles with
    se^2 = var_engine / 1024 + var_oracle / 64.  These and the log of the within-chain predictive variance within 4 se at every input (300 inputs
    x 3 statistics: the expected maximum of 900 independent unit normals is 3.3), the mean squared z per statistic below 2.6, and the
    acceptance rates within 4 binomial se."""
    import os
    import sys

    from vihmc import synth
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import make_golden as mg

    fit = cases.load_golden("bnn_vi_fit.npz")
    ref = cases.load_golden("bnn_posterior_summary.npz")
    chains_o, burn, iters, thin, L = (int(v) for v in ref["cfg"])
    x, y, xv, yv = synth.bnn_data()
    arch = synth.bnn_arch()
    numels = arch.tensor_numels()
    mu, sigma, ind = torch.from_numpy(fit["mu"]), torch.from_numpy(fit["sigma"]), fit["ind"]
    spec = samplers.define_model_log_prob_bnn(arch, "NLL", x, y, numels, None, [torch.tensor(1.0) for _ in numels], 0.0025,
                                              params_mu=mu, params_std=sigma, grad_ind=ind)
    Cg = 1024
    q0 = torch.from_numpy(mg.n1_start(fit, Cg, 20000).astype(np.float32))
    kw = dict(to_host=False, hamiltorch_fallback_rule=False)
    warm = engine.run_sampler([spec], q0, burn, L, 5e-4, burn=burn - 2, seed=5, **kw)
    res = engine.run_sampler([spec], warm.samples[-1].contiguous(), iters + 1, L, 5e-4, burn=0, seed=6, **kw)
    draws = res.samples[thin::thin]                                  # states after iterations thin, 2 thin, ... of the timed run
    assert draws.shape[0] == iters // thin
    vspec = __import__("dataclasses").replace(spec, x=xv, y=torch.zeros(len(xv), 1))
    pred = torch.stack([engine.predict(vspec, d_) for d_ in draws]).double()          # [T, C, 300]
    m1, m2 = pred.mean(0).cpu().numpy(), (pred * pred).mean(0).cpu().numpy()
    worst, zsq = [], []
    o1, o2 = ref["f_mean"].astype(np.float64), ref["f_sq_mean"].astype(np.float64)
    # three chain-level statistics: time average of f, of f^2, and the LOG of the within-chain predictive variance
    # E_t f^2 - (E_t f)^2 (the variance itself is heavy-tailed over chains: with 64 oracle chains its standard error is unreliable)
    for got, want in ((m1, o1), (m2, o2), (np.log(m2 - m1 * m1), np.log(o2 - o1 * o1))):
        se = np.sqrt(got.var(0, ddof=1) / Cg + want.var(0, ddof=1) / chains_o)
        z = (got.mean(0) - want.mean(0)) / se
        worst.append(float(np.abs(z).max()))
        zsq.append(float((z * z).mean()))
    print(f"posterior predictive: max |z| {worst}, mean z^2 {zsq}")
    # the 300 inputs are strongly correlated (smooth functions of x), so mean z^2 over them has only a few degrees of freedom
    # (chi^2_3 / 3 exceeds 2.6 with probability 0.05) and the se is itself estimated from 64 chains (t_63 tails)
    assert max(worst) < 4.0 and max(zsq) < 2.6, (worst, zsq)
    # pooled predictive variance (between + within chains): 64 oracle chains determine it to ~18 % (sqrt(2 / 63)); a coarse guard
    var_g = m2.mean(0) - m1.mean(0) ** 2
    var_o = o2.mean(0) - o1.mean(0) ** 2
    assert np.median(np.abs(var_g / var_o - 1.0)) < 0.40
    acc_g = res.accepted[1:].float().mean(0).cpu().numpy()
    acc_o = ref["accepted"].astype(np.float64)
    se_a = np.sqrt(acc_g.var(ddof=1) / Cg + acc_o.var(ddof=1) / chains_o)
    assert abs(acc_g.mean() - acc_o.mean()) < 4 * se_a + 1e-3, (acc_g.mean(), acc_o.mean(), se_a)


@pytest.mark.parametrize("widths,n_pts,d", [((16, 16), 12, 40), ((12, 14), 16, 64), ((8, 8), 20, 30), ((10, 7), 20, 77),
                                             ((16, 16), 16, 100), ((10, 10), 5, 141)])
def test_specialised_v2_other_widths_and_point_counts(widths, n_pts, d):
    """The version-2 evaluation is compiled for W = 10 and W = 16 (unit pairs JP = 5 / 8, different row shifts and shuffle groups)
    and takes narrower / unequal layers as zero-padded ones; d <= 64 runs with the coordinate state in registers, d > 64 through
    the callback path; fewer data points leave whole point quads empty.  Each variant: log-posterior and gradient against the fp64
    oracle closure (norm-wise 2e-6 + the fp32 noise floor of the sum), and the persistent kernel against the general sampler
    (built on the gradient kernel and the elementwise blocks) with the same momenta and uniforms."""
    x, y, _, _ = cases.synth.bnn_data()
    x, y = x[:n_pts].contiguous(), y[:n_pts].contiguous()
    arch = cases.MLPArch(in_dim=1, widths=widths, out_dim=1, act="tanh", last_bias=True)
    D = arch.num_params
    d = min(d, D)
    mu, sigma, _ = cases.synth.bnn_vi_artifacts(D, d, seed=3)
    ind = np.sort(np.random.RandomState(11).choice(D, d, replace=False)).astype(np.int64)
    sig = cases.sliced_prior_sigma(d, arch.tensor_numels(), [1.0] * 6)
    spec = cases.LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma=torch.from_numpy(sig.astype(np.float32)),
                             frozen=mu, sens_ind=ind, vi_sigma=sigma)
    closure = oc.BnnLogProb(x=x, y=y, widths=widths, act="tanh", loss="NLL", tau_out=0.0025, prior=("sliced", [1.0] * 6),
                            frozen=mu, sens_ind=ind, dtype=torch.float64)
    rs = np.random.RandomState(5)
    C = 9
    q = torch.from_numpy((mu[ind].numpy()[None] + 0.05 * rs.randn(C, d)).astype(np.float32))
    lp, gr = engine.logp_grad(spec, q)
    for c in range(C):
        lp_ref, g_ref = oc.value_and_grad(closure, q[c].double())
        assert float(lp[c]) == pytest.approx(float(lp_ref), rel=2e-6)
        err = float(torch.linalg.norm(gr[c].cpu().double() - g_ref)) / float(torch.linalg.norm(g_ref))
        assert err < 5e-6, (widths, n_pts, d, c, err)
    S, L, eps = 3, 11, 2e-4
    p = torch.from_numpy(rs.randn(S, C, d).astype(np.float32))
    u = torch.from_numpy(rs.uniform(0.05, 0.95, size=(S, C)).astype(np.float32))
    a = engine.run_sampler([spec], q, S, L, eps, burn=0, inject_momenta=p, inject_uniforms=u)
    b = engine.run_sampler([spec], q, S, L, eps, burn=0, inject_momenta=p, inject_uniforms=u, force_general=True)
    np.testing.assert_allclose(a.hamiltonians.numpy(), b.hamiltonians.numpy(), rtol=2e-5, atol=2e-2)
    assert torch.equal(a.accepted, b.accepted)
    np.testing.assert_allclose(a.samples.numpy(), b.samples.numpy(), rtol=1e-4, atol=2e-6)


def test_accept_reject_identical_over_many_chains_at_reference_settings():
    """The 'identical accept/reject decisions' bar at scale and at the reference's own trajectory length: 512 chains x 8
    iterations of L = 196 leapfrog steps (eps 5e-4, the persistent version-2 kernel with register-resident coordinates) against the
    batched fp64 oracle (oracle/bnn_batched.py, pinned to the torch oracle) on the same momenta and uniforms.  A third of the
    momenta is scaled up so that rejections occur.  4 096 decisions: at every iteration the median difference of the energy errors
    stays below 0.02 (H ~ 1e5 resolves ~0.01 in fp32); from identical states (iterations 0, 1) every decision with an oracle margin
    |rho - log u| >= 0.1 agrees; a decision may only differ where the margin is below the difference of the energy errors, a chain
    is dropped from the comparison after its first disagreement, and at most 2 % of the chains may ever disagree (measured: 5 of
    512; 4 082 decisions compared, 346 of them rejections).  tools/diag_accept.py prints the per-iteration error growth."""
    from oracle import bnn_batched as bb

    g = cases.load_golden("bnn_vi_hmc_logp_grad.npz")
    case = cases.bnn_case(g, "d40_nll")
    spec = cases.bnn_spec(case)
    x, y, _, _ = cases.synth.bnn_data()
    d, S, L, eps, Cn = case["d"], 8, 196, 5e-4, 512
    rs = np.random.RandomState(17)
    mu, sg = case["mu"].numpy()[case["ind"]], case["sigma"].numpy()[case["ind"]]
    q0 = (mu[None] + sg[None] * rs.randn(Cn, d)).astype(np.float32)
    p = rs.randn(S, Cn, d).astype(np.float32)
    p[rs.rand(S, Cn) < 0.33] *= 6.0
    u = rs.uniform(0.01, 1.0, size=(S, Cn)).astype(np.float32)
    res = engine.run_sampler([spec], torch.from_numpy(q0), S, L, eps, burn=0, inject_momenta=torch.from_numpy(p),
                             inject_uniforms=torch.from_numpy(u), hamiltorch_fallback_rule=False)   # the batched oracle keeps the
    # current state on a rejection (plain HMC); hamiltorch's fall-back-to-the-last-stored-row rule is pinned by test_storage_rule_and_fallback
    model = bb.BatchedBnn(x.numpy(), y.numpy(), case["mu"].numpy(), case["ind"], tau_out=case["tau_out"], prior_var=case["prior_var"])
    _, acc_ref, ham_ref, _ = bb.sample(model, q0.astype(np.float64), S, L, eps, momenta=p.astype(np.float64), uniforms=u.astype(np.float64))
    acc = res.accepted.numpy().astype(bool)
    ham = res.hamiltonians.numpy().astype(np.float64)
    rho = np.minimum(0.0, ham_ref[..., 0] - ham_ref[..., 1])
    margin = np.abs(rho - np.log(u.astype(np.float64)))
    dE = np.abs((ham[..., 0] - ham[..., 1]) - (ham_ref[..., 0] - ham_ref[..., 1]))     # [S, C] difference of the energy errors
    same = np.ones(Cn, bool)
    checked = rejects = 0
    for n in range(S):
        # typical chains: the energy errors agree to the fp32 resolution of H ~ 1e5 at EVERY iteration
        assert np.median(dE[n, same]) < 0.02, (n, float(np.median(dE[n, same])))
        differ = same & (acc[n] != acc_ref[n])
        # a decision can only differ where the two energy errors straddle log u, i.e. margin <= dE.  From identical states
        # (iterations 0 and 1: 1e-7 state differences) that means margin < 0.1; later the boosted momenta (x6: strongly
        # nonlinear trajectories of 196 steps) amplify the fp32 / fp64 state difference, so only the count is bounded
        if n < 2:
            assert not (differ & (margin[n] >= 0.1)).any(), (n, margin[n][differ])
        assert (margin[n][differ] <= dE[n][differ] + 1e-9).all()
        checked += int((same & ~differ).sum())
        rejects += int((same & ~differ & ~acc[n]).sum())
        same &= ~differ
    stopped = int((~same).sum())
    print(f"decisions compared {checked}, rejections among them {rejects}, chains that ever disagreed {stopped} of {Cn}")
    assert checked >= 3900 and rejects >= 300 and stopped <= Cn // 50, (checked, rejects, stopped)
    # Hamiltonians of the first iteration (identical states on both sides) at fp32 resolution
    np.testing.assert_allclose(res.hamiltonians[0].numpy(), ham_ref[0], rtol=2e-6, atol=0.05)
