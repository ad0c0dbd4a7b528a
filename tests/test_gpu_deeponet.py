"""GPU parity tests for the dense path (DeepONet and wide MLP) and the large-d building blocks, through the C ABI.

Tolerance: log-posterior and gradient within rtol 1e-5 (fp32), gradient taken relative to its largest component."""
import zlib

import numpy as np
import pytest
import torch

import cases
from oracle import closures as oc
from oracle import hamiltorch_restated as hr
from vihmc import engine, samplers, synth
from vihmc.spec import LogProbSpec, MLPArch

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _close(got, ref, rtol=RTOL):
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max())


@pytest.mark.parametrize("name", ["small", "full"])
def test_deeponet_logp_grad_matches_reference_golden(name):
    g = cases.load_golden("deeponet_logp_grad.npz")
    inp = cases.don_inputs(name)
    # VI-HMC closure (reduced vector)
    logp, grad = engine.logp_grad(cases.don_spec(inp, "vi"), torch.from_numpy(g[f"{name}/vi/q"]))
    np.testing.assert_allclose(logp.cpu().numpy(), g[f"{name}/vi/logp"], rtol=RTOL)
    for i in range(grad.shape[0]):
        _close(grad[i].cpu().numpy(), g[f"{name}/vi/grad"][i])
    # full HMC closure and the M=2 split closures
    q = torch.from_numpy(g[f"{name}/full/q"])
    logp, grad = engine.logp_grad(cases.don_spec(inp, "full"), q)
    np.testing.assert_allclose(logp.cpu().numpy(), g[f"{name}/full/logp"], rtol=RTOL)
    for i in range(grad.shape[0]):
        _close(grad[i].cpu().numpy(), g[f"{name}/full/grad"][i])
    for si, sp in enumerate(cases.don_spec(inp, "split")):
        logp, grad = engine.logp_grad(sp, q)
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"{name}/split{si}/logp"], rtol=RTOL)
        for i in range(grad.shape[0]):
            _close(grad[i].cpu().numpy(), g[f"{name}/split{si}/grad"][i])


def test_deeponet_many_chains_and_ragged_tiles_vs_fp64_oracle():
    """N and P chosen off the 128-tile grid; 9 chains; relu variant as well."""
    from vihmc.spec import DeepONetArch

    for act in ("tanh", "relu"):
        arch = DeepONetArch(width_branch=20, width_trunk=24, in_branch=7, depth_branch=3, depth_trunk=3, output_neurons=12, act=act)
        x1, x2, y, theta = synth.burgers_like(arch, n_train=131, n_t=13, n_x=11, seed=3)
        spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
        closure = oc.DeepONetLogProb(x1=x1.unsqueeze(1), x2=x2.unsqueeze(0), y=y, width_branch=20, width_trunk=24, in_branch=7,
                                     depth_branch=3, depth_trunk=3, output_neurons=12, act=act, dtype=torch.float64)
        rs = np.random.RandomState(0)
        q = (theta.numpy()[None] + 0.02 * rs.randn(9, arch.num_params)).astype(np.float32)
        logp, grad = engine.logp_grad(spec, torch.from_numpy(q))
        for c in (0, 4, 8):
            lp, gr = oc.value_and_grad(closure, torch.from_numpy(q[c]).double())
            assert abs(float(logp[c]) - float(lp)) <= RTOL * abs(float(lp))
            _close(grad[c].cpu().numpy(), gr.numpy(), rtol=2e-5)


def test_deeponet_predict_matches_oracle():
    inp = cases.don_inputs("small")
    spec = cases.don_spec(inp, "vi")
    g = cases.load_golden("deeponet_logp_grad.npz")
    q = torch.from_numpy(g["small/vi/q"])
    pred = engine.predict(spec, q).cpu()
    closure = cases.don_oracle(inp, "vi", dtype=torch.float64)
    for i in range(len(q)):
        ref = closure.forward(q[i].double())
        np.testing.assert_allclose(pred[i].numpy(), ref.detach().numpy(), rtol=2e-5, atol=2e-6)
    # predict_model on "validation" data = a different slice of operator data
    vx1, vx2, vy, _ = synth.burgers_like(inp["arch"], n_train=4, n_t=5, n_x=7, seed=9)
    out, logps = samplers.predict_model(spec, q, data=(vx1.unsqueeze(1), vx2.unsqueeze(0), vy))
    assert out.shape == (3, 4, 35) and len(logps) == 3


def test_general_sampler_deeponet_leapfrog_vs_oracle():
    inp = cases.don_inputs("small")
    spec = cases.don_spec(inp, "vi")
    closure = cases.don_oracle(inp, "vi", dtype=torch.float64)
    d, S, L, eps = spec.d, 4, 6, 1e-3
    rs = np.random.RandomState(1)
    q0 = torch.from_numpy((inp["mu"].numpy()[inp["ind"]][None] + 0.01 * rs.randn(3, d)).astype(np.float32))
    p = torch.from_numpy(rs.randn(S, 3, d).astype(np.float32))
    u = torch.from_numpy(rs.uniform(0.2, 1.0, size=(S, 3)).astype(np.float32))
    res = engine.run_sampler([spec], q0, S, L, eps, inject_momenta=p, inject_uniforms=u)
    for c in range(3):
        tr = {}
        out = hr.sample(closure, q0[c].double(), num_samples=S, num_steps_per_sample=L, step_size=eps, momenta=p[:, c].double(),
                        uniforms=u[:, c], trace=tr)
        np.testing.assert_allclose(res.hamiltonians[:, c, 0].numpy(), tr["H0"], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(res.hamiltonians[:, c, 1].numpy(), tr["H1"], rtol=1e-5, atol=1e-3)
        assert [bool(a) for a in res.accepted[:, c].numpy()] == tr["accept"]
        np.testing.assert_allclose(res.samples[:, c].numpy(), torch.stack(out).numpy(), rtol=1e-4, atol=1e-5)


def test_split_integrator_vs_oracle():
    """Integrator.SPLITTING with M = 2 closures (main_HMC_splitting.py:367-369) against the restated integrator."""
    inp = cases.don_inputs("small")
    specs = cases.don_spec(inp, "split")
    closures = cases.don_oracle(inp, "split", dtype=torch.float64)
    D, S, L, eps = specs[0].d, 3, 4, 5e-4
    rs = np.random.RandomState(4)
    q0 = torch.from_numpy((inp["theta"].numpy()[None] + 0.01 * rs.randn(2, D)).astype(np.float32))
    p = torch.from_numpy(rs.randn(S, 2, D).astype(np.float32))
    u = torch.from_numpy(rs.uniform(0.2, 1.0, size=(S, 2)).astype(np.float32))
    out = samplers.sample(specs, q0, num_samples=S, num_steps_per_sample=L, step_size=eps, integrator=samplers.Integrator.SPLITTING,
                          inject_momenta=p, inject_uniforms=u, return_result=True)
    for c in range(2):
        tr = {}
        ref = hr.sample(closures, q0[c].double(), num_samples=S, num_steps_per_sample=L, step_size=eps,
                        integrator=hr.Integrator.SPLITTING, momenta=p[:, c].double(), uniforms=u[:, c], trace=tr)
        np.testing.assert_allclose(out.hamiltonians[:, c, 0].numpy(), tr["H0"], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(out.hamiltonians[:, c, 1].numpy(), tr["H1"], rtol=1e-5, atol=1e-3)
        assert [bool(a) for a in out.accepted[:, c].numpy()] == tr["accept"]
        np.testing.assert_allclose(out.samples[:, c].numpy(), torch.stack(ref).numpy(), rtol=1e-4, atol=1e-5)


def test_wide_mlp_dense_path_vs_oracle():
    """An MLP wider than the small-net kernel's range (64 > 32) goes through the GEMM path; cfg5's shape in miniature."""
    arch = MLPArch(in_dim=1, widths=(64, 48, 64), out_dim=1, act="tanh", last_bias=True)
    x, y = synth.wide_bnn_data(n=333, seed=0)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma_scalar=1.0)
    closure = oc.BnnLogProb(x=x, y=y, widths=(64, 48, 64), loss="NLL", tau_out=0.0025,
                            prior=("sliced", [1.0] * len(arch.tensor_numels())), dtype=torch.float64)
    q = synth.default_linear_init(arch, seed=1).unsqueeze(0).repeat(3, 1)
    q[1:] += 0.05 * torch.from_numpy(np.random.RandomState(2).randn(2, arch.num_params).astype(np.float32))
    logp, grad = engine.logp_grad(spec, q)
    for c in range(3):
        lp, gr = oc.value_and_grad(closure, q[c].double())
        assert abs(float(logp[c]) - float(lp)) <= RTOL * abs(float(lp))
        _close(grad[c].cpu().numpy(), gr.numpy(), rtol=2e-5)
    pred = engine.predict(spec, q).cpu()
    np.testing.assert_allclose(pred[0].numpy(), closure.forward(q[0].double()).detach().numpy()[:, 0], rtol=2e-5, atol=2e-5)


def test_wide_mlp_many_rows_output_layer_gradient_on_the_transposed_product():
    """cfg5's shape with enough rows (K = 3000 >= 512) that the single-row products take the tensor-core kernel: the output
    layer's weight gradient dW[1, in] runs as the transposed problem (launch_gemm, M == 1), the first layer's dW[out, 1] as is,
    and the long reductions go through the in-CTA slice sums.  Every tensor of the gradient against the fp64 oracle."""
    arch = MLPArch(in_dim=1, widths=(64, 64), out_dim=1, act="tanh", last_bias=True)
    x, y = synth.wide_bnn_data(n=3000, seed=3)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma_scalar=1.0)
    closure = oc.BnnLogProb(x=x, y=y, widths=(64, 64), loss="NLL", tau_out=0.0025,
                            prior=("sliced", [1.0] * len(arch.tensor_numels())), dtype=torch.float64)
    q = synth.default_linear_init(arch, seed=4).unsqueeze(0).repeat(2, 1)
    q[1] += 0.05 * torch.from_numpy(np.random.RandomState(5).randn(arch.num_params).astype(np.float32))
    logp, grad = engine.logp_grad(spec, q)
    numels = arch.tensor_numels()
    for c in range(2):
        lp, gr = oc.value_and_grad(closure, q[c].double())
        assert abs(float(logp[c]) - float(lp)) <= RTOL * abs(float(lp))
        got, off = grad[c].cpu().double().numpy(), 0
        for n_el in numels:      # per tensor, norm-wise: the output layer's weights are tensor 4 of 6
            a_, b_ = got[off:off + n_el], gr.numpy()[off:off + n_el]
            assert np.linalg.norm(a_ - b_) <= 1e-5 * np.linalg.norm(b_) + 1e-7, (c, off, n_el)
            off += n_el


def test_long_run_posterior_predictive_small_deeponet_matches_oracle_chains():
    """North-star tier 3 for the operator network (judge row N1, second half): 512 engine chains (general sampler on the dense
    path: tcgen05 forward / backward, fused leapfrog and accept kernels) against the 256 fp64 oracle chains of
    tests/golden/deeponet_posterior_summary.npz (oracle/make_golden.py::deeponet_posterior_summary): small DeepONet VI-HMC problem
    (D = 1393, d = 348, 6 functions x 35 trunk points), same start law, eps 0.04, L 10 (acceptance ~0.77), 100 burn-in + 1000
    iterations, every 10th draw.  As in the BNN test the unit of replication is the chain: per-chain time averages of f, f^2 and
    the log within-chain variance at the 210 outputs, se^2 = var_engine / 512 + var_oracle / 256; max |z| < 4 per output, and mean z^2
    per statistic below 6.6 (the outputs are so strongly correlated that a statistic has about one degree of freedom: the 1 % point
    of chi^2_1), acceptance rates within 4 se."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import make_golden as mg

    ref = cases.load_golden("deeponet_posterior_summary.npz")
    chains_o, burn, iters, thin, L = (int(v) for v in ref["cfg"])
    eps = float(ref["eps"])
    arch, x1, x2, y, mu, sigma, ind = mg.n1_don_problem()
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1, frozen=mu, sens_ind=ind,
                       vi_sigma=sigma)
    Cg = 512
    q0 = torch.from_numpy(mg.n1_don_start(mu, sigma, ind, Cg, 20000).astype(np.float32))
    kw = dict(to_host=False, hamiltorch_fallback_rule=False)
    warm = engine.run_sampler([spec], q0, burn, L, eps, burn=burn - 2, seed=5, **kw)
    res = engine.run_sampler([spec], warm.samples[-1].contiguous(), iters + 1, L, eps, burn=0, seed=6, **kw)
    draws = res.samples[thin::thin]
    assert draws.shape[0] == iters // thin
    pred = torch.stack([engine.predict(spec, d_).reshape(Cg, -1) for d_ in draws]).double()      # [T, C, N * P]
    m1, m2 = pred.mean(0).cpu().numpy(), (pred * pred).mean(0).cpu().numpy()
    o1, o2 = ref["f_mean"].astype(np.float64), ref["f_sq_mean"].astype(np.float64)
    worst, zsq = [], []
    for got, want in ((m1, o1), (m2, o2), (np.log(m2 - m1 * m1), np.log(o2 - o1 * o1))):
        se = np.sqrt(got.var(0, ddof=1) / Cg + want.var(0, ddof=1) / chains_o)
        z = (got.mean(0) - want.mean(0)) / se
        worst.append(float(np.abs(z).max()))
        zsq.append(float((z * z).mean()))
    print(f"deeponet posterior predictive: max |z| {worst}, mean z^2 {zsq}")
    assert max(worst) < 4.0 and max(zsq) < 6.6, (worst, zsq)
    acc_g = res.accepted[1:].float().mean(0).cpu().numpy()
    acc_o = ref["accepted"].astype(np.float64)
    se_a = np.sqrt(acc_g.var(ddof=1) / Cg + acc_o.var(ddof=1) / chains_o)
    print(f"acceptance engine {acc_g.mean():.4f} oracle {acc_o.mean():.4f} se {se_a:.4f}")
    assert abs(acc_g.mean() - acc_o.mean()) < 4 * se_a + 1e-3, (acc_g.mean(), acc_o.mean(), se_a)


# ---------------------------------------------------------------------------------------------
# building blocks of the large-d path
# ---------------------------------------------------------------------------------------------
def test_scatter_and_redraw():
    rs = np.random.RandomState(0)
    D, d, Cn = 1001, 77, 5
    frozen = torch.from_numpy(rs.randn(D).astype(np.float32))
    ind = np.sort(rs.choice(D, d, replace=False)).astype(np.int64)
    q = torch.from_numpy(rs.randn(Cn, d).astype(np.float32))
    W = engine.scatter_vi(frozen, ind, q).cpu()
    ref = frozen[None].repeat(Cn, 1)
    ref[:, ind] = q
    assert torch.equal(W, ref)
    # VI redraw hook (my_make_func.py:45-46): w = mu + sigma * z with the Philox stream 2
    from oracle import philox_ref
    sigma = torch.from_numpy((0.01 + rs.rand(D)).astype(np.float32))
    Wr = engine.vi_redraw_philox(11, 3, 2, Cn, frozen, sigma).cpu().numpy()
    z = philox_ref.normals(11, 2, Cn, 3, D, stream=2)
    np.testing.assert_allclose(Wr, frozen.numpy()[None] + sigma.numpy()[None] * z, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("d", [1, 40, 4097, 172401])
def test_leapfrog_update_bitwise_vs_torch(d):
    """p += kick*eps*g; q += drift*eps*p with torch's rounding (separate mul and add), ragged d, per-chain eps."""
    rs = np.random.RandomState(d)
    Cn = 3
    q = torch.from_numpy(rs.randn(Cn, d).astype(np.float32)).cuda()
    p = torch.from_numpy(rs.randn(Cn, d).astype(np.float32)).cuda()
    g = torch.from_numpy((100 * rs.randn(Cn, d)).astype(np.float32)).cuda()
    eps = torch.tensor([1e-4, 3e-4, 7e-5], device="cuda")
    for kick, drift in ((0.5, 1.0), (1.0, 1.0), (1.0, 0.0), (-0.5, 0.0)):
        q_ref, p_ref = q.clone(), p.clone()
        p_ref = p_ref + (kick * eps)[:, None] * g
        if drift != 0.0:
            q_ref = q_ref + (drift * eps)[:, None] * p_ref
        ke = engine.leapfrog_update(q, p, g, 0.0, kick, drift, eps_per_chain=eps, want_ke=True)
        assert torch.equal(p, p_ref) and torch.equal(q, q_ref)
        np.testing.assert_allclose(ke.cpu().numpy(), 0.5 * (p_ref.double() ** 2).sum(1).cpu().numpy(), rtol=2e-6)


def test_mh_accept_select():
    rs = np.random.RandomState(0)
    Cn, d = 6, 5000
    H0 = torch.tensor([1.0, 1.0, 1.0, float("nan"), 1.0, 5.0], device="cuda")
    H1 = torch.tensor([0.5, 2.0, 2.0, 1.0, float("inf"), 5.0], device="cuda")
    u = torch.tensor([0.9, 0.9, 0.1, 0.5, 0.5, 1.0 - 6e-8], device="cuda")  # rho: 0, -1, -1, nan, -inf, 0
    want = [1, 0, 1, 0, 0, 1]
    q_prop = torch.from_numpy(rs.randn(Cn, d).astype(np.float32)).cuda()
    q_cur = torch.zeros_like(q_prop)
    q_fb = torch.from_numpy(rs.randn(Cn, d).astype(np.float32)).cuda()
    fb0 = q_fb.clone()
    stored = torch.empty_like(q_prop)
    acc = torch.empty(Cn, dtype=torch.uint8, device="cuda")
    engine.mh_accept(H0, H1, u, q_prop, q_cur, q_fb, stored=stored, accepted=acc)
    assert acc.cpu().tolist() == want
    for c, a in enumerate(want):
        exp = q_prop[c] if a else fb0[c]
        assert torch.equal(q_cur[c], exp) and torch.equal(stored[c], exp) and torch.equal(q_fb[c], exp)


def test_deeponet_tensor_core_shapes_vs_fp64_oracle():
    """Shipped 172 401-parameter net at N=200, P=2601: every GEMM (layers, head with M=200 / N=2601, both backward
    products, the K=2601 weight-gradient GEMMs, which take the split-K route) runs on the tcgen05 3xTF32 kernel.
    Checked against the fp64 oracle at rtol 2e-5 (the fp32 reference itself sits ~1e-5 from fp64 at this size)."""
    from vihmc.spec import DeepONetArch

    arch = DeepONetArch()
    x1, x2, y, theta = synth.burgers_like(arch, n_train=200, n_t=51, n_x=51, seed=5)
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    closure = oc.DeepONetLogProb(x1=x1.unsqueeze(1), x2=x2.unsqueeze(0), y=y, dtype=torch.float64)
    rs = np.random.RandomState(0)
    q = (theta.numpy()[None] + 0.02 * rs.randn(2, arch.num_params)).astype(np.float32)
    logp, grad = engine.logp_grad(spec, torch.from_numpy(q))
    for c in range(2):
        lp, gr = oc.value_and_grad(closure, torch.from_numpy(q[c]).double())
        assert abs(float(logp[c]) - float(lp)) <= RTOL * abs(float(lp)), (float(logp[c]), float(lp))
        _close(grad[c].cpu().numpy(), gr.numpy(), rtol=2e-5)


def test_data_sharded_sampler_equals_general_sampler():
    """cfg5 mode (rows sharded, gradient all-reduce) in a single process: the Python-orchestrated loop over the C-ABI
    building blocks must reproduce vihmc_sample on the same problem; and two half-data shards with prior_scale = 2
    must sum to the full log-posterior (what the all-reduce computes across ranks)."""
    from vihmc import dist as vd

    arch = MLPArch(in_dim=1, widths=(64, 64), out_dim=1, act="tanh", last_bias=True)
    x, y = synth.wide_bnn_data(n=500, seed=1)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma_scalar=1.0)
    q0 = synth.default_linear_init(arch, seed=2).unsqueeze(0).repeat(3, 1)
    q0[1:] += 0.02 * torch.from_numpy(np.random.RandomState(3).randn(2, arch.num_params).astype(np.float32))
    kw = dict(num_samples=4, num_steps=5, step_size=2e-5, burn=1, seed=9)
    a = vd.sample_data_sharded(vd.shard_spec_rows(spec, 0, 1), q0, **kw)
    assert a["graphs"]                          # the leapfrog steps replayed from CUDA graphs ...
    e = vd.sample_data_sharded(vd.shard_spec_rows(spec, 0, 1), q0, use_graphs=False, **kw)
    assert not e["graphs"]                      # ... are the eager loop bit for bit
    for k in ("samples", "accepted", "hamiltonians"):
        assert torch.equal(a[k], e[k]), k
    b = engine.run_sampler([spec], q0, force_general=True, **kw)
    assert torch.equal(a["accepted"].cpu(), b.accepted)
    np.testing.assert_allclose(a["samples"].cpu().numpy(), b.samples.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a["hamiltonians"].cpu().numpy(), b.hamiltonians.numpy(), rtol=1e-6)
    # shard sum == full (the quantity the all-reduce produces)
    lp_full, g_full = engine.logp_grad(spec, q0)
    parts = [engine.logp_grad(vd.shard_spec_rows(spec, r, 2), q0) for r in range(2)]
    np.testing.assert_allclose((parts[0][0] + parts[1][0]).cpu().numpy(), lp_full.cpu().numpy(), rtol=2e-6)
    gs = (parts[0][1] + parts[1][1]).cpu().numpy()
    np.testing.assert_allclose(gs, g_full.cpu().numpy(), rtol=2e-5, atol=2e-5 * np.abs(gs).max())


def _gemm_case(rs, batch, M, N, K, a_layout, b_layout, dev):
    """a_layout / b_layout: 'k' = reduction dimension contiguous, 'mn' = M (resp. N) contiguous, 'odd' = K-contiguous
    with a row stride that is not a multiple of 4 floats (scalar staging), 'shared' = one matrix for the whole batch."""
    def make(rows, cols, layout):   # logical [batch, rows, cols(K)] for A; B is built as [batch, N, K] and transposed
        if layout == "k":
            return torch.from_numpy(rs.randn(batch, rows, cols).astype(np.float32)).to(dev)
        if layout == "mn":
            return torch.from_numpy(rs.randn(batch, cols, rows).astype(np.float32)).to(dev).transpose(1, 2)
        if layout == "odd":
            return torch.from_numpy(rs.randn(batch, rows, cols + 1).astype(np.float32)).to(dev)[:, :, :cols]
        if layout == "shared":
            return torch.from_numpy(rs.randn(1, rows, cols).astype(np.float32)).to(dev).expand(batch, -1, -1)
        raise ValueError(layout)
    A = make(M, K, a_layout)
    B = make(N, K, b_layout).transpose(1, 2)
    return A, B


@pytest.mark.parametrize("a_layout,b_layout", [("k", "k"), ("k", "mn"), ("mn", "mn"), ("mn", "k"), ("odd", "k"), ("k", "odd"),
                                               ("shared", "mn"), ("mn", "shared")])
@pytest.mark.parametrize("shape", [(3, 35, 16, 16), (2, 300, 100, 100), (2, 100, 100, 1030), (1, 130, 37, 2500), (2, 64, 8, 600)])
def test_tensor_core_gemm_every_staging_mode_vs_fp64(a_layout, b_layout, shape):
    """vihmc_gemm_batched on the tcgen05 3xTF32 kernel against an fp64 matmul, for K-major, MN-major and scalar operand
    staging, ragged tiles (M, N, K off the 128 / 16 grid, M or N not a multiple of 4) and split-K (K > 2048)."""
    batch, M, N, K = shape
    dev = torch.device("cuda:0")
    A, B = _gemm_case(np.random.RandomState(zlib.crc32(repr((a_layout, b_layout, shape)).encode())), batch, M, N, K, a_layout, b_layout, dev)
    ref = torch.matmul(A.double(), B.double())
    for tc in (True, False):
        got = engine.gemm_batched(A, B, tensor_cores=tc)
        err = (got.double() - ref).abs().max().item()
        scale = (A.double().abs() @ B.double().abs()).max().item()   # sum_k |a||b|: the natural error scale of a dot product
        assert err <= 2e-6 * scale, (tc, err, scale)


_NOFUSE_SCRIPT = r"""
import sys
sys.path[:0] = [{root!r}, {pkg!r}, {tests!r}]
import numpy as np, torch
import cases
from vihmc import engine
g = cases.load_golden("deeponet_logp_grad.npz")
out = {{}}
for name in ("small", "full"):
    inp = cases.don_inputs(name)
    logp, grad = engine.logp_grad(cases.don_spec(inp, "full"), torch.from_numpy(g[name + "/full/q"]))
    out[name + "/logp"], out[name + "/grad"] = logp.cpu().numpy(), grad.cpu().numpy()
np.savez({dst!r}, **out)
"""


def test_fused_stack_kernels_agree_with_the_per_layer_path(tmp_path):
    """The fused forward / backward stack kernels (csrc/fused_stack.cuh) against the per-layer GEMM path (VIHMC_DENSE_NOFUSE=1,
    a separate process because the switch is read once): both meet the reference golden vectors and each other at 1e-5."""
    import os, subprocess, sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for nofuse in ("0", "1"):
        dst = str(tmp_path / f"nofuse{nofuse}.npz")
        script = _NOFUSE_SCRIPT.format(root=root, pkg=os.path.join(root, "vi-hmc_b200"), tests=os.path.join(root, "tests"), dst=dst)
        out = subprocess.run([sys.executable, "-c", script], env=dict(os.environ, VIHMC_DENSE_NOFUSE=nofuse), capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        res[nofuse] = np.load(dst)
    g = cases.load_golden("deeponet_logp_grad.npz")
    for name in ("small", "full"):
        for nofuse in ("0", "1"):
            np.testing.assert_allclose(res[nofuse][f"{name}/logp"], g[f"{name}/full/logp"], rtol=RTOL)
            for i in range(res[nofuse][f"{name}/grad"].shape[0]):
                _close(res[nofuse][f"{name}/grad"][i], g[f"{name}/full/grad"][i])
        for i in range(res["0"][f"{name}/grad"].shape[0]):
            _close(res["0"][f"{name}/grad"][i], res["1"][f"{name}/grad"][i])


def test_validation_step_mse_from_the_log_likelihood_matches_the_predictions(tmp_path):
    """main_VI_HMC_burgers.py:290-301: per-draw MSE and expected log probability.  vihmc.validate gets the MSE from the value-only
    log-likelihood kernel (no S x N x P prediction tensor); here it is checked against the predictions themselves."""
    from vihmc import validate

    inp = cases.don_inputs("small")
    spec = cases.don_spec(inp, "vi")
    g = torch.Generator().manual_seed(3)
    q = inp["mu"][inp["ind"]][None] + 0.02 * torch.randn(7, len(inp["ind"]), generator=g)
    data = (inp["x1"].unsqueeze(1), inp["x2"].unsqueeze(0), inp["y"])
    pred, logp = samplers.predict_model(spec, q, data=data)
    want = ((pred - inp["y"][None]) ** 2).mean(dim=(1, 2))
    lp, mse = validate.sample_log_prob_and_mse(spec, q, data=data)
    np.testing.assert_allclose(mse.numpy(), want.numpy(), rtol=2e-5)
    np.testing.assert_allclose(lp.numpy(), torch.stack(logp).numpy(), rtol=1e-6)
    out = validate.validate(spec, list(q), burn=2, data=data, out_dir=str(tmp_path), uid="t")
    saved = np.load(tmp_path / "sample_mse_t.npy")
    assert saved.shape == (5,) and saved.dtype == np.float32
    assert out["final_mse"] == pytest.approx(float(want[-1]), rel=2e-5) and out["min_mse"] == pytest.approx(float(want[2:].min()), rel=2e-5)
    assert out["expected_log_prob"] == pytest.approx(float(torch.stack(logp)[2:].mean()), rel=1e-6)


def test_trunk_subsampling_closure_calls_match_the_restated_sampler():
    """cfg.sample_data (main_VI_HMC_burgers.py:127-137): every closure call -- 2 Hamiltonians + L + 1 gradients per iteration --
    redraws cfg.p trunk points from Python's global generator.  The engine composes the iteration on the host from its building
    blocks; with the same random.seed, injected momenta and uniforms it must follow the restated hamiltorch sampler running the
    oracle closure (same subsets in the same order), accept / reject included."""
    import dataclasses
    import random

    inp = cases.don_inputs("small")
    spec = dataclasses.replace(cases.don_spec(inp, "vi"), trunk_subsample=9)
    arch = inp["arch"]
    d, S, L, eps = spec.d, 5, 3, 1e-4
    rs = np.random.RandomState(3)
    q0 = inp["mu"][inp["ind"]].clone()
    p = torch.from_numpy(rs.randn(S, 1, d).astype(np.float32))
    p[2] *= 30.0                                   # one overshooting trajectory
    u = torch.from_numpy(rs.uniform(0.2, 1.0, size=(S, 1)).astype(np.float32))
    random.seed(11)
    res = samplers.sample(spec, q0, num_samples=S, num_steps_per_sample=L, step_size=eps, burn=1, inject_momenta=p, inject_uniforms=u,
                          return_result=True)
    oracle = cases.oc.DeepONetLogProb(x1=inp["x1"].unsqueeze(1), x2=inp["x2"].unsqueeze(0), y=inp["y"], frozen=inp["mu"],
                                      sens_ind=inp["ind"], sample_p=9, **cases._don_kwargs(arch, torch.float64))
    random.seed(11)
    trace = {}
    ref = hr.sample(oracle, q0.double(), num_samples=S, num_steps_per_sample=L, step_size=eps, burn=1, momenta=p[:, 0].double(),
                    uniforms=u[:, 0], trace=trace)
    assert res.samples.shape == (S - 1, 1, d)
    np.testing.assert_allclose(res.hamiltonians[:, 0, 0].numpy(), trace["H0"], rtol=1e-5)
    np.testing.assert_allclose(res.hamiltonians[:, 0, 1].numpy(), trace["H1"], rtol=1e-5)
    assert [bool(a) for a in res.accepted[:, 0]] == trace["accept"]
    np.testing.assert_allclose(res.samples[:, 0].numpy(), torch.stack(ref).numpy(), rtol=1e-4, atol=1e-5)
    # a different seed of the global generator gives different subsets, hence different Hamiltonians
    random.seed(12)
    res2 = samplers.sample(spec, q0, num_samples=2, num_steps_per_sample=L, step_size=eps, inject_momenta=p[:2], inject_uniforms=u[:2],
                           return_result=True)
    assert not np.allclose(res2.hamiltonians[0, 0].numpy(), res.hamiltonians[0, 0].numpy(), rtol=1e-7)


def test_vi_redraw_hook_deeponet_matches_oracle():
    """a9 for the operator network (Operator_network/VI_HMC/my_make_func.py:38-42): per-sample redraw of all frozen weights in the
    general sampler (every chain evaluates the closure on its own draw: vihmc_problem.frozen_chain_stride), against the oracle."""
    inp = cases.don_inputs("small")
    spec = cases.don_spec(inp, "vi")
    d, D, S, L, eps, Cn = len(inp["ind"]), inp["arch"].num_params, 3, 3, 1e-4, 2
    rs = np.random.RandomState(5)
    q0 = (inp["mu"][inp["ind"]][None] + 0.01 * torch.from_numpy(rs.randn(Cn, d).astype(np.float32)))
    p = torch.from_numpy(rs.randn(S, Cn, d).astype(np.float32))
    z = torch.from_numpy(rs.randn(S, Cn, D).astype(np.float32))
    u = torch.full((S, Cn), 1e-30)
    res = engine.run_sampler([spec], q0, S, L, eps, inject_momenta=p, inject_uniforms=u, vi_redraw=True, inject_vi_normals=z)
    want = inp["mu"][None, None] + inp["sigma"][None, None] * z
    np.testing.assert_allclose(res.vi_params.numpy(), want.numpy(), rtol=1e-6, atol=1e-7)
    for c in range(Cn):
        closure = cases.don_oracle(inp, "vi", dtype=torch.float64)
        q = q0[c].double()
        for n in range(S):
            closure.frozen = inp["mu"].double() + inp["sigma"].double() * z[n, c].double()
            H0 = float(hr.hamiltonian(q, p[n, c].double(), closure))
            q, p1 = hr.leapfrog(q, p[n, c].double(), closure, L, eps)
            H1 = float(hr.hamiltonian(q, p1, closure))
            assert abs(float(res.hamiltonians[n, c, 0]) - H0) <= 2e-5 * abs(H0) + 1e-3
            assert abs(float(res.hamiltonians[n, c, 1]) - H1) <= 2e-5 * abs(H1) + 1e-3
            if n >= 1:
                np.testing.assert_allclose(res.samples[n, c].numpy(), q.numpy(), rtol=1e-4, atol=1e-5)
