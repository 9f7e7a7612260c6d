"""CPU: pins the oracle restatement against golden vectors produced by the reference itself
(``oracle/make_golden.py``).  The reference ships no tests/golden vectors."""
import numpy as np
import pytest
import torch

from oracle import psld_oracle as O
from oracle.weights import fill_state_dict, inpaint_draws, inpaint_inputs, noise_bank, prior
from psld_b200 import mid_config, tiny_config
from _net import fake_score
from test_gpu_sampler import _golden_cfg


def test_scalars(golden_dir):
    g = np.load(f"{golden_dir}/scalars.npz")
    cols = {c: i for i, c in enumerate(g["columns"])}
    for row in g["table"]:
        v = lambda k: row[cols[k]]
        cfg = tiny_config()
        cfg.model.sde.update(nu=v("nu"), gamma=v("gamma"), beta_min=v("beta0"), beta_max=v("beta1"),
                             decomp_mode="upper" if v("upper") else "lower")
        s = O.PSLDScalars(cfg)
        a, c = s.half_step(v("t"), v("dt") / 2)
        np.testing.assert_allclose(a, [v("a_xx"), v("a_xm"), v("a_mx"), v("a_mm")], rtol=1e-13, atol=1e-16)
        np.testing.assert_allclose(c, [v("c11"), v("c12"), v("c21"), v("c22")], rtol=1e-10, atol=1e-16)
        cov = s.cov(0.0, s.mm_0, 1.0 - v("t"))
        np.testing.assert_allclose(cov, [v("XX"), v("XM"), v("MM")], rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(s.get_inv_coeff(cov), [v("i11"), v("i12"), v("i21"), v("i22")],
                                   rtol=1e-10, atol=1e-16)
        assert s.m_inv == v("m_inv") and s.m == v("m") and s.mm_0 == v("mm_0")


def test_survey_known_answers():
    """SURVEY.md §4 golden scalars (nu=4.01, gamma=0.01, beta=8, h=5e-4)."""
    s = O.PSLDScalars(tiny_config())
    a, c = s.half_step(0.0, 5e-4)
    np.testing.assert_allclose(a, [0.9999720216609386, -0.007967904555066366, 0.0019919761387665914,
                                   0.9920041171058721], rtol=1e-13)
    np.testing.assert_allclose([c[0], c[2], c[3]], [0.006331273141002044, -0.00250690112338919,
                                                    0.06302147561799819], rtol=1e-10)
    ic = s.get_inv_coeff(s.cov(0.0, s.mm_0, 1.0))
    np.testing.assert_allclose(ic, [1.000007398133583, -1.3064602171464342e-05, 0, 2.000011531445317],
                               rtol=1e-10, atol=1e-18)
    ic = s.get_inv_coeff(s.cov(0.0, s.mm_0, 1e-3))
    np.testing.assert_allclose(ic, [109.63789430671079, -20.20633787397167, 0, 7.6699136770391005],
                               rtol=1e-9, atol=1e-18)
    assert s.m_inv == 4.0 and s.m == 0.25 and abs(s.mm_0 - 0.01) < 1e-18
    # L^T L^{-T} = I
    for tau in (1.0, 0.3, 1e-3):
        var = s.cov(0.0, s.mm_0, tau)
        l = s.get_coeff(var); li = s.get_inv_coeff(var)
        Lm = np.array([[l[0], l[1]], [l[2], l[3]]]); Li = np.array([[li[0], li[1]], [li[2], li[3]]])
        np.testing.assert_allclose(Lm.T @ Li, np.eye(2), atol=1e-12)


def test_upfirdn(golden_dir):
    g = np.load(f"{golden_dir}/upfirdn.npz")
    x = torch.from_numpy(g["x"])
    for name in ["down", "up", "pad", "generic", "crop"]:
        up, down, p0, p1, gain = g["arg_" + name]
        y = O.upfirdn2d(x, g["k"] * gain, up=int(up), down=int(down), pad=(int(p0), int(p1)))
        np.testing.assert_allclose(y.numpy(), g["y_" + name], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name,cfg", [("tiny", tiny_config()), ("mid", mid_config())])
def test_forward(golden_dir, name, cfg):
    g = np.load(f"{golden_dir}/forward_{name}.npz")
    from psld_b200 import NCSNpp
    shapes = {k: tuple(v.shape) for k, v in NCSNpp(cfg).state_dict().items()}
    sd = fill_state_dict(shapes, int(g["seed"]))
    y = O.ncsnpp_forward(cfg, sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=1e-5 * np.abs(g["y"]).max())


def test_modules(golden_dir):
    """Single ResnetBlockBigGANpp (plain/down/up/cat), AttnBlockpp and pyramid Downsample."""
    g = np.load(f"{golden_dir}/modules_tiny.npz")
    cfg = tiny_config()
    from psld_b200 import NCSNpp
    shapes = {k: tuple(v.shape) for k, v in NCSNpp(cfg).state_dict().items()}
    sd = fill_state_dict(shapes, 0)
    sf = cfg.model.score_fn
    t = torch.from_numpy(g["t"])
    xp = torch.log(t)[:, None] * sd["all_modules.0.W"][None] * 2 * np.pi
    temb = torch.cat([torch.sin(xp), torch.cos(xp)], -1)
    F = torch.nn.functional
    temb = F.linear(temb, sd["all_modules.1.weight"], sd["all_modules.1.bias"])
    temb = F.linear(F.silu(temb), sd["all_modules.2.weight"], sd["all_modules.2.bias"])
    for kind in g["kinds"]:
        idx, cls, ud = str(kind).split(":")
        x = torch.from_numpy(g[f"in_{idx}"])
        p = f"all_modules.{idx}"
        if cls == "ResnetBlockBigGANpp":
            y = O.resblock(sd, p, x, temb, sf, up=ud[0] == "1", down=ud[1] == "1")
        elif cls == "AttnBlockpp":
            y = O.attnblock(sd, p, x)
        else:
            y = O.conv_downsample_2d(x, sd[p + ".Conv2d_0.weight"], sf.fir_kernel) \
                + sd[p + ".Conv2d_0.bias"].reshape(1, -1, 1, 1)
        ref = g[f"out_{idx}"]
        np.testing.assert_allclose(y.numpy(), ref, rtol=0, atol=2e-6 * np.abs(ref).max())


@pytest.mark.parametrize("tag", ["sscs_fake_uniform", "em_fake_uniform", "sscs_fake_quad",
                                 "em_fake_quad", "sscs_fake_nodenoise"])
def test_sampler_algebra(golden_dir, tag):
    g = np.load(f"{golden_dir}/sampler_{tag}.npz")
    cfg = _golden_cfg(tag)
    kind = cfg.evaluation.sampler.name
    ts, n = O.time_grid(cfg)
    assert n == int(g["n"])
    np.testing.assert_allclose(ts, g["ts"], rtol=0, atol=3e-16)
    B = int(g["B"])
    sde = O.PSLDScalars(cfg)
    u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
    nb = noise_bank((2 if kind == "sscs_sde" else 1) * n, (B, 6, 8, 8), 2)
    states = {}
    fn = O.sscs_sample if kind == "sscs_sde" else O.em_sample
    out = fn(cfg, fake_score, u0, ts, n, nb, denoise=cfg.evaluation.denoise,
             record=lambda i, u: states.__setitem__(i, u.clone()))
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-7 * np.abs(ref).max()
    for i in g["probe"]:
        r = g[f"state_{int(i)}"]
        assert np.abs(states[int(i)].numpy()[: r.shape[0]] - r).max() <= 5e-7 * np.abs(r).max()


@pytest.mark.parametrize("kind,fname", [("em_sde", "sampler_tiny_em100.npz"),
                                        ("sscs_sde", "sampler_tiny_sscs100.npz")])
def test_sampler_with_network(golden_dir, kind, fname):
    """BASELINE.json configs[0] through the oracle (a few seconds of CPU)."""
    g = np.load(f"{golden_dir}/{fname}")
    cfg = tiny_config(sampler=kind)
    from psld_b200 import NCSNpp
    shapes = {k: tuple(v.shape) for k, v in NCSNpp(cfg).state_dict().items()}
    sd = fill_state_dict(shapes, 0)
    ts, n = O.time_grid(cfg)
    B = int(g["B"])
    u0 = prior((B, 3, 32, 32), 0.5, 1)
    nb = noise_bank((2 if kind == "sscs_sde" else 1) * n, (B, 6, 32, 32), 2)
    fn = O.sscs_sample if kind == "sscs_sde" else O.em_sample
    out = fn(cfg, O.OracleScoreFn(cfg, sd), u0, ts, n, nb)
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-6 * np.abs(ref).max()


def inpaint_cfg(mode):
    cfg = tiny_config(sampler="ip_em_sde", n_discrete_steps=30)
    cfg.training.mode = mode
    cfg.data.image_size = 8
    return cfg


@pytest.mark.parametrize("mode", ["hsm", "dsm"])
def test_inpaint_sampler_algebra(golden_dir, mode):
    """Oracle restatement of ES3EulerMaruyamaInpainter (sde.py:125-224) vs the reference's output."""
    g = np.load(f"{golden_dir}/sampler_ip_em_fake_{mode}.npz")
    cfg = inpaint_cfg(mode)
    ts, n = O.time_grid(cfg)
    assert n == int(g["n"])
    B = int(g["B"])
    sde = O.PSLDScalars(cfg)
    u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
    x_0, mask = inpaint_inputs(B, 8, 3)
    out = O.inpaint_em_sample(cfg, fake_score, x_0, mask, u0, ts, n, inpaint_draws(n, B, 8, 2),
                              denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-7 * np.abs(ref).max()
    # the known region of the result is the perturbation mean of x_0 at tau = T - fl32(T - eps)
    axx = O.mean_coeffs(sde, float(np.float32(1.0) - np.float32(1.0 - cfg.evaluation.eval_eps)))[0]
    if mode == "hsm":
        known = (mask.numpy() == 1)
        np.testing.assert_allclose(out.numpy()[:, :3][known], (axx * x_0.double().numpy())[known],
                                   rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("tag,kw", [("uniform", {}), ("quad", dict(stride_type="quadratic"))])
def test_vp_sampler_algebra(golden_dir, tag, kw):
    """EM on the VP-SDE baseline (vpsde.py:42-74 under sde.py:16-58) vs the reference's output."""
    from _net import vp_config
    g = np.load(f"{golden_dir}/sampler_vp_em_fake_{tag}.npz")
    cfg = vp_config(**kw)
    ts, n = O.time_grid(cfg)
    assert n == int(g["n"])
    B = int(g["B"])
    x0 = noise_bank(1, (B, 3, 8, 8), 1)[0]
    out = O.vp_em_sample(cfg, fake_score, x0, ts, n, noise_bank(n, (B, 3, 8, 8), 2),
                         denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-7 * np.abs(ref).max()


@pytest.mark.parametrize("tag,kw", [("uniform", {}), ("quad_nodenoise", dict(stride_type="quadratic", denoise=False))])
def test_cc_sampler_algebra(golden_dir, tag, kw):
    """Oracle restatement of ClassCondEulerMaruyamaSampler (sde.py:61-122) vs the reference's output
    (stand-in score network and stand-in differentiable classifier)."""
    from _net import cc_config
    from oracle.weights import fake_classifier
    g = np.load(f"{golden_dir}/sampler_cc_em_fake_{tag}.npz")
    cfg = cc_config(**kw)
    ts, n = O.time_grid(cfg)
    assert n == int(g["n"])
    B = int(g["B"])
    sde = O.PSLDScalars(cfg)
    u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
    ev = cfg.clf.evaluation
    out = O.cc_em_sample(cfg, fake_score, fake_classifier, u0, ts, n, noise_bank(n, (B, 6, 8, 8), 2),
                         ev.label_to_sample, ev.clf_temp, denoise=cfg.evaluation.denoise,
                         eps=cfg.evaluation.eval_eps)
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-7 * np.abs(ref).max()
    # the guidance term matters: without it the end state differs visibly
    plain = O.em_sample(cfg, fake_score, u0, ts, n, noise_bank(n, (B, 6, 8, 8), 2),
                        denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
    assert np.abs(plain.numpy() - ref).max() >= 1e-3 * np.abs(ref).max()


@pytest.mark.parametrize("tag,tol,den", [("tol1e-5", 1e-5, True), ("tol1e-4_nodenoise", 1e-4, False)])
def test_bb_ode_sampler(golden_dir, tag, tol, den):
    """Oracle restatement of BBODESampler (ode.py:41-76: probability-flow drift + the solve_ivp call
    torchdiffeq's scipy_solver makes) vs the reference's own output and NFE count."""
    from _net import ode_config
    g = np.load(f"{golden_dir}/sampler_bb_ode_gauss_{tag}.npz")
    cfg = ode_config(tol, denoise=den)
    B = int(g["B"])
    sde = O.PSLDScalars(cfg)
    u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
    from oracle.weights import gaussian_score_fn
    out, nfe = O.bb_ode_sample(cfg, gaussian_score_fn(cfg), u0, tol, tol, denoise=den, eps=cfg.evaluation.eval_eps)
    ref = g["final"]
    assert nfe == int(g["nfe"])
    # every evaluation sees the float32 rounding of (t, y): two runs agree to float32 resolution
    assert np.abs(out.double().numpy() - ref).max() <= 5e-6 * np.abs(ref).max()


VP_ODE_CASES = [("tol1e-5", 1e-5, True, torch.float32), ("tol1e-4_f64_nodenoise", 1e-4, False, torch.float64)]


@pytest.mark.parametrize("tag,tol,den,dt", VP_ODE_CASES)
def test_bb_ode_sampler_vp(golden_dir, tag, tol, den, dt):
    """bb_ode over the VP-SDE baseline (sample_uncond_vpsde_ode.sh): oracle vs the reference's BBODESampler +
    VPSDE (float32 and float64 batches)."""
    from _net import vp_config
    from oracle.weights import vp_gaussian_score_fn
    g = np.load(f"{golden_dir}/sampler_bb_ode_vp_gauss_{tag}.npz")
    cfg = vp_config(sampler=dict(name="bb_ode", solver="RK45", rtol=tol, atol=tol), denoise=den)
    B = int(g["B"])
    x0 = prior((B, 3, 8, 8), 1.0, 1)[:, :3].contiguous().to(dt)
    out, nfe = O.bb_ode_sample(cfg, vp_gaussian_score_fn(cfg), x0, tol, tol, denoise=den, eps=cfg.evaluation.eval_eps)
    ref = g["final"]
    assert nfe == int(g["nfe"])
    assert np.abs(out.double().numpy() - ref).max() <= 5e-6 * np.abs(ref).max()


# ---------------------------------------------------------------- classifier-free guidance (configs[4])
def _guided_oracle(cfg, w):
    from psld_b200 import NCSNpp
    shapes = {k: tuple(v.shape) for k, v in NCSNpp(cfg).state_dict().items()}
    return O.GuidedScoreFn(O.OracleScoreFn(cfg, fill_state_dict(shapes, 0)),
                           O.OracleScoreFn(cfg, fill_state_dict(shapes, 1)), w)


def test_guided_forward(golden_dir):
    """eps = (1 + w) eps_c - w eps_u of two reference NCSN++ networks (oracle/make_golden.py::golden_guidance)."""
    g = np.load(f"{golden_dir}/forward_cfg_mid.npz")
    fn = _guided_oracle(mid_config(), float(g["weight"]))
    y = fn(torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=1e-5 * np.abs(g["y"]).max())
    y0 = O.GuidedScoreFn(fn.cond, fn.uncond, 0.0)(torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    assert torch.equal(y0, fn.cond(torch.from_numpy(g["x"]), torch.from_numpy(g["t"])))   # w = 0: identity


@pytest.mark.parametrize("kind", ["sscs_sde", "em_sde"])
def test_guided_sampler(golden_dir, kind):
    """The reference's own SSCS / EM sampler driven by the guided score_fn vs the oracle loop."""
    g = np.load(f"{golden_dir}/sampler_cfg_tiny_{kind.split('_')[0]}40.npz")
    cfg = tiny_config(sampler=kind, n_discrete_steps=40)
    ts, n = O.time_grid(cfg)
    B = int(g["B"])
    u0 = prior((B, 3, 32, 32), 0.5, 1)
    nb = noise_bank((2 if kind == "sscs_sde" else 1) * n, (B, 6, 32, 32), 2)
    fn = O.sscs_sample if kind == "sscs_sde" else O.em_sample
    out = fn(cfg, _guided_oracle(cfg, 1.5), u0, ts, n, nb)
    ref = g["final"]
    assert np.abs(out.numpy() - ref).max() <= 5e-6 * np.abs(ref).max()
