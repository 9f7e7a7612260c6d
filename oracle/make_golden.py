"""TEST INFRASTRUCTURE ONLY — generates ``tests/golden/*.npz`` by running the UNMODIFIED
reference (``/root/reference``) on CPU in the build container.

    python -m oracle.make_golden            # regenerates every fixture

The reference ships no golden vectors, so these fixtures are what pins
both the oracle restatement (``oracle/psld_oracle.py``) and the CUDA path.  Inputs are
reproducible from seeds alone (``oracle/weights.py``), so fixtures hold outputs only
(plus the small inputs, for convenience).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import NoiseBank, load_reference, reference_time_grid  # noqa: E402
from oracle.weights import inpaint_draws, inpaint_inputs, fill_state_dict, noise_bank, prior  # noqa: E402
from psld_b200.config import celeba64_config, cifar10_config, mid_config, tiny_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def ref_net(R, cfg, seed=0):
    net = R.NCSNpp(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = fill_state_dict(shapes, seed)
    net.load_state_dict(sd)
    return net, sd


# ---------------------------------------------------------------- scalar known answers
def golden_scalars(R):
    rows = []
    combos = [(4.01, 0.01, 8.0, 8.0, "lower"), (4.02, 0.02, 8.0, 8.0, "lower"),
              (4.005, 0.005, 8.0, 8.0, "lower"), (4.0, 0.0, 8.0, 8.0, "lower"),
              (4.02, 0.02, 0.5, 12.0, "lower"), (4.01, 0.01, 8.0, 8.0, "upper")]
    for nu, ga, b0, b1, mode in combos:
        cfg = tiny_config()
        cfg.model.sde.update(nu=nu, gamma=ga, beta_min=b0, beta_max=b1, decomp_mode=mode)
        sde = R.PSLD(cfg)
        S = R.get_module("samplers", "sscs_sde")(cfg, sde, None)
        for t, dt in [(0.0, 1e-3), (0.25, 2e-3), (0.5, 1e-2), (0.9, 5e-4), (0.998, 1e-3)]:
            tt = torch.tensor([t], dtype=torch.float64)
            h = torch.tensor(dt / 2, dtype=torch.float64)
            # mean 2x2 from unit vectors
            e = torch.zeros(1, 2, 1, 1, dtype=torch.float64)
            ex = e.clone(); ex[0, 0] = 1.0
            em = e.clone(); em[0, 1] = 1.0
            mx = S._mean(ex, tt, h).flatten()      # (a_xx, a_mx)
            mm = S._mean(em, tt, h).flatten()      # (a_xm, a_mm)
            var = S._var(tt, h)
            c = sde.get_coeff(var)
            tau = torch.tensor([1.0 - t], dtype=torch.float64)
            cov = sde._cov(0, sde.mm_0, tau)
            ic = sde.get_inv_coeff(cov)
            rows.append([nu, ga, b0, b1, 0.0 if mode == "lower" else 1.0, t, dt,
                         mx[0].item(), mm[0].item(), mx[1].item(), mm[1].item(),
                         *[float(v) for v in var], *[float(v) for v in c],
                         *[float(v) for v in cov], *[float(v) for v in ic],
                         float(sde.beta_t(tau)), sde.m_inv, sde.m, sde.mm_0])
    cols = ("nu gamma beta0 beta1 upper t dt a_xx a_xm a_mx a_mm hxx hxm hmm c11 c12 c21 c22 "
            "XX XM MM i11 i12 i21 i22 beta_tau m_inv m mm_0").split()
    _save("scalars.npz", table=np.asarray(rows, dtype=np.float64), columns=np.asarray(cols))


# ---------------------------------------------------------------- upfirdn2d
def golden_upfirdn(R):
    r = np.random.default_rng(7)
    x = torch.from_numpy(r.standard_normal((2, 5, 8, 8)).astype(np.float32))
    k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32)
    k /= k.sum()
    kt = torch.from_numpy(k)
    out = {"x": x.numpy(), "k": k}
    for name, (kk, up, down, p0, p1) in {
        "down": (kt, 1, 2, 1, 1), "up": (kt * 4, 2, 1, 2, 1), "pad": (kt, 1, 1, 2, 2),
        "generic": (kt, 2, 3, 3, 0), "crop": (kt, 1, 1, -1, 2),
    }.items():
        y = R.upfirdn2d_native(x, kk, up, up, down, down, p0, p1, p0, p1)
        out["y_" + name] = y.numpy()
        out["arg_" + name] = np.asarray([up, down, p0, p1, 4.0 if name == "up" else 1.0])
    _save("upfirdn.npz", **out)


# ---------------------------------------------------------------- network forwards
def golden_forward(R, cfg, fname, B, seed=0, times=(0.731, 0.0123, 1.0, 1e-3)):
    net, _ = ref_net(R, cfg, seed)
    r = np.random.default_rng([11, seed])
    H = cfg.data.image_size
    x = torch.from_numpy((r.standard_normal((B, 6, H, H)) * 1.5).astype(np.float32))
    t = torch.from_numpy(np.asarray(list(times)[:B], dtype=np.float32))
    with torch.no_grad():
        y = net(x, t)
    _save(fname, x=x.numpy(), t=t.numpy(), y=y.numpy(), seed=np.asarray(seed))


def golden_modules(R):
    """Single-module outputs (ResnetBlockBigGANpp plain/down/up/cat, AttnBlockpp,
    pyramid Downsample) on seeded inputs, taken with forward hooks from the tiny net."""
    cfg = tiny_config()
    net, _ = ref_net(R, cfg, 0)
    want = {}
    lp = R.layerspp
    seen = set()
    for i, m in enumerate(net.all_modules):
        if isinstance(m, (lp.ResnetBlockBigGANpp, lp.AttnBlockpp, lp.Downsample)):
            kind = (type(m).__name__, getattr(m, "up", 0), getattr(m, "down", 0),
                    getattr(m, "in_ch", 0) != getattr(m, "out_ch", 0))
            if kind not in seen:         # one instance of every distinct kind
                seen.add(kind)
                want[i] = m
    rec = {}

    def mk(i):
        def hook(mod, inp, out):
            rec[f"in_{i}"] = inp[0].detach().numpy().copy()
            rec[f"out_{i}"] = out.detach().numpy().copy()
        return hook

    hs = [m.register_forward_hook(mk(i)) for i, m in want.items()]
    r = np.random.default_rng(5)
    x = torch.from_numpy(r.standard_normal((1, 6, 32, 32)).astype(np.float32))
    t = torch.tensor([0.3])
    with torch.no_grad():
        net(x, t)
    for h in hs:
        h.remove()
    rec["x"] = x.numpy(); rec["t"] = t.numpy()
    rec["kinds"] = np.asarray([f"{i}:{type(m).__name__}:{int(getattr(m, 'up', False))}{int(getattr(m, 'down', False))}"
                               for i, m in want.items()])
    _save("modules_tiny.npz", **rec)


# ---------------------------------------------------------------- samplers
def golden_sampler(R, cfg, fname, B, seed_w=0, seed_p=1, seed_n=2, keep=2, score="net"):
    sde = R.PSLD(cfg)
    H = cfg.data.image_size
    name = cfg.evaluation.sampler.name
    if score == "net":
        net, _ = ref_net(R, cfg, seed_w)
    elif callable(score):
        net = score
    else:
        net = fake_score
    ts, n = reference_time_grid(cfg)
    per = 2 if name == "sscs_sde" else 1
    nb = noise_bank(per * n, (B, 6, H, H), seed_n)
    u0 = prior((B, 3, H, H), float(np.sqrt(sde.m)), seed_p)
    S = R.get_module("samplers", name)(cfg, sde, net)
    stats, states = [], {}
    probe = sorted(set([0, 1, n // 2, n - 1]))
    orig = S.predictor_update_fn
    ctr = {"i": 0}

    def wrapped(u, t, dt):
        r = orig(u, t, dt)
        un = r[0] if isinstance(r, tuple) else r
        i = ctr["i"]; ctr["i"] += 1
        stats.append([un.sum().item(), un.abs().sum().item(), un.abs().max().item(),
                      (un.double() ** 2).sum().item()])
        if i in probe:
            states[f"state_{i}"] = un[:keep].double().numpy().copy()
        return r

    S.predictor_update_fn = wrapped
    with NoiseBank(nb):
        out = S.sample(u0.clone(), ts, n, denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
    _save(fname, final=out.double().numpy(), stats=np.asarray(stats), ts=ts.numpy(),
          n=np.asarray(n), seeds=np.asarray([seed_w, seed_p, seed_n]), B=np.asarray(B),
          probe=np.asarray(probe), **states)


def golden_inpaint(R, cfg, fname, B, seed_w=0, seed_p=1, seed_n=2, seed_x=3, score="fake"):
    sde = R.PSLD(cfg)
    H = cfg.data.image_size
    net = fake_score if score == "fake" else ref_net(R, cfg, seed_w)[0]
    ts, n = reference_time_grid(cfg)
    d = inpaint_draws(n, B, H, seed_n)
    u0 = prior((B, 3, H, H), float(np.sqrt(sde.m)), seed_p)
    x_0, mask = inpaint_inputs(B, H, seed_x)
    order = [d["m0"][0], d["eps"][0]]
    for i in range(n + 1):
        order += [d["pred"][i], d["m0"][i + 1], d["eps"][i + 1]]
    S = R.get_module("samplers", "ip_em_sde")(cfg, sde, net)
    sde.prior_sampling = lambda shape: u0.clone()    # the only non-randn_like draw (psld.py:366-370)
    with NoiseBank(order):
        out = S.sample((x_0, mask), ts, n, denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
    _save(fname, final=out.double().numpy(), ts=ts.numpy(), n=np.asarray(n), B=np.asarray(B),
          seeds=np.asarray([seed_w, seed_p, seed_n, seed_x]))


def fake_score(u, t):
    """Deterministic smooth stand-in score network (tests the sampler algebra alone)."""
    return (torch.tanh(u * 0.3) * 0.7 + 0.1 * torch.sin(torch.roll(u, 1, 1))) * t.view(-1, 1, 1, 1)


def golden_full_size(R):
    """BASELINE.json configs[1] / configs[3] at full architecture size (``init_scale=1`` so that the
    network is numerically visible): the SSCS trajectory of the benchmarked
    configuration (B=2, 50 NFE, per-step probe states) and forwards at B=3 / B=2 over t in
    {1, 0.5, 1e-3}."""
    c = cifar10_config(n_discrete_steps=50, batch_size=2, n_samples=2)
    c.model.score_fn.init_scale = 1.0
    golden_sampler(R, c, "sampler_cifar10_sscs50.npz", B=2, keep=2)
    golden_forward(R, c, "forward_cifar10_b3.npz", B=3, times=(1.0, 0.5, 1e-3))
    c = celeba64_config(n_discrete_steps=20, batch_size=2, n_samples=2)
    c.model.score_fn.init_scale = 1.0
    golden_forward(R, c, "forward_celeba64_b2.npz", B=2, times=(1.0, 1e-3))
    golden_sampler(R, c, "sampler_celeba64_sscs20.npz", B=2, keep=1)


def golden_full_length(R):
    """BASELINE.json configs[1] at FULL length: the benchmarked CIFAR-10 NCSN++ through all 1000 SSCS steps of
    the unmodified reference (B=2, init_scale=1, pre-drawn noise; ~6 min of CPU): per-step energies, probe
    states and the final samples, i.e. the trajectory the bench times."""
    c = cifar10_config(n_discrete_steps=1000, batch_size=2, n_samples=2)
    c.model.score_fn.init_scale = 1.0
    if "--celeba" not in sys.argv:
        golden_sampler(R, c, "sampler_cifar10_sscs1000.npz", B=2, keep=2)
        return
    # configs[3] at full length (CelebA-64 NCSN++, B=1; ~25 min of CPU): `--only-full-length --celeba`
    c = celeba64_config(n_discrete_steps=1000, batch_size=1, n_samples=1)
    c.model.score_fn.init_scale = 1.0
    golden_sampler(R, c, "sampler_celeba64_sscs1000.npz", B=1, keep=1)


def golden_state_dict_contract(R):
    """Names and shapes of the REFERENCE module's state dict for the shipped architectures: the
    checkpoint contract (``ema_score_fn.all_modules.<i>...``, wrapper.py:30-31) the drop-in must keep."""
    import json
    out = {}
    for name, mk in (("tiny", tiny_config), ("cifar10", cifar10_config), ("celeba64", celeba64_config)):
        net = R.NCSNpp(mk())
        out[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
        print(name, len(out[name]), "tensors", sum(int(np.prod(v)) for v in out[name].values()), "elements")
    with open(os.path.join(OUT, "state_dict_contract.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


def main():
    torch.set_num_threads(os.cpu_count())
    R = load_reference()
    if "--only-full-size" in sys.argv:
        return golden_full_size(R)
    if "--only-full-length" in sys.argv:
        return golden_full_length(R)
    if "--only-contract" in sys.argv:
        return golden_state_dict_contract(R)
    if "--only-cc" in sys.argv:
        return golden_cc(R)
    if "--only-ode" in sys.argv:
        return golden_ode(R)
    if "--only-ode-vp" in sys.argv:
        return golden_ode_vp(R)
    if "--only-guidance" in sys.argv:
        return golden_guidance(R)
    if "--only-inpaint" in sys.argv:
        return golden_inpaint_all(R)
    if "--only-vp" in sys.argv:
        return golden_vp(R)
    golden_scalars(R)
    golden_upfirdn(R)
    golden_modules(R)
    golden_forward(R, tiny_config(), "forward_tiny.npz", B=4)
    golden_forward(R, mid_config(), "forward_mid.npz", B=2)
    c = cifar10_config(); c.model.score_fn.init_scale = 1.0
    golden_forward(R, c, "forward_cifar10.npz", B=1)
    c = celeba64_config(); c.model.score_fn.init_scale = 1.0
    golden_forward(R, c, "forward_celeba64.npz", B=1)
    # BASELINE.json configs[0]: tiny net, EM, 100 steps, 8 samples
    golden_sampler(R, tiny_config(), "sampler_tiny_em100.npz", B=8)
    golden_sampler(R, tiny_config(sampler="sscs_sde"), "sampler_tiny_sscs100.npz", B=8)
    # sampler algebra alone (fake score), incl. varying beta, quadratic stride, score_m mode
    for tag, kw, sd in [
        ("sscs_fake_uniform", dict(sampler="sscs_sde", n_discrete_steps=50), {}),
        ("em_fake_uniform", dict(sampler="em_sde", n_discrete_steps=50), {}),
        ("sscs_fake_quad", dict(sampler="sscs_sde", n_discrete_steps=40, stride_type="quadratic"),
         dict(nu=4.02, gamma=0.02, beta_min=0.5, beta_max=12.0)),
        ("em_fake_quad", dict(sampler="em_sde", n_discrete_steps=40, stride_type="quadratic"),
         dict(nu=4.02, gamma=0.02, beta_min=0.5, beta_max=12.0)),
        ("sscs_fake_nodenoise", dict(sampler="sscs_sde", n_discrete_steps=30, denoise=False), {}),
    ]:
        cfg = tiny_config(**kw)
        cfg.model.sde.update(sd)
        cfg.data.image_size = 8
        golden_sampler(R, cfg, f"sampler_{tag}.npz", B=3, keep=3, score="fake")
    golden_inpaint_all(R)
    golden_vp(R)
    golden_full_size(R)
    golden_state_dict_contract(R)
    golden_cc(R)
    golden_ode(R)
    golden_ode_vp(R)
    golden_guidance(R)


def vp_config(**ev):
    e = dict(sampler="em_sde", n_discrete_steps=40)
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.model.sde.update(name="vpsde", beta_min=0.1, beta_max=20.0)
    cfg.model.score_fn.update(in_ch=3, out_ch=3)
    cfg.data.image_size = 8
    return cfg


def golden_vp(R):
    """EM on the VP-SDE baseline (sample_uncond_vpsde.sh): sampler algebra with the stand-in score."""
    for tag, kw in [("uniform", {}), ("quad", dict(stride_type="quadratic"))]:
        cfg = vp_config(**kw)
        sde = R.get_module("sde", "vpsde")(cfg)
        ts, n = reference_time_grid(cfg)
        B = 3
        nb = noise_bank(n, (B, 3, 8, 8), 2)
        x0 = noise_bank(1, (B, 3, 8, 8), 1)[0]
        S = R.get_module("samplers", "em_sde")(cfg, sde, fake_score)
        with NoiseBank(nb):
            out = S.sample(x0.clone(), ts, n, denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
        _save(f"sampler_vp_em_fake_{tag}.npz", final=out.double().numpy(), ts=ts.numpy(), n=np.asarray(n),
              B=np.asarray(B))


def cc_config(**ev):
    e = dict(sampler="cc_em_sde", n_discrete_steps=40)
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.data.image_size = 8
    cfg["clf"] = dict(evaluation=dict(label_to_sample=3, clf_temp=2.5))
    from psld_b200.config import Cfg
    return Cfg(cfg)


def golden_cc(R):
    """Classifier-guided EM (cc_em_sde, sde.py:61-122): sampler algebra with the stand-in score network
    and a stand-in differentiable classifier, uniform and quadratic stride."""
    from oracle.weights import fake_classifier
    for tag, kw in [("uniform", {}), ("quad_nodenoise", dict(stride_type="quadratic", denoise=False))]:
        cfg = cc_config(**kw)
        sde = R.PSLD(cfg)
        ts, n = reference_time_grid(cfg)
        B = 3
        nb = noise_bank(n, (B, 6, 8, 8), 2)
        u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
        S = R.get_module("samplers", "cc_em_sde")(cfg, sde, fake_score, fake_classifier)
        with NoiseBank(nb):
            out = S.sample(u0.clone(), ts, n, denoise=cfg.evaluation.denoise, eps=cfg.evaluation.eval_eps)
        _save(f"sampler_cc_em_fake_{tag}.npz", final=out.double().numpy(), ts=ts.numpy(), n=np.asarray(n),
              B=np.asarray(B))


def ode_config(**ev):
    e = dict(sampler=dict(name="bb_ode", solver="RK45", rtol=1e-5, atol=1e-5))
    e.update(ev)
    cfg = tiny_config(**e)
    cfg.data.image_size = 8
    return cfg


def golden_ode(R):
    """bb_ode (ode.py:41-76) through the restated torchdiffeq scipy_solver wrapper (ref_loader.py): exact
    Gaussian-data score (stable reverse dynamics), the tolerances of the shipped scripts (1e-5 CelebA /
    ablations, 1e-4 CIFAR)."""
    from oracle.weights import gaussian_score_fn
    for tag, tol, den in [("tol1e-5", 1e-5, True), ("tol1e-4_nodenoise", 1e-4, False)]:
        cfg = ode_config(sampler=dict(name="bb_ode", solver="RK45", rtol=tol, atol=tol), denoise=den)
        sde = R.PSLD(cfg)
        B = 3
        u0 = prior((B, 3, 8, 8), float(np.sqrt(sde.m)), 1)
        S = R.get_module("samplers", "bb_ode")(cfg, sde, gaussian_score_fn(cfg))
        out = S.sample(u0.clone(), None, 0, denoise=den, eps=cfg.evaluation.eval_eps)
        print(tag, "nfe", S.nfe, out.dtype, "std of x", float(out[:, :3].std()))
        _save(f"sampler_bb_ode_gauss_{tag}.npz", final=out.double().numpy(), nfe=np.asarray(S.nfe),
              B=np.asarray(B), tol=np.asarray(tol))


def golden_ode_vp(R):
    """bb_ode over the VP-SDE baseline (scripts_psld/ablations/uncond/cifar10/sample_uncond_vpsde_ode.sh):
    the reference's own BBODESampler + VPSDE, exact Gaussian-data score, float32 and float64 batches."""
    from oracle.weights import vp_gaussian_score_fn
    for tag, tol, den, dt in [("tol1e-5", 1e-5, True, torch.float32), ("tol1e-4_f64_nodenoise", 1e-4, False, torch.float64)]:
        cfg = vp_config(sampler=dict(name="bb_ode", solver="RK45", rtol=tol, atol=tol), denoise=den)
        sde = R.get_module("sde", "vpsde")(cfg)
        B = 3
        x0 = prior((B, 3, 8, 8), 1.0, 1)[:, :3].contiguous().to(dt)
        S = R.get_module("samplers", "bb_ode")(cfg, sde, vp_gaussian_score_fn(cfg))
        out = S.sample(x0.clone(), None, 0, denoise=den, eps=cfg.evaluation.eval_eps)
        print("vp", tag, "nfe", S.nfe, out.dtype, "std of x", float(out.std()))
        _save(f"sampler_bb_ode_vp_gauss_{tag}.npz", final=out.double().numpy(), nfe=np.asarray(S.nfe),
              B=np.asarray(B), tol=np.asarray(tol))


def golden_inpaint_all(R):
    # inpainting sampler (ip_em_sde), HSM and DSM perturbation, sampler algebra alone
    for mode in ("hsm", "dsm"):
        cfg = tiny_config(sampler="ip_em_sde", n_discrete_steps=30)
        cfg.training.mode = mode
        cfg.data.image_size = 8
        golden_inpaint(R, cfg, f"sampler_ip_em_fake_{mode}.npz", B=3)




# ---------------------------------------------------------------- classifier-free guidance (configs[4])
CFG_WEIGHT = 1.5


def guided(net_c, net_u, w):
    """eps = (1 + w) eps_c - w eps_u in float32, this order (the reference has no CFG sampler: the
    definition is this composition of two UNMODIFIED reference networks at its score_fn call site)."""
    a, b = torch.tensor(1.0 + w, dtype=torch.float32), torch.tensor(-w, dtype=torch.float32)

    def fn(u, t):
        with torch.no_grad():
            return a * net_c(u, t) + b * net_u(u, t)
    return fn


def golden_guidance(R):
    # forward: two reference NCSN++ (weight seeds 0 / 1) of the tcgen05-eligible mid config
    cfg = mid_config()
    fc, fu = ref_net(R, cfg, 0)[0], ref_net(R, cfg, 1)[0]
    r = np.random.default_rng([11, 7])
    x = torch.from_numpy((r.standard_normal((2, 6, 32, 32)) * 1.5).astype(np.float32))
    t = torch.from_numpy(np.asarray([0.731, 0.0123], dtype=np.float32))
    y = guided(fc, fu, CFG_WEIGHT)(x, t)
    with torch.no_grad():
        yc = fc(x, t)
    _save("forward_cfg_mid.npz", x=x.numpy(), t=t.numpy(), y=y.numpy(), y_cond=yc.numpy(),
          seeds=np.asarray([0, 1]), weight=np.asarray(CFG_WEIGHT))
    if "--full-size" in sys.argv:
        # configs[4] at full architecture size: two CIFAR-10 NCSN++ (init_scale = 1, weight seeds 0 / 1), SSCS,
        # 100 NFE, B = 1 through the reference's own sampler (~3 min of CPU): `--only-guidance --full-size`
        c = cifar10_config(n_discrete_steps=100, batch_size=1, n_samples=1)
        c.model.score_fn.init_scale = 1.0
        fc, fu = ref_net(R, c, 0)[0], ref_net(R, c, 1)[0]
        golden_sampler(R, c, "sampler_cfg_cifar10_sscs100.npz", B=1, keep=1, score=guided(fc, fu, CFG_WEIGHT))
        return
    # trajectory: the reference's own SSCS / EM samplers driven by the guided score_fn
    for name in ("sscs_sde", "em_sde"):
        cfg = tiny_config(sampler=name, n_discrete_steps=40)
        fc, fu = ref_net(R, cfg, 0)[0], ref_net(R, cfg, 1)[0]
        golden_sampler(R, cfg, f"sampler_cfg_tiny_{name.split('_')[0]}40.npz", B=4,
                       score=guided(fc, fu, CFG_WEIGHT))


if __name__ == "__main__":
    main()
