"""The CPU oracle (oracle/vaegan.py) against fixtures produced by the UNMODIFIED reference modules
(tests/golden/*.npz, written by oracle/make_golden.py in the build container).

Tolerances: the fixtures are fp64 runs of the reference; the oracle in fp64 must agree to 1e-9 relative (same
arithmetic, different operator grouping only); in fp32 the forward tensors / losses must agree to 1e-4 (the fp32 noise
floor of the reference itself is ~1e-6 forward, ~1e-3 on end-to-end gradients through ReLU-mask flips).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vaegan as O
from oracle.golden_util import summarize, summary_error

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _cases(kind):
    fs = sorted(glob.glob(os.path.join(GOLD, f"{kind}_B*_s*.npz")))
    assert fs, f"no golden fixtures for {kind}"
    return fs


@pytest.mark.parametrize("path", _cases("stage1_vaegan") + _cases("stage1_vaegan100") + _cases("stage1_betavae")
                         + _cases("stage1_dcgan") + _cases("stage1_vae"))
def test_stage1_vaegan_fp64_matches_reference(path):
    """64x64 fixtures and the reference's ACTIVE 100x100 / latent-512 configuration (configs/models_config.py:13-21:
    stride-2 first discriminator conv, output_padding [False, True, True], odd 13/25/50-pixel feature maps)."""
    g = np.load(path)
    B, seed = int(g["B"]), int(g["seed"])
    cfg = O.CFG100 if "vaegan100" in os.path.basename(path) else O.CFG64
    z = cfg["latent_dim"]
    P, S = O.make_vaegan(cfg, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, size=cfg["image_size"], seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, z, seed=seed)]
    kw = dict(mode="beta-vae", beta=float(g["beta"])) if "betavae" in os.path.basename(path) else {}   # :359-365
    if "mode" in g.files:                                                     # 'dcgan' / 'vae' pixel-NLE mixes, :374-387
        kw = dict(mode=str(g["mode"]))
    out = O.stage1_vaegan_step(P, S, x, eps, z_p, cfg=cfg, **kw)
    assert out["train_dis"] == bool(g["train_dis"]) and out["train_dec"] == bool(g["train_dec"])
    for k in ("mu", "logvar", "kl", "mse", "bce_o", "bce_p", "bce_s", "disc_class", "loss_encoder", "loss_decoder",
              "loss_discriminator"):
        assert _rel(out[k].numpy(), g[k]) < 1e-9, k
    assert _rel(out["nle"].sum().numpy(), g["nle_sum"]) < 1e-9
    assert summary_error(summarize(out["x_tilde"]), g["x_tilde"]) < 1e-9
    assert summary_error(summarize(out["disc_layer"]), g["disc_layer"]) < 1e-9
    n = 0
    for k in g.files:
        if k.startswith("grad:"):
            assert summary_error(summarize(out["grads"][k[5:]]), g[k]) < 1e-8, k
            n += 1
        elif k.startswith("delta:"):
            name = k[6:]
            assert summary_error(summarize(out["params"][name] - P[name]), g[k]) < 1e-6, k
        elif k.startswith("buf:"):
            assert summary_error(summarize(S[k[4:]]), g[k]) < 1e-9, k
    n_expected = len(P) - (sum(k.startswith("encoder.") for k in P) if kw.get("mode") == "dcgan" else 0)  # :375-377
    assert n == n_expected
    # inference path (VaeGan.forward in eval mode, models/vae_gan.py:288-295): updated weights, the running statistics the
    # step produced, eval-mode BatchNorm
    P2 = out["params"]
    mu_e, lv_e = O.encoder(P2, S, x, cfg, train=False)
    x_e = O.decoder(P2, S, O.reparameterize(mu_e, lv_e, eps), cfg, train=False)
    assert summary_error(summarize(x_e), g["eval_x_tilde"]) < 1e-9


@pytest.mark.parametrize("path", _cases("stage1_vaegan")[:1])
def test_stage1_vaegan_fp32_forward_close(path):
    g = np.load(path)
    B, seed = int(g["B"]), int(g["seed"])
    P, S = O.make_vaegan(O.CFG64, seed=seed, dtype=torch.float32)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    out = O.stage1_vaegan_step(P, S, x, eps, z_p, update=False)
    for k in ("mu", "logvar", "kl", "mse", "bce_o", "bce_p", "bce_s", "disc_class", "loss_encoder",
              "loss_discriminator"):
        assert _rel(out[k].numpy(), g[k]) < 1e-4, k


@pytest.mark.parametrize("path", _cases("stage1_waegan"))
def test_stage1_waegan_fp64_matches_reference(path):
    g = np.load(path)
    B, seed = int(g["B"]), int(g["seed"])
    P, S = O.make_waegan(O.CFG64, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, seed=seed).double()
    z_fake = (O.synthetic_noise(B, 128, seed=seed)[0] * 0.5).double()
    out = O.stage1_waegan_step(P, S, x, z_fake)
    for k in ("z_real", "d_real", "d_fake", "d_real_g", "loss_discriminator_fake", "loss_discriminator_real",
              "loss_reconstruction", "loss_penalty"):
        assert _rel(out[k].numpy(), g[k]) < 1e-9, k
    assert summary_error(summarize(out["x_recon"]), g["x_recon"]) < 1e-9
    n = 0
    for k in g.files:
        if k.startswith("grad:"):
            assert summary_error(summarize(out["grads"][k[5:]]), g[k]) < 1e-8, k
            n += 1
        elif k.startswith("delta:"):
            name = k[6:]
            assert summary_error(summarize(out["params"][name] - P[name]), g[k]) < 1e-6, k
        elif k.startswith("buf:"):
            assert summary_error(summarize(S[k[4:]]), g[k]) < 1e-9, k
    assert n == len(out["grads"]) == len(P) - 2  # l_var.{weight,bias} receive no gradient in the WAE


@pytest.mark.parametrize("stage", [2, 3])
def test_cognitive_stages_fp64_match_reference(stage):
    g = np.load(os.path.join(GOLD, f"stage{stage}_cognitive_B4_s4711.npz"))
    B, seed = int(g["B"]), int(g["seed"])
    P, S = O.make_cognitive(O.CFG64, seed=seed, dtype=torch.float64)
    fmri = O.synthetic_fmri(B, seed=seed).double()
    image = O.synthetic_images(B, seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, 128, seed=seed)]
    eps_t = O.synthetic_noise(B, 128, seed=seed + 1)[0].double()
    out = O.cognitive_vaegan_step(P, S, fmri, image, eps, eps_t, z_p, stage)
    assert out["train_dis"] == bool(g["train_dis"]) and out["train_dec"] == bool(g["train_dec"])
    for k in ("mu", "logvar", "kl", "mse", "bce_o", "bce_p", "bce_s", "disc_class", "loss_encoder", "loss_decoder",
              "loss_discriminator"):
        assert _rel(out[k].numpy(), g[k]) < 1e-9, k
    for k in ("x_tilde", "gt_x", "disc_layer"):
        assert summary_error(summarize(out[k]), g[k]) < 1e-9, k
    n = 0
    for k in g.files:
        if k.startswith("grad:"):
            assert summary_error(summarize(out["grads"][k[5:]]), g[k]) < 1e-8, k
            n += 1
        elif k.startswith("delta:"):
            assert summary_error(summarize(out["params"][k[6:]] - P[k[6:]]), g[k]) < 1e-6, k
        elif k.startswith("buf:"):
            assert summary_error(summarize(S[k[4:]]), g[k]) < 1e-9, k
    assert n == len(out["grads"]) > 0


@pytest.mark.parametrize("stage", [2, 3])
def test_cognitive_wae_stages_fp64_match_reference(stage):
    g = np.load(os.path.join(GOLD, f"stage{stage}_cognitive_wae_B4_s3131.npz"))
    B, seed = int(g["B"]), int(g["seed"])
    P, S = O.make_cognitive_wae(O.CFG64, seed=seed, dtype=torch.float64)
    fmri = O.synthetic_fmri(B, seed=seed).double()
    image = O.synthetic_images(B, seed=seed).double()
    out = O.cognitive_wae_step(P, S, fmri, image, stage)
    for k in ("z_fake", "z_real", "d_real", "d_fake", "d_real_g", "loss_discriminator_fake", "loss_discriminator_real",
              "loss_reconstruction", "loss_penalty"):
        assert _rel(out[k].numpy(), g[k]) < 1e-9, k
    assert summary_error(summarize(out["x_recon"]), g["x_recon"]) < 1e-9
    n = 0
    for k in g.files:
        if k.startswith("grad:"):
            assert summary_error(summarize(out["grads"][k[5:]]), g[k]) < 1e-8, k
            n += 1
        elif k.startswith("delta:"):
            assert summary_error(summarize(out["params"][k[6:]] - P[k[6:]]), g[k]) < 1e-6, k
        elif k.startswith("buf:"):
            assert summary_error(summarize(S[k[4:]]), g[k]) < 1e-9, k
    assert n == len(out["grads"]) > 0


def test_gate_table():
    # train/train_vgan_stage1.py:396-404
    assert O.gate(0.5, 0.7) == (True, True)
    assert O.gate(0.2, 0.7) == (False, True)
    assert O.gate(0.5, 1.2) == (True, False)
    assert O.gate(0.2, 1.2) == (True, True)


def test_optimizer_restatements_match_torch_optim():
    g0 = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g0, dtype=torch.float64)
    grads = [torch.randn(1000, generator=g0, dtype=torch.float64) for _ in range(3)]
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.RMSprop([p], lr=1e-3, alpha=0.9, eps=1e-8)
    q, sq = p0.clone(), torch.zeros_like(p0)
    for g in grads:
        p.grad = g.clone()
        opt.step()
        q, sq = O.rmsprop_update(q, g, sq, 1e-3)
    assert _rel(q.numpy(), p.detach().numpy()) < 1e-12
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.5, 0.999))
    q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for t, g in enumerate(grads, 1):
        p.grad = g.clone()
        opt.step()
        q, m, v = O.adam_update(q, g, m, v, t, 1e-3)
    assert _rel(q.numpy(), p.detach().numpy()) < 1e-12


def test_dual_stage3_composite_equals_its_pinned_halves():
    """BASELINE.json configs[3] is a composite, so the reference holds no single fixture for it; its two
    halves are the golden-pinned steps above. The composite must reproduce each half exactly on shared weights."""
    B, seed = 4, 77
    Pw, Sw = O.make_cognitive_wae(O.CFG64, seed=seed, dtype=torch.float64)
    Pd, Sd = O.make_dual_stage3(O.CFG64, seed=seed, dtype=torch.float64)
    for k in Pw:   # share the cognitive encoder, the teacher encoder and the latent discriminator
        if k.startswith("encoder.") or k.startswith("teacher_net.encoder."):
            Pd[k] = Pw[k].clone()
        elif k.startswith("discriminator."):
            Pd["latent_" + k] = Pw[k].clone()
    fmri, image = O.synthetic_fmri(B, seed=seed).double(), O.synthetic_images(B, seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, 128, seed=seed)]
    wae = O.cognitive_wae_step(Pw, Sw, fmri, image, 3)
    Sc = {k: v.clone() for k, v in Sd.items()}
    img = O.cognitive_vaegan_step(Pd, Sc, fmri, image, eps, None, z_p, 3)
    dual = O.dual_stage3_step(Pd, Sd, fmri, image, eps, z_p)
    for k in ("z_real", "d_real", "d_fake", "loss_discriminator_fake", "loss_discriminator_real"):
        assert _rel(dual[k].numpy(), wae[k].numpy()) < 1e-12, k
    for k, g in wae["grads"].items():
        if k.startswith("discriminator."):
            assert _rel(dual["grads"]["latent_" + k].numpy(), g.numpy()) < 1e-12, k
            assert _rel(dual["params"]["latent_" + k].numpy(), wae["params"][k].numpy()) < 1e-12, k
    for k in ("x_tilde", "disc_layer", "disc_class", "loss_decoder", "loss_discriminator"):
        assert _rel(dual[k].numpy(), img[k].numpy()) < 1e-12, k
    for k, g in img["grads"].items():
        assert _rel(dual["grads"][k].numpy(), g.numpy()) < 1e-12, k
        assert _rel(dual["params"][k].numpy(), img["params"][k].numpy()) < 1e-12, k


def test_mmd_statement_known_answers():
    """oracle/mmd.py (parity unpinned: no reference MMD code). Closed form at B=2 and basic estimator properties."""
    from oracle.mmd import SCALES, mmd_imq

    q = torch.tensor([[0.0, 0.0], [1.0, 0.0]], dtype=torch.float64)
    p = torch.tensor([[0.0, 1.0], [1.0, 1.0]], dtype=torch.float64)
    # d_qq = d_pp = 1 (off-diagonal), d_qp = [[1, 2], [2, 1]], Z = 2, sigma2 = 1 -> C_s = 4 s
    want = sum(2 * (4 * s / (4 * s + 1)) * 2 / 2 - 2 * (2 * 4 * s / (4 * s + 1) + 2 * 4 * s / (4 * s + 2)) / 4 for s in SCALES)
    assert abs(mmd_imq(q, p, 1.0).item() - want) < 1e-12
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(256, 8, generator=g, dtype=torch.float64), torch.randn(256, 8, generator=g, dtype=torch.float64)
    same, shifted = mmd_imq(a, b, 1.0).item(), mmd_imq(a + 2.0, b, 1.0).item()
    assert abs(same) < 0.05 and shifted > 0.5
    assert abs(mmd_imq(a, b, 1.0).item() - mmd_imq(b, a, 1.0).item()) < 1e-12   # symmetric in its arguments


def test_dual_stage1_fp64_matches_reference():
    """train/wae_vgan_stage1.py:282-441, torch >= 2 semantics: the reference's VaeGan + a second
    WaeGan's latent discriminator run by oracle/make_golden.py against oracle.dual_stage1_step."""
    g = np.load(os.path.join(GOLD, "stage1_dual_B4_s606.npz"))
    B, seed, lam = int(g["B"]), int(g["seed"]), float(g["lam"])
    P, S = O.make_dual_stage1(O.CFG64, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, 128, seed=seed)]
    z_fake = (O.synthetic_noise(B, 128, seed=seed + 7)[0] * 0.5).double()
    out = O.dual_stage1_step(P, S, x, eps, z_p, z_fake, lam=lam)
    assert out["train_dis"] == bool(g["train_dis"]) and out["train_dec"] == bool(g["train_dec"])
    for k in ("mu", "kl", "mse", "loss_encoder", "loss_decoder", "loss_discriminator", "d_real", "d_fake", "d_real_g",
              "loss_discriminator_fake", "loss_discriminator_real", "loss_penalty"):
        assert _rel(out[k].numpy(), g[k]) < 1e-9, k
    assert summary_error(summarize(out["x_tilde"]), g["x_tilde"]) < 1e-9
    n = 0
    for k in g.files:
        if k.startswith("grad:"):
            assert summary_error(summarize(out["grads"][k[5:]]), g[k]) < 1e-8, k
            n += 1
        elif k.startswith("delta:"):
            assert summary_error(summarize(out["params"][k[6:]] - P[k[6:]]), g[k]) < 1e-6, k
        elif k.startswith("buf:"):
            assert summary_error(summarize(S[k[4:]]), g[k]) < 1e-9, k
    assert n == len(P)     # every parameter of the four networks receives a gradient (l_var through the KL term)
