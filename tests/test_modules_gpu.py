"""The nn.Module drop-in surface (models/vae_gan.py) driven exactly like the reference's training scripts, against the
CPU oracle: strict state_dict loading with the reference's key names, the Stage-I VAE/GAN script sequence INCLUDING the
interleaved backward(retain_graph=True) / optimizer.step() order of train_vgan_stage1.py:408-432 (which stock torch >= 1.5
rejects on the reference's own modules), the WAE Stage-I script sequence (train_wae_stage1.py:263-311), and eval mode.

Tolerances (rel-L2): fp32 exact path -- forward 1e-4, gradient buckets max(5e-3, 3 x the oracle's fp32-vs-fp64 noise);
bf16 tensor path -- forward 2e-2, gradient buckets reported and bounded at 0.5 (ReLU-mask flips).
"""
import contextlib

import pytest
import torch

import configs.models_config as mc
from oracle import vaegan as O
from thesis_fmri_reconstruction_b200 import autograd as ag

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@contextlib.contextmanager
def patched_randn(value):
    orig = torch.randn
    torch.randn = lambda *a, **k: value.clone()
    try:
        yield
    finally:
        torch.randn = orig


@contextlib.contextmanager
def compute(dtype):
    old = ag.compute_dtype()
    ag.set_compute_dtype(dtype)
    try:
        yield
    finally:
        ag.set_compute_dtype(old)


def bucket_err(grads, ref, pre):
    a = torch.cat([grads[k].reshape(-1).cpu() for k in ref if k.startswith(pre)])
    r = torch.cat([ref[k].reshape(-1) for k in ref if k.startswith(pre)])
    return rel(a, r)


def build_vaegan(P, S):
    mc.use_resolution(64)
    from models.vae_gan import VaeGan

    model = VaeGan(device="cuda", z_size=128)
    missing, unexpected = model.load_state_dict({**P, **S}, strict=True)   # the reference's key names, strictly
    assert not missing and not unexpected
    return model


@pytest.mark.parametrize("dtype,B", [(torch.float32, 8), (torch.bfloat16, 16)])
def test_stage1_script_sequence(dtype, B):
    from models.vae_gan import VaeGan

    seed = 31
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_vaegan_step(P, S_ref, x, eps, z_p)
    ref64 = O.stage1_vaegan_step({k: v.double() for k, v in P.items()},
                                 {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in S.items()},
                                 x.double(), eps.double(), z_p.double(), update=False)
    hp = O.HP_VGAN
    with compute(dtype):
        model = build_vaegan(P, S)
        model.train()
        eps_d = eps.cuda()
        model.reparameterize = lambda mu, lv: ag.reparameterize(mu, lv, eps_d)
        opt = {b: torch.optim.RMSprop(getattr(model, b).parameters(), lr=hp["lr"], alpha=0.9, eps=1e-8)
               for b in ("encoder", "decoder", "discriminator")}
        with patched_randn(z_p):
            x_tilde, disc_class, disc_layer, mus, lv = model(x)                      # train_vgan_stage1.py:330
        dl_o, dl_p, dl_s = disc_layer[:B], disc_layer[B:-B], disc_layer[-B:]
        dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
        nle, kld, mse, bo, bp, bs = VaeGan.loss(x.cuda(), x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, lv)
        loss_encoder = torch.sum(kld) + torch.sum(mse)
        loss_discriminator = torch.sum(bo) + torch.sum(bp) + torch.sum(bs)
        loss_decoder = torch.sum(hp["lambda_mse"] * mse) - (1.0 - hp["lambda_mse"]) * loss_discriminator
        grads = {}
        # the script's own order, optimizer steps interleaved (train_vgan_stage1.py:408-432)
        model.zero_grad()
        loss_encoder.backward(retain_graph=True)
        grads.update({"encoder." + k: p.grad.clone() for k, p in model.encoder.named_parameters()})
        opt["encoder"].step()
        model.zero_grad()
        loss_decoder.backward(retain_graph=True)
        grads.update({"decoder." + k: p.grad.clone() for k, p in model.decoder.named_parameters()})
        opt["decoder"].step()
        model.discriminator.zero_grad()
        loss_discriminator.backward()
        grads.update({"discriminator." + k: p.grad.clone() for k, p in model.discriminator.named_parameters()})
        opt["discriminator"].step()
        torch.cuda.synchronize()
    ftol = 1e-4 if dtype == torch.float32 else 2e-2
    fwd = dict(x_tilde=rel(x_tilde, ref["x_tilde"]), disc_layer=rel(disc_layer, ref["disc_layer"]),
               disc_class=rel(disc_class, ref["disc_class"]), mu=rel(mus, ref["mu"]), logvar=rel(lv, ref["logvar"]),
               kl=rel(kld, ref["kl"]), mse=rel(mse, ref["mse"]), bce_o=rel(bo, ref["bce_o"]),
               loss_encoder=rel(loss_encoder, ref["loss_encoder"]), loss_decoder=rel(loss_decoder, ref["loss_decoder"]))
    print(dtype, "forward", fwd)
    assert max(fwd.values()) < ftol, fwd
    for pre in ("encoder.", "decoder.", "discriminator."):
        e = bucket_err(grads, ref64["grads"], pre)
        noise = bucket_err(ref["grads"], ref64["grads"], pre)
        print(dtype, pre, "grad rel-L2 vs fp64 oracle", e, "oracle fp32 noise", noise)
        assert e < (max(5e-3, 3 * noise) if dtype == torch.float32 else 0.5), (pre, e, noise)
    sd = model.state_dict()
    for k, v in S_ref.items():   # BN running statistics incl. the discriminator's double update, num_batches_tracked
        if v.dtype.is_floating_point:
            assert rel(sd[k], v) < ftol, k
        else:
            assert int(sd[k]) == int(v), k


@pytest.mark.parametrize("dtype,B", [(torch.float32, 8), (torch.bfloat16, 16)])
def test_wae_stage1_script_sequence(dtype, B):
    mc.use_resolution(64)
    from models.vae_gan import WaeGan

    seed = 57
    P, S = O.make_waegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    z_fake = O.synthetic_noise(B, 128, seed=seed)[0] * 0.5
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_waegan_step(P, S_ref, x, z_fake)
    hp = O.HP_WAE

    def freeze(m, on):
        for p in m.parameters():
            p.requires_grad = not on

    with compute(dtype):
        model = WaeGan(device="cuda", z_size=128)
        model.load_state_dict({**P, **S}, strict=True)
        model.train()
        xd, zf = x.cuda(), z_fake.cuda()
        opt_e = torch.optim.Adam(model.encoder.parameters(), lr=hp["lr"], betas=(0.5, 0.999))
        opt_d = torch.optim.Adam(model.decoder.parameters(), lr=hp["lr"], betas=(0.5, 0.999))
        opt_c = torch.optim.Adam(model.discriminator.parameters(), lr=0.5 * hp["lr"], betas=(0.5, 0.999))
        model.encoder.zero_grad(); model.decoder.zero_grad(); model.discriminator.zero_grad()
        freeze(model.decoder, True); freeze(model.encoder, True); freeze(model.discriminator, False)
        z_real, _ = model.encoder(xd)
        d_real = model.discriminator(z_real)
        d_fake = model.discriminator(zf)
        loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))
        loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))
        loss_fake.backward(retain_graph=True)
        loss_real.backward(retain_graph=True)
        grads = {"discriminator." + k: p.grad.clone() for k, p in model.discriminator.named_parameters()}
        opt_c.step()
        freeze(model.encoder, False); freeze(model.decoder, False); freeze(model.discriminator, True)
        z_real2, _ = model.encoder(xd)
        x_recon = model.decoder(z_real2)
        d_real2 = model.discriminator(z_real2)
        loss_rec = torch.sum(torch.sum(0.5 * (x_recon - xd) ** 2, 1))
        loss_pen = -10 * torch.sum(torch.log(d_real2 + 1e-3))
        loss_rec.backward(retain_graph=True)
        loss_pen.backward()
        assert model.encoder.l_var.weight.grad is None          # logvar is unused: Adam must skip it as in the reference
        grads.update({"encoder." + k: p.grad.clone() for k, p in model.encoder.named_parameters() if p.grad is not None})
        grads.update({"decoder." + k: p.grad.clone() for k, p in model.decoder.named_parameters()})
        opt_e.step(); opt_d.step()
        torch.cuda.synchronize()
    ftol = 1e-4 if dtype == torch.float32 else 2e-2
    fwd = dict(z_real=rel(z_real, ref["z_real"]), d_real=rel(d_real, ref["d_real"]), d_fake=rel(d_fake, ref["d_fake"]),
               x_recon=rel(x_recon, ref["x_recon"]), loss_rec=rel(loss_rec, ref["loss_reconstruction"]),
               loss_pen=rel(loss_pen, ref["loss_penalty"]), loss_fake=rel(loss_fake, ref["loss_discriminator_fake"]))
    print(dtype, "forward", fwd)
    assert max(fwd.values()) < ftol, fwd
    for pre in ("encoder.", "decoder.", "discriminator."):
        e = bucket_err(grads, ref["grads"], pre)
        print(dtype, pre, "grad rel-L2 vs fp32 oracle", e)
        assert e < (5e-3 if dtype == torch.float32 else 0.5), (pre, e)
    assert int(model.state_dict()["encoder.conv.0.bn.num_batches_tracked"]) == 2   # two encoder forwards per step


def test_eval_mode_and_cpu_refusal():
    from thesis_fmri_reconstruction_b200.lib import FmriError

    seed = 5
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    for k in S:  # non-trivial running statistics
        if k.endswith("running_mean"):
            S[k] = 0.05 * torch.randn(S[k].shape, generator=torch.Generator().manual_seed(1))
        if k.endswith("running_var"):
            S[k] = 1.0 + 0.2 * torch.rand(S[k].shape, generator=torch.Generator().manual_seed(2))
    x = O.synthetic_images(4, seed=seed)
    mu, lv = O.encoder(P, {k: v.clone() for k, v in S.items()}, x, O.CFG64, train=False)
    img = O.decoder(P, {k: v.clone() for k, v in S.items()}, mu, O.CFG64, train=False)
    with compute(torch.float32):
        model = build_vaegan(P, S)
        model.eval()
        with torch.no_grad():
            m2, l2 = model.encoder(x.cuda())
            im2 = model.decoder(m2)
        assert rel(m2, mu) < 1e-4 and rel(l2, lv) < 1e-4 and rel(im2, img) < 1e-4
        assert int(model.state_dict()["encoder.conv.0.bn.num_batches_tracked"]) == 0
        with pytest.raises(FmriError):
            model.encoder(x)  # CPU input: no fallback path


@pytest.mark.parametrize("stage", [2, 3])
@pytest.mark.parametrize("dtype,B", [(torch.float32, 8), (torch.bfloat16, 16)])
def test_cognitive_stage_script_sequence(stage, dtype, B):
    """VaeGanCognitive + CognitiveEncoder driven like train_vgan_stage2.py:321-407 / train_vgan_stage3.py:324-411 (teacher
    distillation in stage 2, frozen nets, gradient clamp to [-1, 1], RMSprop) against the oracle."""
    mc.use_resolution(64)
    from models.vae_gan import CognitiveEncoder, Decoder, Discriminator, VaeGan, VaeGanCognitive

    seed = 77
    P, S = O.make_cognitive(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    eps_t = O.synthetic_noise(B, 128, seed=seed + 1)[0]
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.cognitive_vaegan_step(P, S_ref, fmri, image, eps, eps_t, z_p, stage)
    hp = O.HP_VGAN
    with compute(dtype):
        teacher = VaeGan(device="cuda", z_size=128)
        cog = CognitiveEncoder(input_size=O.NUM_VOXELS, z_size=128).cuda()
        if stage == 2:   # decoder / discriminator are the teacher's modules (train_vgan_stage2.py:216-232)
            model = VaeGanCognitive(device="cuda", encoder=cog, decoder=teacher.decoder,
                                    discriminator=teacher.discriminator, teacher_net=teacher, stage=2, z_size=128)
        else:
            model = VaeGanCognitive(device="cuda", encoder=cog, decoder=Decoder(z_size=128, size=256).cuda(),
                                    discriminator=Discriminator().cuda(), teacher_net=teacher, stage=3, z_size=128)
        sd = model.state_dict()
        for k, v in {**P, **S}.items():
            sd[k].copy_(v)
        model.train()
        draws = [eps.cuda(), eps_t.cuda()]
        model.reparameterize = lambda mu, lv: ag.reparameterize(mu, lv, draws.pop(0))
        opt = {b: torch.optim.RMSprop(getattr(model, b).parameters(), lr=hp["lr"], alpha=0.9, eps=1e-8)
               for b in ("encoder", "decoder", "discriminator")}
        frozen = model.decoder if stage == 2 else model.encoder
        for p_ in frozen.parameters():
            p_.requires_grad = False
        with patched_randn(z_p):
            x_gt, x_tilde, disc_class, disc_layer, mus, lv = model({"fmri": fmri, "image": image})
        dl_o, dl_p, dl_s = disc_layer[:B], disc_layer[B:-B], disc_layer[-B:]
        dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
        nle, kld, mse, bo, bp, bs = VaeGanCognitive.loss(x_gt, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, lv)
        loss_encoder = torch.sum(kld) + torch.sum(mse)
        loss_discriminator = torch.sum(bo) + torch.sum(bp) + torch.sum(bs)
        loss_decoder = torch.sum(hp["lambda_mse"] * mse) - (1.0 - hp["lambda_mse"]) * loss_discriminator
        grads = {}
        first, first_loss = ("encoder", loss_encoder) if stage == 2 else ("decoder", loss_decoder)
        model.zero_grad()
        first_loss.backward(retain_graph=True)
        grads.update({first + "." + k: p_.grad.clone() for k, p_ in getattr(model, first).named_parameters()})
        [p_.grad.data.clamp_(-1, 1) for p_ in getattr(model, first).parameters()]
        opt[first].step()
        model.zero_grad() if stage == 2 else model.discriminator.zero_grad()
        loss_discriminator.backward()
        grads.update({"discriminator." + k: p_.grad.clone() for k, p_ in model.discriminator.named_parameters()})
        [p_.grad.data.clamp_(-1, 1) for p_ in model.discriminator.parameters()]
        opt["discriminator"].step()
        torch.cuda.synchronize()
    ftol = 1e-4 if dtype == torch.float32 else 2e-2
    fwd = dict(x_tilde=rel(x_tilde, ref["x_tilde"]), gt_x=rel(x_gt, ref["gt_x"]), disc_layer=rel(disc_layer, ref["disc_layer"]),
               disc_class=rel(disc_class, ref["disc_class"]), mu=rel(mus, ref["mu"]), kl=rel(kld, ref["kl"]),
               mse=rel(mse, ref["mse"]), loss_decoder=rel(loss_decoder, ref["loss_decoder"]))
    print(stage, dtype, "forward", fwd)
    assert max(fwd.values()) < ftol, fwd
    for pre in (first + ".", "discriminator."):
        e = bucket_err(grads, {k: v for k, v in ref["grads"].items()}, pre)
        print(stage, dtype, pre, "grad rel-L2 vs fp32 oracle", e)
        assert e < (5e-3 if dtype == torch.float32 else 0.5), (pre, e)
    sd = model.state_dict()
    for k, v in S_ref.items():
        if v.dtype.is_floating_point:
            assert rel(sd[k], v) < ftol, k
        else:
            assert int(sd[k]) == int(v), k   # decoder BN: 3 updates in stage 2 (x_tilde, teacher reconstruction, x_p)


@pytest.mark.parametrize("stage", [2, 3])
@pytest.mark.parametrize("dtype,B", [(torch.float32, 8), (torch.bfloat16, 16)])
def test_cognitive_wae_script_sequence(stage, dtype, B):
    """WaeGanCognitive + CognitiveEncoder + a WaeGan teacher driven like train_wae_stage2.py:274-328 /
    train_wae_stage3.py:295-347 (nn.MSELoss on the module output, mean penalty, Adam) against the oracle."""
    mc.use_resolution(64)
    from models.vae_gan import CognitiveEncoder, WaeGan, WaeGanCognitive

    seed = 91
    P, S = O.make_cognitive_wae(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.cognitive_wae_step(P, S_ref, fmri, image, stage)

    def freeze(m, on):
        for p_ in m.parameters():
            p_.requires_grad = not on

    with compute(dtype):
        trained_model = WaeGan(device="cuda", z_size=128)
        cog = CognitiveEncoder(input_size=O.NUM_VOXELS, z_size=128).cuda()
        model = WaeGanCognitive(device="cuda", encoder=cog, decoder=trained_model.decoder, z_size=128)
        msd, tsd = model.state_dict(), trained_model.state_dict()
        for k, v in {**P, **S}.items():
            (tsd[k[len("teacher_net."):]] if k.startswith("teacher_net.") else msd[k]).copy_(v)
        opt_e = torch.optim.Adam(model.encoder.parameters(), lr=0.001, betas=(0.5, 0.999))
        opt_d = torch.optim.Adam(model.decoder.parameters(), lr=0.001, betas=(0.5, 0.999))
        opt_c = torch.optim.Adam(model.discriminator.parameters(), lr=0.0005, betas=(0.5, 0.999))
        x_fmri, x_image = fmri.cuda(), image.cuda()
        model.train()
        if stage == 2:
            freeze(model.decoder, True)
            model.encoder.zero_grad(); model.discriminator.zero_grad()
            z, _ = trained_model.encoder(x_image)
            trained_model.decoder(z)                      # the script's unused x_gt forward (BN side effects)
            freeze(model.encoder, True); freeze(model.discriminator, False)
        else:
            freeze(model.encoder, True)
            model.decoder.zero_grad(); model.discriminator.zero_grad()
            freeze(model.decoder, True); freeze(model.discriminator, False)
        z_fake, _ = model.encoder(x_fmri)
        z_real, _ = trained_model.encoder(x_image)
        d_real = model.discriminator(z_real)
        d_fake = model.discriminator(z_fake)
        loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))
        loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))
        loss_fake.backward(retain_graph=True)
        loss_real.backward(retain_graph=True)
        grads = {"discriminator." + k: p_.grad.clone() for k, p_ in model.discriminator.named_parameters()}
        opt_c.step()
        freeze(model.discriminator, True)
        freeze(model.encoder if stage == 2 else model.decoder, False)
        z2, _ = model.encoder(x_fmri)
        x_recon = model.decoder(z2)
        d2 = model.discriminator(z2)
        loss_rec = torch.nn.MSELoss()(x_recon, x_image)
        loss_pen = -10 * torch.mean(torch.log(d2 + 1e-3))
        loss_rec.backward(retain_graph=True)
        if stage == 2:
            loss_pen.backward()
            grads.update({"encoder." + k: p_.grad.clone() for k, p_ in model.encoder.named_parameters() if p_.grad is not None})
            opt_e.step()
        else:
            grads.update({"decoder." + k: p_.grad.clone() for k, p_ in model.decoder.named_parameters()})
            opt_d.step()
        torch.cuda.synchronize()
    ftol = 1e-4 if dtype == torch.float32 else 2e-2
    fwd = dict(z_fake=rel(z_fake, ref["z_fake"]), z_real=rel(z_real, ref["z_real"]), d_real=rel(d_real, ref["d_real"]),
               x_recon=rel(x_recon, ref["x_recon"]), loss_rec=rel(loss_rec, ref["loss_reconstruction"]),
               loss_pen=rel(loss_pen, ref["loss_penalty"]))
    print(stage, dtype, "forward", fwd)
    assert max(fwd.values()) < ftol, fwd
    for pre in (ref["trained"] + ".", "discriminator."):
        e = bucket_err(grads, ref["grads"], pre)
        print(stage, dtype, pre, "grad rel-L2 vs fp32 oracle", e)
        assert e < (5e-3 if dtype == torch.float32 else 0.5), (pre, e)
    sd = {**{("teacher_net." + k): v for k, v in trained_model.state_dict().items()}, **model.state_dict()}
    for k, v in S_ref.items():
        if v.dtype.is_floating_point:
            assert rel(sd[k], v) < ftol, k
        else:
            assert int(sd[k]) == int(v), k


def test_pixel_nle_modes_and_dcgan_module_fp32():
    """The 'vae' and 'dcgan' loss mixes of train/train_vgan_stage1.py:374-388 (pixel NLE instead of the feature MSE) and the
    DCGan container (models/vae_gan.py:581-622, the same generated batch passed as `predicted` AND `sampled`) through the
    module path, against the oracle's functional nets differentiated by torch autograd on the CPU (fp32 exact path)."""
    from models.vae_gan import DCGan, VaeGan

    B, seed = 8, 53
    lam = O.HP_VGAN["lambda_mse"]
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    # ---------------- oracle side
    W = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    Sr = {k: v.clone() for k, v in S.items()}
    mu, lv = O.encoder(W, Sr, x, O.CFG64)
    xt = O.decoder(W, Sr, O.reparameterize(mu, lv, eps), O.CFG64)
    xp = O.decoder(W, Sr, z_p, O.CFG64)
    dl = O.discriminator(W, Sr, x, xt, xp, O.CFG64, "REC")
    dc = O.discriminator(W, Sr, x, xt, xp, O.CFG64, "GAN")
    nle, kl, mse, bo, bp, bs = O.vaegan_loss(x, xt, dl[:B], dl[B:-B], dc[:B], dc[B:-B], dc[-B:], mu, lv)
    ref_enc = kl.sum() + nle.sum()                                      # 'vae' :384 / 'dcgan' :378
    ref_dis = bo.sum() + bs.sum()                                       # :379 / :385
    ref_dec = (lam * nle).sum() - (1.0 - lam) * ref_dis                 # 'dcgan' :380
    names = {b: [k for k in W if k.startswith(b + ".")] for b in ("encoder", "decoder", "discriminator")}
    g_ref = {}
    for b, loss in (("encoder", ref_enc), ("decoder", ref_dec), ("discriminator", ref_dis)):
        g_ref.update(zip(names[b], torch.autograd.grad(loss, [W[n] for n in names[b]], retain_graph=True)))
    # ---------------- module side
    with compute(torch.float32):
        model = build_vaegan(P, S)
        model.train()
        eps_d = eps.cuda()
        model.reparameterize = lambda m, l: ag.reparameterize(m, l, eps_d)
        with patched_randn(z_p):
            x_tilde, disc_class, disc_layer, mus, lvs = model(x)
        n2, k2, m2, o2, p2, s2 = VaeGan.loss(x.cuda(), x_tilde, disc_layer[:B], disc_layer[B:-B], disc_layer[-B:],
                                             disc_class[:B], disc_class[B:-B], disc_class[-B:], mus, lvs)
        loss_enc = torch.sum(k2) + torch.sum(n2)
        loss_dis = torch.sum(o2) + torch.sum(s2)
        loss_dec = torch.sum(lam * n2) - (1.0 - lam) * loss_dis
        g = {}
        for b, loss in (("encoder", loss_enc), ("decoder", loss_dec), ("discriminator", loss_dis)):
            ps = dict(getattr(model, b).named_parameters())
            gs = torch.autograd.grad(loss, list(ps.values()), retain_graph=True)
            g.update({b + "." + k: v for k, v in zip(ps, gs)})
        fwd = dict(nle=rel(n2, nle), loss_enc=rel(loss_enc, ref_enc), loss_dec=rel(loss_dec, ref_dec), loss_dis=rel(loss_dis, ref_dis))
        gerr = {b: bucket_err(g, g_ref, b + ".") for b in names}
        print("pixel-NLE modes: forward", fwd, "grad buckets", gerr)
        assert max(fwd.values()) < 1e-4, fwd
        assert max(gerr.values()) < 5e-3, gerr
        # ---------------- DCGan container: decoder(z_p) judged against the real batch, generated batch in two slots
        dcg = DCGan(device="cuda", decoder=model.decoder, discriminator=model.discriminator, z_size=128)
        dcg.train()
        with patched_randn(z_p):
            out = dcg(x)
        torch.cuda.synchronize()
    # train-mode outputs depend on the weights and the batch statistics only (not on the running statistics)
    Sr2 = {k: v.clone() for k, v in S.items()}
    xg = O.decoder(P, Sr2, z_p, O.CFG64)
    dl2 = O.discriminator(P, Sr2, x, xg, xg, O.CFG64, "REC")
    dc2 = O.discriminator(P, Sr2, x, xg, xg, O.CFG64, "GAN")
    gt_x, x_gen, d_class, d_layer = out                                  # models/vae_gan.py:613
    assert torch.equal(gt_x.cpu(), x)
    errs = dict(x_tilde=rel(x_gen, xg), disc_class=rel(d_class, dc2), disc_layer=rel(d_layer, dl2))
    print("DCGan forward", errs)
    assert max(errs.values()) < 1e-4, errs


def test_module_path_frees_saved_state_without_gc():
    """the decoder Function once kept its own output on the saved state (output -> grad_fn -> ctx -> output), a
    reference cycle only the cyclic GC frees. With the collector disabled, repeated script-style steps must not grow the
    allocation: refcounting alone releases every call's saved activations and operand packs."""
    import gc

    from models.vae_gan import VaeGan

    B, seed = 16, 5
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed).cuda()
    with compute(torch.bfloat16):
        model = build_vaegan(P, S)
        model.train()
        opt = torch.optim.RMSprop(model.parameters(), lr=1e-4, alpha=0.9, eps=1e-8)

        def one():
            x_tilde, disc_class, disc_layer, mus, lv = model(x)
            nle, kld, mse, bo, bp, bs = VaeGan.loss(x, x_tilde, disc_layer[:B], disc_layer[B:-B], disc_layer[-B:],
                                                    disc_class[:B], disc_class[B:-B], disc_class[-B:], mus, lv)
            loss = torch.sum(kld) + torch.sum(mse) + torch.sum(bo) + torch.sum(bp) + torch.sum(bs)
            model.zero_grad()
            loss.backward()
            opt.step()

        def census():
            import warnings

            n = 0
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")   # isinstance() on lazily-deprecated module attributes warns
                for o in gc.get_objects():
                    try:
                        if isinstance(o, torch.Tensor) and o.is_cuda:
                            n += 1
                    except Exception:
                        pass
            return n

        gc.collect()
        gc.disable()
        try:
            for _ in range(2):
                one()
            torch.cuda.synchronize()
            m0, n0 = torch.cuda.memory_allocated(), census()
            for _ in range(6):
                one()
            torch.cuda.synchronize()
            m1, n1 = torch.cuda.memory_allocated(), census()
        finally:
            gc.enable()
    print("allocated after 2 steps", m0, "after 8 steps", m1, "live CUDA tensors", n0, n1)
    # no tensor survives a step (a trapped cycle would add its tensors every step) ...
    assert n1 == n0, (n0, n1)
    # ... and the allocation does not grow: one step's saved state is > 100 MB, the caching allocator's block rounding
    # moves memory_allocated() by a few MB from step to step (measured +-3.7 MB with an unchanged tensor census)
    assert m1 <= m0 + (8 << 20), (m0, m1)


@pytest.mark.parametrize("level", [1, 2])
def test_discriminator_recon_level_1_and_2_fp32(level):
    """Discriminator(recon_level=1|2) (reference vae_gan.py:166-175): "REC" returns the raw conv output of block `level`
    and stops there (later blocks neither run nor update their BatchNorm statistics). Output, input gradients, parameter
    gradients and BN buffers against the oracle; fp32 exact path, 1e-4 forward / 5e-3 gradients."""
    from models.vae_gan import Discriminator

    B, seed = 4, 17
    mc.use_resolution(64)
    P, S = O.make_net("discriminator.", O.discriminator_spec(O.CFG64), seed)
    xs = [O.synthetic_images(B, seed=seed + i) for i in range(3)]
    W = {k: v.double().requires_grad_(True) for k, v in P.items()}
    S64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in S.items()}
    xr = [t.double().requires_grad_(True) for t in xs]
    ref = O.discriminator(W, S64, xr[0], xr[1], xr[2], O.CFG64, "REC", recon_level=level)
    g = torch.Generator().manual_seed(seed)
    up = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    names = list(W)
    gs = torch.autograd.grad((ref * up).sum(), [W[n] for n in names] + xr, allow_unused=True)
    with compute(torch.float32):
        m = Discriminator(channel_in=3, recon_level=level).cuda()
        m.load_state_dict({k[len("discriminator."):]: v for k, v in {**P, **S}.items()}, strict=True)
        m.train()
        xd = [t.cuda().requires_grad_(True) for t in xs]
        out = m(xd[0], xd[1], xd[2], "REC")
        assert out.shape == ref.shape
        (out * up.float().cuda()).sum().backward()
        torch.cuda.synchronize()
    assert rel(out, ref) < 1e-4
    for i in range(3):
        assert rel(xd[i].grad, gs[len(names) + i]) < 5e-3, i
    mp = dict(m.named_parameters())
    for n, gr in zip(names, gs[:len(names)]):
        k = n[len("discriminator."):]
        if gr is None:      # blocks above the tap and the head take no part in this pass
            assert mp[k].grad is None or float(mp[k].grad.abs().max()) == 0.0, k
        else:
            assert rel(mp[k].grad, gr) < 5e-3, k
    sd = m.state_dict()
    for k, v in S64.items():
        kk = k[len("discriminator."):]
        if v.dtype.is_floating_point:
            assert rel(sd[kk], v) < 1e-4, kk
        else:
            assert int(sd[kk]) == int(v), kk      # num_batches_tracked: 1 up to the tap, 0 above it


@pytest.mark.parametrize("kind,cin,cout,dtype", [("enc", 32, 128, torch.bfloat16), ("enc", 3, 64, torch.bfloat16),
                                                 ("enc", 64, 128, torch.float32), ("dec", 128, 32, torch.bfloat16),
                                                 ("dec", 256, 128, torch.float32)])
def test_standalone_blocks(kind, cin, cout, dtype):
    """EncoderBlock.forward(ten, out=True) / DecoderBlock.forward(ten) called ON THEIR OWN, as the reference's modules can be
    (vae_gan.py:23-35, 56-60), against plain PyTorch: outputs, the raw-conv feature tap, input / parameter gradients, BN
    running statistics. A 3-channel input is not tiled by the tensor path and runs on the exact fp32 path."""
    import torch.nn.functional as F

    from models.vae_gan import DecoderBlock, EncoderBlock

    N, H = 4, 16
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    exact = dtype == torch.float32 or cin == 3
    rnd = (lambda t: t) if exact else (lambda t: t.bfloat16().float())
    x = rnd(torch.randn(N, cin, H, H, generator=g))
    with compute(dtype):
        m = (EncoderBlock(cin, cout) if kind == "enc" else DecoderBlock(cin, cout, out=True)).cuda()
        with torch.no_grad():
            m.conv.weight.copy_(rnd(m.conv.weight.cpu()))
            m.bn.weight.add_(0.1 * torch.randn(cout, generator=g).cuda())
            m.bn.bias.add_(0.1 * torch.randn(cout, generator=g).cuda())
        w, gamma, beta = (t.detach().cpu().double().requires_grad_(True) for t in (m.conv.weight, m.bn.weight, m.bn.bias))
        xr = x.double().requires_grad_(True)
        if kind == "enc":
            raw = F.conv2d(xr, w, stride=2, padding=2)
        else:
            raw = F.conv_transpose2d(xr, w, stride=2, padding=2, output_padding=1)
        rm, rv = torch.zeros(cout, dtype=torch.float64), torch.ones(cout, dtype=torch.float64)
        ref = torch.relu(F.batch_norm(raw, rm, rv, gamma, beta, True, 0.9, 1e-5))
        up = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        up_raw = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        m.train()
        xd = x.cuda().requires_grad_(True)
        if kind == "enc":
            y, tap = m(xd, True)
            loss_ref = (ref * up).sum() + (raw * up_raw).sum()
            (y * up.float().cuda()).sum().add((tap * up_raw.float().cuda()).sum()).backward()
        else:
            y, tap = m(xd), None
            loss_ref = (ref * up).sum()
            (y * up.float().cuda()).sum().backward()
        gx, gw, gg, gb = torch.autograd.grad(loss_ref, (xr, w, gamma, beta))
        torch.cuda.synchronize()
    ft, gt = (1e-4, 2e-3) if exact else (4e-3, 2e-2)
    assert rel(y, ref) < ft
    if tap is not None:
        assert rel(tap, raw) < ft
    errs = dict(dx=rel(xd.grad, gx), dw=rel(m.conv.weight.grad, gw), dgamma=rel(m.bn.weight.grad, gg), dbeta=rel(m.bn.bias.grad, gb))
    print(kind, cin, cout, dtype, errs)
    assert max(errs.values()) < gt, errs
    assert rel(m.bn.running_mean, rm) < max(ft, 1e-3) and rel(m.bn.running_var, rv) < max(ft, 1e-3)
    assert int(m.bn.num_batches_tracked) == 1
    m.eval()
    with compute(dtype), torch.no_grad():
        ye = m(xd) if kind == "dec" else m(xd, False)
    ref_e = torch.relu(F.batch_norm(raw.detach(), rm, rv, gamma.detach(), beta.detach(), False, 0.9, 1e-5))
    assert rel(ye, ref_e) < max(ft, 1e-3)


def test_vaegan_cognitive_wae_mode_fp32():
    """VaeGanCognitive(mode='wae') (reference vae_gan.py:379-387): x_tilde = decoder(mu_cog), gt_x = decoder(mu_teacher) with
    NO sampling, then the same discriminator passes. Forward tensors and the cognitive-encoder gradient of the feature-matching
    loss against the oracle's functions; fp32 exact path."""
    from models.vae_gan import CognitiveEncoder, VaeGan, VaeGanCognitive

    B, seed = 4, 23
    mc.use_resolution(64)
    P, S = O.make_cognitive(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    z_p = O.synthetic_noise(B, 128, seed=seed)[1]
    W = {k: v.double().requires_grad_(True) for k, v in P.items()}
    S64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in S.items()}
    mu, lv = O.cognitive_encoder(W, S64, fmri.double())
    x_t = O.decoder(W, S64, mu, O.CFG64)
    mu_t, _ = O.encoder(W, S64, image.double(), O.CFG64, pre="teacher_net.encoder.")
    gt = O.decoder(W, S64, mu_t, O.CFG64)
    x_p = O.decoder(W, S64, z_p.double(), O.CFG64)
    dl = O.discriminator(W, S64, gt, x_t, x_p, O.CFG64, "REC")
    dc = O.discriminator(W, S64, gt, x_t, x_p, O.CFG64, "GAN")
    mse_ref = torch.sum(0.5 * (dl[:B] - dl[B:2 * B]) ** 2)
    enc_names = [k for k in W if k.startswith("encoder.")]
    g_ref = dict(zip(enc_names, torch.autograd.grad(mse_ref, [W[k] for k in enc_names], allow_unused=True)))
    with compute(torch.float32):
        teacher = VaeGan(device="cuda", z_size=128)
        teacher.load_state_dict({k[len("teacher_net."):]: v for k, v in {**P, **S}.items() if k.startswith("teacher_net.")},
                                strict=False)
        cog = CognitiveEncoder(input_size=fmri.shape[1], z_size=128).cuda()
        model = VaeGanCognitive(device="cuda", encoder=cog, decoder=teacher.decoder, discriminator=teacher.discriminator,
                                teacher_net=teacher, stage=2, z_size=128, mode="wae").cuda()
        own = {k: v for k, v in {**P, **S}.items() if not k.startswith("teacher_net.")}
        sd = model.state_dict()
        sd.update({k: v for k, v in own.items() if k in sd})
        sd.update({k: v for k, v in {**P, **S}.items() if k in sd and k.startswith("teacher_net.encoder.")})
        # decoder / discriminator are shared objects: teacher_net.decoder.* aliases decoder.*
        for k in list(sd):
            for a, b in (("teacher_net.decoder.", "decoder."), ("teacher_net.discriminator.", "discriminator.")):
                if k.startswith(a) and b + k[len(a):] in own:
                    sd[k] = own[b + k[len(a):]]
        model.load_state_dict(sd, strict=True)
        model.train()
        with patched_randn(z_p):
            gt_x, x_tilde, disc_class, disc_layer, mus, log_variances = model({"fmri": fmri, "image": image})
        mse = ag.row_sq_diff(disc_layer[:B], disc_layer[B:2 * B], 0.5).sum()
        model.zero_grad()
        mse.backward()
        torch.cuda.synchronize()
    fwd = dict(gt_x=rel(gt_x, gt), x_tilde=rel(x_tilde, x_t), disc_layer=rel(disc_layer, dl), disc_class=rel(disc_class, dc),
               mus=rel(mus, mu), logvar=rel(log_variances, lv), mse=rel(mse, mse_ref))
    print("VaeGanCognitive wae mode", fwd)
    assert max(fwd.values()) < 1e-4, fwd
    # l_mu.bias is skipped: decoder.fc is a Linear followed by train-mode BatchNorm, which removes any shift common to the
    # whole batch, so d mse / d l_mu.bias is zero in exact arithmetic (the oracle returns 1e-17) and has no relative error
    scale = max(float(g.norm()) for g in g_ref.values() if g is not None)
    ge = {k: rel(dict(model.named_parameters())[k].grad, g) for k, g in g_ref.items()
          if g is not None and float(g.norm()) > 1e-9 * scale}
    print(ge)
    assert ge and max(ge.values()) < 5e-3, ge
