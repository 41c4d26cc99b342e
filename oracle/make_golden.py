"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (TEST INFRASTRUCTURE ONLY).

Run in the build container only:   python oracle/make_golden.py      (needs /root/reference; ~1 min on 8 cores)

What runs on the reference side (imported from /root/reference, nothing copied):
  * models.vae_gan.VaeGan / WaeGan built under the 64x64 configuration (configs/models_config.py:23-31 + :9),
    .double(), weights loaded with load_state_dict from oracle.vaegan.make_vaegan / make_waegan (deterministic);
  * VaeGan.forward (train) + VaeGan.loss + the loss mix, gate and backward order of train/train_vgan_stage1.py:330-432,
    with torch.optim.RMSprop(alpha=0.9, eps=1e-8); the three optimizer steps are applied after the three backward sweeps
    (equivalent under the torch-1.4 semantics the script was written for; torch >= 1.5 raises on the
    interleaved order);
  * train/train_wae_stage1.py:263-311 verbatim order (D-phase, Adam step, G-phase, Adam steps) with torch.optim.Adam.
Noise (eps, z_p, z_fake) is injected by patching VaeGan.reparameterize / torch.randn for the duration of the call.
"""
from __future__ import annotations

import os
import sys
from contextlib import contextmanager

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import vaegan as O  # noqa: E402
from oracle.golden_util import summarize, summarize_dict  # noqa: E402


def import_reference(cfg=None):
    """cfg: an oracle architecture dict (O.CFG64 default, O.CFG100 = the reference's ACTIVE block,
    configs/models_config.py:13-21); its values are written into the imported reference config module."""
    cfg = cfg or O.CFG64
    saved = list(sys.path)
    sys.path[:] = [REF] + [q for q in sys.path if os.path.abspath(q or ".") != ROOT]  # reference's models/ is a namespace pkg
    for m in [k for k in sys.modules if k == "configs" or k.startswith("configs.") or k == "models"
              or k.startswith("models.")]:
        del sys.modules[m]
    import configs.models_config as mc

    assert mc.__file__.startswith(REF), mc.__file__
    for k in ("image_size", "fc_input", "fc_output", "fc_input_gan", "fc_output_gan", "stride_gan", "latent_dim"):
        setattr(mc, k, cfg[k])
    mc.output_pad_dec = list(cfg["output_pad_dec"])
    mc.encoder_channels = list(cfg["encoder_channels"])
    mc.decoder_channels = list(cfg["decoder_channels"])
    mc.discrim_channels = list(cfg["discrim_channels"])
    import models.vae_gan as ref

    assert ref.__file__.startswith(REF), ref.__file__
    sys.path[:] = saved
    return ref


@contextmanager
def patched_randn(value):
    orig = torch.randn

    def fake(*a, **k):
        return value.clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = orig


def load(model, P, S):
    sd = {**P, **S}
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def buffers_of(model):
    return {k: v for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}


def golden_stage1_vaegan(ref, B, seed, cfg=None, mode="vae-gan", beta=1.0):
    cfg = cfg or O.CFG64
    z = cfg["latent_dim"]
    torch.manual_seed(0)
    P, S = O.make_vaegan(cfg, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, size=cfg["image_size"], seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, z, seed=seed)]
    model = ref.VaeGan(device="cpu", z_size=z).double()
    load(model, P, S)
    model.train()
    model.reparameterize = lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu  # models/vae_gan.py:266-269, eps injected
    hp = O.HP_VGAN
    opts = {b: torch.optim.RMSprop(getattr(model, b).parameters(), lr=hp["lr"], alpha=0.9, eps=1e-8, weight_decay=0,
                                   momentum=0, centered=False) for b in ("encoder", "decoder", "discriminator")}
    with patched_randn(z_p):
        x_tilde, disc_class, disc_layer, mus, lv = model(x)                       # train_vgan_stage1.py:330
    dl_o, dl_p, dl_s = disc_layer[:B], disc_layer[B:-B], disc_layer[-B:]          # :333-335
    dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]          # :337-339
    nle, kld, mse, bo, bp, bs = ref.VaeGan.loss(x, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, lv)  # :342
    train_dis, train_dec, train_enc = True, True, True                            # :352-356
    if mode == "beta-vae":
        kld_weight = 1 / B                                                        # :361
        loss_encoder = torch.sum(kld) * beta * kld_weight + torch.sum(mse)        # :362
    elif mode in ("vae-gan",):
        loss_encoder = torch.sum(kld) + torch.sum(mse)                            # :369
    if mode in ("vae-gan", "beta-vae"):
        loss_discriminator = torch.sum(bo) + torch.sum(bp) + torch.sum(bs)        # :370
        loss_decoder = torch.sum(hp["lambda_mse"] * mse) - (1.0 - hp["lambda_mse"]) * loss_discriminator  # :372
    if mode == "dcgan":                                                           # :374-380
        train_enc = False
        for param in model.encoder.parameters():
            param.requires_grad = False
        loss_encoder = torch.sum(kld) + torch.sum(nle)
        loss_discriminator = torch.sum(bo) + torch.sum(bs)
        loss_decoder = torch.sum(hp["lambda_mse"] * nle) - (1.0 - hp["lambda_mse"]) * loss_discriminator
    if mode == "vae":                                                             # :382-387
        loss_encoder = torch.sum(kld) + torch.sum(nle)
        loss_discriminator = torch.sum(bo) + torch.sum(bs)
        loss_decoder = torch.sum(hp["lambda_mse"] * nle)
        train_dis = False
    if torch.mean(bo).item() < hp["equilibrium"] - hp["margin"] or torch.mean(bp).item() < hp["equilibrium"] - hp["margin"]:
        train_dis = False
    if torch.mean(bo).item() > hp["equilibrium"] + hp["margin"] or torch.mean(bp).item() > hp["equilibrium"] + hp["margin"]:
        train_dec = False
    if train_dec is False and train_dis is False:
        train_dis = True
        train_dec = True
    def grab(mod):   # a parameter the loss does not reach keeps grad None (torch >= 2): recorded as zeros, RMSprop skips it
        return {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in mod.named_parameters()}

    grads = {}
    model.zero_grad()                                                             # :408
    if train_enc:                                                                 # :410
        loss_encoder.backward(retain_graph=True)                                  # :412
        grads.update({"encoder." + k: v for k, v in grab(model.encoder).items()})
        model.zero_grad()                                                         # :418
    loss_decoder.backward(retain_graph=True)                                      # :422 (the script guards it with train_dec)
    grads.update({"decoder." + k: v for k, v in grab(model.decoder).items()})
    model.discriminator.zero_grad()                                               # :426
    loss_discriminator.backward()                                                 # :430
    grads.update({"discriminator." + k: v for k, v in grab(model.discriminator).items()})
    for b, on in (("encoder", train_enc), ("decoder", train_dec), ("discriminator", train_dis)):
        if not on:
            continue
        for k, p in getattr(model, b).named_parameters():
            p.grad = grads[b + "." + k].clone()
        opts[b].step()
    newP = {k: v.detach() for k, v in model.named_parameters()}
    delta = {k: newP[k] - P[k] for k in P}
    fx = dict(B=np.array(B), seed=np.array(seed), train_dis=np.array(train_dis), train_dec=np.array(train_dec),
              mu=mus.detach().numpy(), logvar=lv.detach().numpy(), kl=kld.detach().numpy(), mse=mse.detach().numpy(),
              bce_o=bo.detach().numpy(), bce_p=bp.detach().numpy(), bce_s=bs.detach().numpy(),
              disc_class=disc_class.detach().numpy(), nle_sum=nle.sum().detach().numpy(),
              loss_encoder=loss_encoder.detach().numpy(), loss_decoder=loss_decoder.detach().numpy(),
              loss_discriminator=loss_discriminator.detach().numpy(),
              x_tilde=summarize(x_tilde), disc_layer=summarize(disc_layer))
    fx.update(summarize_dict(grads, "grad:"))
    fx.update(summarize_dict(delta, "delta:"))
    fx.update(summarize_dict({k: v for k, v in buffers_of(model).items()}, "buf:"))
    # inference path (inference/inference_gan.py drives VaeGan.forward in eval mode, models/vae_gan.py:288-292): updated
    # weights, the running statistics the step just produced, eval-mode BatchNorm, the same injected eps
    model.eval()
    with torch.no_grad():
        x_eval = model(x)
    fx["eval_x_tilde"] = summarize(x_eval)
    return fx


def golden_stage1_waegan(ref, B, seed):
    torch.manual_seed(0)
    P, S = O.make_waegan(O.CFG64, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, seed=seed).double()
    z_fake = (O.synthetic_noise(B, 128, seed=seed)[0] * 0.5).double()
    model = ref.WaeGan(device="cpu", z_size=128).double()
    load(model, P, S)
    model.train()
    hp = O.HP_WAE
    opt_e = torch.optim.Adam(model.encoder.parameters(), lr=hp["lr"], betas=(0.5, 0.999))           # wae1 :221
    opt_d = torch.optim.Adam(model.decoder.parameters(), lr=hp["lr"], betas=(0.5, 0.999))           # :222
    opt_c = torch.optim.Adam(model.discriminator.parameters(), lr=0.5 * hp["lr"], betas=(0.5, 0.999))  # :223-224

    def freeze(m, on):
        for p in m.parameters():
            p.requires_grad = not on

    model.encoder.zero_grad(); model.decoder.zero_grad(); model.discriminator.zero_grad()   # :263-265
    freeze(model.decoder, True); freeze(model.encoder, True); freeze(model.discriminator, False)  # :271-273
    z_real, _ = model.encoder(x)                                                  # :275
    d_real = model.discriminator(z_real)                                          # :278
    d_fake = model.discriminator(z_fake)                                          # :279
    loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))                         # :281
    loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))                     # :282
    loss_fake.backward(retain_graph=True)                                         # :283
    loss_real.backward(retain_graph=True)                                         # :284
    grads = {"discriminator." + k: p.grad.clone() for k, p in model.discriminator.named_parameters()}
    opt_c.step()                                                                  # :288
    freeze(model.encoder, False); freeze(model.decoder, False); freeze(model.discriminator, True)  # :292-294
    z_real2, _ = model.encoder(x)                                                 # :296
    x_recon = model.decoder(z_real2)                                              # :297
    d_real2 = model.discriminator(z_real2)                                        # :298
    loss_rec = torch.sum(torch.sum(0.5 * (x_recon - x) ** 2, 1))                  # :301
    loss_pen = -10 * torch.sum(torch.log(d_real2 + 1e-3))                         # :303
    loss_rec.backward(retain_graph=True)                                          # :306
    loss_pen.backward()                                                           # :307
    # l_var gets no gradient (the WAE ignores logvar): p.grad stays None and Adam skips it
    grads.update({"encoder." + k: p.grad.clone() for k, p in model.encoder.named_parameters() if p.grad is not None})
    grads.update({"decoder." + k: p.grad.clone() for k, p in model.decoder.named_parameters()})
    opt_e.step()                                                                  # :310
    opt_d.step()                                                                  # :311
    newP = {k: v.detach() for k, v in model.named_parameters()}
    delta = {k: newP[k] - P[k] for k in P}
    fx = dict(B=np.array(B), seed=np.array(seed), z_real=z_real.detach().numpy(), d_real=d_real.detach().numpy(),
              d_fake=d_fake.detach().numpy(), d_real_g=d_real2.detach().numpy(),
              loss_discriminator_fake=loss_fake.detach().numpy(), loss_discriminator_real=loss_real.detach().numpy(),
              loss_reconstruction=loss_rec.detach().numpy(), loss_penalty=loss_pen.detach().numpy(),
              x_recon=summarize(x_recon))
    fx.update(summarize_dict(grads, "grad:"))
    fx.update(summarize_dict(delta, "delta:"))
    fx.update(summarize_dict(buffers_of(model), "buf:"))
    return fx


def golden_cognitive(ref, B, seed, stage):
    """train/train_vgan_stage2.py:210-232, 321-407 (stage 2) / train_vgan_stage3.py:225-262, 324-411 (stage 3)."""
    torch.manual_seed(0)
    P, S = O.make_cognitive(O.CFG64, seed=seed, dtype=torch.float64)
    fmri = O.synthetic_fmri(B, seed=seed).double()
    image = O.synthetic_images(B, seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, 128, seed=seed)]
    eps_t = O.synthetic_noise(B, 128, seed=seed + 1)[0].double()
    teacher = ref.VaeGan(device="cpu", z_size=128).double()
    cog = ref.CognitiveEncoder(input_size=O.NUM_VOXELS, z_size=128).double()
    if stage == 2:   # decoder / discriminator are the teacher's own modules (train_vgan_stage2.py:216-232)
        model = ref.VaeGanCognitive(device="cpu", encoder=cog, decoder=teacher.decoder, discriminator=teacher.discriminator,
                                    teacher_net=teacher, stage=2, z_size=128)
    else:            # fresh decoder / discriminator objects, teacher only carried along (train_vgan_stage3.py:230-241)
        model = ref.VaeGanCognitive(device="cpu", encoder=cog, decoder=ref.Decoder(z_size=128, size=256).double(),
                                    discriminator=ref.Discriminator().double(), teacher_net=teacher, stage=3, z_size=128)
    sd = model.state_dict()
    for k, v in {**P, **S}.items():
        sd[k].copy_(v)
        if stage == 2 and (k.startswith("decoder.") or k.startswith("discriminator.")):
            assert torch.equal(sd["teacher_net." + k], sd[k])   # same tensors in stage 2
    model.train()
    draws = [eps, eps_t]
    model.reparameterize = lambda mu, logvar: draws.pop(0) * torch.exp(0.5 * logvar) + mu
    hp = O.HP_VGAN
    opt = {b: torch.optim.RMSprop(getattr(model, b).parameters(), lr=hp["lr"], alpha=0.9, eps=1e-8)
           for b in ("encoder", "decoder", "discriminator")}
    if stage == 2:
        for p_ in model.decoder.parameters():
            p_.requires_grad = False                                                  # stage2 :328-329
    else:
        for p_ in model.encoder.parameters():
            p_.requires_grad = False                                                  # stage3 :329-330
    with patched_randn(z_p):
        x_gt, x_tilde, disc_class, disc_layer, mus, lv = model({"fmri": fmri, "image": image})
    dl_o, dl_p, dl_s = disc_layer[:B], disc_layer[B:-B], disc_layer[-B:]
    dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
    nle, kld, mse, bo, bp, bs = ref.VaeGanCognitive.loss(x_gt, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, lv)
    loss_encoder = torch.sum(kld) + torch.sum(mse)
    loss_discriminator = torch.sum(bo) + torch.sum(bp) + torch.sum(bs)
    loss_decoder = torch.sum(hp["lambda_mse"] * mse) - (1.0 - hp["lambda_mse"]) * loss_discriminator
    grads = {}
    if stage == 2:
        train_dis, train_dec = True, False                                            # stage2 :375-376
        model.zero_grad()
        loss_encoder.backward(retain_graph=True)                                      # :389
        grads.update({"encoder." + k: p_.grad.clone() for k, p_ in model.encoder.named_parameters()})
        model.zero_grad()                                                             # :393
        loss_discriminator.backward()                                                 # :405
        grads.update({"discriminator." + k: p_.grad.clone() for k, p_ in model.discriminator.named_parameters()})
        steps = [("encoder", True), ("discriminator", True)]
    else:
        train_dis, train_dec = True, True
        m, e = hp["margin"], hp["equilibrium"]
        if torch.mean(bo).item() < e - m or torch.mean(bp).item() < e - m:
            train_dis = False
        if torch.mean(bo).item() > e + m or torch.mean(bp).item() > e + m:
            train_dec = False
        if train_dec is False and train_dis is False:
            train_dis = True
            train_dec = True
        model.zero_grad()                                                             # stage3 :392
        loss_decoder.backward(retain_graph=True)                                      # :400
        grads.update({"decoder." + k: p_.grad.clone() for k, p_ in model.decoder.named_parameters()})
        model.discriminator.zero_grad()                                               # :405
        loss_discriminator.backward()                                                 # :409
        grads.update({"discriminator." + k: p_.grad.clone() for k, p_ in model.discriminator.named_parameters()})
        steps = [("decoder", train_dec), ("discriminator", train_dis)]
    for b, on in steps:
        if not on:
            continue
        for k, p_ in getattr(model, b).named_parameters():
            p_.grad = grads[b + "." + k].clone().clamp_(-1, 1)                        # stage2 :391,406 / stage3 :402,410
        opt[b].step()
    newP = {k: v.detach() for k, v in model.state_dict().items() if k in P}
    delta = {k: newP[k] - P[k] for k in P}
    bufs = {k: v for k, v in model.state_dict().items() if k in S}
    fx = dict(B=np.array(B), seed=np.array(seed), stage=np.array(stage), train_dis=np.array(train_dis),
              train_dec=np.array(train_dec), mu=mus.detach().numpy(), logvar=lv.detach().numpy(),
              kl=kld.detach().numpy(), mse=mse.detach().numpy(), bce_o=bo.detach().numpy(), bce_p=bp.detach().numpy(),
              bce_s=bs.detach().numpy(), disc_class=disc_class.detach().numpy(),
              loss_encoder=loss_encoder.detach().numpy(), loss_decoder=loss_decoder.detach().numpy(),
              loss_discriminator=loss_discriminator.detach().numpy(), x_tilde=summarize(x_tilde), gt_x=summarize(x_gt),
              disc_layer=summarize(disc_layer))
    fx.update(summarize_dict(grads, "grad:"))
    fx.update(summarize_dict(delta, "delta:"))
    fx.update(summarize_dict(bufs, "buf:"))
    return fx


def golden_cognitive_wae(ref, B, seed, stage):
    """train/train_wae_stage2.py:195-203, 274-328 (stage 2) / train/train_wae_stage3.py:295-347 (stage 3)."""
    torch.manual_seed(0)
    P, S = O.make_cognitive_wae(O.CFG64, seed=seed, dtype=torch.float64)
    fmri = O.synthetic_fmri(B, seed=seed).double()
    image = O.synthetic_images(B, seed=seed).double()
    trained_model = ref.WaeGan(device="cpu", z_size=128).double()
    cog = ref.CognitiveEncoder(input_size=O.NUM_VOXELS, z_size=128).double()
    model = ref.WaeGanCognitive(device="cpu", encoder=cog, decoder=trained_model.decoder, z_size=128).double()
    msd, tsd = model.state_dict(), trained_model.state_dict()
    for k, v in {**P, **S}.items():
        if k.startswith("teacher_net."):
            tsd[k[len("teacher_net."):]].copy_(v)
        else:
            msd[k].copy_(v)
    hp = O.HP_WAE23
    opt_e = torch.optim.Adam(model.encoder.parameters(), lr=0.001, betas=(0.5, 0.999))
    opt_d = torch.optim.Adam(model.decoder.parameters(), lr=0.001, betas=(0.5, 0.999))
    opt_c = torch.optim.Adam(model.discriminator.parameters(), lr=0.0005, betas=(0.5, 0.999))

    def freeze(m, on):
        for p_ in m.parameters():
            p_.requires_grad = not on

    model.train()
    if stage == 2:
        freeze(model.decoder, True)
        model.encoder.zero_grad(); model.discriminator.zero_grad()
        z, _ = trained_model.encoder(image)                                           # wae2 :284
        x_gt = trained_model.decoder(z)                                               # :285 (unused)
        freeze(model.encoder, True); freeze(model.discriminator, False)
    else:
        freeze(model.encoder, True)
        model.decoder.zero_grad(); model.discriminator.zero_grad()
        freeze(model.decoder, True); freeze(model.discriminator, False)
    z_fake, _ = model.encoder(fmri)
    z_real, _ = trained_model.encoder(image)
    d_real = model.discriminator(z_real)
    d_fake = model.discriminator(z_fake)
    loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))
    loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))
    loss_fake.backward(retain_graph=True)
    loss_real.backward(retain_graph=True)
    grads = {"discriminator." + k: p_.grad.clone() for k, p_ in model.discriminator.named_parameters()}
    opt_c.step()
    freeze(model.discriminator, True)
    if stage == 2:
        freeze(model.encoder, False)
    else:
        freeze(model.decoder, False)
    z2, _ = model.encoder(fmri)
    x_recon = model.decoder(z2)
    d2 = model.discriminator(z2)
    loss_rec = torch.nn.MSELoss()(x_recon, image)
    loss_pen = -10 * torch.mean(torch.log(d2 + 1e-3))
    loss_rec.backward(retain_graph=True)
    if stage == 2:
        loss_pen.backward()
        grads.update({"encoder." + k: p_.grad.clone() for k, p_ in model.encoder.named_parameters() if p_.grad is not None})
        opt_e.step()
    else:
        grads.update({"decoder." + k: p_.grad.clone() for k, p_ in model.decoder.named_parameters()})
        opt_d.step()
    newP = {k: v.detach() for k, v in model.state_dict().items() if k in P}
    delta = {k: newP[k] - P[k] for k in newP}
    bufs = {k: v for k, v in model.state_dict().items() if k in S}
    bufs.update({"teacher_net." + k: v for k, v in trained_model.state_dict().items() if "teacher_net." + k in S})
    fx = dict(B=np.array(B), seed=np.array(seed), stage=np.array(stage), z_fake=z_fake.detach().numpy(),
              z_real=z_real.detach().numpy(), d_real=d_real.detach().numpy(), d_fake=d_fake.detach().numpy(),
              d_real_g=d2.detach().numpy(), loss_discriminator_fake=loss_fake.detach().numpy(),
              loss_discriminator_real=loss_real.detach().numpy(), loss_reconstruction=loss_rec.detach().numpy(),
              loss_penalty=loss_pen.detach().numpy(), x_recon=summarize(x_recon))
    fx.update(summarize_dict(grads, "grad:"))
    fx.update(summarize_dict(delta, "delta:"))
    fx.update(summarize_dict(bufs, "buf:"))
    return fx


def golden_dual_stage1(ref, B, seed, lam=1.0):
    """train/wae_vgan_stage1.py:282-441 (mode 'vae-gan') with the reference's VaeGan + a second WaeGan's latent discriminator,
    under THIS torch (>= 2: zero_grad() leaves None gradients, so the decoder step of :417 is a no-op -- asserted below). As in
    golden_stage1_vaegan the three VAE/GAN optimizer steps are applied after the three backward sweeps (torch-1.4 order)."""
    torch.manual_seed(0)
    P, S = O.make_dual_stage1(O.CFG64, seed=seed, dtype=torch.float64)
    x = O.synthetic_images(B, seed=seed).double()
    eps, z_p = [t.double() for t in O.synthetic_noise(B, 128, seed=seed)]
    z_fake = (O.synthetic_noise(B, 128, seed=seed + 7)[0] * 0.5).double()      # :385 torch.randn_like(z_real) * 0.5
    model = ref.VaeGan(device="cpu", z_size=128).double()
    load(model, {k: v for k, v in P.items() if not k.startswith("latent_")}, S)
    model_wae = ref.WaeGan(device="cpu", z_size=128).double()                  # :200; only its discriminator is used
    model_wae.discriminator.load_state_dict({k[len("latent_discriminator."):]: v for k, v in P.items()
                                             if k.startswith("latent_discriminator.")}, strict=True)
    model.train()
    model.reparameterize = lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu
    hp = O.HP_VGAN
    mk = lambda ps: torch.optim.RMSprop(ps, lr=hp["lr"], alpha=0.9, eps=1e-8, weight_decay=0, momentum=0, centered=False)
    opts = {b: mk(getattr(model, b).parameters()) for b in ("encoder", "decoder", "discriminator")}   # :238-246
    opt_wae = mk(model_wae.discriminator.parameters())                                                  # :248

    def freeze(m, on):
        for p in m.parameters():
            p.requires_grad = not on

    with patched_randn(z_p):
        x_tilde, disc_class, disc_layer, mus, lv = model(x)                       # :290
    dl_o, dl_p, dl_s = disc_layer[:B], disc_layer[B:-B], disc_layer[-B:]
    dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
    nle, kld, mse, bo, bp, bs = ref.VaeGan.loss(x, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, lv)  # :302
    loss_encoder = torch.sum(kld) + torch.sum(mse)                                # :329
    loss_discriminator = torch.sum(bo) + torch.sum(bp) + torch.sum(bs)
    loss_decoder = torch.sum(hp["lambda_mse"] * mse) - (1.0 - hp["lambda_mse"]) * loss_discriminator
    train_dis, train_dec = True, True                                             # :356-364
    if torch.mean(bo).item() < hp["equilibrium"] - hp["margin"] or torch.mean(bp).item() < hp["equilibrium"] - hp["margin"]:
        train_dis = False
    if torch.mean(bo).item() > hp["equilibrium"] + hp["margin"] or torch.mean(bp).item() > hp["equilibrium"] + hp["margin"]:
        train_dec = False
    if train_dec is False and train_dis is False:
        train_dis = True
        train_dec = True
    model.zero_grad()                                                             # :368
    model.encoder.zero_grad(); model.decoder.zero_grad(); model.discriminator.zero_grad()   # :372-374
    # ---------- latent discriminator (:380-397)
    freeze(model.decoder, True); freeze(model.encoder, True); freeze(model_wae.discriminator, False)
    z_real, _ = model.encoder(x)
    d_real = model_wae.discriminator(z_real)
    d_fake = model_wae.discriminator(z_fake)
    loss_fake = -lam * torch.sum(torch.log(d_fake + 1e-3))
    loss_real = -lam * torch.sum(torch.log(1 - d_real + 1e-3))
    loss_fake.backward(retain_graph=True)
    loss_real.backward(retain_graph=True)
    grads = {"latent_discriminator." + k: p.grad.clone() for k, p in model_wae.discriminator.named_parameters()}
    opt_wae.step()
    # ---------- penalty (:401-417)
    freeze(model.encoder, False); freeze(model.decoder, False); freeze(model_wae.discriminator, True)
    z_real2, _ = model.encoder(x)
    x_recon = model.decoder(z_real2)                                              # unused (BatchNorm side effects)
    d_real2 = model_wae.discriminator(z_real2)
    loss_penalty = -lam * torch.sum(torch.log(d_real2 + 1e-3))
    loss_penalty.backward()
    assert all(p.grad is None for p in model.decoder.parameters()), "torch >= 2 semantics expected (SURVEY 8a a16)"
    opts["decoder"].step()                                                        # :417 -- a no-op here
    # ---------- the three VAE/GAN sweeps (:419-441), steps deferred
    loss_encoder.backward(retain_graph=True)                                      # accumulates onto the penalty gradient
    grads.update({"encoder." + k: p.grad.clone() for k, p in model.encoder.named_parameters() if p.grad is not None})
    model.zero_grad()
    loss_decoder.backward(retain_graph=True)
    grads.update({"decoder." + k: p.grad.clone() for k, p in model.decoder.named_parameters()})
    model.discriminator.zero_grad()
    loss_discriminator.backward()
    grads.update({"discriminator." + k: p.grad.clone() for k, p in model.discriminator.named_parameters()})
    for b, on in (("encoder", True), ("decoder", train_dec), ("discriminator", train_dis)):
        if not on:
            continue
        for k, p in getattr(model, b).named_parameters():
            p.grad = grads[b + "." + k].clone() if (b + "." + k) in grads else None
        opts[b].step()
    newP = {k: v.detach() for k, v in model.named_parameters()}
    newP.update({"latent_discriminator." + k: v.detach() for k, v in model_wae.discriminator.named_parameters()})
    delta = {k: newP[k] - P[k] for k in P}
    fx = dict(B=np.array(B), seed=np.array(seed), lam=np.array(lam), train_dis=np.array(train_dis), train_dec=np.array(train_dec),
              mu=mus.detach().numpy(), kl=kld.detach().numpy(), mse=mse.detach().numpy(),
              loss_encoder=loss_encoder.detach().numpy(), loss_decoder=loss_decoder.detach().numpy(),
              loss_discriminator=loss_discriminator.detach().numpy(), d_real=d_real.detach().numpy(),
              d_fake=d_fake.detach().numpy(), d_real_g=d_real2.detach().numpy(),
              loss_discriminator_fake=loss_fake.detach().numpy(), loss_discriminator_real=loss_real.detach().numpy(),
              loss_penalty=loss_penalty.detach().numpy(), x_tilde=summarize(x_tilde))
    fx.update(summarize_dict(grads, "grad:"))
    fx.update(summarize_dict(delta, "delta:"))
    fx.update(summarize_dict({k: v for k, v in buffers_of(model).items()}, "buf:"))
    return fx


def main():
    ref = import_reference()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for B, seed in ((4, 12345), (6, 777)):
        np.savez_compressed(os.path.join(out, f"stage1_vaegan_B{B}_s{seed}.npz"), **golden_stage1_vaegan(ref, B, seed))
        np.savez_compressed(os.path.join(out, f"stage1_waegan_B{B}_s{seed}.npz"), **golden_stage1_waegan(ref, B, seed))
        print("wrote goldens for", B, seed)
    np.savez_compressed(os.path.join(out, "stage1_betavae_B4_s2024.npz"),
                        **golden_stage1_vaegan(ref, 4, 2024, mode="beta-vae", beta=4.0), beta=np.array(4.0))
    print("wrote the beta-vae golden")
    for m in ("dcgan", "vae"):
        np.savez_compressed(os.path.join(out, f"stage1_{m}_B4_s31{len(m)}.npz"),
                            **golden_stage1_vaegan(ref, 4, 310 + len(m), mode=m), mode=np.array(m))
        print("wrote the", m, "golden")
    np.savez_compressed(os.path.join(out, "stage1_dual_B4_s606.npz"), **golden_dual_stage1(ref, 4, 606))
    print("wrote the dual WAE/GAN Stage-I golden")
    ref100 = import_reference(O.CFG100)   # the reference's active 100x100 / latent-512 block: odd 13/25/50-pixel grids
    np.savez_compressed(os.path.join(out, "stage1_vaegan100_B2_s99.npz"), **golden_stage1_vaegan(ref100, 2, 99, O.CFG100))
    print("wrote the 100x100 golden")
    ref = import_reference()
    for stage in (2, 3):
        np.savez_compressed(os.path.join(out, f"stage{stage}_cognitive_B4_s4711.npz"), **golden_cognitive(ref, 4, 4711, stage))
        print("wrote cognitive golden for stage", stage)
        np.savez_compressed(os.path.join(out, f"stage{stage}_cognitive_wae_B4_s3131.npz"),
                            **golden_cognitive_wae(ref, 4, 3131, stage))
        print("wrote cognitive WAE golden for stage", stage)


if __name__ == "__main__":
    main()
