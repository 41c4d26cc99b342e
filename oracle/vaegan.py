"""Functional torch-CPU restatement of the reference's VAE/GAN / WAE/GAN training step.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).  Works in fp32 or fp64 (dtype follows the parameters).

Parameters live in a flat dict keyed by the reference's state_dict names ("encoder.conv.0.conv.weight", ...), BatchNorm
buffers in a second dict with the same naming ("....bn.running_mean" / "running_var" / "num_batches_tracked").
All paths below are relative to /root/reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------- architecture
# configs/models_config.py:3-31 -- the active block is the 100x100 / latent-512 set, the commented block the 64x64 one.
CFG64 = dict(image_size=64, fc_input=8, fc_output=1024, fc_input_gan=8, fc_output_gan=512, stride_gan=1,
             latent_dim=128, output_pad_dec=[True, True, True], encoder_channels=[64, 128, 256],
             decoder_channels=[256, 128, 32, 3], discrim_channels=[32, 128, 256, 256, 512])
CFG100 = dict(image_size=100, fc_input=13, fc_output=1024, fc_input_gan=7, fc_output_gan=256, stride_gan=2,
              latent_dim=512, output_pad_dec=[False, True, True], encoder_channels=[64, 128, 256],
              decoder_channels=[256, 128, 64, 3], discrim_channels=[32, 128, 256, 256, 512])
NUM_VOXELS = 3620  # configs/data_config.py:62-73 (sum of rois_max)
ROIS_MAX = [522, 455, 279, 86, 172, 696, 597, 335, 278, 200]  # configs/data_config.py:62-71, listed order

BN_MOMENTUM = 0.9  # models/vae_gan.py:21,54,80,108,158,200
BN_EPS = 1e-5      # torch default


def encoder_spec(cfg, z):
    """models/vae_gan.py:63-85, in nn.Module.parameters() order. Returns (params [(name, shape)], bn prefixes)."""
    ps, bns, cin = [], [], 3
    for i, c in enumerate(cfg["encoder_channels"]):
        ps += [(f"conv.{i}.conv.weight", (c, cin, 5, 5)), (f"conv.{i}.bn.weight", (c,)), (f"conv.{i}.bn.bias", (c,))]
        bns.append((f"conv.{i}.bn.", c))
        cin = c
    fi = cfg["fc_input"] ** 2 * cin
    ps += [("fc.0.weight", (cfg["fc_output"], fi)), ("fc.1.weight", (cfg["fc_output"],)),
           ("fc.1.bias", (cfg["fc_output"],)), ("l_mu.weight", (z, cfg["fc_output"])), ("l_mu.bias", (z,)),
           ("l_var.weight", (z, cfg["fc_output"])), ("l_var.bias", (z,))]
    bns.append(("fc.1.", cfg["fc_output"]))
    return ps, bns


def decoder_spec(cfg, z, size=256):
    """models/vae_gan.py:99-123."""
    dc = cfg["decoder_channels"]
    fo = cfg["fc_input"] ** 2 * size
    ps = [("fc.0.weight", (fo, z)), ("fc.1.weight", (fo,)), ("fc.1.bias", (fo,))]
    bns = [("fc.1.", fo)]
    chans = [(size, size), (size, dc[1]), (dc[1], dc[2])]
    for i, (ci, co) in enumerate(chans):
        ps += [(f"conv.{i}.conv.weight", (ci, co, 5, 5)), (f"conv.{i}.bn.weight", (co,)), (f"conv.{i}.bn.bias", (co,))]
        bns.append((f"conv.{i}.bn.", co))
    ps += [("conv.3.0.weight", (dc[3], dc[2], 5, 5)), ("conv.3.0.bias", (dc[3],))]
    return ps, bns


def discriminator_spec(cfg):
    """models/vae_gan.py:135-161."""
    ch = cfg["discrim_channels"]
    ps = [("conv.0.0.weight", (ch[0], 3, 5, 5)), ("conv.0.0.bias", (ch[0],))]
    bns = []
    for i in (1, 2, 3):
        ps += [(f"conv.{i}.conv.weight", (ch[i], ch[i - 1], 5, 5)), (f"conv.{i}.bn.weight", (ch[i],)),
               (f"conv.{i}.bn.bias", (ch[i],))]
        bns.append((f"conv.{i}.bn.", ch[i]))
    fi = cfg["fc_input_gan"] ** 2 * ch[3]
    fo = cfg["fc_output_gan"]
    ps += [("fc.0.weight", (fo, fi)), ("fc.1.weight", (fo,)), ("fc.1.bias", (fo,)), ("fc.3.weight", (1, fo)),
           ("fc.3.bias", (1,))]
    bns.append(("fc.1.", fo))
    return ps, bns


def cognitive_encoder_spec(z, input_size=NUM_VOXELS):
    """models/vae_gan.py:190-207."""
    ps = [("fc1.0.weight", (1024, input_size)), ("fc1.1.weight", (1024,)), ("fc1.1.bias", (1024,)),
          ("l_mu.weight", (z, 1024)), ("l_mu.bias", (z,)), ("l_var.weight", (z, 1024)), ("l_var.bias", (z,))]
    return ps, [("fc1.1.", 1024)]


def wae_discriminator_spec(z, dim_h=512):
    """models/vae_gan.py:499-521."""
    dims = [(dim_h, z), (dim_h, dim_h), (dim_h, dim_h), (dim_h, dim_h), (1, dim_h)]
    ps = []
    for i, (o, k) in zip((0, 2, 4, 6, 8), dims):
        ps += [(f"main.{i}.weight", (o, k)), (f"main.{i}.bias", (o,))]
    return ps, []


def make_net(prefix, spec, seed, dtype=torch.float32, jitter=True, wae_disc=False):
    """Deterministic parameters + fresh BN buffers for one sub-network.

    Weights follow VaeGan.init_parameters (models/vae_gan.py:252-264): U(-s, s), s = 1/sqrt(prod(shape[1:]))/sqrt(3),
    biases 0, BatchNorm gamma 1 / beta 0.  WaeDiscriminator: N(0, 0.0099999), bias 0 (models/vae_gan.py:522-525).
    ``jitter`` perturbs gamma/beta/biases so that tests exercise them (a trained net has non-trivial values there).
    """
    params, bns = spec
    g = torch.Generator().manual_seed(seed)
    P, S = OrderedDict(), OrderedDict()
    for name, shape in params:
        is_bn = any(name.startswith(b) for b, _ in bns)
        if len(shape) >= 2:
            if wae_disc:
                t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.0099999
            else:
                s = 1.0 / math.sqrt(math.prod(shape[1:])) / math.sqrt(3.0)
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * s
        elif is_bn and name.endswith("weight"):
            t = torch.ones(shape, dtype=torch.float64)
            if jitter:
                t = t + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            t = torch.zeros(shape, dtype=torch.float64)
            if jitter:
                t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        P[prefix + name] = t.to(dtype)
    for b, c in bns:
        S[prefix + b + "running_mean"] = torch.zeros(c, dtype=dtype)
        S[prefix + b + "running_var"] = torch.ones(c, dtype=dtype)
        S[prefix + b + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return P, S


def make_vaegan(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """encoder + decoder + discriminator of VaeGan (models/vae_gan.py:240-250)."""
    z = z or cfg["latent_dim"]
    P, S = OrderedDict(), OrderedDict()
    for i, (pre, spec) in enumerate((("encoder.", encoder_spec(cfg, z)), ("decoder.", decoder_spec(cfg, z)),
                                     ("discriminator.", discriminator_spec(cfg)))):
        p, s = make_net(pre, spec, seed + i, dtype, jitter)
        P.update(p)
        S.update(s)
    return P, S


def make_waegan(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """encoder + decoder + latent discriminator of WaeGan (models/vae_gan.py:440-450)."""
    z = z or cfg["latent_dim"]
    P, S = OrderedDict(), OrderedDict()
    for i, (pre, spec, wd) in enumerate((("encoder.", encoder_spec(cfg, z), False),
                                         ("decoder.", decoder_spec(cfg, z), False),
                                         ("discriminator.", wae_discriminator_spec(z), True))):
        p, s = make_net(pre, spec, seed + i, dtype, jitter, wae_disc=wd)
        P.update(p)
        S.update(s)
    return P, S


def bucket(P, prefix):
    """Parameter names of one optimizer bucket, in nn.Module.parameters() order."""
    return [k for k in P if k.startswith(prefix)]


# ---------------------------------------------------------------------------------------------------- synthetic data
def synthetic_images(B, size=64, seed=1234):
    """images in [-1, 1] (the reference normalises with mean = std = 0.5, configs/gan_config.py:40-41)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, size, size, generator=g) * 2 - 1


def synthetic_fmri(B, seed=1234):
    """Standardised voxel vectors zero-padded per ROI to rois_max (data_preprocessing/roi_extraction.py:128)."""
    g = torch.Generator().manual_seed(seed + 7)
    v = torch.randn(B, NUM_VOXELS, generator=g)
    o = 0
    for L in ROIS_MAX:
        keep = int(math.floor(0.85 * L))
        v[:, o + keep:o + L] = 0
        o += L
    return v


def synthetic_noise(B, z, seed=1234):
    g = torch.Generator().manual_seed(seed + 13)
    return torch.randn(B, z, generator=g), torch.randn(B, z, generator=g)


# ---------------------------------------------------------------------------------------------------- layers
def _tap(taps, key, t):
    """Record an intermediate tensor for the teacher-forced backward tests (tests/test_teacher_forced_gpu.py)."""
    if taps is not None:
        taps[key] = t.detach()


def _bn(P, S, pre, x, train, taps=None):
    """nn.BatchNorm{1,2}d(momentum=0.9): batch statistics in train mode, in-place running-stat update."""
    if taps is not None and train:
        dims = [d for d in range(x.dim()) if d != 1]
        _tap(taps, pre + "raw", x)
        _tap(taps, pre + "mean", x.mean(dims))
        _tap(taps, pre + "invstd", (x.var(dims, unbiased=False) + BN_EPS).rsqrt())
    y = F.batch_norm(x, S[pre + "running_mean"], S[pre + "running_var"], P[pre + "weight"], P[pre + "bias"], train,
                     BN_MOMENTUM, BN_EPS)
    if train:
        S[pre + "num_batches_tracked"] += 1
    return y


def _enc_block(P, S, pre, x, train, taps=None):
    """EncoderBlock.forward (models/vae_gan.py:23-35): conv 5x5 s2 p2 (no bias) -> BN -> ReLU; also returns the raw conv."""
    _tap(taps, pre + "in", x)
    raw = F.conv2d(x, P[pre + "conv.weight"], None, stride=2, padding=2)
    return F.relu(_bn(P, S, pre + "bn.", raw, train, taps)), raw


def encoder(P, S, x, cfg, train=True, pre="encoder.", taps=None):
    """Encoder.forward (models/vae_gan.py:87-93)."""
    h = x
    for i in range(3):
        h, _ = _enc_block(P, S, f"{pre}conv.{i}.", h, train, taps)
    h = h.reshape(len(h), -1)
    _tap(taps, pre + "fc.in", h)
    h = F.relu(_bn(P, S, pre + "fc.1.", F.linear(h, P[pre + "fc.0.weight"]), train, taps))
    _tap(taps, pre + "h", h)
    mu = F.linear(h, P[pre + "l_mu.weight"], P[pre + "l_mu.bias"])
    logvar = F.linear(h, P[pre + "l_var.weight"], P[pre + "l_var.bias"])
    return mu, logvar


def cognitive_encoder(P, S, v, train=True, pre="encoder.", taps=None):
    """CognitiveEncoder.forward (models/vae_gan.py:224-229)."""
    h = F.relu(_bn(P, S, pre + "fc1.1.", F.linear(v, P[pre + "fc1.0.weight"]), train, taps))
    _tap(taps, pre + "h", h)
    return (F.linear(h, P[pre + "l_mu.weight"], P[pre + "l_mu.bias"]),
            F.linear(h, P[pre + "l_var.weight"], P[pre + "l_var.bias"]))


def decoder(P, S, z, cfg, train=True, pre="decoder.", taps=None):
    """Decoder.forward (models/vae_gan.py:125-129) with DecoderBlock.forward (:56-60)."""
    _tap(taps, pre + "fc.in", z)
    h = F.relu(_bn(P, S, pre + "fc.1.", F.linear(z, P[pre + "fc.0.weight"]), train, taps))
    h = h.reshape(len(h), -1, cfg["fc_input"], cfg["fc_input"])
    for i in range(3):
        _tap(taps, f"{pre}conv.{i}.in", h)
        h = F.conv_transpose2d(h, P[f"{pre}conv.{i}.conv.weight"], None, stride=2, padding=2,
                               output_padding=1 if cfg["output_pad_dec"][i] else 0)
        h = F.relu(_bn(P, S, f"{pre}conv.{i}.bn.", h, train, taps))
    _tap(taps, pre + "conv.3.in", h)
    return torch.tanh(F.conv2d(h, P[pre + "conv.3.0.weight"], P[pre + "conv.3.0.bias"], stride=1, padding=2))


def discriminator(P, S, x_orig, x_pred, x_samp, cfg, mode="REC", train=True, recon_level=3, pre="discriminator.",
                  taps=None):
    """Discriminator.forward (models/vae_gan.py:163-183): "REC" returns the raw conv output of block `recon_level`
    flattened, anything else the sigmoid class score."""
    h = torch.cat((x_orig, x_pred, x_samp), 0)
    h = F.relu(F.conv2d(h, P[pre + "conv.0.0.weight"], P[pre + "conv.0.0.bias"], stride=cfg["stride_gan"], padding=2))
    for i in (1, 2, 3):
        h, raw = _enc_block(P, S, f"{pre}conv.{i}.", h, train, taps)
        if mode == "REC" and i == recon_level:
            return raw.reshape(len(raw), -1)
    h = h.reshape(len(h), -1)
    _tap(taps, pre + "fc.in", h)
    h = F.relu(_bn(P, S, pre + "fc.1.", F.linear(h, P[pre + "fc.0.weight"]), train, taps))
    _tap(taps, pre + "h", h)
    return torch.sigmoid(F.linear(h, P[pre + "fc.3.weight"], P[pre + "fc.3.bias"]))


def wae_discriminator(P, zs, pre="discriminator.", taps=None):
    """WaeDiscriminator.forward (models/vae_gan.py:510-529)."""
    h = zs
    for i in (0, 2, 4, 6):
        _tap(taps, f"{pre}main.{i}.in", h)
        h = F.relu(F.linear(h, P[f"{pre}main.{i}.weight"], P[f"{pre}main.{i}.bias"]))
    _tap(taps, pre + "h", h)
    return torch.sigmoid(F.linear(h, P[pre + "main.8.weight"], P[pre + "main.8.bias"]))


def reparameterize(mu, logvar, eps):
    """VaeGan.reparameterize (models/vae_gan.py:266-269) with the noise passed in."""
    return eps * torch.exp(0.5 * logvar) + mu


def vaegan_loss(x, x_tilde, dl_o, dl_p, dc_o, dc_p, dc_s, mu, logvar):
    """VaeGan.loss (models/vae_gan.py:302-320) == VaeGanCognitive.loss (:411-432)."""
    nle = 0.5 * (x.reshape(len(x), -1) - x_tilde.reshape(len(x_tilde), -1)) ** 2
    kl = -0.5 * torch.sum(-logvar.exp() - mu ** 2 + logvar + 1, 1)
    mse = torch.sum(0.5 * (dl_o - dl_p) ** 2, 1)
    bce_o = -torch.log(dc_o + 1e-3)
    bce_p = -torch.log(1 - dc_p + 1e-3)
    bce_s = -torch.log(1 - dc_s + 1e-3)
    return nle, kl, mse, bce_o, bce_p, bce_s


# ---------------------------------------------------------------------------------------------------- optimizers
def rmsprop_update(p, g, sq, lr, alpha=0.9, eps=1e-8):
    """torch.optim.RMSprop(alpha=0.9, eps=1e-8, momentum=0, centered=False) (train/train_vgan_stage1.py:275-283)."""
    sq = alpha * sq + (1 - alpha) * g * g
    return p - lr * g / (sq.sqrt() + eps), sq


def adam_update(p, g, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """torch.optim.Adam(betas=(0.5, 0.999)) (train/train_wae_stage1.py:221-224)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    return p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps), m, v


def gate(bce_o_mean, bce_p_mean, margin=0.35, equilibrium=0.68, train_dis_in=True):
    """Equilibrium gate (train/train_vgan_stage1.py:396-404). Returns (train_dis, train_dec). train_dis_in: the value the
    loss-mix block left in train_dis ('vae' mode sets it False, :387, before these lines run)."""
    train_dis, train_dec = train_dis_in, True
    if bce_o_mean < equilibrium - margin or bce_p_mean < equilibrium - margin:
        train_dis = False
    if bce_o_mean > equilibrium + margin or bce_p_mean > equilibrium + margin:
        train_dec = False
    if not train_dec and not train_dis:
        train_dis = train_dec = True
    return train_dis, train_dec


HP_VGAN = dict(lr=1e-4, alpha=0.9, eps=1e-8, lambda_mse=1e-6, margin=0.35, equilibrium=0.68)  # configs/gan_config.py:18-31
HP_WAE = dict(lr=1e-4, beta1=0.5, beta2=0.999, eps=1e-8)                                      # configs/wae_config.py:17


def _leaf(P):
    return OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in P.items())


# ---------------------------------------------------------------------------------------------------- Stage I VAE/GAN
def stage1_vaegan_step(P, S, x, eps, z_p, cfg=CFG64, hp=HP_VGAN, sq=None, update=True, force_gate=None, mode="vae-gan",
                       beta=1.0, taps=None):
    """One iteration of train/train_vgan_stage1.py:316-432 (mode 'vae-gan').

    Forward = VaeGan.forward train branch (models/vae_gan.py:276-286): encoder, reparameterize, decoder(z), decoder(z_p),
    discriminator "REC" then "GAN" on the concatenated 3B batch (two passes -> two BN running-stat updates).
    Backward: under torch-1.4 semantics the three optimizers consume g_enc = d loss_encoder / d encoder,
    g_dec = d loss_decoder / d decoder, g_dis = d loss_discriminator / d discriminator, all at the pre-step weights
    (verified bit-identical to the script's backward/step/zero_grad sequence).
    S (BN buffers) is updated in place; P is not modified -- the updated parameters are returned.
    """
    W = _leaf(P)
    B = len(x)
    tp = (lambda k: taps.setdefault(k, {})) if taps is not None else (lambda k: None)
    mu, logvar = encoder(W, S, x, cfg, taps=tp("enc"))
    z = reparameterize(mu, logvar, eps)
    x_tilde = decoder(W, S, z, cfg, taps=tp("dec1"))
    x_p = decoder(W, S, z_p, cfg, taps=tp("dec2"))
    disc_layer = discriminator(W, S, x, x_tilde, x_p, cfg, "REC")
    disc_class = discriminator(W, S, x, x_tilde, x_p, cfg, "GAN", taps=tp("dis"))
    dl_o, dl_p = disc_layer[:B], disc_layer[B:-B]
    dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
    nle, kl, mse, bce_o, bce_p, bce_s = vaegan_loss(x, x_tilde, dl_o, dl_p, dc_o, dc_p, dc_s, mu, logvar)
    if mode not in ("vae-gan", "beta-vae", "dcgan", "vae"):
        raise ValueError("mode must be one of the script's four loss mixes (train_vgan_stage1.py:359-388)")
    train_enc, pre_dis = True, True
    if mode in ("vae-gan", "beta-vae"):
        kl_w = beta / B if mode == "beta-vae" else 1.0                         # :360-362 (kld_weight = 1 / batch_size)
        loss_enc = kl.sum() * kl_w + mse.sum()                                 # :369 / :362
        loss_dis = bce_o.sum() + bce_p.sum() + bce_s.sum()                     # :370
        loss_dec = (hp["lambda_mse"] * mse).sum() - (1.0 - hp["lambda_mse"]) * loss_dis  # :372
    elif mode == "dcgan":                                                      # :374-380: pixel NLE, encoder not trained
        train_enc = False
        loss_enc = kl.sum() + nle.sum()
        loss_dis = bce_o.sum() + bce_s.sum()
        loss_dec = (hp["lambda_mse"] * nle).sum() - (1.0 - hp["lambda_mse"]) * loss_dis
    else:                                                                      # 'vae' :382-387
        loss_enc = kl.sum() + nle.sum()
        loss_dis = bce_o.sum() + bce_s.sum()
        loss_dec = (hp["lambda_mse"] * nle).sum()
        pre_dis = False                                                        # :387 train_dis = False before the gate
    train_dis, train_dec = gate(bce_o.mean().item(), bce_p.mean().item(), hp["margin"], hp["equilibrium"], pre_dis)
    if force_gate is not None:
        train_dis, train_dec = force_gate
    names = {b: bucket(W, b + ".") for b in ("encoder", "decoder", "discriminator")}
    grads = OrderedDict()
    for b, loss in (("encoder", loss_enc), ("decoder", loss_dec), ("discriminator", loss_dis)):
        if b == "encoder" and not train_enc:
            continue
        gs = torch.autograd.grad(loss, [W[n] for n in names[b]], retain_graph=True, allow_unused=True)
        grads.update((n, g if g is not None else torch.zeros_like(W[n])) for n, g in zip(names[b], gs))
    out = dict(x_tilde=x_tilde, x_p=x_p, disc_layer=disc_layer, disc_class=disc_class, mu=mu, logvar=logvar, z=z,
               nle=nle, kl=kl, mse=mse, bce_o=bce_o, bce_p=bce_p, bce_s=bce_s, loss_encoder=loss_enc,
               loss_decoder=loss_dec, loss_discriminator=loss_dis, train_dis=train_dis, train_dec=train_dec,
               grads=grads)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    if update:
        sq = sq if sq is not None else OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        newP, newsq = OrderedDict(P), OrderedDict(sq)
        active = dict(encoder=train_enc, decoder=train_dec, discriminator=train_dis)
        for b in names:
            if not active[b]:
                continue
            for n in names[b]:
                newP[n], newsq[n] = rmsprop_update(P[n], grads[n], sq[n], hp["lr"], hp["alpha"], hp["eps"])
        out["params"], out["square_avg"] = newP, newsq
    return out


# ---------------------------------------------------------------------------------------------------- Stage I WAE/GAN
def stage1_waegan_step(P, S, x, z_fake, cfg=CFG64, hp=HP_WAE, opt=None, step=1, update=True):
    """One iteration of train/train_wae_stage1.py:259-311.

    D-phase (:271-288): z_real = encoder(x) (encoder/decoder frozen), d_real / d_fake from the latent discriminator,
    L_fake = -10 sum log(d_fake + 1e-3), L_real = -10 sum log(1 - d_real + 1e-3), Adam(lr/2) on the discriminator.
    G-phase (:292-311): encoder forward AGAIN (second BN running update), x_recon = decoder(z_real), d_real from the
    UPDATED discriminator; L_rec = sum 0.5 (x_recon - x)^2, L_pen = -10 sum log(d_real + 1e-3); Adam on encoder
    (grad of L_rec + L_pen) and decoder (grad of L_rec).
    """
    names = {b: bucket(P, b + ".") for b in ("encoder", "decoder", "discriminator")}
    if opt is None:
        opt = dict(m=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()),
                   v=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()))
    W = _leaf(P)
    z_real, _ = encoder(W, S, x, cfg)
    d_real = wae_discriminator(W, z_real.detach())
    d_fake = wae_discriminator(W, z_fake)
    loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))       # :281
    loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))   # :282
    g_dis = torch.autograd.grad(loss_fake + loss_real, [W[n] for n in names["discriminator"]])
    grads = OrderedDict(zip(names["discriminator"], g_dis))
    newP, newm, newv = OrderedDict(P), OrderedDict(opt["m"]), OrderedDict(opt["v"])
    for n in names["discriminator"]:
        newP[n], newm[n], newv[n] = adam_update(P[n], grads[n], opt["m"][n], opt["v"][n], step, 0.5 * hp["lr"],
                                                hp["beta1"], hp["beta2"], hp["eps"])
    # generator phase, with the updated discriminator (or the old one when update=False, for gradient-only checks)
    W2 = _leaf(newP if update else P)
    z_real2, _ = encoder(W2, S, x, cfg)
    x_recon = decoder(W2, S, z_real2, cfg)
    d_real2 = wae_discriminator(W2, z_real2)
    loss_rec = torch.sum(torch.sum(0.5 * (x_recon - x) ** 2, 1))  # :301
    loss_pen = -10 * torch.sum(torch.log(d_real2 + 1e-3))        # :303
    # l_var receives no gradient (logvar is ignored, :296): p.grad stays None in the script and Adam skips it
    g_enc = torch.autograd.grad(loss_rec + loss_pen, [W2[n] for n in names["encoder"]], retain_graph=True,
                                allow_unused=True)
    g_dec = torch.autograd.grad(loss_rec, [W2[n] for n in names["decoder"]])
    grads.update((n, g) for n, g in zip(names["encoder"], g_enc) if g is not None)
    grads.update(zip(names["decoder"], g_dec))
    for b in ("encoder", "decoder"):
        for n in names[b]:
            if n not in grads:
                continue
            newP[n], newm[n], newv[n] = adam_update(P[n], grads[n], opt["m"][n], opt["v"][n], step, hp["lr"],
                                                    hp["beta1"], hp["beta2"], hp["eps"])
    out = dict(z_real=z_real, x_recon=x_recon, d_real=d_real, d_fake=d_fake, d_real_g=d_real2,
               loss_discriminator_fake=loss_fake, loss_discriminator_real=loss_real, loss_reconstruction=loss_rec,
               loss_penalty=loss_pen, grads=grads)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    if update:
        out["params"], out["adam"] = newP, dict(m=newm, v=newv)
    return out


def stage1_wae_mmd_step(P, S, x, z_fake, cfg=CFG64, hp=HP_WAE, opt=None, step=1, lambda_mmd=10.0, sigma2=0.25):
    """WAE-MMD variant of the Stage-I WAE step (BASELINE.json configs[1] names an MMD latent loss). PARITY UNPINNED: the
    reference has no MMD code; this is train/train_wae_stage1.py:292-311 with the latent-discriminator
    penalty replaced by lambda_mmd * B * MMD_IMQ(z_real, z_fake) (oracle/mmd.py; the factor B keeps the reference's
    batch-SUM loss convention, :301-303) and no discriminator phase: ONE encoder forward, L_rec = sum 0.5 (x_recon - x)^2,
    Adam(lr) on the encoder (grad of L_rec + L_pen; l_var receives none) and on the decoder (grad of L_rec)."""
    from .mmd import mmd_imq

    names = {b: bucket(P, b + ".") for b in ("encoder", "decoder")}
    if opt is None:
        opt = dict(m=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()),
                   v=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()))
    W = _leaf(P)
    z_real, _ = encoder(W, S, x, cfg)
    x_recon = decoder(W, S, z_real, cfg)
    loss_rec = torch.sum(torch.sum(0.5 * (x_recon - x) ** 2, 1))
    loss_pen = lambda_mmd * len(x) * mmd_imq(z_real, z_fake, sigma2)
    g_enc = torch.autograd.grad(loss_rec + loss_pen, [W[n] for n in names["encoder"]], retain_graph=True, allow_unused=True)
    g_dec = torch.autograd.grad(loss_rec, [W[n] for n in names["decoder"]])
    grads = OrderedDict((n, g) for n, g in zip(names["encoder"], g_enc) if g is not None)
    grads.update(zip(names["decoder"], g_dec))
    newP, newm, newv = OrderedDict(P), OrderedDict(opt["m"]), OrderedDict(opt["v"])
    for n in grads:
        newP[n], newm[n], newv[n] = adam_update(P[n], grads[n], opt["m"][n], opt["v"][n], step, hp["lr"], hp["beta1"],
                                                hp["beta2"], hp["eps"])
    out = dict(z_real=z_real, x_recon=x_recon, loss_reconstruction=loss_rec, loss_penalty=loss_pen, grads=grads)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    out["params"], out["adam"] = newP, dict(m=newm, v=newv)
    return out


# ---------------------------------------------------------------------------------------------------- Stages II / III
def make_cognitive(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """CognitiveEncoder + Decoder + Discriminator + teacher visual Encoder of VaeGanCognitive (models/vae_gan.py:323-345;
    train/train_vgan_stage2.py:210-232: decoder / discriminator ARE the teacher's modules, so their teacher_net.* duplicates
    in the state_dict are the same tensors). Keys: encoder.* (cognitive), decoder.*, discriminator.*, teacher_net.encoder.*."""
    z = z or cfg["latent_dim"]
    P, S = OrderedDict(), OrderedDict()
    for i, (pre, spec) in enumerate((("encoder.", cognitive_encoder_spec(z)), ("decoder.", decoder_spec(cfg, z)),
                                     ("discriminator.", discriminator_spec(cfg)),
                                     ("teacher_net.encoder.", encoder_spec(cfg, z)))):
        p, s = make_net(pre, spec, seed + 10 + i, dtype, jitter)
        if pre == "encoder.":  # CognitiveEncoder keeps torch's default Linear init (its init_parameters is commented out,
            g = torch.Generator().manual_seed(seed + 99)  # models/vae_gan.py:208-222): U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            for k, v in p.items():
                if v.dim() == 2:
                    b = 1.0 / math.sqrt(v.shape[1])
                    p[k] = ((torch.rand(v.shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
        P.update(p)
        S.update(s)
    return P, S


def cognitive_vaegan_step(P, S, fmri, image, eps, eps_t, z_p, stage, cfg=CFG64, hp=HP_VGAN, sq=None, update=True,
                          force_gate=None):
    """One iteration of train/train_vgan_stage2.py:321-407 (stage=2) or train/train_vgan_stage3.py:324-411 (stage=3).

    Forward = VaeGanCognitive.forward, mode 'vae' (models/vae_gan.py:361-393): cognitive encoder -> reparameterize ->
    decoder; stage 2 only: the "real" image is replaced by the teacher's reconstruction decoder(reparameterize(
    teacher.encoder(image))) (teacher encoder in train-mode BN, its parameters frozen); decoder(z_p); discriminator REC + GAN.
    Stage 2 trains {cognitive encoder, discriminator} (decoder frozen, no gate: train_dis=True, train_dec=False, :375-376);
    stage 3 trains {decoder, discriminator} with the gate (:382-388), encoder frozen. Both clamp gradients to [-1, 1]
    before the RMSprop step (stage2 :391,406; stage3 :402,410)."""
    W = _leaf(P)
    B = len(fmri)
    mu, logvar = cognitive_encoder(W, S, fmri, pre="encoder.")
    z = reparameterize(mu, logvar, eps)
    x_tilde = decoder(W, S, z, cfg)
    gt_x = image
    if stage == 2:
        mu_t, lv_t = encoder(W, S, image, cfg, pre="teacher_net.encoder.")
        gt_x = decoder(W, S, reparameterize(mu_t, lv_t, eps_t), cfg)
    x_p = decoder(W, S, z_p, cfg)
    disc_layer = discriminator(W, S, gt_x, x_tilde, x_p, cfg, "REC")
    disc_class = discriminator(W, S, gt_x, x_tilde, x_p, cfg, "GAN")
    dl_o, dl_p = disc_layer[:B], disc_layer[B:-B]
    dc_o, dc_p, dc_s = disc_class[:B], disc_class[B:-B], disc_class[-B:]
    nle, kl, mse, bce_o, bce_p, bce_s = vaegan_loss(gt_x, x_tilde, dl_o, dl_p, dc_o, dc_p, dc_s, mu, logvar)
    loss_enc = kl.sum() + mse.sum()
    loss_dis = bce_o.sum() + bce_p.sum() + bce_s.sum()
    loss_dec = (hp["lambda_mse"] * mse).sum() - (1.0 - hp["lambda_mse"]) * loss_dis
    if stage == 2:
        train_dis, train_dec = True, False
        trained = [("encoder", loss_enc), ("discriminator", loss_dis)]
    else:
        train_dis, train_dec = gate(bce_o.mean().item(), bce_p.mean().item(), hp["margin"], hp["equilibrium"])
        if force_gate is not None:
            train_dis, train_dec = force_gate
        trained = [("decoder", loss_dec), ("discriminator", loss_dis)]
    names = {b: bucket(W, b + ".") for b, _ in trained}
    grads = OrderedDict()
    for b, loss in trained:
        gs = torch.autograd.grad(loss, [W[n] for n in names[b]], retain_graph=True)
        grads.update(zip(names[b], gs))
    out = dict(gt_x=gt_x, x_tilde=x_tilde, x_p=x_p, disc_layer=disc_layer, disc_class=disc_class, mu=mu, logvar=logvar,
               kl=kl, mse=mse, bce_o=bce_o, bce_p=bce_p, bce_s=bce_s, nle=nle, loss_encoder=loss_enc,
               loss_decoder=loss_dec, loss_discriminator=loss_dis, train_dis=train_dis, train_dec=train_dec, grads=grads)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    if update:
        sq = sq if sq is not None else OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        newP, newsq = OrderedDict(P), OrderedDict(sq)
        active = dict(encoder=stage == 2, decoder=stage == 3 and train_dec, discriminator=train_dis)
        for b in names:
            if not active[b]:
                continue
            for n in names[b]:
                newP[n], newsq[n] = rmsprop_update(P[n], grads[n].clamp(-1, 1), sq[n], hp["lr"], hp["alpha"], hp["eps"])
        out["params"], out["square_avg"] = newP, newsq
    return out


# ---------------------------------------------------------------------------------------------------- WAE Stages II / III
def make_cognitive_wae(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """WaeGanCognitive (models/vae_gan.py:532-546) + the Stage-I WaeGan teacher's visual encoder
    (train/train_wae_stage2.py:195-203). Keys: encoder.* (cognitive), decoder.*, discriminator.main.*, teacher_net.encoder.*."""
    z = z or cfg["latent_dim"]
    P, S = OrderedDict(), OrderedDict()
    for i, (pre, spec, wd) in enumerate((("encoder.", cognitive_encoder_spec(z), False),
                                         ("decoder.", decoder_spec(cfg, z), False),
                                         ("discriminator.", wae_discriminator_spec(z), True),
                                         ("teacher_net.encoder.", encoder_spec(cfg, z), False))):
        p, s = make_net(pre, spec, seed + 20 + i, dtype, jitter, wae_disc=wd)
        P.update(p)
        S.update(s)
    return P, S


HP_WAE23 = dict(lr=1e-3, lr_dis=5e-4, beta1=0.5, beta2=0.999, eps=1e-8)  # train_wae_stage2.py:237-239 (hard-coded)


def cognitive_wae_step(P, S, fmri, image, stage, cfg=CFG64, hp=HP_WAE23, opt=None, step=1):
    """One iteration of train/train_wae_stage2.py:274-328 (stage=2) or train/train_wae_stage3.py:295-347 (stage=3).

    stage 2: an unused teacher reconstruction forward (:284-285, BN side effects only), D-phase on z_fake = cognitive
    encoder(fmri) vs z_real = teacher.encoder(image), Adam(5e-4) on the latent discriminator; G-phase trains the cognitive
    encoder with nn.MSELoss(x_recon, image) + (-10 * mean log(d + 1e-3)) (:320-321), Adam(1e-3); decoder frozen.
    stage 3: same D-phase; G-phase trains the DECODER on nn.MSELoss only (:338-347), cognitive encoder frozen."""
    names = {b: bucket(P, b + ".") for b in ("encoder", "decoder", "discriminator")}
    if opt is None:
        opt = dict(m=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()),
                   v=OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()))
    W = _leaf(P)
    if stage == 2:
        zt, _ = encoder(W, S, image, cfg, pre="teacher_net.encoder.")   # x_gt = decoder(teacher.encoder(image)): unused
        decoder(W, S, zt, cfg)
    z_fake, _ = cognitive_encoder(W, S, fmri, pre="encoder.")
    z_real, _ = encoder(W, S, image, cfg, pre="teacher_net.encoder.")
    d_real = wae_discriminator(W, z_real.detach())
    d_fake = wae_discriminator(W, z_fake.detach())
    loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))
    loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))
    g_dis = torch.autograd.grad(loss_fake + loss_real, [W[n] for n in names["discriminator"]])
    grads = OrderedDict(zip(names["discriminator"], g_dis))
    newP, newm, newv = OrderedDict(P), OrderedDict(opt["m"]), OrderedDict(opt["v"])
    for n in names["discriminator"]:
        newP[n], newm[n], newv[n] = adam_update(P[n], grads[n], opt["m"][n], opt["v"][n], step, hp["lr_dis"], hp["beta1"],
                                                hp["beta2"], hp["eps"])
    W2 = _leaf(newP)
    z2, _ = cognitive_encoder(W2, S, fmri, pre="encoder.")
    x_recon = decoder(W2, S, z2, cfg)
    d2 = wae_discriminator(W2, z2)
    loss_rec = F.mse_loss(x_recon, image)
    loss_pen = -10 * torch.mean(torch.log(d2 + 1e-3))
    if stage == 2:
        g = torch.autograd.grad(loss_rec + loss_pen, [W2[n] for n in names["encoder"]], allow_unused=True)
        trained = "encoder"
    else:
        g = torch.autograd.grad(loss_rec, [W2[n] for n in names["decoder"]])
        trained = "decoder"
    grads.update((n, gg) for n, gg in zip(names[trained], g) if gg is not None)
    for n in names[trained]:
        if n in grads:
            newP[n], newm[n], newv[n] = adam_update(P[n], grads[n], opt["m"][n], opt["v"][n], step, hp["lr"], hp["beta1"],
                                                    hp["beta2"], hp["eps"])
    out = dict(z_fake=z_fake, z_real=z_real, d_real=d_real, d_fake=d_fake, x_recon=x_recon, d_real_g=d2,
               loss_discriminator_fake=loss_fake, loss_discriminator_real=loss_real, loss_reconstruction=loss_rec,
               loss_penalty=loss_pen, grads=grads, trained=trained)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    out["params"], out["adam"] = newP, dict(m=newm, v=newv)
    return out


# ---------------------------------------------------------------------------------------------------- Stage III "dual" composite
def make_dual_stage3(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """BASELINE.json configs[3] ("Stage III cognitive WAE/Dual-GAN with two discriminators and fixed cognitive encoder").
    NOT a single reference script: make_cognitive()'s nets plus the latent WaeDiscriminator of
    WaeGanCognitive (models/vae_gan.py:542, kept N(0, 0.0099999) init) under the prefix latent_discriminator.*."""
    z = z or cfg["latent_dim"]
    P, S = make_cognitive(cfg, z, seed, dtype, jitter)
    p, s = make_net("latent_discriminator.", wae_discriminator_spec(z), seed + 40, dtype, jitter, wae_disc=True)
    P.update(p)
    S.update(s)
    return P, S


def dual_stage3_step(P, S, fmri, image, eps, z_p, cfg=CFG64, hp=HP_VGAN, hp_lat=HP_WAE23, sq=None, opt=None, step=1,
                     force_gate=None):
    """The composite step of configs[3]: the image-side update of train/train_vgan_stage3.py:324-411
    (cognitive_vaegan_step(stage=3): decoder + image discriminator trained with the gate and the gradient clamp, cognitive
    encoder fixed) PLUS the latent-discriminator phase of train/train_wae_stage3.py:308-326 on
    z_fake = cognitive_encoder(fmri) (the same forward, so its BN statistics are updated once) and
    z_real = teacher.encoder(image) (train-mode BN, frozen): L_fake = -10 sum log(d(z_fake) + 1e-3),
    L_real = -10 sum log(1 - d(z_real) + 1e-3), Adam(5e-4, betas (0.5, 0.999)) on the latent discriminator.
    Each half is pinned to the reference through its own golden fixture (stage3_cognitive_*, stage3_cognitive_wae_*)."""
    out = cognitive_vaegan_step(P, S, fmri, image, eps, None, z_p, 3, cfg, hp, sq=sq, force_gate=force_gate)
    names = bucket(P, "latent_discriminator.")
    if opt is None:
        opt = dict(m=OrderedDict((k, torch.zeros_like(P[k])) for k in names),
                   v=OrderedDict((k, torch.zeros_like(P[k])) for k in names))
    W = _leaf(P)
    z_fake = out["mu"]
    z_real, _ = encoder(W, S, image, cfg, pre="teacher_net.encoder.")
    d_real = wae_discriminator(W, z_real.detach(), pre="latent_discriminator.")
    d_fake = wae_discriminator(W, z_fake.detach(), pre="latent_discriminator.")
    loss_fake = -10 * torch.sum(torch.log(d_fake + 1e-3))
    loss_real = -10 * torch.sum(torch.log(1 - d_real + 1e-3))
    g = torch.autograd.grad(loss_fake + loss_real, [W[n] for n in names])
    newP, newm, newv = out["params"], OrderedDict(opt["m"]), OrderedDict(opt["v"])
    for n, gg in zip(names, g):
        out["grads"][n] = gg.detach()
        newP[n], newm[n], newv[n] = adam_update(P[n], gg.detach(), opt["m"][n], opt["v"][n], step, hp_lat["lr_dis"],
                                                hp_lat["beta1"], hp_lat["beta2"], hp_lat["eps"])
    out.update(z_real=z_real.detach(), d_real=d_real.detach(), d_fake=d_fake.detach(),
               loss_discriminator_fake=loss_fake.detach(), loss_discriminator_real=loss_real.detach(),
               adam=dict(m=newm, v=newv))
    return out


# ---------------------------------------------------------------------------------------------------- dual WAE/GAN Stage I
def make_dual_stage1(cfg=CFG64, z=None, seed=12345, dtype=torch.float32, jitter=True):
    """VaeGan (encoder / decoder / image discriminator) plus the latent WaeDiscriminator of a second WaeGan model
    (train/wae_vgan_stage1.py:196-200) under the prefix latent_discriminator.*."""
    z = z or cfg["latent_dim"]
    P, S = make_vaegan(cfg, z, seed, dtype, jitter)
    p, s = make_net("latent_discriminator.", wae_discriminator_spec(z), seed + 50, dtype, jitter, wae_disc=True)
    P.update(p)
    S.update(s)
    return P, S


def dual_stage1_step(P, S, x, eps, z_p, z_fake, cfg=CFG64, hp=HP_VGAN, lam=1.0, sq=None, force_gate=None):
    """One iteration of train/wae_vgan_stage1.py:282-441 (mode 'vae-gan'), TORCH >= 2 SEMANTICS (SURVEY.md 8a row a16:
    `zero_grad()` sets gradients to None, so the decoder `step()` of :417 on never-written gradients is a no-op; torch 1.4
    decayed the decoder's square_avg there).

      (1) the Stage-I VAE/GAN forward and losses (:290-366, = stage1_vaegan_step);
      (2) latent discriminator phase (:380-397): z_real = encoder(x) (second encoder forward, frozen), L_fake = -lam sum
          log(d(z_fake) + 1e-3), L_real = -lam sum log(1 - d(z_real) + 1e-3), RMSprop on the latent discriminator;
      (3) penalty (:401-417): third encoder forward, an unused decoder(z_real) forward (BatchNorm side effects), d from the
          UPDATED latent discriminator, L_pen = -lam sum log(d + 1e-3) backpropagated into the encoder only;
      (4) the three Stage-I updates (:419-441) with the penalty gradient ACCUMULATED into the encoder's (:421).
    All gradients are taken at the pre-step weights of encoder / decoder / image discriminator (torch-1.4 order)."""
    vg = {k: v for k, v in P.items() if not k.startswith("latent_discriminator.")}
    out = stage1_vaegan_step(vg, S, x, eps, z_p, cfg, hp, update=False, force_gate=force_gate)
    lat = bucket(P, "latent_discriminator.")
    sq = sq if sq is not None else OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    W = _leaf(P)
    # (2)
    z_real, _ = encoder(W, S, x, cfg)
    d_real = wae_discriminator(W, z_real.detach(), pre="latent_discriminator.")
    d_fake = wae_discriminator(W, z_fake, pre="latent_discriminator.")
    loss_fake = -lam * torch.sum(torch.log(d_fake + 1e-3))
    loss_real = -lam * torch.sum(torch.log(1 - d_real + 1e-3))
    g_lat = torch.autograd.grad(loss_fake + loss_real, [W[n] for n in lat])
    newP, newsq = OrderedDict(P), OrderedDict(sq)
    for n, g in zip(lat, g_lat):
        out["grads"][n] = g.detach()
        newP[n], newsq[n] = rmsprop_update(P[n], g.detach(), sq[n], hp["lr"], hp["alpha"], hp["eps"])
    # (3)
    W2 = _leaf(newP)
    z2, _ = encoder(W2, S, x, cfg)
    decoder(W2, S, z2, cfg)                                   # x_recon: unused, BatchNorm running statistics only
    d2 = wae_discriminator(W2, z2, pre="latent_discriminator.")
    loss_pen = -lam * torch.sum(torch.log(d2 + 1e-3))
    enc = bucket(P, "encoder.")
    g_pen = torch.autograd.grad(loss_pen, [W2[n] for n in enc], allow_unused=True)
    # (4)
    for n, g in zip(enc, g_pen):
        if g is not None:
            out["grads"][n] = out["grads"][n] + g.detach()
    active = dict(encoder=True, decoder=out["train_dec"], discriminator=out["train_dis"])
    for b in ("encoder", "decoder", "discriminator"):
        if not active[b]:
            continue
        for n in bucket(P, b + "."):
            newP[n], newsq[n] = rmsprop_update(P[n], out["grads"][n], sq[n], hp["lr"], hp["alpha"], hp["eps"])
    out.update(z_real=z_real.detach(), d_real=d_real.detach(), d_fake=d_fake.detach(), d_real_g=d2.detach(),
               loss_discriminator_fake=loss_fake.detach(), loss_discriminator_real=loss_real.detach(),
               loss_penalty=loss_pen.detach(), params=newP, square_avg=newsq)
    return out
