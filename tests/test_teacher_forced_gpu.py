"""T2 of SURVEY.md 8c: TEACHER-FORCED BACKWARD. The bf16 tensor-path backward kernels are fed the ORACLE's saved forward
state -- every layer input, every pre-BatchNorm activation, the BatchNorm batch statistics, the ReLU masks, the latent
heads, the discriminator scores -- and the three gradient buckets they produce are compared, per tensor, with the gradients
of the pinned oracle itself (oracle.stage1_vaegan_step in fp64; that function is pinned to the unmodified reference by
tests/golden/*, tests/test_oracle_cpu.py). Bound: 2e-2 rel-L2 PER TENSOR (north_star's bf16 tolerance).

Why teacher-forced: an end-to-end bf16 gradient differs from the fp32 reference by 0.06..0.2 because bf16 storage of the
forward activations flips ~4e-4 of the ReLU masks per layer (the reference's own modules under CPU bf16
autocast show 0.06 / 0.21 / 0.24); that number says nothing about the backward kernels. Here the forward state is the
oracle's on both sides, so everything that remains is the backward arithmetic: bf16 operand rounding of activations /
weights / gradients, fp32 accumulation order, the BatchNorm-backward reductions, the masks as the kernels recompute them.

How the state is fed: the engine's own forward builds the saved-state objects (descriptors, packs, workspaces), then every
saved tensor is OVERWRITTEN in place with the oracle's fp32 tensor, converted to the layout / dtype the kernels store
(NHWC bf16 activations, fp32 linear pre-activations). The kernels recompute a BatchNorm+ReLU mask from the stored pre-BN
value; to feed the oracle's MASK, the few stored values whose bf16 rounding would change the sign of bn(x) (~4e-4 of a
layer) are moved to the neighbouring bf16 value on the oracle's side of the threshold (`consistent_round`; the count is
reported).
"""
import json
import os

import pytest
import torch

from oracle import vaegan as O
from thesis_fmri_reconstruction_b200 import engine, hp
from thesis_fmri_reconstruction_b200 import lib as L
from thesis_fmri_reconstruction_b200 import nets as NN

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _chan_view(v, ndim):
    shape = [1] * ndim
    shape[1] = -1
    return v.view(shape)


def consistent_round(raw, mean, invstd, gamma, beta):
    """raw: fp32 [N, C, ...] (oracle layout, channel dim 1) on the GPU. Returns (bf16 tensor, #moved): round to bf16, then
    move the elements whose ReLU decision bn(x) > 0 changed by the rounding one bf16 step back across the threshold."""
    n = raw.dim()
    sc, m, b = _chan_view(gamma * invstd, n), _chan_view(mean, n), _chan_view(beta, n)
    want = ((raw - m) * sc + b) > 0
    r = raw.to(BF)
    moved = 0
    for it in range(6):
        got = ((r.float() - m) * sc + b) > 0
        bad = got != want
        nb = int(bad.sum())
        if it == 0:
            moved = nb
        if nb == 0:
            break
        up = (want == (sc > 0))                       # the value must increase
        i = r.view(torch.int16).to(torch.int32)
        pos = r.float() > 0
        neg = r.float() < 0
        step = torch.where(up == pos, 1, -1)          # positive numbers grow with the bit pattern, negative ones shrink
        i2 = i + step
        i2 = torch.where(~pos & ~neg, torch.where(up, 0x0080, 0x8080 - 65536), i2)   # +-0 -> smallest normal
        i = torch.where(bad, i2, i)
        r = i.to(torch.int16).view(BF)
    assert int((((r.float() - m) * sc + b) > 0).ne(want).sum()) == 0
    return r, moved


def put(dst, src):
    """Overwrite a saved kernel-side tensor with an oracle tensor (NCHW -> NHWC when 4-D)."""
    s = src.cuda() if not src.is_cuda else src
    if dst.dim() == 4 and s.dim() == 4 and dst.shape != s.shape:
        s = s.permute(0, 2, 3, 1)
    dst.copy_(s.reshape(dst.shape).to(dst.dtype))


def put_flat(dst, src, chw):
    """A flattened conv output feeding an fc layer: the oracle flattens NCHW, the kernels keep (h, w, c) columns when the
    layer's packs are permuted (nets.LinearOp chw)."""
    C_, h, w = chw
    s = src.cuda() if not src.is_cuda else src
    dst.copy_(s.reshape(-1, C_, h, w).permute(0, 2, 3, 1).reshape(dst.shape).to(dst.dtype))


class Forcer:
    def __init__(self, P):
        self.P = {k: v.cuda() for k, v in P.items()}
        self.moved, self.total = 0, 0
        self.enc_chw = self.dis_chw = None   # set by the test from nets.*.fc.lin.chw

    def bn(self, c, taps, pre):
        """c: nets BatchNorm ctx (raw, mean, invstd); taps[pre + 'raw' / 'mean' / 'invstd'] from the oracle."""
        raw, mean, invstd = (taps[pre + k].float().cuda() for k in ("raw", "mean", "invstd"))
        put(c.mean, mean)
        put(c.invstd, invstd)
        if c.raw.dtype == BF:
            r, moved = consistent_round(raw, mean, invstd, self.P[pre + "weight"], self.P[pre + "bias"])
            self.moved += moved
            self.total += raw.numel()
            put(c.raw, r)
        else:
            put(c.raw, raw)

    def encoder(self, ce, taps, pre="encoder."):
        self.bn(ce.c0, taps, pre + "conv.0.bn.")
        for i, c in enumerate(ce.blocks, start=1):
            put(c.x, taps[f"{pre}conv.{i}.in"])
            self.bn(c.bn, taps, f"{pre}conv.{i}.bn.")
        if self.enc_chw is not None:
            put_flat(ce.fc.x, taps[pre + "fc.in"], self.enc_chw)
        else:
            put(ce.fc.x, taps[pre + "fc.in"])
        self.bn(ce.fc.bn, taps, pre + "fc.1.")
        put(ce.heads.h, taps[pre + "h"])

    def decoder(self, cd, taps, img, pre="decoder."):
        put(cd.fc.x, taps[pre + "fc.in"])
        self.bn(cd.fc.bn, taps, pre + "fc.1.")
        for i, c in enumerate(cd.blocks):
            put(c.x, taps[f"{pre}conv.{i}.in"])
            self.bn(c.bn, taps, f"{pre}conv.{i}.bn.")
        put(cd.a3, taps[pre + "conv.3.in"])
        put(cd.img, img)

    def discriminator(self, cc, taps, imgs, p, pre="discriminator."):
        for dst, src in zip(cc.imgs[1:], imgs[1:]):
            put(dst, src)
        put(cc.y0, taps[pre + "conv.1.in"])
        if cc.mask0 is not None:
            OH, OW = cc.hw0
            L.relu_bitmask(cc.y0, cc.N * OH * OW, cc.y0.shape[-1], cc.mask0)
        for i, c in enumerate(cc.blocks, start=1):
            if i > 1:
                put(c.x, taps[f"{pre}conv.{i}.in"])
            self.bn(c.bn, taps, f"{pre}conv.{i}.bn.")
        if self.dis_chw is not None:
            put_flat(cc.fc.x, taps[pre + "fc.in"], self.dis_chw)
        else:
            put(cc.fc.x, taps[pre + "fc.in"])
        self.bn(cc.fc.bn, taps, pre + "fc.1.")
        put(cc.hfc, taps[pre + "h"])
        put(cc.p, p.reshape(-1))


@pytest.mark.parametrize("B", [16, 64])
def test_stage1_teacher_forced_backward_bf16(B):
    seed = 2718
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    taps = {}
    ref32 = O.stage1_vaegan_step(P, {k: v.clone() for k, v in S.items()}, x, eps, z_p, update=False, taps=taps)
    P64 = {k: v.double() for k, v in P.items()}
    S64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in S.items()}
    ref = O.stage1_vaegan_step(P64, S64, x.double(), eps.double(), z_p.double(), update=False)

    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, BF)
    st = tr.forward(x.cuda(), eps.cuda(), z_p.cuda())
    f = Forcer(P)
    f.enc_chw, f.dis_chw = tr.enc.fc.lin.chw, tr.dis.fc.lin.chw
    f.encoder(st.ce, taps["enc"])
    put(st.mu, ref32["mu"])
    put(st.lv, ref32["logvar"])
    f.decoder(st.cd1, taps["dec1"], ref32["x_tilde"])
    f.decoder(st.cd2, taps["dec2"], ref32["x_p"])
    f.discriminator(st.cc, taps["dis"], [x, ref32["x_tilde"], ref32["x_p"]], ref32["disc_class"])
    assert st.raw3.data_ptr() == st.cc.blocks[2].bn.raw.data_ptr()   # the feature tap IS block 3's saved pre-BN tensor
    tr.backward(st)
    torch.cuda.synchronize()
    grads = tr.named_grads()
    # primary reference: the fp32 oracle run whose state (and masks) was fed; the fp64 run is reported beside it (the two
    # oracle precisions differ by single mask flips)
    per = {k: rel(grads[k], ref32["grads"][k]) for k in grads}
    buckets, buckets64 = {}, {}
    for b in ("encoder.", "decoder.", "discriminator."):
        ks = [k for k in grads if k.startswith(b)]
        ours = torch.cat([grads[k].reshape(-1) for k in ks])
        buckets[b] = rel(ours, torch.cat([ref32["grads"][k].reshape(-1) for k in ks]))
        buckets64[b] = rel(ours, torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    noise = {}
    for b in ("encoder.", "decoder.", "discriminator."):   # the oracle's own fp32-vs-fp64 deviation, for scale
        ks = [k for k in grads if k.startswith(b)]
        noise[b] = rel(torch.cat([ref32["grads"][k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    worst = max(per.items(), key=lambda t: t[1])
    rep = dict(test="T2 teacher-forced backward, Stage-I VAE/GAN, bf16 tensor path", B=B, tolerance=TOL,
               masks_moved=f.moved, mask_elements=f.total, bucket_rel_l2=buckets, bucket_rel_l2_vs_fp64_oracle=buckets64,
               oracle_fp32_vs_fp64=noise,
               worst_tensor=worst, per_tensor=per)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_T2_stage1_B{B}_bf16.json", "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps({k: v for k, v in rep.items() if k != "per_tensor"}, indent=1))
    print("per tensor:", {k: f"{v:.2e}" for k, v in sorted(per.items(), key=lambda t: -t[1])[:12]})
    assert max(buckets.values()) < TOL, buckets
    assert worst[1] < TOL, worst


def test_cognitive_encoder_teacher_forced_backward_bf16():
    """CognitiveEncoder (fMRI voxel MLP, K = 3620): backward kernels on the oracle's saved state vs the oracle's autograd."""
    B, z, seed = 64, 128, 99
    spec = O.cognitive_encoder_spec(z)
    P, S = O.make_net("encoder.", spec, seed)
    v = O.synthetic_fmri(B, seed=seed)
    g = torch.Generator().manual_seed(seed)
    dmu, dlv = torch.randn(B, z, generator=g), torch.randn(B, z, generator=g)
    taps = {}
    W = {k: t.double().requires_grad_(True) for k, t in P.items()}
    S64 = {k: (t.double() if t.dtype.is_floating_point else t.clone()) for k, t in S.items()}
    mu, lv = O.cognitive_encoder(W, S64, v.double(), taps=taps)
    names = list(W)
    gref = dict(zip(names, torch.autograd.grad((mu * dmu.double()).sum() + (lv * dlv.double()).sum(), [W[n] for n in names])))
    net = NN.CognitiveEncoderNet(spec[0][0][1][1], z, BF)
    Pd = {k[len("encoder."):]: t.cuda() for k, t in P.items()}
    Sd = {k[len("encoder."):]: t.cuda() for k, t in S.items()}
    net.refresh(Pd)
    ycat, c = net.forward(Pd, Sd, v.cuda(), True, 1, {})
    f = Forcer(P)
    xs = torch.zeros_like(c.fc.x)
    xs[:, :v.shape[1]] = v.cuda().to(BF)
    c.fc.x.copy_(xs)
    f.bn(c.fc.bn, taps, "encoder.fc1.1.")
    put(c.heads.h, taps["encoder.h"])
    G = {k: torch.zeros_like(t) for k, t in Pd.items()}
    dycat = torch.cat([dmu, dlv], 1).cuda().to(BF)
    net.backward(Pd, c, dycat, G, False, True, True)
    torch.cuda.synchronize()
    # the upstream gradient the kernels saw is the bf16-rounded one: compare against the same
    per = {k: rel(G[k[len("encoder."):]], gref[k]) for k in names}
    print("cognitive encoder T2 per tensor:", {k: f"{e:.2e}" for k, e in per.items()})
    assert max(per.values()) < TOL, per


def test_wae_discriminator_teacher_forced_backward_bf16():
    """WaeDiscriminator (5-layer latent MLP): backward kernels on the oracle's saved activations vs the oracle's autograd."""
    B, z, seed = 256, 128, 77
    P, _ = O.make_net("discriminator.", O.wae_discriminator_spec(z), seed, wae_disc=True)
    P = {k: (t * 8 if k.endswith("weight") else t) for k, t in P.items()}   # N(0, 0.08): non-degenerate activations
    g = torch.Generator().manual_seed(seed)
    zin, gp = torch.randn(B, z, generator=g), torch.randn(B, generator=g)
    taps = {}
    W = {k: t.double().requires_grad_(True) for k, t in P.items()}
    zl = zin.double().requires_grad_(True)
    p = O.wae_discriminator(W, zl, taps=taps)
    names = list(W)
    gs = torch.autograd.grad((p.reshape(-1) * gp.double()).sum(), [W[n] for n in names] + [zl])
    gref, dz_ref = dict(zip(names, gs[:-1])), gs[-1]
    net = NN.WaeDiscriminatorNet(z, BF)
    Pd = {k[len("discriminator."):]: t.cuda() for k, t in P.items()}
    net.refresh(Pd)
    pk, c = net.forward(Pd, zin.cuda())
    for i, key in enumerate(("main.0.in", "main.2.in", "main.4.in", "main.6.in")):
        put(c.acts[i], taps["discriminator." + key])
    put(c.acts[4], taps["discriminator.h"])
    put(c.p, p.detach().reshape(-1))
    G = {k: torch.zeros_like(t) for k, t in Pd.items()}
    dz = net.backward(Pd, c, gp.cuda(), G, False, True, True)
    torch.cuda.synchronize()
    per = {k: rel(G[k[len("discriminator."):]], gref[k]) for k in names}
    per["dz"] = rel(dz, dz_ref)
    print("WAE discriminator T2 per tensor:", {k: f"{e:.2e}" for k, e in per.items()})
    assert max(per.values()) < TOL, per
