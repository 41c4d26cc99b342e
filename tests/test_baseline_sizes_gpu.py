"""Whole-step parity against the CPU oracle AT THE BATCH SIZES BASELINE.json NAMES ("What's weak" #2): Stage-I
VAE/GAN at 64 (configs[0]) and 256, Stage-I WAE/GAN at 256 (configs[1]), Stage-II cognitive VAE/GAN at 256 (configs[2]),
Stage-III cognitive VAE/GAN and the configs[3] composite at 512. At these sizes every tensor-path launch runs the
persistent / wave-split / parity-merged kernels (>= 296 tiles), so this is the oracle -- not the library itself -- checking
the large-batch code paths end to end. The fp32 oracle runs once per case (seconds to ~1 minute on the host cores) and both
compute dtypes are compared with it.

Tolerances (rel-L2 per tensor vs the fp32 oracle): bf16 tensor path -- forward tensors, per-sample losses, loss sums,
BatchNorm buffers 2e-2; fp32 exact path -- 1e-4. End-to-end gradient buckets: fp32 path 5e-3 (single ReLU-mask flips,
SURVEY.md 0-9); bf16 path REPORTED with the mask-flip explanation and bounded at 0.5 -- the bf16 backward kernels themselves
are certified at 2e-2 by tests/test_teacher_forced_gpu.py (T2) and tests/test_fullsize_kernels_gpu.py.
"""
import json
import os

import pytest
import torch

from oracle import vaegan as O
from thesis_fmri_reconstruction_b200 import engine, hp

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nchw_flat(raw_nhwc):
    return raw_nhwc.float().permute(0, 3, 1, 2).reshape(raw_nhwc.shape[0], -1)


def _buckets(grads, ref_grads):
    out = {}
    for b in sorted({k.split(".")[0] + "." for k in ref_grads}):
        ks = [k for k in ref_grads if k.startswith(b) and k in grads]
        if ks:
            out[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref_grads[k].reshape(-1) for k in ks]))
    return out


def _buffers(tr, S_ref):
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    return max(berr.items(), key=lambda t: t[1]), nbt_ok


def _report(name, rep):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_{name}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))


def _check(rep, adt):
    ft, gt = (1e-4, 5e-3) if adt == F32 else (2e-2, 0.5)
    assert max(rep["forward"].values()) < ft, rep["forward"]
    assert max(rep["grad_bucket"].values()) < gt, rep["grad_bucket"]
    assert rep["bn_worst"][1] < ft and rep["nbt_ok"], rep["bn_worst"]


@pytest.mark.parametrize("B", [64, 256])
def test_stage1_vaegan_baseline_batch(B):
    seed = 6400 + B
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_vaegan_step(P, S_ref, x, eps, z_p)
    for adt in (BF, F32):
        tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, adt)
        out = tr.forward_backward(x.cuda(), eps.cuda(), z_p.cuda())
        grads = {k: v.clone() for k, v in tr.named_grads().items()}
        tr.update(B)
        torch.cuda.synchronize()
        lo = tr.losses()
        fwd = dict(mu=rel(out["mu"], ref["mu"]), logvar=rel(out["logvar"], ref["logvar"]), x_tilde=rel(out["x_tilde"], ref["x_tilde"]),
                   x_p=rel(out["x_p"], ref["x_p"]), disc_layer=rel(nchw_flat(out["disc_layer_nhwc"]), ref["disc_layer"]),
                   disc_class=rel(out["disc_class"], ref["disc_class"].reshape(-1)), kl=rel(out["kl"], ref["kl"]),
                   mse=rel(out["mse"], ref["mse"]),
                   bce=rel(out["bce"], torch.cat([ref["bce_o"], ref["bce_p"], ref["bce_s"]]).reshape(-1)))
        for k in ("loss_encoder", "loss_decoder", "loss_discriminator"):
            fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
        bn_worst, nbt_ok = _buffers(tr, S_ref)
        rep = dict(case=f"Stage-I VAE/GAN B={B}", dtype=str(adt), forward=fwd, grad_bucket=_buckets(grads, ref["grads"]),
                   gate_ok=(lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"]), bn_worst=bn_worst,
                   nbt_ok=nbt_ok)
        _report(f"baseline_stage1_B{B}_{str(adt).split('.')[-1]}", rep)
        _check(rep, adt)
        assert rep["gate_ok"]
        del tr, out, grads


def test_stage1_waegan_baseline_batch():
    B, seed = 256, 2560
    P, S = O.make_waegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    z_fake = O.synthetic_noise(B, 128, seed=seed)[0] * 0.5
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_waegan_step(P, S_ref, x, z_fake)
    for adt in (BF, F32):
        tr = engine.WaeGanStage1(P, S, hp.CFG64, 128, adt)
        out = tr.step(x.cuda(), z_fake.cuda())
        torch.cuda.synchronize()
        lo = tr.losses()
        fwd = {k: rel(out[k], ref[k].reshape(out[k].shape)) for k in ("z_real", "x_recon", "d_real", "d_fake", "d_real_g")}
        for k in ("loss_discriminator_fake", "loss_discriminator_real", "loss_reconstruction", "loss_penalty"):
            fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
        bn_worst, nbt_ok = _buffers(tr, S_ref)
        rep = dict(case=f"Stage-I WAE/GAN B={B}", dtype=str(adt), forward=fwd, grad_bucket=_buckets(tr.named_grads(), ref["grads"]),
                   bn_worst=bn_worst, nbt_ok=nbt_ok)
        _report(f"baseline_wae1_B{B}_{str(adt).split('.')[-1]}", rep)
        _check(rep, adt)
        del tr, out


@pytest.mark.parametrize("stage,B", [(2, 256), (3, 512)])
def test_cognitive_vaegan_baseline_batch(stage, B):
    seed = 900 + B
    P, S = O.make_cognitive(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    eps_t = O.synthetic_noise(B, 128, seed=seed + 1)[0]
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.cognitive_vaegan_step(P, S_ref, fmri, image, eps, eps_t, z_p, stage)
    if stage == 3:
        P = {k: v for k, v in P.items() if not k.startswith("teacher_net.")}
        S = {k: v for k, v in S.items() if not k.startswith("teacher_net.")}
    for adt in (BF, F32):
        tr = engine.VaeGanCognitiveStage(P, S, hp.CFG64, stage, 128, adt)
        out = tr.forward_backward(fmri.cuda(), image.cuda(), eps.cuda(), eps_t.cuda(), z_p.cuda())
        grads = {k: v.clone() for k, v in tr.named_grads().items()}
        tr.update(B)
        torch.cuda.synchronize()
        lo = tr.losses()
        fwd = dict(mu=rel(out["mu"], ref["mu"]), x_tilde=rel(out["x_tilde"], ref["x_tilde"]), gt_x=rel(out["gt_x"], ref["gt_x"]),
                   disc_layer=rel(nchw_flat(out["disc_layer_nhwc"]), ref["disc_layer"]),
                   disc_class=rel(out["disc_class"], ref["disc_class"].reshape(-1)), kl=rel(out["kl"], ref["kl"]),
                   mse=rel(out["mse"], ref["mse"]))
        for k in ("loss_encoder", "loss_decoder", "loss_discriminator"):
            fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
        bn_worst, nbt_ok = _buffers(tr, S_ref)
        rep = dict(case=f"Stage-{stage} cognitive VAE/GAN B={B}", dtype=str(adt), forward=fwd,
                   grad_bucket=_buckets(grads, ref["grads"]),
                   gate_ok=(lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"]), bn_worst=bn_worst,
                   nbt_ok=nbt_ok)
        _report(f"baseline_stage{stage}_B{B}_{str(adt).split('.')[-1]}", rep)
        _check(rep, adt)
        assert rep["gate_ok"]
        del tr, out, grads


def test_stage3_dual_baseline_batch():
    """BASELINE.json configs[3] (composite, engine.DualCognitiveStage3) at its batch 512, bf16 tensor path."""
    B, seed = 512, 5120
    P, S = O.make_dual_stage3(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.dual_stage3_step(P, S_ref, fmri, image, eps, z_p)
    adt = BF
    tr = engine.DualCognitiveStage3(P, S, hp.CFG64, 128, adt)
    out = tr.forward_backward(fmri.cuda(), image.cuda(), eps.cuda(), z_p.cuda())
    grads = {k: v.clone() for k, v in tr.named_grads().items()}
    tr.update(B)
    torch.cuda.synchronize()
    lo = tr.losses()
    fwd = dict(mu=rel(out["mu"], ref["mu"]), x_tilde=rel(out["x_tilde"], ref["x_tilde"]),
               disc_layer=rel(nchw_flat(out["disc_layer_nhwc"]), ref["disc_layer"]),
               disc_class=rel(out["disc_class"], ref["disc_class"].reshape(-1)), mse=rel(out["mse"], ref["mse"]),
               z_real=rel(out["z_real"], ref["z_real"]), d_real=rel(out["d_real"], ref["d_real"].reshape(-1)),
               d_fake=rel(out["d_fake"], ref["d_fake"].reshape(-1)))
    for k in ("loss_decoder", "loss_discriminator", "loss_discriminator_fake", "loss_discriminator_real"):
        fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    bn_worst, nbt_ok = _buffers(tr, S_ref)
    rep = dict(case=f"configs[3] dual Stage III B={B}", dtype=str(adt), forward=fwd, grad_bucket=_buckets(grads, ref["grads"]),
               gate_ok=(lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"]), bn_worst=bn_worst, nbt_ok=nbt_ok)
    _report(f"baseline_stage3_dual_B{B}_bfloat16", rep)
    _check(rep, adt)
    assert rep["gate_ok"]
