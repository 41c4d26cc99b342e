"""Whole-step parity of the fused Stage-I VAE/GAN engine (CUDA kernels behind the C ABI) against the CPU oracle on the
same seeded inputs and weights.

Tolerances (rel-L2 per tensor, vs the fp32 oracle on CPU):
  fp32 exact path : forward tensors / losses 1e-4; BN buffers 1e-4; gradient buckets vs the fp64 oracle:
                    max(5e-3, 3 x the oracle's own fp32-vs-fp64 deviation measured in the same run) -- end-to-end gradients
                    are dominated by single ReLU-mask flips (1e-3..2e-3 for the reference against itself).
  bf16 tensor path: forward tensors / losses 2e-2 (north_star); end-to-end gradient buckets are REPORTED and bounded
                    loosely (0.5) because thousands of ReLU masks flip under bf16 rounding -- CPU bf16 autocast of the
                    reference itself shows 6e-2 .. 2.4e-1.
"""
import json
import os

import pytest
import torch

from oracle import vaegan as O
from thesis_fmri_reconstruction_b200 import engine, hp

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nchw_flat(raw_nhwc):
    return raw_nhwc.float().permute(0, 3, 1, 2).reshape(raw_nhwc.shape[0], -1)


def run_case(B, adt, seed=4242):
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_vaegan_step(P, S_ref, x, eps, z_p)
    P64 = {k: v.double() for k, v in P.items()}
    S64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in S.items()}
    ref64 = O.stage1_vaegan_step(P64, S64, x.double(), eps.double(), z_p.double(), update=False)
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, adt)
    out = tr.forward_backward(x.cuda(), eps.cuda(), z_p.cuda())
    grads = {k: v.clone() for k, v in tr.named_grads().items()}
    tr.update(B)
    torch.cuda.synchronize()
    errs = {}
    errs["mu"] = rel(out["mu"], ref["mu"])
    errs["logvar"] = rel(out["logvar"], ref["logvar"])
    errs["x_tilde"] = rel(out["x_tilde"], ref["x_tilde"])
    errs["disc_layer"] = rel(nchw_flat(out["disc_layer_nhwc"]), ref["disc_layer"])
    errs["disc_class"] = rel(out["disc_class"], ref["disc_class"].reshape(-1))
    errs["kl"] = rel(out["kl"], ref["kl"])
    errs["mse"] = rel(out["mse"], ref["mse"])
    errs["bce"] = rel(out["bce"], torch.cat([ref["bce_o"], ref["bce_p"], ref["bce_s"]]).reshape(-1))
    lo = tr.losses()
    for k in ("loss_encoder", "loss_decoder", "loss_discriminator"):
        errs[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    gate_ok = (lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"])
    gerr, gnoise = {}, {}
    for b in ("encoder.", "decoder.", "discriminator."):
        a = torch.cat([grads[k].reshape(-1) for k in grads if k.startswith(b)])
        r = torch.cat([ref["grads"][k].reshape(-1) for k in grads if k.startswith(b)])
        r64 = torch.cat([ref64["grads"][k].reshape(-1) for k in grads if k.startswith(b)])
        gerr[b] = rel(a, r64)
        gnoise[b] = rel(r, r64)
    gten = {k: rel(grads[k], ref["grads"][k]) for k in grads}
    newP = tr.named_parameters()
    derr = {}
    for b in ("encoder.", "decoder.", "discriminator."):
        a = torch.cat([(newP[k].cpu() - P[k]).reshape(-1) for k in newP if k.startswith(b)])
        r = torch.cat([(ref["params"][k] - P[k]).reshape(-1) for k in newP if k.startswith(b)])
        derr[b] = rel(a, r)
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    rep = dict(B=B, dtype=str(adt), forward=errs, grad_bucket=gerr, grad_bucket_oracle_fp32_noise=gnoise, grad_tensor_worst=max(gten.items(), key=lambda t: t[1]),
               delta_bucket=derr, bn_worst=max(berr.items(), key=lambda t: t[1]), gate_ok=gate_ok, nbt_ok=nbt_ok)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_stage1_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


def test_stage1_vaegan_fp32_exact_path():
    rep = run_case(8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    for b, e in rep["grad_bucket"].items():
        assert e < max(5e-3, 3 * rep["grad_bucket_oracle_fp32_noise"][b]), (b, e, rep["grad_bucket_oracle_fp32_noise"])
    assert rep["bn_worst"][1] < 1e-4 and rep["gate_ok"] and rep["nbt_ok"]


def test_stage1_vaegan_bf16_tensor_path():
    rep = run_case(16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["bn_worst"][1] < 2e-2 and rep["gate_ok"] and rep["nbt_ok"]


def run_wae_case(B, adt, seed=777):
    P, S = O.make_waegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    z_fake = O.synthetic_noise(B, 128, seed=seed)[0] * 0.5
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_waegan_step(P, S_ref, x, z_fake)
    tr = engine.WaeGanStage1(P, S, hp.CFG64, 128, adt)
    out = tr.step(x.cuda(), z_fake.cuda())
    torch.cuda.synchronize()
    lo = tr.losses()
    fwd = dict(z_real=rel(out["z_real"], ref["z_real"]), x_recon=rel(out["x_recon"], ref["x_recon"]),
               d_real=rel(out["d_real"], ref["d_real"].reshape(-1)), d_fake=rel(out["d_fake"], ref["d_fake"].reshape(-1)),
               d_real_g=rel(out["d_real_g"], ref["d_real_g"].reshape(-1)))
    for k in ("loss_discriminator_fake", "loss_discriminator_real", "loss_reconstruction", "loss_penalty"):
        fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    grads = tr.named_grads()
    gerr = {}
    for b in ("encoder.", "decoder."):   # the discriminator's flat_g holds the D-phase gradient, also checked
        ks = [k for k in ref["grads"] if k.startswith(b)]
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    ks = [k for k in ref["grads"] if k.startswith("discriminator.")]
    gerr["discriminator."] = rel(torch.cat([grads[k].reshape(-1) for k in ks]),
                                 torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    newP = tr.named_parameters()
    # Adam's first step is lr * sign(g): compare the updated discriminator through its effect (d_real_g above) and the
    # untouched l_var head exactly
    lvar_same = all(torch.equal(newP[k].cpu(), P[k]) for k in ("encoder.l_var.weight", "encoder.l_var.bias"))
    nbt = int(tr.named_buffers()["encoder.conv.0.bn.num_batches_tracked"])
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    rep = dict(B=B, dtype=str(adt), forward=fwd, grad_bucket=gerr, lvar_same=lvar_same, nbt=nbt,
               bn_worst=max(berr.items(), key=lambda t: t[1]))
    with open(f"gpurun_out/parity_wae1_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


def test_stage1_waegan_fp32_exact_path():
    rep = run_wae_case(8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 5e-3, rep["grad_bucket"]
    assert rep["lvar_same"] and rep["nbt"] == 2 and rep["bn_worst"][1] < 1e-4


def test_stage1_waegan_bf16_tensor_path():
    rep = run_wae_case(16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["lvar_same"] and rep["nbt"] == 2 and rep["bn_worst"][1] < 2e-2


def run_cog_case(stage, B, adt, seed=901):
    P, S = O.make_cognitive(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    eps_t = O.synthetic_noise(B, 128, seed=seed + 1)[0]
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.cognitive_vaegan_step(P, S_ref, fmri, image, eps, eps_t, z_p, stage)
    if stage == 3:  # stage 3 carries no teacher encoder
        P = {k: v for k, v in P.items() if not k.startswith("teacher_net.")}
        S = {k: v for k, v in S.items() if not k.startswith("teacher_net.")}
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
    gerr = {}
    for b in sorted({k.split(".")[0] + "." for k in ref["grads"]}):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    sqerr = {}
    for pre, bk in tr.buckets.items():   # RMSprop state = 0.1 * clamp(g)^2 after one step: checks the fused clamp
        ks = [k for k in ref["grads"] if k.startswith(pre)]
        if ks and (lo["train_dis"] if pre == "discriminator." else True):
            got = torch.cat([bk.state_view(0, k[len(pre):]).reshape(-1) for k in ks])
            want = torch.cat([ref["square_avg"][k].reshape(-1) for k in ks])
            sqerr[pre] = rel(got, want)
    gate_ok = (lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"])
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    rep = dict(stage=stage, B=B, dtype=str(adt), forward=fwd, grad_bucket=gerr, square_avg=sqerr, gate_ok=gate_ok,
               nbt_ok=nbt_ok, bn_worst=max(berr.items(), key=lambda t: t[1]))
    with open(f"gpurun_out/parity_stage{stage}_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


@pytest.mark.parametrize("stage", [2, 3])
def test_cognitive_stage_fp32_exact_path(stage):
    rep = run_cog_case(stage, 8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 5e-3, rep["grad_bucket"]
    assert max(rep["square_avg"].values()) < 1e-2, rep["square_avg"]
    assert rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 1e-4


@pytest.mark.parametrize("stage", [2, 3])
def test_cognitive_stage_bf16_tensor_path(stage):
    rep = run_cog_case(stage, 16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 2e-2


@pytest.mark.parametrize("adt,B", [(torch.float32, 4), (torch.bfloat16, 16)])
def test_stage1_vaegan_100x100_config(adt, B):
    """The reference's ACTIVE configuration (configs/models_config.py:13-21): 100x100 images, latent 512, stride_gan 2,
    output_pad [False, True, True], odd feature maps 13/25/50 -- every kernel's ragged-tile / odd-parity path."""
    seed = 100
    P, S = O.make_vaegan(O.CFG100, seed=seed)
    x = O.synthetic_images(B, size=100, seed=seed)
    eps, z_p = O.synthetic_noise(B, 512, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_vaegan_step(P, S_ref, x, eps, z_p, cfg=O.CFG100, update=False)
    tr = engine.VaeGanStage1(P, S, hp.CFG100, 512, adt)
    out = tr.forward_backward(x.cuda(), eps.cuda(), z_p.cuda())
    torch.cuda.synchronize()
    grads = tr.named_grads()
    fwd = dict(mu=rel(out["mu"], ref["mu"]), x_tilde=rel(out["x_tilde"], ref["x_tilde"]),
               disc_layer=rel(nchw_flat(out["disc_layer_nhwc"]), ref["disc_layer"]),
               disc_class=rel(out["disc_class"], ref["disc_class"].reshape(-1)), kl=rel(out["kl"], ref["kl"]),
               mse=rel(out["mse"], ref["mse"]))
    gerr = {}
    for b in ("encoder.", "decoder.", "discriminator."):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    print(adt, "100x100 forward", fwd, "grad buckets", gerr)
    assert max(fwd.values()) < (1e-4 if adt == torch.float32 else 2e-2), fwd
    assert max(gerr.values()) < (5e-3 if adt == torch.float32 else 0.5), gerr


def test_stage1_multi_step_stays_finite():
    """Regression: at batch sizes where CTAs walk many tiles (persistent kernels, halo double buffering) no kernel may
    manufacture a NaN (an unpaired filter tap once multiplied zero weights with uninitialised shared memory)."""
    B = 96
    P, S = O.make_vaegan(O.CFG64, seed=5, jitter=False)
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
    x = O.synthetic_images(B, seed=5).cuda()
    eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128, seed=5)]
    for step in range(3):
        out = tr.forward_backward(x, eps, z_p)
        torch.cuda.synchronize()
        for k, v in out.items():
            assert torch.isfinite(v.float()).all(), (step, k)
        for k, v in tr.named_grads().items():
            assert torch.isfinite(v).all(), (step, k)
        tr.update(B)
        for k, v in tr.named_parameters().items():
            assert torch.isfinite(v).all(), (step, k)


def run_dual_case(B, adt, seed=606):
    """BASELINE.json configs[3] composite (engine.DualCognitiveStage3) against oracle.dual_stage3_step."""
    P, S = O.make_dual_stage3(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.dual_stage3_step(P, S_ref, fmri, image, eps, z_p)
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
    gerr = {}
    for b in ("decoder.", "discriminator.", "latent_discriminator."):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        assert ks
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    # Adam state of the latent discriminator after one step: exp_avg = 0.5 g, exp_avg_sq = 0.001 g^2
    bk = tr.buckets["latent_discriminator."]
    ks = [k for k in ref["grads"] if k.startswith("latent_discriminator.")]
    adam = {}
    for i, name in enumerate(("m", "v")):
        got = torch.cat([bk.state_view(i, k[len("latent_discriminator."):]).reshape(-1) for k in ks])
        want = torch.cat([ref["adam"][name][k].reshape(-1) for k in ks])
        adam[name] = rel(got, want)
    newP = tr.named_parameters()
    frozen_same = all(torch.equal(newP[k].cpu(), P[k].float()) for k in newP
                      if k.startswith("encoder.") or k.startswith("teacher_net."))
    gate_ok = (lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"])
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    rep = dict(B=B, dtype=str(adt), forward=fwd, grad_bucket=gerr, adam=adam, frozen_same=frozen_same, gate_ok=gate_ok,
               nbt_ok=nbt_ok, bn_worst=max(berr.items(), key=lambda t: t[1]))
    with open(f"gpurun_out/parity_stage3_dual_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


def test_stage3_dual_fp32_exact_path():
    rep = run_dual_case(8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 5e-3, rep["grad_bucket"]
    assert max(rep["adam"].values()) < 1e-2, rep["adam"]
    assert rep["frozen_same"] and rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 1e-4


def test_stage3_dual_bf16_tensor_path():
    rep = run_dual_case(16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["frozen_same"] and rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 2e-2


@pytest.mark.parametrize("adt,B", [(torch.float32, 8), (torch.bfloat16, 16)])
def test_stage1_wae_mmd_variant(adt, B):
    """WaeGanStage1(penalty="mmd") against oracle.stage1_wae_mmd_step. EXTENSION, parity unpinned: the reference has no
    MMD; the bar is agreement with our own published-estimator statement (oracle/mmd.py)."""
    seed = 313
    P, S = O.make_waegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    z_fake = O.synthetic_noise(B, 128, seed=seed)[0] * 0.5
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.stage1_wae_mmd_step(P, S_ref, x, z_fake)
    tr = engine.WaeGanStage1(P, S, hp.CFG64, 128, adt, penalty="mmd")
    out = tr.step(x.cuda(), z_fake.cuda())
    torch.cuda.synchronize()
    lo = tr.losses()
    fwd = dict(z_real=rel(out["z_real"], ref["z_real"]), x_recon=rel(out["x_recon"], ref["x_recon"]))
    fwd["loss_reconstruction"] = abs(lo["loss_reconstruction"] - ref["loss_reconstruction"].item()) / abs(ref["loss_reconstruction"].item())
    # the penalty is a difference of kernel sums of magnitude ~ 7 * lambda * B: error measured against that magnitude
    fwd["loss_penalty"] = abs(lo["loss_penalty"] - ref["loss_penalty"].item()) / (7 * 10.0 * B)
    grads = tr.named_grads()
    gerr = {}
    for b in ("encoder.", "decoder."):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    nbt = int(tr.named_buffers()["encoder.conv.0.bn.num_batches_tracked"])
    print(adt, "wae-mmd forward", fwd, "grad buckets", gerr, "penalty", lo["loss_penalty"], ref["loss_penalty"].item())
    assert max(fwd.values()) < (1e-4 if adt == torch.float32 else 2e-2), fwd
    assert max(gerr.values()) < (5e-3 if adt == torch.float32 else 0.5), gerr
    assert nbt == 1


def run_wae_cog_case(stage, B, adt, seed=515):
    """engine.WaeCognitiveStage against oracle.cognitive_wae_step (pinned to tests/golden/stage{2,3}_cognitive_wae_*)."""
    P, S = O.make_cognitive_wae(O.CFG64, seed=seed)
    fmri, image = O.synthetic_fmri(B, seed=seed), O.synthetic_images(B, seed=seed)
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.cognitive_wae_step(P, S_ref, fmri, image, stage)
    tr = engine.WaeCognitiveStage(P, S, hp.CFG64, stage, 128, adt)
    out = tr.step(fmri.cuda(), image.cuda())
    torch.cuda.synchronize()
    lo = tr.losses()
    fwd = {k: rel(out[k], ref[k].reshape(out[k].shape)) for k in ("z_fake", "z_real", "x_recon", "d_real", "d_fake", "d_real_g")}
    for k in ("loss_discriminator_fake", "loss_discriminator_real", "loss_reconstruction", "loss_penalty"):
        fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    grads = tr.named_grads()
    trained = ref["trained"] + "."
    gerr = {}
    for b in ("discriminator.", trained):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        assert ks
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    # Adam state after one step (m = (1 - beta1) g, v = (1 - beta2) g^2) of the trained bucket
    bk = tr.buckets[trained]
    ks = [k for k in ref["grads"] if k.startswith(trained)]
    adam = {}
    for i, name in enumerate(("m", "v")):
        got = torch.cat([bk.state_view(i, k[len(trained):]).reshape(-1) for k in ks])
        want = torch.cat([ref["adam"][name][k].reshape(-1) for k in ks])
        adam[name] = rel(got, want)
    newP = tr.named_parameters()
    frozen = [p for p in ("encoder.", "decoder.", "teacher_net.encoder.") if p != trained]
    frozen_same = all(torch.equal(newP[k].cpu(), P[k].float()) for k in newP if any(k.startswith(f) for f in frozen))
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    rep = dict(stage=stage, B=B, dtype=str(adt), forward=fwd, grad_bucket=gerr, adam=adam, frozen_same=frozen_same,
               nbt_ok=nbt_ok, bn_worst=max(berr.items(), key=lambda t: t[1]))
    with open(f"gpurun_out/parity_wae_stage{stage}_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


@pytest.mark.parametrize("stage", [2, 3])
def test_wae_cognitive_stage_fp32_exact_path(stage):
    rep = run_wae_cog_case(stage, 8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 5e-3, rep["grad_bucket"]
    assert max(rep["adam"].values()) < 1e-2, rep["adam"]
    assert rep["frozen_same"] and rep["nbt_ok"] and rep["bn_worst"][1] < 1e-4


@pytest.mark.parametrize("stage", [2, 3])
def test_wae_cognitive_stage_bf16_tensor_path(stage):
    rep = run_wae_cog_case(stage, 16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["frozen_same"] and rep["nbt_ok"] and rep["bn_worst"][1] < 2e-2


def test_graphed_step_matches_eager():
    """engine.GraphedStep (CUDA-graph replay of the Stage-I step) == the eager step. Two runs drift apart chaotically over
    several steps (fp32 atomics in the weight-gradient splits; RMSprop's first steps move every element by lr * sign(g) /
    sqrt(0.1) however small g is), so the comparison is ONE step from an identical, already warmed-up state: the eager
    trainer's checkpoint is loaded into the captured trainer (the graph reads the same flat buffers)."""
    B, seed = 8, 31
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed).cuda()
    eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128, seed=seed)]
    a = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
    b = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
    g = engine.GraphedStep(b, x, eps, z_p, warmup=2)
    for _ in range(3):
        a.step(x, eps, z_p)
    b.load_state_dict(a.state_dict())
    a.step(x, eps, z_p)
    out = g(x, eps, z_p)
    torch.cuda.synchronize()
    pa, pb = a.named_parameters(), b.named_parameters()
    worst = max(rel(pb[k], pa[k].cpu()) for k in pa)
    la, lb = a.losses(), b.losses()
    print("graphed vs eager, one step from the same state: worst parameter rel", worst, la["loss_encoder"], lb["loss_encoder"],
          "launches/replay", g.launches)
    # 1e-3: an element whose gradient is below the fp32-atomics noise can change sign, and both optimizers move such an
    # element by ~lr whatever its magnitude; a replay that lost state or inputs is off by >= 1e-2
    assert worst < 1e-3, worst
    assert abs(la["loss_encoder"] - lb["loss_encoder"]) <= 1e-5 * abs(la["loss_encoder"])
    assert (la["train_dis"], la["train_dec"]) == (lb["train_dis"], lb["train_dec"])
    sa, sb = a.named_buffers(), b.named_buffers()
    assert all(int(sa[k]) == int(sb[k]) for k in sa if not sa[k].dtype.is_floating_point)
    assert max(rel(sb[k], sa[k].cpu()) for k in sa if sa[k].dtype.is_floating_point) < 1e-5
    assert g.launches > 100 and torch.isfinite(out["x_tilde"].float()).all()
    for _ in range(2):   # further replays keep running and stay finite
        out = g(x, eps, z_p)
    torch.cuda.synchronize()
    assert torch.isfinite(out["x_tilde"].float()).all() and all(torch.isfinite(v).all() for v in b.named_parameters().values())


def test_graphed_step_follows_epoch_schedule():
    """learning rates and gate constants are host scalars baked into the captured launches; after end_epoch()
    (ExponentialLR 0.98, train_vgan_stage1.py:446-457) a replay must use the NEW values, i.e. equal the eager step."""
    B, seed = 8, 32
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed).cuda()
    eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128, seed=seed)]
    a = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
    b = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
    g = engine.GraphedStep(b, x, eps, z_p, warmup=2)
    for _ in range(2):
        a.step(x, eps, z_p)
    for t in (a, b):
        for _ in range(25):                      # 0.98^25 = 0.60: a stale learning rate would be far outside the bound
            t.end_epoch(decay_margin=0.99, decay_equilibrium=0.99)
    b.load_state_dict(a.state_dict())
    before = {k: v.clone() for k, v in a.named_parameters().items()}
    a.step(x, eps, z_p)
    g(x, eps, z_p)
    torch.cuda.synchronize()
    assert g.recaptures >= 1
    pa, pb = a.named_parameters(), b.named_parameters()
    da = torch.cat([(pa[k] - before[k]).reshape(-1) for k in pa])
    db = torch.cat([(pb[k] - before[k]).reshape(-1) for k in pa])
    print("graphed vs eager parameter DELTA after end_epoch x25: rel", rel(db, da), "recaptures", g.recaptures)
    assert rel(db, da) < 5e-2      # deltas scale with lr: a replay at the stale lr would be off by 0.67
    assert (a.losses()["train_dis"], a.losses()["train_dec"]) == (b.losses()["train_dis"], b.losses()["train_dec"])


def test_stage1_beta_vae_mode_fp32():
    """The 'beta-vae' loss mix of train/train_vgan_stage1.py:359-365 (KL weighted by beta / batch_size): engine vs the oracle
    (pinned to the reference by tests/golden/stage1_betavae_*). Only the encoder bucket and loss_encoder differ from 'vae-gan'."""
    B, seed, beta = 8, 77, 4.0
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    P64 = {k: v.double() for k, v in P.items()}
    ref = O.stage1_vaegan_step(P64, {k: v.clone().double() if v.dtype.is_floating_point else v.clone() for k, v in S.items()},
                               x.double(), eps.double(), z_p.double(), update=False, mode="beta-vae", beta=beta)
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32, mode="beta-vae", beta=beta)
    tr.forward_backward(x.cuda(), eps.cuda(), z_p.cuda())
    torch.cuda.synchronize()
    lo, grads = tr.losses(), tr.named_grads()
    ks = [k for k in ref["grads"] if k.startswith("encoder.")]
    gerr = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    lerr = abs(lo["loss_encoder"] - ref["loss_encoder"].item()) / abs(ref["loss_encoder"].item())
    print("beta-vae encoder bucket", gerr, "loss_encoder", lerr)
    assert gerr < 5e-3 and lerr < 1e-4


def run_dual1_case(B, adt, seed=808):
    """engine.DualWaeVaeGanStage1 (train/wae_vgan_stage1.py:282-441, row a16) against oracle.dual_stage1_step, which is pinned
    to the reference by tests/golden/stage1_dual_*."""
    P, S = O.make_dual_stage1(O.CFG64, seed=seed)
    x = O.synthetic_images(B, seed=seed)
    eps, z_p = O.synthetic_noise(B, 128, seed=seed)
    z_fake = O.synthetic_noise(B, 128, seed=seed + 7)[0] * 0.5
    S_ref = {k: v.clone() for k, v in S.items()}
    ref = O.dual_stage1_step(P, S_ref, x, eps, z_p, z_fake)
    tr = engine.DualWaeVaeGanStage1(P, S, hp.CFG64, 128, adt)
    out = tr.forward_backward(x.cuda(), eps.cuda(), z_p.cuda(), z_fake.cuda())
    grads = {k: v.clone() for k, v in tr.named_grads().items()}
    tr.update(B)
    torch.cuda.synchronize()
    lo = tr.losses()
    fwd = dict(mu=rel(out["mu"], ref["mu"]), x_tilde=rel(out["x_tilde"], ref["x_tilde"]),
               disc_class=rel(out["disc_class"], ref["disc_class"].reshape(-1)), z_real=rel(out["z_real"], ref["z_real"]),
               d_real=rel(out["d_real"], ref["d_real"].reshape(-1)), d_fake=rel(out["d_fake"], ref["d_fake"].reshape(-1)),
               d_real_g=rel(out["d_real_g"], ref["d_real_g"].reshape(-1)))
    for k in ("loss_encoder", "loss_decoder", "loss_discriminator", "loss_discriminator_fake", "loss_discriminator_real",
              "loss_penalty"):
        fwd[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    gerr = {}
    for b in ("encoder.", "decoder.", "discriminator.", "latent_discriminator."):
        ks = [k for k in ref["grads"] if k.startswith(b)]
        gerr[b] = rel(torch.cat([grads[k].reshape(-1) for k in ks]), torch.cat([ref["grads"][k].reshape(-1) for k in ks]))
    newP = tr.named_parameters()
    ks = [k for k in P if k.startswith("latent_discriminator.")]
    lat_delta = rel(torch.cat([(newP[k].cpu() - P[k]).reshape(-1) for k in ks]),
                    torch.cat([(ref["params"][k] - P[k]).reshape(-1) for k in ks]))
    gate_ok = (lo["train_dis"], lo["train_dec"]) == (ref["train_dis"], ref["train_dec"])
    berr = {k: rel(v, S_ref[k]) for k, v in tr.named_buffers().items() if v.dtype.is_floating_point}
    nbt_ok = all(int(v) == int(S_ref[k]) for k, v in tr.named_buffers().items() if not v.dtype.is_floating_point)
    rep = dict(B=B, dtype=str(adt), forward=fwd, grad_bucket=gerr, latent_delta=lat_delta, gate_ok=gate_ok, nbt_ok=nbt_ok,
               bn_worst=max(berr.items(), key=lambda t: t[1]))
    with open(f"gpurun_out/parity_stage1_dual_B{B}_{str(adt).split('.')[-1]}.json", "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


def test_dual_stage1_fp32_exact_path():
    rep = run_dual1_case(8, torch.float32)
    assert max(rep["forward"].values()) < 1e-4, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 5e-3, rep["grad_bucket"]
    assert rep["latent_delta"] < 1e-2 and rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 1e-4


def test_dual_stage1_bf16_tensor_path():
    rep = run_dual1_case(16, torch.bfloat16)
    assert max(rep["forward"].values()) < 2e-2, rep["forward"]
    assert max(rep["grad_bucket"].values()) < 0.5, rep["grad_bucket"]
    assert rep["gate_ok"] and rep["nbt_ok"] and rep["bn_worst"][1] < 2e-2


@pytest.mark.parametrize("kind", ["vaegan", "waegan"])
def test_trainer_checkpoint_resume(kind):
    """state_dict() / load_state_dict(): a resumed trainer continues exactly like the original (parameters, BatchNorm buffers,
    RMSprop / Adam state, Adam step count), and the `model` part loads strictly into the drop-in nn.Module."""
    B, seed = 8, 11
    x = O.synthetic_images(B, seed=seed).cuda()
    eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128, seed=seed)]
    if kind == "vaegan":
        P, S = O.make_vaegan(O.CFG64, seed=seed)
        mk = lambda: engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
        args = (x, eps, z_p)
    else:
        P, S = O.make_waegan(O.CFG64, seed=seed)
        mk = lambda: engine.WaeGanStage1(P, S, hp.CFG64, 128, torch.float32)
        args = (x, eps * 0.5)
    a = mk()
    for _ in range(2):
        a.step(*args)
    a.end_epoch(epoch=30)
    sd = a.state_dict()
    b = mk()
    b.load_state_dict(sd)
    a.step(*args)
    b.step(*args)
    torch.cuda.synchronize()
    pa, pb = a.named_parameters(), b.named_parameters()
    worst = max(rel(pb[k], pa[k].cpu()) for k in pa)
    sa, sb = a.named_buffers(), b.named_buffers()
    # 1e-3, not 1e-5: elements with a gradient below the fp32-atomics noise may change sign and move by ~lr under RMSprop /
    # Adam; a resume that lost the optimizer state is off by >= 1e-2 (first-step updates are lr * sign(g))
    assert worst < 1e-3, worst
    assert all(int(sa[k]) == int(sb[k]) for k in sa if not sa[k].dtype.is_floating_point)
    assert max(rel(sb[k], sa[k].cpu()) for k in sa if sa[k].dtype.is_floating_point) < 1e-5
    assert a.lr == b.lr and getattr(a, "t", 0) == getattr(b, "t", 0)
    if kind == "vaegan":
        import configs.models_config as mc
        mc.use_resolution(64)
        from models.vae_gan import VaeGan
        m = VaeGan(device="cuda", z_size=128)
        missing, unexpected = m.load_state_dict(sd["model"], strict=True)
        assert not missing and not unexpected


def test_full_size_batch_tiling_property():
    """BASELINE.json's full per-GPU batch (4096) through a size-independent property: a batch made of 64 copies of a 64-sample
    batch has the same BatchNorm statistics, so every replica's forward tensors equal the 64-sample step's and the loss sums /
    gradient buckets are 64x larger. bf16 tensor path.

    What bounds the comparison (measured layer by layer, scripts/layerdiff.py -> profiles/parity/r2_layerdiff_B64_vs_4096.txt):
    a perturbation d entering a bf16 rounding step leaves it as ~sqrt(d * 2^-8) (a fraction d / ulp of the elements flips by one
    ulp), so ANY difference -- the 1e-7 of a different fp32 summation order in a split-K GEMM or a BatchNorm mean -- grows
    1e-7 -> 2e-5 -> 3e-4 -> 1e-3 -> 2e-3 -> 3e-3 across five layers and saturates at the bf16 storage noise (~4e-3). The same
    64-sample step run TWICE differs by 2.9e-3 in x_tilde for that reason (non-deterministic order of fp32 atomics in the
    encoder's split-K fc layer). The large-batch kernel paths are therefore certified per launch against torch
    (tests/test_fullsize_kernels_gpu.py) and per step against the oracle (tests/test_baseline_sizes_gpu.py); this test checks the
    property at the level it can hold: the tiled step differs from the small one by no more than the bf16 storage noise
    (forward < 1e-2), loss sums agree to 1e-3, and all replicas of the tiled batch are bit-identical. The run-to-run spread of
    the 64-sample step is printed beside it (it varies from 1e-6 to 3e-3 with the order the atomics happen to land in)."""
    B0, R, seed = 64, 64, 404
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x0 = O.synthetic_images(B0, seed=seed).cuda()
    eps0, zp0 = [t.cuda() for t in O.synthetic_noise(B0, 128, seed=seed)]

    def small():
        t = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
        o = t.forward_backward(x0, eps0, zp0)
        return o, {k: v.clone() for k, v in t.named_grads().items()}, t.losses()

    oa, ga, la = small()
    oa2, ga2, _ = small()
    b = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
    ob = b.forward_backward(x0.repeat(R, 1, 1, 1), eps0.repeat(R, 1), zp0.repeat(R, 1))
    gb = b.named_grads()
    lb = b.losses()
    torch.cuda.synchronize()
    fwd, noise, spread = {}, {}, {}
    for k in ("x_tilde", "mu", "kl", "mse"):
        noise[k] = rel(oa2[k], oa[k].cpu())
        fwd[k + "_first"] = rel(ob[k][:B0], oa[k].cpu())
        fwd[k + "_last"] = rel(ob[k][-B0:], oa[k].cpu())
        spread[k] = rel(ob[k][-B0:], ob[k][:B0].cpu())
    lerr = {k: abs(lb[k] / R - la[k]) / abs(la[k]) for k in ("loss_encoder", "loss_decoder", "loss_discriminator", "kl", "mse")}
    gerr, gnoise = {}, {}
    for pre in ("encoder.", "decoder.", "discriminator."):
        ks = [k for k in ga if k.startswith(pre)]
        cat = lambda g: torch.cat([g[k].reshape(-1) for k in ks])
        gerr[pre] = rel(cat(gb) / R, cat(ga).cpu())
        gnoise[pre] = rel(cat(ga2), cat(ga).cpu())
    print("tiling property: forward", fwd, "run-to-run", noise, "replica spread", spread, "losses", lerr, "grad buckets", gerr,
          "grad run-to-run", gnoise)
    assert max(spread.values()) == 0.0, spread          # every replica of the tiled batch computes the same values
    # forward: at the bf16 storage-noise floor (measured 2e-3 .. 6e-3, the same size as the run-to-run spread of x_tilde and as
    # the bf16-vs-oracle error), well inside north_star's 2e-2
    assert max(fwd.values()) < 1e-2, (fwd, noise)
    assert max(lerr.values()) < 1e-3, lerr
    # gradients: end-to-end bf16 gradients of two runs that differ in ~4e-4 of their ReLU masks (measured 0.16 / 0.13 / 0.02;
    # the backward kernels themselves are held to 2e-2 by the teacher-forced test)
    assert max(gerr.values()) < 0.3, (gerr, gnoise)
    assert all(torch.isfinite(v).all() for v in gb.values())
