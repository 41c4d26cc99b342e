"""__graft_entry__.smoke(): one small Stage-I VAE/GAN training step on cuda:0 through the product path (the sm_100a
kernels behind the C ABI), checked against the CPU oracle (the only place outside tests/ and bench.py's CPU legs that
touches oracle/ -- as the checker, never as the thing run)."""
from __future__ import annotations

import torch


def _rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(B=16, verbose=True):
    from oracle import vaegan as O

    from . import engine, hp, lib

    if not lib.load().fmri_tensor_path_available():
        raise lib.FmriError("smoke(): the tcgen05 tensor path needs an sm_100 device")
    P, S = O.make_vaegan(O.CFG64, seed=2024)
    x = O.synthetic_images(B, seed=2024)
    eps, z_p = O.synthetic_noise(B, 128, seed=2024)
    ref = O.stage1_vaegan_step(P, {k: v.clone() for k, v in S.items()}, x, eps, z_p, update=False)
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
    lib.launch_count(reset=True)
    out = tr.step(x.cuda(), eps.cuda(), z_p.cuda())
    torch.cuda.synchronize()
    n = lib.launch_count()
    lo = tr.losses()
    errs = dict(mu=_rel(out["mu"], ref["mu"]), x_tilde=_rel(out["x_tilde"], ref["x_tilde"]),
                disc_class=_rel(out["disc_class"], ref["disc_class"].reshape(-1)), kl=_rel(out["kl"], ref["kl"]),
                mse=_rel(out["mse"], ref["mse"]))
    for k in ("loss_encoder", "loss_decoder", "loss_discriminator"):
        errs[k] = abs(lo[k] - ref[k].item()) / abs(ref[k].item())
    if verbose:
        print(f"smoke: Stage-I VAE/GAN step B={B} bf16 tensor path, {n} kernel launches; rel. error vs CPU oracle: "
              + ", ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    bad = {k: v for k, v in errs.items() if not v < 2e-2}  # north_star: 2e-2 relative for bf16
    if bad or n == 0:
        raise AssertionError(f"smoke(): parity outside the bf16 tolerance 2e-2: {bad} (launches={n})")
    return errs
