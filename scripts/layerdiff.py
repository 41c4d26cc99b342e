"""Layer-by-layer forward diff of the Stage-I networks: a 64-sample batch against the SAME 64 samples tiled R times
(identical BatchNorm statistics by construction), i.e. the one-tile-per-CTA kernel paths against the persistent /
large-batch ones (VERDICT r1 "What's weak" #2). Also runs the 64-sample forward twice (run-to-run noise floor).

    python scripts/layerdiff.py [R]          # on a B200; prints one line per saved tensor
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vaegan as O  # noqa: E402  (seeded weights / inputs only)
from thesis_fmri_reconstruction_b200 import engine, hp  # noqa: E402
from thesis_fmri_reconstruction_b200 import lib as L  # noqa: E402


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def collect(tr, x, eps, z_p):
    """Forward only; returns {name: tensor} of every saved per-layer tensor (first B0 samples are compared later)."""
    B, z = x.shape[0], tr.z
    be, bd, bc = tr.buckets["encoder."], tr.buckets["decoder."], tr.buckets["discriminator."]
    Se, Sd, Sc = ({k: v.clone() for k, v in tr.Ssub[p].items()} for p in ("encoder.", "decoder.", "discriminator."))
    out = {}
    ycat, ce = tr.enc.forward(be.P, Se, x, True, 1, {})
    out["enc.raw0"], out["enc.mean0"], out["enc.invstd0"] = ce.c0.raw, ce.c0.mean, ce.c0.invstd
    for i, c in enumerate(ce.blocks):
        out[f"enc.x{i + 1}"] = c.x
        out[f"enc.raw{i + 1}"], out[f"enc.mean{i + 1}"], out[f"enc.invstd{i + 1}"] = c.bn.raw, c.bn.mean, c.bn.invstd
    out["enc.fc_in"], out["enc.fc_raw"], out["enc.fc_mean"], out["enc.fc_invstd"] = ce.fc.x, ce.fc.bn.raw, ce.fc.bn.mean, ce.fc.bn.invstd
    out["enc.h"] = ce.heads.h
    out["enc.ycat"] = ycat
    zz, kl = torch.empty(B, z, device="cuda"), torch.empty(B, device="cuda")
    L.reparam_kl_fwd(ycat[:, :z], ycat[:, z:], eps, zz, kl, B, z, ld=2 * z)
    out["z"] = zz
    img, cd = tr.dec.forward(bd.P, Sd, zz, True, 1, {})
    out["dec.fc_raw"], out["dec.fc_mean"], out["dec.fc_invstd"] = cd.fc.bn.raw, cd.fc.bn.mean, cd.fc.bn.invstd
    for i, c in enumerate(cd.blocks):
        out[f"dec.x{i}"] = c.x
        out[f"dec.raw{i}"], out[f"dec.mean{i}"], out[f"dec.invstd{i}"] = c.bn.raw, c.bn.mean, c.bn.invstd
    out["dec.a3"], out["dec.img"] = cd.a3, img
    x_p, _ = tr.dec.forward(bd.P, Sd, z_p, True, 1, {})
    raw3, p, cc = tr.dis.forward(bc.P, Sc, [x, img, x_p], True, 1, True, {})
    out["dis.y0"] = cc.y0
    for i, c in enumerate(cc.blocks):
        out[f"dis.raw{i + 1}"], out[f"dis.mean{i + 1}"], out[f"dis.invstd{i + 1}"] = c.bn.raw, c.bn.mean, c.bn.invstd
    out["dis.fc_raw"], out["dis.hfc"], out["dis.p"] = cc.fc.bn.raw, cc.hfc, p
    torch.cuda.synchronize()
    return {k: v.float().clone() for k, v in out.items()}


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B0, seed = 64, 404
    P, S = O.make_vaegan(O.CFG64, seed=seed)
    x0 = O.synthetic_images(B0, seed=seed).cuda()
    eps0, zp0 = [t.cuda() for t in O.synthetic_noise(B0, 128, seed=seed)]
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16)
    a = collect(tr, x0, eps0, zp0)
    a2 = collect(tr, x0, eps0, zp0)
    b = collect(tr, x0.repeat(R, 1, 1, 1), eps0.repeat(R, 1), zp0.repeat(R, 1))
    print(f"{'tensor':18s} {'run-to-run':>11s} {'first64':>11s} {'last64':>11s} {'replica spread':>14s}")
    for k in a:
        va, vb = a[k], b[k]
        if va.dim() == 1 and va.shape == vb.shape:      # per-channel statistics
            print(f"{k:18s} {rel(a2[k], va):11.2e} {rel(vb, va):11.2e}")
            continue
        n = va.shape[0]
        if k.startswith("dis.") and k != "dis.p":       # discriminator batch = [x | x_tilde | x_p] blocks of B
            blocks_b = vb.reshape(3, R, n // 3, *vb.shape[1:])
            first, last = blocks_b[:, 0].reshape(va.shape), blocks_b[:, -1].reshape(va.shape)
        elif k == "dis.p":
            blocks_b = vb.reshape(3, R, n // 3)
            first, last = blocks_b[:, 0].reshape(-1), blocks_b[:, -1].reshape(-1)
        else:
            first, last = vb[:n], vb[-n:]
        print(f"{k:18s} {rel(a2[k], va):11.2e} {rel(first, va):11.2e} {rel(last, va):11.2e} {rel(last, first):14.2e}")


if __name__ == "__main__":
    main()
