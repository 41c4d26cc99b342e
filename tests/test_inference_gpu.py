"""Inference fast path and input pipeline on the GPU:
  * fmri_pearson / fmri_ssim vs the oracle restatement of train/train_utils.py (itself pinned to reference goldens), and vs
    the golden values directly: 1e-5 absolute (fp32 window sums, fp64 accumulation);
  * the BatchNorm-folded eval forward (inference.Reconstructor) vs the oracle's eval-mode modules after one training step
    (non-trivial running statistics): fp32 1e-4, bf16 2e-2; also vs this repo's own nn.Module eval forward;
  * fmri_image_pipeline vs the oracle image pipeline: bit-exact; Prefetcher delivers every batch once, in order.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import metrics as M
from oracle import vaegan as O
from oracle.make_golden_metrics import inputs
from thesis_fmri_reconstruction_b200 import data, hp, inference

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "metrics_*.npz")))


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_metric_kernels_match_reference_goldens(path):
    g = np.load(path)
    N, C, H, W, seed = (int(v) for v in g["shape"])
    a, b = inputs(N, C, H, W, seed)
    p = float(inference.pcc(a.cuda(), b.cuda()))
    s = float(inference.ssim(a.cuda(), b.cuda()))
    print(os.path.basename(path), "pcc", p, float(g["pcc"]), "ssim", s, float(g["ssim"]))
    assert abs(p - float(g["pcc64"])) < 1e-5 and abs(p - float(M.pearson(a.double(), b.double()))) < 1e-5
    assert abs(s - float(g["ssim"])) < 1e-5 and abs(s - float(M.ssim(a, b))) < 1e-5


def test_ssim_rejects_images_smaller_than_the_window():
    from thesis_fmri_reconstruction_b200.lib import FmriError

    a = torch.rand(2, 3, 9, 9).cuda()
    with pytest.raises(FmriError):
        inference.ssim(a, a)


@pytest.mark.parametrize("adt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfgname", ["64", "100"])
def test_folded_eval_forward_matches_oracle(adt, cfgname):
    cfg_o, cfg, z, size = (O.CFG64, hp.CFG64, 128, 64) if cfgname == "64" else (O.CFG100, hp.CFG100, 512, 100)
    B, seed = 8, 55
    P, S = O.make_vaegan(cfg_o, seed=seed)
    x = O.synthetic_images(B, size=size, seed=seed)
    eps, z_p = O.synthetic_noise(B, z, seed=seed)
    step = O.stage1_vaegan_step(P, S, x, eps, z_p, cfg=cfg_o)          # S now holds non-trivial running statistics
    P1 = step["params"]
    xe = O.synthetic_images(B, size=size, seed=seed + 1)
    mu, lv = O.encoder(P1, S, xe, cfg_o, train=False)
    ref = O.decoder(P1, S, O.reparameterize(mu, lv, eps), cfg_o, train=False)
    ref_mean = O.decoder(P1, S, mu, cfg_o, train=False)
    ref_gen = O.decoder(P1, S, z_p, cfg_o, train=False)
    r = inference.Reconstructor({**P1, **S}, cfg, z=z, kind="vaegan", adt=adt)
    got_mu, got_lv = r.encode(xe.cuda())
    got = r(xe.cuda(), eps=eps.cuda())
    got_mean = r(xe.cuda(), sample=False)
    got_gen = r.generate(z_p=z_p)
    torch.cuda.synchronize()
    errs = dict(mu=rel(got_mu, mu), logvar=rel(got_lv, lv), x_hat=rel(got, ref), x_hat_mean=rel(got_mean, ref_mean),
                generated=rel(got_gen, ref_gen))
    print(cfgname, adt, errs)
    assert max(errs.values()) < (1e-4 if adt == torch.float32 else 2e-2), errs
    pcc, ssim, mse = (float(v) for v in r.metrics(got, xe.cuda()))
    assert abs(pcc - float(M.pearson(got.cpu().double(), xe.double()))) < 1e-5
    assert abs(ssim - float(M.ssim(got.cpu(), xe))) < 1e-5
    assert abs(mse - float(torch.nn.functional.mse_loss(got.cpu(), xe))) < 1e-6


def test_folded_cognitive_eval_forward_and_module_agreement():
    """Reconstructor(kind='cognitive') vs the oracle, and vs this repo's own VaeGanCognitive module in eval mode."""
    import configs.models_config as mc
    from thesis_fmri_reconstruction_b200 import autograd as ag

    B, seed = 8, 66
    P, S = O.make_cognitive(O.CFG64, seed=seed)
    S = {k: (v + 0.05 * torch.randn(v.shape, generator=torch.Generator().manual_seed(1)).abs() if v.dtype.is_floating_point else v)
         for k, v in S.items()}
    fmri = O.synthetic_fmri(B, seed=seed)
    mu, lv = O.cognitive_encoder(P, S, fmri, train=False)
    ref = O.decoder(P, S, mu, O.CFG64, train=False)
    for adt in (torch.float32, torch.bfloat16):
        r = inference.Reconstructor({**P, **S}, hp.CFG64, z=128, kind="cognitive", adt=adt)
        got = r(fmri.cuda(), sample=False)
        torch.cuda.synchronize()
        e = rel(got, ref)
        print("cognitive folded eval", adt, e)
        assert e < (1e-4 if adt == torch.float32 else 2e-2)
    mc.use_resolution(64)
    from models.vae_gan import CognitiveEncoder, Decoder

    old = ag.compute_dtype()
    ag.set_compute_dtype(torch.float32)
    try:
        enc = CognitiveEncoder(input_size=fmri.shape[1], z_size=128).cuda()
        dec = Decoder(z_size=128, size=256).cuda()
        enc.load_state_dict({k[len("encoder."):]: v for k, v in {**P, **S}.items() if k.startswith("encoder.")}, strict=True)
        dec.load_state_dict({k[len("decoder."):]: v for k, v in {**P, **S}.items() if k.startswith("decoder.")}, strict=True)
        enc.eval(), dec.eval()
        with torch.no_grad():
            m_mu, _ = enc(fmri.cuda())
            m_img = dec(m_mu)
        r = inference.Reconstructor({**P, **S}, hp.CFG64, z=128, kind="cognitive", adt=torch.float32)
        assert rel(r(fmri.cuda(), sample=False), m_img) < 1e-4
    finally:
        ag.set_compute_dtype(old)


def test_image_pipeline_kernel_and_prefetcher():
    g = torch.Generator().manual_seed(9)
    mean, std = (0.5, 0.4, 0.3), (0.5, 0.25, 0.2)
    for C in (3, 1):
        u8 = torch.randint(0, 256, (6, 64, 64, C), generator=g, dtype=torch.uint8)
        flip = torch.tensor([0, 1, 1, 0, 1, 0], dtype=torch.int32)
        sh = torch.randint(-5, 6, (6, 2), generator=g, dtype=torch.int32)
        pipe = data.DevicePipeline(mean, std)
        got = pipe(u8.cuda(), flip, sh)
        torch.cuda.synchronize()
        want = M.image_pipeline(u8, flip, sh, mean, std)
        assert (got.cpu() - want).abs().max().item() < 1e-6
        plain = pipe(u8.cuda())
        assert (plain.cpu() - M.image_pipeline(u8, None, None, mean, std)).abs().max().item() < 1e-6
    # prefetcher: every batch once, in order, extras delivered, augmentation parameters drawn per batch
    batches = [(torch.full((4, 16, 16, 3), i * 10, dtype=torch.uint8), {"fmri": torch.full((4, 7), float(i))}) for i in range(5)]
    pipe = data.DevicePipeline((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), random_flip=True, max_shift=3,
                               generator=torch.Generator().manual_seed(0))
    seen = []
    for x, ex in data.Prefetcher(batches, pipe):
        assert x.shape == (4, 3, 16, 16) and x.is_cuda and ex["fmri"].is_cuda
        seen.append((round(float(x.mean()) * 255), round(float(ex["fmri"].mean()), 3)))
    assert seen == [(i * 10, float(i)) for i in range(5)], seen
