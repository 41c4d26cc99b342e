"""The algorithmic-FLOP accounting behind bench.py's roofline numbers, recomputed from the
layer shapes of the 64x64 configuration: 2 x MACs of the NECESSARY contractions only (one discriminator forward, no discarded
gradient). If a constant in bench.ALG_MFLOP drifts from the formulas, `roofline.whole_step_frac` would be wrong."""
import importlib.util
import os

from oracle import vaegan as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def conv(cin, cout, oh, ow):            # 5x5 kernel, per image, forward (a transposed conv counts its INPUT grid)
    return 2.0 * oh * ow * cin * cout * 25


def lin(k, n):
    return 2.0 * k * n


def layer_flops(cfg=O.CFG64, z=128):
    e = cfg["encoder_channels"]
    enc_convs = [conv(3, e[0], 32, 32), conv(e[0], e[1], 16, 16), conv(e[1], e[2], 8, 8)]
    enc_fc = lin(8 * 8 * e[2], 1024)
    E = sum(enc_convs) + enc_fc + 2 * lin(1024, z)
    d = cfg["decoder_channels"]
    dec_fc = lin(z, 8 * 8 * 256)
    dec_convs = [conv(256, 256, 8, 8), conv(256, d[1], 16, 16), conv(d[1], d[2], 32, 32), conv(d[2], 3, 64, 64)]
    D = dec_fc + sum(dec_convs)
    c = cfg["discrim_channels"]
    dis_convs = [conv(3, c[0], 64, 64), conv(c[0], c[1], 32, 32), conv(c[1], c[2], 16, 16), conv(c[2], c[3], 8, 8)]
    S = sum(dis_convs) + lin(8 * 8 * c[3], 512) + lin(512, 1)
    C = lin(O.NUM_VOXELS, 1024) + 2 * lin(1024, z)
    W = lin(z, 512) + 3 * lin(512, 512) + lin(512, 1)
    return dict(E=E, D=D, S=S, C=C, W=W, enc_c0=enc_convs[0], dec_fc=dec_fc, dis=dis_convs)


def test_forward_flops_per_image_match_the_survey():
    f = layer_flops()
    for k, want in (("E", 253.624), ("D", 862.716), ("S", 875.300), ("C", 7.938), ("W", 1.705)):
        assert abs(f[k] / 1e6 - want) < 0.01, (k, f[k] / 1e6)


def test_step_flops_match_bench_constants():
    f = layer_flops()
    E, D, S, C, W = (f[k] for k in "EDSCW")
    c0, c1, c2, c3 = f["dis"]
    fc = f["dec_fc"]
    # Stage I: forward E + 2D + 3S; backward = discriminator class sweep [wgrad 3S + dgrad 3(S - c0) + 2 c0]
    # + feature-tap sweep, dgrad only [2 c3 + 3 (c2 + c1) + c0] + decoder sweep A on two calls [2 (2D - fc)] + decoder sweep B,
    # dgrad only [D] + encoder [2E - c0_enc]
    stage1 = (E + 2 * D + 3 * S) + (3 * S + 3 * (S - c0) + 2 * c0) + (2 * c3 + 3 * (c2 + c1) + c0) + 2 * (2 * D - fc) + D + \
             (2 * E - f["enc_c0"])
    b = _bench()
    assert abs(stage1 / 1e6 - b.ALG_MFLOP["stage1_vaegan"]) < 0.5, stage1 / 1e6
    # Stage III proper: forward C + 2D + 3S; backward = the two discriminator sweeps + decoder sweep A
    stage3 = (C + 2 * D + 3 * S) + (3 * S + 3 * (S - c0) + 2 * c0) + (2 * c3 + 3 * (c2 + c1) + c0) + 2 * (2 * D - fc)
    assert abs(stage3 / 1e6 - b.ALG_MFLOP["stage3_cognitive"]) < 0.5, stage3 / 1e6
    # Stage II (C3): forward C + E + 3D + 3S; backward [3S + 3(S - c0)] + feature-tap sweep + D (dgrad through the frozen
    # decoder) + (C + heads)
    stage2 = (C + E + 3 * D + 3 * S) + (3 * S + 3 * (S - c0)) + (2 * c3 + 3 * (c2 + c1) + c0) + D + C
    assert abs(stage2 / 1e6 - b.ALG_MFLOP["stage2_cognitive"]) < 1.0, stage2 / 1e6
    # Stage I WAE/GAN (C2): forward E + D + 3W; backward 4W + W + 2D + (2E - c0_enc)
    wae1 = (E + D + 3 * W) + 5 * W + 2 * D + (2 * E - f["enc_c0"])
    assert abs(wae1 / 1e6 - b.ALG_MFLOP["stage1_waegan"]) < 1.0, wae1 / 1e6
