"""Kernel-level parity at BASELINE-sized launches: the code paths that only engage at large batch (persistent implicit
GEMM with double-buffered TMEM accumulators, parity-merged scatter with per-tap column ranges, wave-split weight gradients,
persistent halo-tile edge kernels, register-accumulated BatchNorm statistics, one-wave BatchNorm backward grids) against a
plain PyTorch fp32 / fp64 reference of the same op ON THE SAME INPUTS -- not against the library itself.

Why this file exists ("What's weak" #2): whole-step comparisons of two bf16 runs cannot be tighter than the bf16
storage noise floor. Any perturbation d (even the 1e-7 of a different fp32 summation order) that passes through a bf16
rounding step comes out as sqrt(d * 2^-8): 1e-7 -> 2e-5 -> 3e-4 -> 1e-3 -> 2e-3 -> 3e-3 over five layers, which is what
scripts/layerdiff.py measures layer by layer, run-to-run at a FIXED batch size as well as between batch 64 and the same 64
samples tiled to 4096 (profiles/parity/r2_layerdiff_B64_vs_4096.txt). So the large-batch paths are certified here, one
launch at a time, where inputs are identical and only one rounding step separates kernel and reference.

Tolerances (rel-L2): bf16-stored outputs 3e-3 (one bf16 rounding: 2^-9 rms ~ 1.1e-3 .. 2e-3), fp32 outputs 2e-4,
BatchNorm sums 1e-5. Shapes: the Stage-I layers at 1024 samples per GPU (3072 discriminator images), plus the largest
launch of the batch-4096 step (12288 x 64 x 64 x 32 -> 128) for 64-bit offsets.
"""
import pytest
import torch
import torch.nn.functional as F

from thesis_fmri_reconstruction_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def setup_module(module):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rnd(x):
    return x.to(BF).float()


def _randn(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return rnd(torch.randn(shape, generator=g, device=DEV) * scale)


# (N, H, W, Cin, Cout): discriminator blocks 1-3 at 3 x 1024 images, encoder blocks at 1024, the largest launch of the step
CONV_FULL = [
    (3072, 64, 64, 32, 128),
    (3072, 32, 32, 128, 256),
    (3072, 16, 16, 256, 256),
    (1024, 32, 32, 64, 128),
    (1024, 16, 16, 128, 256),
    (12288, 64, 64, 32, 128),
]


@pytest.mark.parametrize("case", CONV_FULL)
def test_conv_s2_fullsize(case):
    N, H, W, Cin, Cout = case
    x = _randn((N, Cin, H, W), 1)
    w = _randn((Cout, Cin, 5, 5), 2, 0.05)
    d = L.conv_desc(N, H, W, Cin, Cout, 2, False, 0, BF)
    OH, OW = L.conv_out_hw(d)
    pack_f = torch.empty(L.conv_pack_elems(d), dtype=BF, device=DEV)
    pack_d = torch.empty(L.conv_pack_elems(d), dtype=BF, device=DEV)
    L.conv_pack_weights(d, w, pack_f, pack_d)
    xs = nhwc(x).to(BF)
    # ---- fprop + BatchNorm statistics
    y = torch.full((N, OH, OW, Cout), float("nan"), dtype=BF, device=DEV)
    ssum = torch.zeros(Cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros(Cout, dtype=torch.float64, device=DEV)
    L.conv_fprop(d, xs, w, pack_f, None, L.ACT_NONE, y, ssum, ssq)
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, stride=2, padding=2)
    e_f = rel(nchw(y), ref)
    del ref
    yf = y.float().reshape(-1, Cout)
    e_s = rel(ssum, yf.double().sum(0))
    e_q = rel(ssq, (yf.double() ** 2).sum(0))
    del yf, y
    # ---- data gradient
    dy = _randn((N, Cout, OH, OW), 3)
    dys = nhwc(dy).to(BF)
    dx = torch.full((N, H, W, Cin), float("nan"), dtype=BF, device=DEV)
    L.conv_dgrad(d, dys, w, pack_d, dx)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=2, padding=2)
    e_d = rel(nchw(dx), ref)
    del ref, dx
    # ---- weight gradient (fp32 output, split over pixel tiles)
    ws = torch.empty(max(1, L.conv_wgrad_workspace(d)), dtype=torch.uint8, device=DEV)
    dw = torch.full((Cout, Cin, 5, 5), float("nan"), dtype=torch.float32, device=DEV)
    L.conv_wgrad(d, xs, dys, dw, False, ws)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, 5, 5), dy, stride=2, padding=2)
    e_w = rel(dw, ref)
    print(f"conv s2 {case}: fprop {e_f:.2e} stats {e_s:.1e}/{e_q:.1e} dgrad {e_d:.2e} wgrad {e_w:.2e}")
    # wgrad bound 5e-4 at the largest case: the fp32 torch reference itself sums 5e7 products per weight in fp32 (measured
    # 2.4e-4 at 12288 images against 6e-5 at 3072; our kernel accumulates 128-pixel tiles in TMEM fp32 and adds them with red.add)
    assert e_f < 3e-3 and e_d < 3e-3 and e_w < (5e-4 if N > 4096 else 2e-4) and e_s < 1e-5 and e_q < 1e-5


# (N, H, W, Cin, Cout): decoder blocks 0-2 at 1024 samples (block 2 = the parity-merged scatter, 32 output channels)
CONVT_FULL = [(1024, 8, 8, 256, 256), (1024, 16, 16, 256, 128), (1024, 32, 32, 128, 32), (4096, 32, 32, 128, 32)]


@pytest.mark.parametrize("case", CONVT_FULL)
def test_convT_fullsize(case):
    N, H, W, Cin, Cout = case
    x = _randn((N, Cin, H, W), 4).requires_grad_(True)
    w = _randn((Cin, Cout, 5, 5), 5, 0.05).requires_grad_(True)
    d = L.conv_desc(N, H, W, Cin, Cout, 2, True, 1, BF)
    OH, OW = L.conv_out_hw(d)
    ref = F.conv_transpose2d(x, w, stride=2, padding=2, output_padding=1)
    dyr = _randn((N, Cout, OH, OW), 6)
    gx, gw = torch.autograd.grad(ref, (x, w), dyr)
    ref = ref.detach()
    pack_f = torch.empty(L.conv_pack_elems(d), dtype=BF, device=DEV)
    pack_d = torch.empty(L.conv_pack_elems(d), dtype=BF, device=DEV)
    wd = w.detach()
    L.conv_pack_weights(d, wd, pack_f, pack_d)
    xs = nhwc(x.detach()).to(BF)
    y = torch.full((N, OH, OW, Cout), float("nan"), dtype=BF, device=DEV)
    ssum = torch.zeros(Cout, dtype=torch.float64, device=DEV)
    ssq = torch.zeros(Cout, dtype=torch.float64, device=DEV)
    L.conv_fprop(d, xs, wd, pack_f, None, L.ACT_NONE, y, ssum, ssq)
    torch.cuda.synchronize()
    e_f = rel(nchw(y), ref)
    yf = y.float().reshape(-1, Cout).double()
    e_s, e_q = rel(ssum, yf.sum(0)), rel(ssq, (yf * yf).sum(0))
    del yf
    dys = nhwc(dyr).to(BF)
    dx = torch.full((N, H, W, Cin), float("nan"), dtype=BF, device=DEV)
    L.conv_dgrad(d, dys, wd, pack_d, dx)
    torch.cuda.synchronize()
    e_d = rel(nchw(dx), gx)
    ws = torch.empty(max(1, L.conv_wgrad_workspace(d)), dtype=torch.uint8, device=DEV)
    dw = torch.full((Cin, Cout, 5, 5), float("nan"), dtype=torch.float32, device=DEV)
    L.conv_wgrad(d, xs, dys, dw, False, ws)
    torch.cuda.synchronize()
    e_w = rel(dw, gw)
    print(f"convT {case}: fprop {e_f:.2e} stats {e_s:.1e}/{e_q:.1e} dgrad {e_d:.2e} wgrad {e_w:.2e}")
    assert e_f < 3e-3 and e_d < 3e-3 and e_w < 2e-4 and e_s < 1e-5 and e_q < 1e-5


@pytest.mark.parametrize("M,N,K", [(4096, 1024, 16384), (12288, 512, 16384), (4096, 16384, 128), (4096, 1024, 3620)])
def test_linear_fullsize(M, N, K):
    x = _randn((M, K), 7)
    w = _randn((N, K), 8, 0.02)
    dy = _randn((M, N), 9)
    d = L.linear_desc(M, N, K, BF)
    Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    xs = torch.zeros(M, Kp, dtype=BF, device=DEV)
    xs[:, :K] = x.to(BF)
    wp = torch.zeros(N, Kp, dtype=BF, device=DEV)
    wpt = torch.zeros(K, Np, dtype=BF, device=DEV)
    L.linear_pack_weights(d, w, wp, Kp, wpt, Np)
    y = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
    L.linear_fprop(d, xs, Kp, w, wp, Kp, None, L.ACT_NONE, y, N)
    dx = torch.full((M, Kp), float("nan"), dtype=BF, device=DEV)
    L.linear_dgrad(d, dy.to(BF), N, w, wpt, Np, dx, Kp)
    dw = torch.full((N, K), float("nan"), dtype=torch.float32, device=DEV)
    L.linear_wgrad(d, xs, Kp, dy.to(BF), N, dw, False)
    torch.cuda.synchronize()
    e_f, e_d, e_w = rel(y, x @ w.t()), rel(dx[:, :K], dy @ w), rel(dw, dy.t() @ x)
    print(f"linear {(M, N, K)}: fprop {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e}")
    assert e_f < 2e-4 and e_d < 3e-3 and e_w < 2e-4


@pytest.mark.parametrize("C,stride,N", [(32, 1, 3072), (64, 2, 1024)])
def test_edge_in_fullsize(C, stride, N):
    """Discriminator.conv[0] (3 -> 32, stride 1, bias + ReLU, three image sources) and Encoder.conv[0] (3 -> 64, stride 2)."""
    H = W = 64
    imgs = [_randn((N // 3 if C == 32 else N, 3, H, W), 10 + i).requires_grad_(True) for i in range(3 if C == 32 else 1)]
    w = _randn((C, 3, 5, 5), 14, 0.1).requires_grad_(True)
    b = _randn((C,), 15) if C == 32 else None
    cat = torch.cat(imgs, 0)
    pre = F.conv2d(cat.double(), w.double(), b.double() if b is not None else None, stride=stride, padding=2)
    d = L.edge_desc(N, H, W, C, stride, BF)
    ws = torch.empty(L.edge_workspace(d), dtype=torch.uint8, device=DEV)
    OH = (H - 1) // stride + 1
    y = torch.full((N, OH, OH, C), float("nan"), dtype=BF, device=DEV)
    act = L.ACT_RELU if C == 32 else L.ACT_NONE
    L.edge_in_fprop(d, [t.detach() for t in imgs], imgs[0].shape[0], w.detach(), b, act, y, ws)
    torch.cuda.synchronize()
    e_f = rel(nchw(y), torch.relu(pre) if C == 32 else pre)
    dyr = _randn(tuple(pre.shape), 16)
    grads = torch.autograd.grad(pre, imgs + [w], dyr.double())
    dys = nhwc(dyr).to(BF)
    dimg = torch.full((N, 3, H, W), float("nan"), dtype=torch.float32, device=DEV)
    L.edge_in_dgrad(d, dys, w.detach(), dimg, ws)
    dw = torch.full((C, 3, 5, 5), float("nan"), dtype=torch.float32, device=DEV)
    db = torch.zeros(C, device=DEV) if b is not None else None
    L.edge_in_wgrad(d, [t.detach() for t in imgs], imgs[0].shape[0], dys, dw, False, ws, db)
    torch.cuda.synchronize()
    e_d, e_w = rel(dimg, torch.cat(grads[:-1], 0)), rel(dw, grads[-1])
    e_b = rel(db, dyr.double().sum((0, 2, 3))) if db is not None else 0.0
    print(f"edge_in C={C} s={stride} N={N}: fprop {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e} dbias {e_b:.2e}")
    assert e_f < 3e-3 and e_d < 2e-4 and e_w < 2e-4 and e_b < 2e-4


def test_edge_out_fullsize():
    """Decoder.conv[3]: Conv2d(32, 3, 5, s1) + bias + tanh at 1024 images."""
    N, C, H, W = 1024, 32, 64, 64
    x = _randn((N, C, H, W), 17).requires_grad_(True)
    w = _randn((3, C, 5, 5), 18, 0.1).requires_grad_(True)
    b = _randn((3,), 19)
    pre = F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=2)
    d = L.edge_desc(N, H, W, C, 1, BF)
    ws = torch.empty(L.edge_workspace(d), dtype=torch.uint8, device=DEV)
    xs = nhwc(x.detach()).to(BF)
    img = torch.full((N, 3, H, W), float("nan"), dtype=torch.float32, device=DEV)
    L.edge_out_fprop(d, xs, w.detach(), b, L.ACT_TANH, img, ws)
    torch.cuda.synchronize()
    e_f = rel(img, torch.tanh(pre))
    dimg = _randn(tuple(pre.shape), 20)
    gx, gw = torch.autograd.grad(pre, (x, w), dimg.double())
    dx = torch.full((N, H, W, C), float("nan"), dtype=BF, device=DEV)
    L.edge_out_dgrad(d, dimg, w.detach(), dx, ws)
    dw = torch.full((3, C, 5, 5), float("nan"), dtype=torch.float32, device=DEV)
    L.edge_out_wgrad(d, xs, dimg, dw, False, ws)
    torch.cuda.synchronize()
    e_d, e_w = rel(nchw(dx), gx), rel(dw, gw)
    print(f"edge_out N={N}: fprop {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e}")
    assert e_f < 2e-4 and e_d < 3e-3 and e_w < 2e-4


@pytest.mark.parametrize("rows,C", [(3072 * 32 * 32, 128), (3072 * 16 * 16, 256), (1024 * 64 * 64, 32), (4096, 16384)])
def test_batchnorm_fullsize(rows, C):
    x = (_randn((rows, C), 21) * 2 + 0.5).to(BF).float().requires_grad_(True)
    g = torch.Generator(device=DEV).manual_seed(22)
    gamma = (torch.rand(C, generator=g, device=DEV) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g, device=DEV) * 0.1).requires_grad_(True)
    ref = torch.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.9, 1e-5))
    xs = x.detach().to(BF)
    s = torch.zeros(C, dtype=torch.float64, device=DEV)
    q = torch.zeros(C, dtype=torch.float64, device=DEV)
    L.colstats(xs, rows, C, s, q)
    mean, invstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    L.bn_finalize(s, q, rows, C, 1e-5, 0.9, mean, invstd, None, None)
    y = torch.empty(rows, C, dtype=BF, device=DEV)
    L.bn_apply(xs, y, rows, C, mean, invstd, gamma.detach(), beta.detach(), True)
    torch.cuda.synchronize()
    e_y = rel(y, ref)
    del y
    dy = _randn((rows, C), 23)
    gx, gg, gb = torch.autograd.grad(ref, (x, gamma, beta), dy)
    del ref
    dx = torch.empty(rows, C, dtype=BF, device=DEV)
    dgamma, dbeta = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ws = torch.empty(3 * C, dtype=torch.float64, device=DEV)
    L.bn_backward(xs, dy.to(BF), dx, rows, C, mean, invstd, gamma.detach(), beta.detach(), True, True, dgamma, dbeta,
                  False, ws)
    torch.cuda.synchronize()
    e_x, e_g, e_b = rel(dx, gx), rel(dgamma, gg), rel(dbeta, gb)
    print(f"batchnorm rows={rows} C={C}: y {e_y:.2e} dx {e_x:.2e} dgamma {e_g:.2e} dbeta {e_b:.2e}")
    assert e_y < 3e-3 and e_x < 4e-3 and e_g < 2e-4 and e_b < 2e-4
