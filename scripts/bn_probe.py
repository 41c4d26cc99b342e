"""Times fmri_bn_apply / fmri_bn_backward at the Stage-I layer shapes of batch 4096 (CUDA events, 5 repetitions) and prints the
effective HBM bandwidth of each (bytes the algorithm has to move / time).  python scripts/bn_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thesis_fmri_reconstruction_b200 import lib as L  # noqa: E402

BF = torch.bfloat16


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rows, C in ((12288 * 32 * 32, 128), (12288 * 16 * 16, 256), (4096 * 64 * 64, 32), (4096 * 32 * 32, 128)):
    x = torch.randn(rows, C, device="cuda").to(BF)
    dy = torch.randn(rows, C, device="cuda").to(BF)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    mean, invstd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ws = torch.empty(3 * C, dtype=torch.float64, device="cuda")
    t_a = timeit(lambda: L.bn_apply(x, y, rows, C, mean, invstd, g, b, True))
    t_b = timeit(lambda: L.bn_backward(x, dy, dx, rows, C, mean, invstd, g, b, True, True, dg, db, False, ws))
    n = rows * C * 2
    print(f"rows={rows} C={C}: apply {t_a:.3f} ms = {2 * n / t_a / 1e6:.0f} GB/s   backward {t_b:.3f} ms = {5 * n / t_b / 1e6:.0f} GB/s")
