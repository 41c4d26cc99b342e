"""Golden vectors for the reconstruction metrics: imports the UNMODIFIED /root/reference/train/train_utils.py (with the
script harness's matplotlib stub on the path -- matplotlib is absent from this image) and records PearsonCorrelation and
StructuralSimilarity on seeded inputs. Run in the build container:  python oracle/make_golden_metrics.py
Writes tests/golden/metrics_<case>.npz (inputs are regenerated from the seed by the tests).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {"64": (6, 3, 64, 64, 11), "100": (4, 3, 100, 100, 12), "ragged": (3, 3, 11, 37, 13), "grey": (5, 1, 32, 48, 14)}


def inputs(N, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(N, C, H, W, generator=g) * 2 - 1
    b = (a + 0.3 * torch.randn(N, C, H, W, generator=g)).clamp(-1, 1)
    return a, b


def main():
    # only here: the tests import `inputs` from this module and must not get the reference tree on their sys.path
    sys.path.insert(0, os.path.join(ROOT, "tests", "script_harness", "stubs"))
    sys.path.insert(1, "/root/reference")
    from train.train_utils import PearsonCorrelation, StructuralSimilarity

    pc, ss = PearsonCorrelation(), StructuralSimilarity()
    out = os.path.join(ROOT, "tests", "golden")
    for name, (N, C, H, W, seed) in CASES.items():
        a, b = inputs(N, C, H, W, seed)
        np.savez(os.path.join(out, f"metrics_{name}.npz"), shape=np.array([N, C, H, W, seed]),
                 pcc=float(pc(a, b)), ssim=float(ss(a, b)), pcc64=float(pc(a.double(), b.double())),
                 mse=float(torch.nn.MSELoss()(a, b)))     # the reference's SSIM window is fp32: no fp64 variant
        print(name, float(pc(a, b)), float(ss(a, b)))


if __name__ == "__main__":
    main()
