#!/usr/bin/env python
"""Per-call device times of one Stage-I VAE/GAN step (CUDA events around every C-ABI entry point, everything on the
compute stream), with the algorithmic FLOPs of the tensor-core calls -> TFLOP/s per call.
usage (GPU box): python scripts/step_profile.py [--batch 4096] > gpurun_out/per_call.txt"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thesis_fmri_reconstruction_b200 import engine, hp, init, lib, nets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
a = ap.parse_args()
B, z = a.batch, 128
P, S = init.init_vaegan(hp.CFG64, z, seed=12345)
tr = engine.VaeGanStage1(P, S, hp.CFG64, z, torch.bfloat16)
g = torch.Generator().manual_seed(1234)
x = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).cuda()
n1, n2 = torch.randn(B, z, generator=g).cuda(), torch.randn(B, z, generator=g).cuda()
nets.WGRAD_SIDE_STREAM = False
for _ in range(3):
    tr.step(x, n1, n2)
torch.cuda.synchronize()
lib.profile_begin()
tr.step(x, n1, n2)
torch.cuda.synchronize()
calls = [(name, fl, e0.elapsed_time(e1)) for name, fl, e0, e1 in lib.PROF]
lib.profile_end()
tot = sum(c[2] for c in calls)
print(f"# Stage-I VAE/GAN step, batch {B}, one GPU: {len(calls)} entry-point calls, {tot:.2f} ms of device time")
print(f"# {'#':>3} {'entry point':32} {'ms':>8} {'GFLOP':>10} {'TFLOP/s':>8}")
for i, (name, fl, ms) in enumerate(calls):
    tf = f"{fl / ms / 1e9:8.1f}" if fl else "        "
    print(f"  {i:3d} {name:32} {ms:8.3f} {fl / 1e9:10.1f} {tf}")
