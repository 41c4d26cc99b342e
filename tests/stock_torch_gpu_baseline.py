"""Stock PyTorch on one B200 (cuDNN / cuBLAS, ATen autograd): the practical bar BASELINE.md section 3 asks to report next to
the hand-written path.  NOT a pytest file and NOT part of the product: it runs the oracle restatement of the reference
step (oracle/vaegan.py -- the same torch.nn.functional calls the reference modules make, models/vae_gan.py:11-187, and
the update order of train/train_vgan_stage1.py:316-432 in its torch-1.4 "needed gradients" form) with every tensor on
cuda:0, in three arithmetic modes: fp32 (TF32 off), fp32 with TF32 allowed, and bf16 autocast (suffix _cl = images
and 4-D weights in channels_last memory format, cuDNN's preferred layout).

  python tests/stock_torch_gpu_baseline.py --workload stage1_vaegan --batch 1024 [--steps 10 --warmup 3]

Prints one JSON line per mode.  Timing: CUDA events around `steps` whole steps (forward, three gradient sweeps, gate,
RMSprop/Adam update), after warm-up (cudnn.benchmark autotuning happens there).
"""
import argparse
import json
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vaegan as O  # noqa: E402


CL = False  # channels_last operands (modes *_cl): cuDNN's preferred NHWC layout for tensor-core convolutions


def _cl(v):
    return v.contiguous(memory_format=torch.channels_last) if CL and v.dim() == 4 else v


def to_dev(d):
    return OrderedDict((k, _cl(v.cuda())) for k, v in d.items())


def build(workload, B):
    if workload == "stage1_vaegan":
        P, S = O.make_vaegan(O.CFG64, seed=12345, jitter=False)
        P, S = to_dev(P), to_dev(S)
        x = _cl(O.synthetic_images(B).cuda())
        eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128)]
        st = dict(P=P, sq=None)

        def one():
            out = O.stage1_vaegan_step(st["P"], S, x, eps, z_p, sq=st["sq"], force_gate=(True, True))
            st["P"], st["sq"] = out["params"], out["square_avg"]
    elif workload == "stage1_waegan":
        P, S = O.make_waegan(O.CFG64, seed=12345, jitter=False)
        P, S = to_dev(P), to_dev(S)
        x = _cl(O.synthetic_images(B).cuda())
        z_fake = (O.synthetic_noise(B, 128)[0] * 0.5).cuda()
        st = dict(P=P, opt=None, t=1)

        def one():
            out = O.stage1_waegan_step(st["P"], S, x, z_fake, opt=st["opt"], step=st["t"])
            st["P"], st["opt"], st["t"] = out["params"], out["adam"], st["t"] + 1
    else:
        stage = 2 if workload == "stage2_cognitive" else 3
        P, S = O.make_cognitive(O.CFG64, seed=12345, jitter=False)
        P, S = to_dev(P), to_dev(S)
        fmri, x = O.synthetic_fmri(B).cuda(), _cl(O.synthetic_images(B).cuda())
        eps, z_p = [t.cuda() for t in O.synthetic_noise(B, 128)]
        eps_t = O.synthetic_noise(B, 128, seed=99)[0].cuda()
        st = dict(P=P, sq=None)

        def one():
            out = O.cognitive_vaegan_step(st["P"], S, fmri, x, eps, eps_t, z_p, stage, sq=st["sq"],
                                          force_gate=(True, True))
            st["P"], st["sq"] = out["params"], out["square_avg"]
    return one


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="stage1_vaegan",
                    choices=["stage1_vaegan", "stage1_waegan", "stage2_cognitive", "stage3_cognitive"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--modes", default="fp32,tf32,bf16")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    global CL
    for mode in a.modes.split(","):
        CL = mode.endswith("_cl")
        torch.backends.cuda.matmul.allow_tf32 = mode.startswith("tf32")
        torch.backends.cudnn.allow_tf32 = mode.startswith("tf32")
        line = dict(impl="stock torch %s (cuDNN %s)" % (torch.__version__, torch.backends.cudnn.version()), mode=mode,
                    workload=a.workload, batch=a.batch, steps=a.steps, warmup=a.warmup)
        one = None
        try:
            one = build(a.workload, a.batch)
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode.startswith("bf16") else torch.autocast("cuda", enabled=False)
            with ctx:
                for _ in range(a.warmup):
                    one()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    one()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            line.update(ms_per_step=ms, samples_per_s=a.batch / ms * 1e3,
                        peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
        except Exception as ex:  # out of memory at a large batch is reported, not fatal
            line.update(error=repr(ex)[:200])
        print(json.dumps(line), flush=True)
        del one
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
