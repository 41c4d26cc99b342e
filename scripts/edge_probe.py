"""Times the halo-tile edge kernels at the batch-4096 step's shapes (CUDA events, 5 repetitions after warm-up). With
FMRI_HC_SKIP=1|2|4 one role of hconv_kernel is switched off (producer copies / epilogue stores / MMAs): the role whose
removal shortens the kernel most is the one that bounds it.  python scripts/edge_probe.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thesis_fmri_reconstruction_b200 import lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
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


def main():
    H = W = 64
    N3 = 3 * B
    imgs = [torch.randn(B, 3, H, W, device="cuda") for _ in range(3)]
    w_in = torch.randn(32, 3, 5, 5, device="cuda") * 0.1
    b_in = torch.randn(32, device="cuda")
    d3 = L.edge_desc(N3, H, W, 32, 1, BF)
    ws = torch.empty(L.edge_workspace(d3), dtype=torch.uint8, device="cuda")
    y = torch.empty(N3, H, W, 32, dtype=BF, device="cuda")
    dimg = torch.empty(N3, 3, H, W, device="cuda")
    dw = torch.empty(32, 3, 5, 5, device="cuda")
    db = torch.zeros(32, device="cuda")
    d1 = L.edge_desc(B, H, W, 32, 1, BF)
    x1 = torch.randn(B, H, W, 32, device="cuda").to(BF)
    w_out = torch.randn(3, 32, 5, 5, device="cuda") * 0.1
    b_out = torch.randn(3, device="cuda")
    img1 = torch.empty(B, 3, H, W, device="cuda")
    dx1 = torch.empty(B, H, W, 32, dtype=BF, device="cuda")
    dwo = torch.empty(3, 32, 5, 5, device="cuda")
    de = L.edge_desc(B, H, W, 64, 2, BF)
    ye = torch.empty(B, 32, 32, 64, dtype=BF, device="cuda")
    w_e = torch.randn(64, 3, 5, 5, device="cuda") * 0.1
    dwe = torch.empty(64, 3, 5, 5, device="cuda")
    rows = [
        ("edge_in_fprop  3->32 s1 (3B imgs)", lambda: L.edge_in_fprop(d3, imgs, B, w_in, b_in, L.ACT_RELU, y, ws), N3 * H * W * (32 * 2 + 12)),
        ("edge_in_dgrad  32->3 s1 (3B imgs)", lambda: L.edge_in_dgrad(d3, y, w_in, dimg, ws), N3 * H * W * (32 * 2 + 12)),
        ("edge_in_wgrad  3->32 s1 (3B imgs)", lambda: L.edge_in_wgrad(d3, imgs, B, y, dw, False, ws, db), N3 * H * W * (32 * 2 + 12)),
        ("edge_out_fprop 32->3 s1 (B imgs)", lambda: L.edge_out_fprop(d1, x1, w_out, b_out, L.ACT_TANH, img1, ws), B * H * W * (32 * 2 + 12)),
        ("edge_out_dgrad 3->32 s1 (B imgs)", lambda: L.edge_out_dgrad(d1, img1, w_out, dx1, ws), B * H * W * (32 * 2 + 12)),
        ("edge_out_wgrad 32->3 s1 (B imgs)", lambda: L.edge_out_wgrad(d1, x1, img1, dwo, False, ws), B * H * W * (32 * 2 + 12)),
        ("edge_in_fprop  3->64 s2 (B imgs)", lambda: L.edge_in_fprop(de, imgs[:1], B, w_e, None, L.ACT_NONE, ye, ws), B * (H * W * 12 + 32 * 32 * 64 * 2)),
        ("edge_in_wgrad  3->64 s2 (B imgs)", lambda: L.edge_in_wgrad(de, imgs[:1], B, ye, dwe, False, ws), B * (H * W * 12 + 32 * 32 * 64 * 2)),
    ]
    print(f"B={B} FMRI_HC_SKIP={os.environ.get('FMRI_HC_SKIP', '0')}")
    for name, fn, nbytes in rows:
        ms = timeit(fn)
        print(f"  {name:36s} {ms:7.3f} ms   algorithmic {nbytes / 1e9:6.2f} GB -> {nbytes / ms / 1e6:7.1f} GB/s")
    if os.environ.get("FMRI_PROBE_KERNELS"):      # per-kernel split of every entry point (CUPTI through torch.profiler)
        from torch.profiler import ProfilerActivity, profile

        for name, fn, _ in rows:
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
            parts = [(e.key[:70], e.device_time_total / 3 / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
            print(f"  {name}: " + "; ".join(f"{k} {v:.3f} ms" for k, v in sorted(parts, key=lambda t: -t[1])))


if __name__ == "__main__":
    main()
