"""Shows whether the gradient all-reduces of the data-parallel Stage-I step overlap the backward kernels: one
step on every rank under torch.profiler (CUPTI), then rank 0 prints, for each NCCL kernel, when it ran relative to the step,
how long it took and how much of that time a compute kernel of this library was running on the other stream.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        scripts/overlap_trace.py [global_batch]
"""
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thesis_fmri_reconstruction_b200 import engine, hp, init  # noqa: E402


def main():
    gb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = gb // world
    P, S = init.init_vaegan(hp.CFG64, 128, seed=12345)
    tr = engine.VaeGanStage1(P, S, hp.CFG64, 128, torch.bfloat16, dist_group=td.group.WORLD, gate=True)
    g = torch.Generator().manual_seed(1234 + rank)
    x = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).cuda()
    eps, zp = torch.randn(B, 128, generator=g).cuda(), torch.randn(B, 128, generator=g).cuda()
    for _ in range(3):
        tr.step(x, eps, zp)
    td.barrier()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    from thesis_fmri_reconstruction_b200 import lib as L

    marker = torch.zeros(1, dtype=torch.int32, device="cuda")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):          # the first profiled steps absorb CUPTI start-up skew between the ranks; the LAST one is reported
            td.barrier()
            torch.cuda.synchronize()
            L.step_increment(marker)    # a marker kernel the RMSprop step never launches: splits the trace into steps
            tr.step(x, eps, zp)
        torch.cuda.synchronize()
    td.barrier()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
        ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
        marks = [i for i, k in enumerate(ks) if "step_increment" in k[2]]
        ks = ks[marks[-1] + 1:]
        t0, t1 = ks[0][0], max(k[1] for k in ks)
        comp = [(a, b) for a, b, n in ks if "fmri::" in n]
        nccl = [(a, b, n) for a, b, n in ks if "nccl" in n.lower()]
        print(f"Stage-I step, global batch {gb} on {world} GPUs ({B} per rank), rank 0: step {1e-3 * (t1 - t0):.3f} ms, "
              f"{len(comp)} kernels of libfmri_b200 ({1e-3 * sum(b - a for a, b in comp):.3f} ms), {len(nccl)} NCCL kernels")
        print(f"{'NCCL kernel':44s} {'start ms':>9s} {'dur ms':>8s} {'overlapped by compute':>22s}")
        tot = ovl = 0.0
        for a, b, n in nccl:
            o = sum(max(0, min(b, d) - max(a, c)) for c, d in comp)
            tot += b - a
            ovl += o
            print(f"{n[:44]:44s} {1e-3 * (a - t0):9.3f} {1e-3 * (b - a):8.3f} {100 * o / max(b - a, 1):21.1f}%")
        last_comp = max(d for _, d in comp)
        print(f"NCCL total {1e-3 * tot:.3f} ms, of which {1e-3 * ovl:.3f} ms ({100 * ovl / max(tot, 1):.1f} %) ran under compute kernels; "
              f"exposed tail after the last compute kernel: {1e-3 * max(0, max(b for _, b, _ in nccl) - last_comp):.3f} ms")
    td.destroy_process_group()


if __name__ == "__main__":
    main()
