"""World-size-2 gloo test of the data-parallel host logic (thesis_fmri_reconstruction_b200/dp.py) on CPU.

Each rank runs the CPU oracle's Stage-I step on its shard of a global batch (per-rank BatchNorm, as on the GPUs), puts
the three gradient buckets into FlatBuckets, SUM-all-reduces them, and reduces the BCE sums for the gate. The result must
equal the single-process emulation: shards processed one after the other, gradients summed.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(rank, world, B):
    from oracle import vaegan as O
    from thesis_fmri_reconstruction_b200 import dp

    P, S = O.make_vaegan(O.CFG64, seed=99)
    x = O.synthetic_images(B, seed=99)
    eps, z_p = O.synthetic_noise(B, 128, seed=99)
    lo, hi = dp.shard_range(B, rank, world)
    out = O.stage1_vaegan_step(P, S, x[lo:hi], eps[lo:hi], z_p[lo:hi], update=False)
    sums = torch.tensor([out["bce_o"].sum().item(), out["bce_p"].sum().item()])
    return P, out["grads"], sums


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as td

    from thesis_fmri_reconstruction_b200 import dp

    torch.set_num_threads(2)
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P, grads, sums = _shard_grads(rank, world, B)
    buckets = {}
    for pre in ("encoder.", "decoder.", "discriminator."):
        named = {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
        b = dp.FlatBucket(pre, named, 1, "cpu")
        for k in b.G:
            b.G[k].copy_(grads[pre + k])
        buckets[pre] = b
    dp.allreduce_sum_([b.flat_g for b in buckets.values()] + [sums])
    gate = dp.gate_from_sums(sums[0].item(), sums[1].item(), B, 0.35, 0.68)
    norms = {pre: b.flat_g.double().norm().item() for pre, b in buckets.items()}
    sample = {pre: b.flat_g[:: max(1, b.numel // 64)].clone() for pre, b in buckets.items()}
    q.put((rank, norms, sample, sums.tolist(), gate))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_allreduce_equals_sequential_shards():
    world, B = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process emulation: the same shards one after the other, gradients summed
    sys.path.insert(0, ROOT)
    from thesis_fmri_reconstruction_b200 import dp

    torch.set_num_threads(2)  # same reduction order as the workers
    tot, sums = None, torch.zeros(2)
    for r in range(world):
        P, grads, s = _shard_grads(r, world, B)
        sums += s
        tot = grads if tot is None else {k: tot[k] + grads[k] for k in grads}
    for pre in ("encoder.", "decoder.", "discriminator."):
        named = {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
        b = dp.FlatBucket(pre, named, 1, "cpu")
        for k in b.G:
            b.G[k].copy_(tot[pre + k])
        want = b.flat_g.double().norm().item()
        for rank, norms, sample, rs, gate in res:
            assert abs(norms[pre] - want) <= 1e-5 * want, (pre, rank, norms[pre], want)
            ref_s = b.flat_g[:: max(1, b.numel // 64)]
            assert (sample[pre] - ref_s).abs().max().item() <= 1e-4 * ref_s.abs().max().item()
    # every rank sees the same global sums and therefore the same gate decision
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]
    assert abs(res[0][3][0] - sums[0].item()) < 1e-4 and res[0][4] == dp.gate_from_sums(sums[0].item(), sums[1].item(), B, 0.35, 0.68)


def test_shard_range_and_bucket_views():
    from thesis_fmri_reconstruction_b200 import dp

    assert dp.shard_range(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        dp.shard_range(10, 0, 4)
    named = {"a": torch.arange(6.0).view(2, 3), "b": torch.ones(5)}
    b = dp.FlatBucket("x.", named, 2, "cpu")
    assert b.numel == 8 + 8 and b.P["a"].shape == (2, 3)
    b.P["b"].mul_(3)
    assert b.flat_p[8:13].tolist() == [3.0] * 5          # views alias the flat buffer
    assert b.state_view(1, "a").shape == (2, 3)
