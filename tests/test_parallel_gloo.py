"""World-size-2 gloo test (CPU) of the utterance sharding used at N > 1 GPUs: sharded results equal the unsharded
ones because nothing on the path mixes utterances (checked here with the CPU oracle as the per-shard function)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from summarymixing_b200 import parallel as P


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = P.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [P.shard_bounds(n, r, world)[1] - P.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import smx_oracle as O
        from tests import _golden as G

        fx = G.Fixture("cell_sm_h4_swish")
        g = torch.Generator().manual_seed(7)
        B, T, D = 5, 37, fx.cfg["enc_dim"]
        x = torch.randn(B, T, D, generator=g)
        lens = torch.tensor([37, 20, 1, 37, 9])
        mask = torch.arange(T)[None] < lens[:, None]

        def fn(xs, ms):
            return O.summary_mixing(xs, fx.sd, mode=fx.cfg["mode"], act=fx.cfg["act"], src_padding_mask=ms)

        y = P.sharded_forward(fn, x, mask, gather=True)
        y_full = fn(x, mask)
        t = P.max_over_ranks(float(rank + 1))
        q.put((rank, float((y - y_full).abs().max()), tuple(y.shape), t))
    finally:
        dist.destroy_process_group()


def test_sharded_forward_equals_unsharded_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape, t in res:
        assert shape == (5, 37, 64) or shape[0] == 5
        assert err == 0.0, f"rank {rank}: sharded != unsharded ({err})"
        assert t == 2.0


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import smx_oracle as O
        from tests import _golden as G

        fx = G.Fixture("cell_sm_h4_swish")
        g = torch.Generator().manual_seed(8)
        B, T, D = 5, 37, fx.cfg["enc_dim"]
        x = torch.randn(B, T, D, generator=g)
        lens = torch.tensor([37, 20, 1, 37, 9])
        mask = torch.arange(T)[None] < lens[:, None]

        def grads(xs, ms):
            sd = {k: torch.nn.Parameter(v.clone()) for k, v in fx.sd.items()}
            if xs.shape[0] > 0:
                y = O.summary_mixing(xs, sd, mode=fx.cfg["mode"], act=fx.cfg["act"], src_padding_mask=ms)
                (y * ms[..., None]).pow(2).sum().backward()
            return sd

        full = grads(x, mask)
        xs, ms = P.shard_batch(x, mask, rank, world)
        mine = grads(xs, ms)
        calls = P.allreduce_gradients(list(mine.values()), bucket_bytes=16 << 10)  # small buckets: several collectives
        err = max(float((mine[k].grad * world - full[k].grad).abs().max() / (1.0 + full[k].grad.abs().max())) for k in full)
        q.put((rank, err, calls))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_equals_full_batch_gradient_world2():
    """Data-parallel training plumbing: per-shard gradients averaged by allreduce_gradients x world == the gradient of
    the whole batch (the loss is a sum over utterances), over several buckets."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, calls in res:
        assert err < 1e-5, f"rank {rank}: averaged shard gradients != full-batch gradient ({err})"
        assert calls >= 2


def _bucketer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import smx_oracle as O
        from tests import _golden as G

        fx = G.Fixture("cell_sm_h4_swish")
        g = torch.Generator().manual_seed(9)
        B, T, D = 5, 37, fx.cfg["enc_dim"]
        x = torch.randn(B, T, D, generator=g)
        lens = torch.tensor([37, 20, 1, 37, 9])
        mask = torch.arange(T)[None] < lens[:, None]
        xs, ms = P.shard_batch(x, mask, rank, world)

        def run(overlapped):
            sd = {k: torch.nn.Parameter(v.clone()) for k, v in fx.sd.items()}
            params = list(sd.values())
            bk = P.GradientBucketer(params, bucket_bytes=16 << 10) if overlapped else None  # hooks fire during backward
            y = O.summary_mixing(xs, sd, mode=fx.cfg["mode"], act=fx.cfg["act"], src_padding_mask=ms)
            (y * ms[..., None]).pow(2).sum().backward()
            calls = bk.finish() if overlapped else P.allreduce_gradients(params, bucket_bytes=16 << 10)
            return [p.grad.clone() for p in params], calls

        ga, ca = run(True)
        gb, cb = run(False)
        err = max(float((a - b).abs().max()) for a, b in zip(ga, gb))
        q.put((rank, err, ca, cb))
    finally:
        dist.destroy_process_group()


def test_overlapped_bucketer_matches_plain_allreduce_world2():
    """GradientBucketer (bucket all-reduces launched from gradient-ready hooks DURING backward) gives the same averaged
    gradients as the after-the-fact allreduce_gradients; several buckets, including one completed only by finish()."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucketer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ca, cb in res:
        assert err < 1e-6, f"rank {rank}: overlapped != plain ({err})"
        assert ca >= 2 and cb >= 2
