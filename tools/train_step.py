"""One data-parallel training step of the SummaryMixing-Conformer encoder (BASELINE configs[1] shape), timed.

    python tools/train_step.py [--steps 5 --warmup 2 --layers 12 --B 32 --T 1000 --D 256 --dtype fp32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py ...

Per step and per rank: forward through the autograd chain of libsmx calls, a masked mean-square loss, backward through
smx_*_bwd, a bucketed NCCL gradient all-reduce (summarymixing_b200.parallel.allreduce_gradients) and an SGD update.  Each
rank holds its own B x T batch (weak scaling, like bench.py).  Times: CUDA events around the K steps, max over ranks.
The backward is the first-correct fp32-math arm: this tool checks the training path end to end (loss goes down, ranks
stay in sync) and puts a number on it; it is not the headline metric (bench.py: encoder forward)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import summarymixing_b200 as S  # noqa: E402
from summarymixing_b200 import _lib as L  # noqa: E402
from summarymixing_b200 import parallel as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--D", type=int, default=256)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--ddp", action="store_true", help="wrap the encoder in torch DistributedDataParallel (what SpeechBrain's "
                    "Brain does) instead of calling parallel.allreduce_gradients")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # same initial weights on every rank
    enc = S.ConformerEncoder(a.layers, a.D, 4 * a.D, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[a.D],
                             local_proj_out_dim=a.D, summary_hid_dim=[a.D], dropout=0.0).to(dev).train()
    params = list(enc.parameters())
    model = enc
    if a.ddp and world > 1:
        model = torch.nn.parallel.DistributedDataParallel(enc, device_ids=[local])
    opt = torch.optim.SGD(params, lr=0.02)
    g = torch.Generator().manual_seed(100 + rank)  # a different batch on every rank
    dt = torch.float32 if a.dtype == "fp32" else torch.bfloat16
    x = torch.randn(a.B, a.T, a.D, generator=g).to(dev).to(dt)
    lens = torch.randint(a.T // 2, a.T + 1, (a.B,), generator=g)
    mask = (torch.arange(a.T)[None] < lens[:, None]).to(dev)
    target = torch.randn(a.B, a.T, a.D, generator=g).to(dev)
    losses, calls, host_ms = [], 0, 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.lib().smx_launch_count()
    for i in range(a.warmup + a.steps):
        if i == a.warmup:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            n0 = L.lib().smx_launch_count()
            ev0.record()
        h0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        y = model(x, src_key_padding_mask=mask)[0]
        loss = ((y.float() - target) * mask[..., None]).pow(2).mean()
        loss.backward()
        calls = -1 if model is not enc else P.allreduce_gradients(params)  # (-1: DDP's own bucketed all-reduce)
        opt.step()
        if i >= a.warmup:
            host_ms += (time.perf_counter() - h0) * 1e3   # host time to ENQUEUE the step (the loss read-back below waits for the GPU)
        losses.append(float(loss.detach()))
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = P.max_over_ranks(ev0.elapsed_time(ev1) / a.steps, device=dev)
    launches = (L.lib().smx_launch_count() - n0) // a.steps
    # ranks must hold identical weights after the synchronised updates
    chk = torch.stack([p.detach().float().sum() for p in params]).sum().reshape(1).double()
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "encoder training step (fwd + bwd + grad all-reduce + SGD)", "n_gpus": world, "io": a.dtype,
                          "layers": a.layers, "B_per_gpu": a.B, "T": a.T, "D": a.D, "ms_per_step": ms, "host_enqueue_ms_per_step": host_ms / a.steps,
                          "frames_per_s": world * a.B * a.T / (ms * 1e-3), "libsmx_launches_per_step": int(launches),
                          "allreduce_calls_per_step": calls, "params": sum(p.numel() for p in params),
                          "loss_first": losses[0], "loss_last": losses[-1],
                          "weights_in_sync": bool(float(hi - lo) == 0.0)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
