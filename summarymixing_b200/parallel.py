"""Multi-GPU host plumbing for the encoder path: utterance sharding, one process per GPU.

The SummaryMixing cell reduces over time inside one utterance only (summary_mixing.py:229-231, dim=1), and every
other op of the encoder block is per-frame or per-utterance (the depthwise conv never crosses utterances), so the
forward path shards over utterances with NO data-path collective.  `torch.distributed` (NCCL over NVLink on the
B200 box, gloo in the CPU tests) is used only to agree on timing / to gather results when the caller wants them
on one rank.  Training adds the one real exchange step of data-parallel SGD — a gradient all-reduce per step
(`allreduce_gradients`, bucketed; the reference gets it implicitly from SpeechBrain's DDP wrapper).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_utts: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n_utts utterances: ranks < n_utts % world get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_utts, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, mask: Optional[torch.Tensor], rank: int, world: int):
    """This rank's utterances of a padded batch x (B,T,D) and its (B,T) padding mask (1/True = valid)."""
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi], (None if mask is None else mask[lo:hi])


def sharded_forward(fn: Callable, x: torch.Tensor, mask: Optional[torch.Tensor], gather: bool = False,
                    group=None, out_dim: Optional[int] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Run `fn(x_shard, mask_shard) -> (b,T,D')` on this rank's utterances.  With gather=True every rank returns the
    full (B,T,D') result (all_gather of equal-size padded shards); otherwise only its own shard.
    `out_dim` / `out_dtype`: feature dim and dtype of fn's result when they differ from x's (a cell with
    summary_out_dim != enc_dim): a rank whose shard is empty never runs fn and needs them to build its (empty) share."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    xs, ms = shard_batch(x, mask, rank, world)
    if xs.shape[0] > 0:
        y = fn(xs, ms)
        if out_dim is not None and y.shape[-1] != out_dim:
            raise ValueError(f"sharded_forward: fn returned feature dim {y.shape[-1]}, out_dim says {out_dim}")
    else:
        y = torch.zeros((0,) + tuple(x.shape[1:-1]) + (out_dim if out_dim is not None else x.shape[-1],),
                        dtype=out_dtype or x.dtype, device=x.device)
    if not gather or world == 1:
        return y
    B = x.shape[0]
    per = (B + world - 1) // world
    pad = y.new_zeros((per,) + tuple(y.shape[1:]))
    pad[: y.shape[0]] = y
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(B, r, world)
        parts.append(outs[r][: hi - lo])
    return torch.cat(parts, dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing helper: the slowest rank's value (bench.py reports device time as the max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def allreduce_gradients(params, world: Optional[int] = None, bucket_bytes: int = 64 << 20, group=None) -> int:
    """Average `.grad` of `params` over the ranks (data-parallel training: every rank back-propagated its own utterance
    shard).  Gradients are packed into flat fp32 buckets of about `bucket_bytes` — NVSwitch all-reduces are latency-, not
    link-bound, so a few large buckets beat one call per tensor (17.5 M parameters for the D=256 encoder = 70 MB = two
    buckets) — reduced with ONE all_reduce each (NCCL on the GPU box, gloo in the CPU tests) and scattered back in place.
    Every rank must pass the same parameters in the same order; a parameter without a gradient counts as zeros (ranks
    whose shard is empty still take part).  Returns the number of collectives issued."""
    if not dist.is_initialized():
        return 0
    world = dist.get_world_size(group) if world is None else world
    if world == 1:
        return 0
    params = [p for p in params if p.requires_grad]
    calls, i = 0, 0
    while i < len(params):
        j, nbytes = i, 0
        while j < len(params) and (j == i or nbytes + params[j].numel() * 4 <= bucket_bytes):
            nbytes += params[j].numel() * 4
            j += 1
        chunk = params[i:j]
        flat = torch.zeros(nbytes // 4, dtype=torch.float32, device=chunk[0].device)
        off = 0
        for p in chunk:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for p in chunk:
            g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        calls += 1
        i = j
    return calls


class GradientBucketer:
    """Data-parallel gradient exchange OVERLAPPED with the backward pass (what SpeechBrain's DistributedDataParallel wrapper does
    for the reference): parameters are grouped into flat fp32 buckets in reverse registration order (the order in
    which backward produces their gradients); a post-accumulate hook copies each gradient into its bucket, and the moment a
    bucket is complete its all-reduce is launched asynchronously (NCCL: on the communicator's own stream, under the rest of the
    backward).  `finish()` waits for the collectives, divides by the world size and hands the averaged gradients back.

        bucketer = GradientBucketer(model.parameters())          # once
        loss.backward(); bucketer.finish(); optimizer.step()     # every step

    Every rank must build it over the same parameters in the same order.  A parameter that received no gradient in a step counts
    as zeros (its bucket is completed by finish()).  Results are identical to `allreduce_gradients` (same sums, same order)."""

    def __init__(self, params, bucket_bytes: int = 32 << 20, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = []   # (flat, [(param, offset)], )
        order = list(reversed(self.params))
        i = 0
        while i < len(order):
            j, n = i, 0
            while j < len(order) and (j == i or (n + order[j].numel()) * 4 <= bucket_bytes):
                n += order[j].numel()
                j += 1
            flat = torch.zeros(n, dtype=torch.float32, device=order[i].device)
            items, off = [], 0
            for p in order[i:j]:
                items.append((p, off))
                off += p.numel()
            self.buckets.append({"flat": flat, "items": items, "ready": 0, "work": None, "seen": set()})
            i = j
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for p, off in b["items"]:
                self._where[id(p)] = (bi, off)
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params] if self.world > 1 else []
        self.calls = 0

    def _launch(self, b):
        b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.calls += 1

    def _hook(self, p):
        bi, off = self._where[id(p)]
        b = self.buckets[bi]
        if id(p) in b["seen"]:
            return
        b["seen"].add(id(p))
        b["flat"][off:off + p.numel()].copy_(p.grad.reshape(-1))
        b["ready"] += 1
        if b["ready"] == len(b["items"]):
            self._launch(b)

    def finish(self) -> int:
        """Wait for the bucket all-reduces (launching those whose parameters got no gradient this step), average, write back."""
        if self.world == 1:
            return 0
        for b in self.buckets:
            if b["work"] is None:
                for p, off in b["items"]:
                    if id(p) not in b["seen"]:
                        if p.grad is None:
                            b["flat"][off:off + p.numel()].zero_()
                        else:
                            b["flat"][off:off + p.numel()].copy_(p.grad.reshape(-1))
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
            b["flat"].div_(self.world)
            for p, off in b["items"]:
                g = b["flat"][off:off + p.numel()].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            b["ready"], b["work"] = 0, None
            b["seen"].clear()
        n, self.calls = self.calls, 0
        return n

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
