"""CUDA-graph replay of a module forward (fixed shapes): the whole encoder step becomes one graph launch.

libsmx is enqueue-only on the caller's stream, allocates nothing and keeps its programmatic-dependent-launch edges under
capture, so `torch.cuda.graph` can record a forward of any module of this package; a replay is bit-identical to the
eager call (tests/test_tc_path_gpu.py::test_cuda_graph_replay_matches_eager) and removes the host-side launch work
(85 launches per 12-layer step)."""
from __future__ import annotations

import torch


class GraphedForward:
    """g = GraphedForward(encoder, x_example, mask_example); y = g(x, mask)

    x / mask are copied into the graph's static input buffers (device-to-device) and the graph is replayed; the returned
    tensor is the graph's static output (overwritten by the next call).  `static_x` / `static_mask` may also be filled
    directly (e.g. by a host-to-device copy on another stream) followed by `replay()`."""

    def __init__(self, module: torch.nn.Module, x: torch.Tensor, mask: torch.Tensor | None = None, warmup: int = 3):
        if not x.is_cuda:
            raise RuntimeError("GraphedForward: CUDA tensors only (the product path has no CPU fallback)")
        self.module = module
        self.static_x = x.clone()
        self.static_mask = mask.clone() if mask is not None else None
        side = torch.cuda.Stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):
                self._call()
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            out = self._call()
        self.static_out = out[0] if isinstance(out, tuple) else out

    def _call(self):
        if self.static_mask is None:
            return self.module(self.static_x)
        return self.module(self.static_x, src_key_padding_mask=self.static_mask)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_out

    def __call__(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        self.static_x.copy_(x, non_blocking=True)
        if self.static_mask is not None and mask is not None:
            self.static_mask.copy_(mask, non_blocking=True)
        return self.replay()
