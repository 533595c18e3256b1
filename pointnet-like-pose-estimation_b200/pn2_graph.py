"""CUDA-graph replay of a whole eval-mode forward for small batches (BASELINE.json configs[0]: B = 24).

At B = 24 the seven product kernels of a pointnet2_cls_ssg forward take 0.39 ms on a B200 but the forward takes
0.54 ms: the rest is launch latency (about 70 launches: the kernels, the FC head's GEMMs / BatchNorms, allocations).
``GraphedForward`` captures the forward once into a ``torch.cuda.CUDAGraph`` and replays it.

What stays outside the graph is the one piece of host state the reference has on this path: the FPS start indices are
drawn with ``torch.randint`` on the CPU default generator, once per sampling layer per forward (pointnet2_utils.py:59).
During capture every draw goes into its own pinned host buffer and is uploaded by a memcpy node of the graph; before
each replay the buffers are refilled with fresh draws from the same generator in the same order, so a seeded sequence of
graphed forwards picks exactly the centroids the eager sequence would.
"""
from __future__ import annotations

import torch

import pointnet2_utils as _pu

__all__ = ["GraphedForward"]


class _PinnedStarts:
    """fps_start_source for capture / replay: one pinned [B] int64 buffer per draw of a forward."""

    def __init__(self, source=None):
        self.buffers = []          # (pinned tensor, N)
        self.capturing = True
        self.source = source       # (B, N) -> int64 CPU tensor [B]; default = the reference's torch.randint draw

    def _draw(self, B, N):
        if self.source is None:
            return torch.randint(0, N, (B,), dtype=torch.long)
        return self.source(B, N).to("cpu", torch.long)

    def __call__(self, B, N):
        assert self.capturing, "the graphed forward draws its start indices through refill()"
        buf = torch.empty(B, dtype=torch.long).pin_memory()
        buf.copy_(self._draw(B, N))
        self.buffers.append((buf, N))
        return buf

    def refill(self):
        for buf, N in self.buffers:            # same order, same generator as the eager path
            buf.copy_(self._draw(buf.shape[0], N))


class GraphedForward:
    """``g = GraphedForward(model, *example_inputs); out = g(*inputs)`` -- eval mode, fixed shapes, CUDA tensors.

    Outputs are the graph's static output tensors (overwritten by the next call); clone them to keep them.
    ``start_source``: optional ``(B, N) -> int64 CPU tensor [B]`` that replaces the torch.randint draw of the FPS start
    indices, e.g. a ``pn2_shard.GlobalStarts`` so that batch-sharded replays pick the centroids of the unsharded run."""

    def __init__(self, model: torch.nn.Module, *example_inputs: torch.Tensor, warmup: int = 3, start_source=None):
        if model.training:
            raise ValueError("GraphedForward captures the eval-mode forward; call model.eval() first")
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise ValueError("GraphedForward needs CUDA example inputs")
        self.model = model
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream(device=example_inputs[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):            # packs the weights, sizes every allocation
                model(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.starts = _PinnedStarts(start_source)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), _pu.fps_start_source(self.starts), torch.cuda.graph(self.graph):
            self.static_out = model(*self.static_in)
        self.starts.capturing = False
        self.done = torch.cuda.Event()
        self.done.record()

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.done.synchronize()                # the previous replay's upload nodes have read the pinned buffers
        self.starts.refill()
        self.graph.replay()
        self.done.record()
        return self.static_out
