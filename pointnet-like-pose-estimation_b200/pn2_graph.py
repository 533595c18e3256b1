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
        buf = torch.zeros(B, dtype=torch.long).pin_memory()      # placeholder: refill() writes the real draw before every replay,
        self.buffers.append((buf, N))                            # so capturing consumes nothing from the generator
        return buf

    def refill(self):
        for buf, N in self.buffers:            # same order, same generator as the eager path
            buf.copy_(self._draw(buf.shape[0], N))


def _weights_fingerprint(model: torch.nn.Module):
    """What the captured kernels' weight pointers were derived from: every parameter / buffer (storage + version), the
    precision mode and the cache generation (pointnet2_utils.invalidate_caches)."""
    items = tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers()))
    return (_pu.get_precision(), _pu._cache_generation, items)


def _derived_buffers(model: torch.nn.Module):
    """The packed-weight tensors the fused kernels read (derived caches of the modules, not parameters).  The graph bakes
    their device pointers in, so it must keep them alive itself: a cache that repacks drops its old tensor."""
    keep = []
    for m in model.modules():
        for v in vars(m).values():
            for c in (v if isinstance(v, (list, tuple)) else (v,)):
                if isinstance(c, (_pu._FoldedMLP, _pu._FactorisedFirstLayer, _pu._FoldedHead)):
                    for t in vars(c).values():
                        keep += [u for u in (t if isinstance(t, (list, tuple)) else (t,)) if isinstance(u, torch.Tensor)]
    return keep


class GraphedForward:
    """``g = GraphedForward(model, *example_inputs); out = g(*inputs)`` -- eval mode, fixed shapes, CUDA tensors.

    Outputs are the graph's static output tensors (overwritten by the next call); clone them to keep them.
    ``start_source``: optional ``(B, N) -> int64 CPU tensor [B]`` that replaces the torch.randint draw of the FPS start
    indices, e.g. a ``pn2_shard.GlobalStarts`` so that batch-sharded replays pick the centroids of the unsharded run.

    Weights: the fused kernels read BN-folded, packed copies of the parameters (derived caches), whose device pointers
    the graph bakes in.  The graph keeps those buffers alive and, before every replay, compares the parameters' storage
    and version counters (and the precision mode) with what it captured: after an optimizer step, ``load_state_dict``,
    ``set_precision`` or ``pointnet2_utils.invalidate_caches()`` it RE-CAPTURES instead of replaying stale weights.
    (Writes through ``param.data`` bump no counter -- call ``invalidate_caches()`` after those.)"""

    def __init__(self, model: torch.nn.Module, *example_inputs: torch.Tensor, warmup: int = 3, start_source=None):
        if model.training:
            raise ValueError("GraphedForward captures the eval-mode forward; call model.eval() first")
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise ValueError("GraphedForward needs CUDA example inputs")
        self.model = model
        self.static_in = [t.clone() for t in example_inputs]
        self.warmup = warmup
        self.start_source = start_source
        self.captures = 0
        self._capture()

    def _capture(self):
        model, warmup = self.model, self.warmup
        if model.training:
            raise ValueError("GraphedForward replays the eval-mode forward; the model is in training mode")
        side = torch.cuda.Stream(device=self.static_in[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad(), _pu.fps_start_source(lambda B, N: torch.zeros(B, dtype=torch.long)):
            for _ in range(warmup):            # packs the weights, sizes every allocation; draws nothing from the CPU generator
                model(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.starts = _PinnedStarts(self.start_source)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), _pu.fps_start_source(self.starts), torch.cuda.graph(self.graph):
            self.static_out = model(*self.static_in)
        self.starts.capturing = False
        self.fingerprint = _weights_fingerprint(model)
        self.keep = _derived_buffers(model)
        self.captures += 1
        self.done = torch.cuda.Event()
        self.done.record()

    def __call__(self, *inputs: torch.Tensor):
        if _weights_fingerprint(self.model) != self.fingerprint:
            self.done.synchronize()
            source = self.starts.source
            self._capture()                    # weights / precision changed since capture: never replay stale packed weights
            self.starts.source = source
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.done.synchronize()                # the previous replay's upload nodes have read the pinned buffers
        self.starts.refill()
        self.graph.replay()
        self.done.record()
        return self.static_out
