"""The full pose pipeline of BASELINE.json configs[3]: the six PointNet++ pose heads (rotation / translation / sign, SSG and
MSG: model/{rotation,translation,sign}_{ssg,msg}.py) plus the PointNet-v1 width head (model/width.py) on one batch of clouds,
the way the reference's deployment script runs its heads side by side on one GPU (mutilthreading/predict_test.py:44-62: one
Python thread per head on the same CUDA tensor).

One shard of the batch per GPU, no collective.  What limits a small shard (128 clouds per GPU at B = 1024 over 8 GPUs) is not
arithmetic but the serial chains: every head starts with a 512-iteration FPS that occupies one CTA per cloud for ~0.3 ms however
few clouds there are, and seven heads run ~50 launches each.  ``PosePipeline`` therefore

  * prepares the input once (provider.normalization + splice_torch + transpose in one launch, pn2_preprocess),
  * replays every head as ONE CUDA graph (pn2_graph.GraphedForward: launch latency removed, FPS start draws still made on the
    CPU generator per layer per forward, in head order, and uploaded by the graph's memcpy nodes), and
  * issues the seven graphs on seven streams, so the FPS chains and ball queries of different heads overlap instead of queueing
    behind each other's persistent tensor-core kernels.

``mode='eager'`` runs the heads one after the other on the current stream (the comparator; bit-identical outputs).
"""
from __future__ import annotations

import torch

import pn2_models
import pn2_preprocess
import pn2_v1
import pointnet2_utils as _pu
from pn2_graph import GraphedForward

__all__ = ["HEADS", "PosePipeline", "build_heads"]

HEADS = ["rotation_ssg", "rotation_msg", "translation_ssg", "translation_msg", "sign_ssg", "sign_msg", "v1/width"]


def build_heads(device, num_category: int = 7, seed: int = 0, names=HEADS):
    """The seven heads with seeded random weights and non-trivial BatchNorm statistics (benchmarks / tests)."""
    heads = {}
    for name in names:
        torch.manual_seed(seed)
        m = pn2_v1.get_v1_model(name[3:], num_category) if name.startswith("v1/") else pn2_models.get_model(name, num_category)
        pn2_models.randomize_bn_(m)
        heads[name] = m.eval().to(device)
    return heads


class PosePipeline:
    """``pipe = PosePipeline(heads); outs = pipe(xyz, normals, category)`` -> {head name: tuple of output tensors}.

    xyz [B,N,3] raw coordinates, normals [B,N,3] (width.py takes xyz + normals + one-hot), category [B] int64; all CUDA.
    ``mode``: 'eager' | 'graphs' (one CUDA graph per head, replayed in order on the current stream) | 'streams' (the graphs on
    one stream per head).  ``start_source``: optional fps_start_source callable shared by all heads (e.g. a
    pn2_shard.GlobalStarts); heads draw in HEADS order, layers in module order, exactly like the eager sequence."""

    def __init__(self, heads: dict, num_category: int = 7, mode: str = "streams", start_source=None):
        if mode not in ("eager", "graphs", "streams"):
            raise ValueError("mode must be 'eager', 'graphs' or 'streams'")
        self.heads, self.num_category, self.mode, self.start_source = heads, num_category, mode, start_source
        self.names = [n for n in HEADS if n in heads]
        self.graphs = {}
        self.streams = {}
        self.joined = {}

    def prepare(self, xyz, normals, category):
        """-> (x [B,3+k,N], mean [B,3], xw [B,6+k,N]): the model inputs (train_rotation.py:108-116, train_translation.py:113)."""
        x = pn2_preprocess.prepare_pose_input(xyz, category, self.num_category)
        mean = xyz[:, :3, :].mean(1)                      # the reference really averages the first three POINTS
        xw = torch.cat([x[:, :3], normals.permute(0, 2, 1), x[:, 3:]], 1) if "v1/width" in self.heads else None
        return x, mean, xw

    def _args(self, name, x, mean, xw):
        if name == "v1/width":
            return (xw,)
        return (x, mean) if name.startswith("translation") else (x,)

    def forward_prepared(self, x, mean, xw=None):
        outs = {}
        if self.mode == "eager":
            ctx = _pu.fps_start_source(self.start_source) if self.start_source is not None else _null()
            with torch.no_grad(), ctx:
                for n in self.names:
                    o = self.heads[n](*self._args(n, x, mean, xw))
                    outs[n] = o if isinstance(o, tuple) else (o,)
            return outs
        main = torch.cuda.current_stream()
        for n in self.names:
            a = self._args(n, x, mean, xw)
            if n not in self.graphs:
                self.graphs[n] = GraphedForward(self.heads[n], *a, start_source=self.start_source)
                self.streams[n] = torch.cuda.Stream(device=x.device)
                self.joined[n] = torch.cuda.Event()
        if self.mode == "graphs":
            for n in self.names:
                o = self.graphs[n](*self._args(n, x, mean, xw))
                outs[n] = o if isinstance(o, tuple) else (o,)
            return outs
        ready = torch.cuda.Event()
        ready.record(main)
        for n in self.names:                               # draws are made here, in head order, on the host
            s = self.streams[n]
            s.wait_event(ready)
            with torch.cuda.stream(s):
                o = self.graphs[n](*self._args(n, x, mean, xw))
                self.joined[n].record(s)
            outs[n] = o if isinstance(o, tuple) else (o,)
        for n in self.names:
            main.wait_event(self.joined[n])
        return outs

    def __call__(self, xyz, normals, category):
        x, mean, xw = self.prepare(xyz, normals, category)
        return self.forward_prepared(x, mean, xw)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False
