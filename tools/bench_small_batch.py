#!/usr/bin/env python
"""Latency of the literal small batches of BASELINE.json configs[0] / [1] (pointnet2_cls_ssg B = 24, pointnet2_cls_msg B = 32)
and of rotation_msg at B = 32: eager forward against the CUDA-graph replay of pn2_graph.GraphedForward (bf16 mode, one GPU)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('pointnet-like-pose-estimation_b200', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu
from pn2_graph import GraphedForward
pu.set_precision("bf16")
res={}
for name,B in [("pointnet2_cls_msg",32),("pointnet2_cls_ssg",24),("rotation_msg",32)]:
    torch.manual_seed(0)
    m=pn2_models.get_model(name,7); pn2_models.randomize_bn_(m); m=m.eval().cuda()
    xyz=cloudgen.uniform_cloud(B,1024,1002)
    x=cloudgen.channels_first(cloudgen.with_category(xyz) if name.startswith("rotation") else xyz).cuda()
    def t(fn,n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
    with torch.no_grad():
        eager=t(lambda: m(x))
    g=GraphedForward(m,x)
    graphed=t(lambda: g(x))
    res[f"{name}/B{B}"]={"eager_ms":round(eager,4),"graph_ms":round(graphed,4),"eager_clouds_per_s":round(B/eager*1e3),"graph_clouds_per_s":round(B/graphed*1e3)}
    print(name,B,res[f"{name}/B{B}"],flush=True)
print(json.dumps(res))
