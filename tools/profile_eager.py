#!/usr/bin/env python
"""Host-side profile of the small-batch eager forward (B = 24, configs[0] batch): where the enqueueing thread's time goes."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_models, pointnet2_utils as pu
pu.set_precision("bf16")
torch.manual_seed(0)
m = pn2_models.get_model("pointnet2_cls_ssg", 7); pn2_models.randomize_bn_(m); m = m.eval().cuda()
x = cloudgen.channels_first(cloudgen.uniform_cloud(24, 1024, 1)).cuda()
with torch.no_grad():
    for _ in range(20): m(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): m(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e6 * (t1 - t0) / 300:.1f} us per forward, with drain {1e6 * (t2 - t0) / 300:.1f} us; torch threads {torch.get_num_threads()}")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(300): m(x)
    pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
