"""Debug helper (not collected by pytest): one fused set-abstraction launch against the oracle, printing where they differ.
Lives under tests/ because only tests/, smoke() and bench.py's CPU legs may use oracle/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("pointnet-like-pose-estimation_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, cloudgen, pn2_oracle as oracle, _pn2_native
from test_gpu_kernels import make_cloud, rand_layers
nat = _pn2_native.Native()

def run(B, N, C, D, S, K, r, order, mlp):
    x = make_cloud(B, N, C, "uni", 3000 + N)
    f = torch.randn(B, N, D, generator=torch.Generator().manual_seed(5)) if D else None
    dims = [C + D] + mlp
    layers = rand_layers(dims, 11)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(8), dtype=torch.long)
    cen = oracle.index_points(x, oracle.farthest_point_sample(x, S, start)).contiguous()
    idx = oracle.query_ball_point(r, K, x, cen)
    rows = oracle.group(x, f, idx, cen, order)
    exp = oracle.mlp_max(rows.reshape(B * S, K, C + D), layers).reshape(B, S, -1)
    d = lambda t: t.cuda()
    packed, _ = nat.pack([tuple(d(t) for t in l) for l in layers], 1e-5, 1, C, order)
    for twin in ((False, True) if D else (False,)):
        out = torch.zeros(B, S, mlp[-1], device="cuda")
        nat.sa_fused(d(x), None if f is None else d(f), d(cen), d(idx.int()), order, dims, packed, 1, out, 0,
                     feature_bf16=d(f).bfloat16().contiguous() if twin else None)
        torch.cuda.synchronize()
        e = (out.cpu() - exp).abs() / exp.abs().max()
        bad = (e.max(-1)[0][0] > 0.02).nonzero().flatten()
        print(f"C={C} D={D} K={K} mlp={mlp} twin={twin}: max err {float(e.max()):.3f} bad groups {len(bad)}/{S}: {bad[:12].tolist()}")
run(1, 512, 3, 0, 128, 64, 0.4, 0, [128])
run(1, 512, 3, 0, 128, 32, 0.4, 0, [128])
run(1, 512, 3, 128, 128, 32, 0.4, 0, [128])
run(1, 512, 3, 128, 128, 64, 0.4, 0, [64, 128])
run(1, 512, 3, 128, 128, 64, 0.4, 0, [128, 128, 256])
run(1, 512, 3, 64, 128, 64, 0.4, 0, [128, 128, 256])
run(1, 512, 10, 128, 128, 64, 0.4, 0, [128, 128, 256])
