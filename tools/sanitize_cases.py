#!/usr/bin/env python
"""Small invocations of every product kernel family, written for `compute-sanitizer --tool memcheck python tools/sanitize_cases.py`.
(The sanitizer is closed on the GPU pool of round 1 -- the wrapper refuses to start it -- so the round only has the plain run:
all six cases pass with the caching allocator off, gpurun_out/sanitize_plain.log.)

Run with PYTORCH_NO_CUDA_MEMORY_CACHING=1 so that every tensor is its own cudaMalloc and an out-of-bounds access of a
kernel cannot land inside the caching allocator's pool unnoticed.  Sizes are small (the sanitizer slows kernels 10-100x) but
each case goes through the same kernel variants as the BASELINE configurations (same N, S, K, channel widths; small B).

    python tools/sanitize_cases.py [case ...]      # default: all cases
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pointnet-like-pose-estimation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import cloudgen  # noqa: E402
import pn2_models  # noqa: E402
import pn2_preprocess  # noqa: E402
import pn2_train  # noqa: E402
import pn2_v1  # noqa: E402
import pointnet2_utils as pu  # noqa: E402

DEV = "cuda:0"


def _cloud(B, N, seed, pose=False):
    xyz = cloudgen.uniform_cloud(B, N, seed)
    return cloudgen.channels_first(cloudgen.with_category(xyz) if pose else xyz).to(DEV)


def _model(name):
    torch.manual_seed(0)
    m = pn2_models.get_model(name, 7)
    pn2_models.randomize_bn_(m)
    return m.to(DEV).eval()


def case_functions():
    """square_distance / index_points / FPS / ball query / grouping as free functions, ragged sizes included."""
    for B, N, S, K, C in [(3, 1024, 512, 32, 3), (2, 515, 100, 16, 3), (2, 512, 128, 64, 10), (1, 8192, 2048, 32, 3), (2, 33, 7, 8, 3)]:
        pts = torch.rand(B, N, C, device=DEV)
        idx = pu.farthest_point_sample(pts, S)
        ctr = pu.index_points(pts, idx)
        pu.query_ball_point(0.4, K, pts, ctr)
        pu.square_distance(ctr, pts)
        feat = torch.rand(B, N, 16, device=DEV)
        pu.sample_and_group(pts, feat, S, K, 0.4)
        pu.sample_and_group_all(pts, feat)
    # strided [B,C,N].permute operand (the layout sa1 receives)
    x = torch.rand(2, 3, 1024, device=DEV).permute(0, 2, 1)
    i = pu.farthest_point_sample(x, 512)
    pu.query_ball_point(0.2, 32, x, pu.index_points(x, i))


def case_models_bf16():
    pu.set_precision("bf16")
    with torch.no_grad():
        _model("pointnet2_cls_ssg")(_cloud(3, 1024, 5))
        _model("pointnet2_cls_msg")(_cloud(2, 1024, 6))
        _model("rotation_ssg")(_cloud(2, 1024, 7, pose=True))
        _model("rotation_msg")(_cloud(2, 1024, 8, pose=True))
        x = _cloud(2, 1024, 9, pose=True)
        _model("translation_msg")(x, x[:, :3, :].mean(2))
        _model("sign_ssg")(x)


def case_models_fp32():
    pu.set_precision("fp32")
    with torch.no_grad():
        _model("pointnet2_cls_ssg")(_cloud(2, 1024, 5))
        _model("pointnet2_cls_msg")(_cloud(1, 1024, 6))
        x = _cloud(2, 1024, 9, pose=True)
        _model("translation_ssg")(x, x[:, :3, :].mean(2))
    pu.set_precision("bf16")


def case_v1():
    torch.manual_seed(0)
    m = pn2_v1.get_v1_model("rotation", 7).to(DEV).eval()
    with torch.no_grad():
        for mode in ("bf16", "fp32"):
            pu.set_precision(mode)
            m(_cloud(3, 1024, 11, pose=True))
    pu.set_precision("bf16")


def case_train():
    torch.manual_seed(0)
    model = pn2_models.get_model("rotation_msg", 7).to(DEV)
    bucket = pn2_train.GradBucket(model.parameters())
    opt = pn2_train.make_optimizer(model, learning_rate=1e-3)
    crit = pn2_train.get_loss("L2_loss", "sum")
    x = _cloud(2, 1024, 21, pose=True)
    tgt = torch.randn(2, 3, device=DEV)
    for _ in range(2):
        pn2_train.train_step(model, crit, opt, bucket, x, tgt)


def case_preprocess():
    g = torch.Generator().manual_seed(1)
    raw = torch.rand(3, 1500, 3, generator=g).to(DEV)
    pn2_preprocess.prepare_pose_input(raw, torch.tensor([0, 3, 6], device=DEV), 7)
    import numpy as np
    pc = np.random.default_rng(0).random((3000, 6))
    np.random.seed(0)
    pn2_preprocess.farthest_point_sample(pc, 1024)
    pn2_preprocess.farthest_point_sample(pc.astype(np.float32), 512)


CASES = {"functions": case_functions, "models_bf16": case_models_bf16, "models_fp32": case_models_fp32, "v1": case_v1,
         "train": case_train, "preprocess": case_preprocess}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    be = pu._backend()
    for n in names:
        t0, l0 = time.time(), be.launches
        CASES[n]()
        torch.cuda.synchronize()
        print(f"case {n}: ok, {be.launches - l0} product launches, {time.time() - t0:.1f} s", flush=True)
    print("all cases ok")
