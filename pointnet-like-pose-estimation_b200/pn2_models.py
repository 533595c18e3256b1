"""The reference's PointNet++ callers, table-driven, for machines where the reference checkout is absent.

The reference's own model files (model/pointnet2_cls_{ssg,msg}.py, model/{rotation,translation,sign}_{ssg,msg}.py)
run unchanged on top of this directory's ``pointnet2_utils``.  The GPU box that runs the benchmark and the
``-m gpu`` tests has no reference checkout, so the same eight architectures are restated here from one spec
table: same attribute names (``sa1..sa3, fc1..fc3, bn1, bn2, drop, mean_fc1, mean_fc2, mean_bn1``), same
construction order (so a seeded construction draws identical initial weights) and therefore interchangeable
``state_dict``s -- tests/test_reference_live.py checks exactly that against the live reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

import pointnet2_utils as _default_sa

# (kind, args) per set-abstraction level; `c` = coordinate channels (3, or 3 + num_category for pose heads)
_SSG_SA = lambda c: [  # pointnet2_cls_ssg.py:9-11, rotation_ssg.py:11-13
    ("ssg", dict(point_number=512, sample_number=32, radius=0.2, in_channel=c, mlp=[64, 64, 128], group_all=False)),
    ("ssg", dict(point_number=128, sample_number=64, radius=0.4, in_channel=128 + c, mlp=[128, 128, 256], group_all=False)),
    ("ssg", dict(point_number=None, sample_number=None, radius=None, in_channel=256 + c, mlp=[256, 512, 1024], group_all=True)),
]
_MSG_SA = lambda c, k: [  # pointnet2_cls_msg.py:9-11, rotation_msg.py:11-13
    ("msg", dict(point_number=512, sample_number_list=[16, 32, 128], radius_list=[0.1, 0.2, 0.4], in_channel=0,
                 mlp_list=[[32, 32, 64], [64, 64, 128], [64, 96, 128]], num_category=k)),
    ("msg", dict(point_number=128, sample_number_list=[32, 64, 128], radius_list=[0.2, 0.4, 0.8], in_channel=320,
                 mlp_list=[[64, 64, 128], [128, 128, 256], [128, 128, 256]], num_category=k)),
    ("ssg", dict(point_number=None, sample_number=None, radius=None, in_channel=640 + c, mlp=[256, 512, 1024], group_all=True)),
]
_SSG2_SA = lambda c: [  # translation_ssg.py:11-12, sign_ssg.py:10-11
    ("ssg", dict(point_number=512, sample_number=32, radius=0.2, in_channel=c, mlp=[64, 64, 128], group_all=False)),
    ("ssg", dict(point_number=None, sample_number=None, radius=None, in_channel=128 + c, mlp=[256, 512, 1024], group_all=True)),
]
_MSG2_SA = lambda c, k: [  # translation_msg.py:11-12, sign_msg.py:10-11
    ("msg", dict(point_number=512, sample_number_list=[16, 32, 64], radius_list=[0.1, 0.2, 0.4], in_channel=0,
                 mlp_list=[[32, 64, 128], [64, 128, 256], [96, 128, 256]], num_category=k)),
    ("ssg", dict(point_number=None, sample_number=None, radius=None, in_channel=640 + c, mlp=[256, 512, 1024], group_all=True)),
]

# name -> (SA spec builder, pose head?, output width (None = num_category), head kind)
ARCH = {
    "pointnet2_cls_ssg": ("ssg3", False, None, "cls"),
    "pointnet2_cls_msg": ("msg3", False, None, "cls"),
    "rotation_ssg": ("ssg3", True, 3, "reg"),
    "rotation_msg": ("msg3", True, 3, "reg"),
    "translation_ssg": ("ssg2", True, 3, "trans"),
    "translation_msg": ("msg2", True, 3, "trans"),
    "sign_ssg": ("ssg2", True, 1, "sign"),
    "sign_msg": ("msg2", True, 1, "sign"),
}


class PointNet2Head(nn.Module):
    """One of the eight PointNet++ models of the reference, selected by name (see ARCH)."""

    def __init__(self, name: str, num_category: int = 7, mean_mlp: str = "True", sa_impl=None):
        """sa_impl: module providing PointNetSetAbstraction / PointNetSetAbstractionMsg (default: this package's
        CUDA-backed pointnet2_utils; the CPU baseline legs of bench.py and the tests pass other providers)."""
        super().__init__()
        sa_impl = sa_impl or _default_sa
        self._sa_impl = sa_impl
        self._head_cache = None
        backbone, pose, out_width, self.kind = ARCH[name]
        self.arch = name
        c = 3 + num_category if pose else 3
        k = num_category if pose else 0
        spec = {"ssg3": lambda: _SSG_SA(c), "msg3": lambda: _MSG_SA(c, k),
                "ssg2": lambda: _SSG2_SA(c), "msg2": lambda: _MSG2_SA(c, k)}[backbone]()
        for i, (kind, kw) in enumerate(spec, 1):
            cls = sa_impl.PointNetSetAbstraction if kind == "ssg" else sa_impl.PointNetSetAbstractionMsg
            setattr(self, f"sa{i}", cls(**kw))
        self.n_sa = len(spec)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, num_category if out_width is None else out_width)
        self.drop = nn.Dropout(0.4)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.mean_mlp = mean_mlp
        if self.kind == "trans" and mean_mlp == "True":      # translation_ssg.py:23-26
            self.mean_fc1 = nn.Linear(3, 6)
            self.mean_fc2 = nn.Linear(6, 3)
            self.mean_bn1 = nn.BatchNorm1d(6)

    def forward(self, points, mean=None):
        B = points.shape[0]
        fused = (not self.training and hasattr(self._sa_impl, "fc_head")
                 and self._sa_impl._FoldedHead.supported((self.fc1, self.fc2, self.fc3)))
        if self.kind == "trans" and self.mean_mlp == "True" and not fused:  # translation_ssg.py:31-33 (fused: inside the head kernel)
            mean = self.mean_fc2(F.relu(self.mean_bn1(self.mean_fc1(mean))))
        xyz, feat = points, None
        for i in range(1, self.n_sa + 1):
            xyz, feat = getattr(self, f"sa{i}")(xyz, feat)
        global_feature = feat
        x = global_feature.view(B, 1024)
        if fused:
            return self._head_fused(x, global_feature, mean)
        x = self.drop(F.relu(self.bn1(self.fc1(x))))
        x = self.drop(F.relu(self.bn2(self.fc2(x))))
        x = self.fc3(x)
        if self.kind == "cls":                                 # pointnet2_cls_ssg.py:36-38
            x = F.log_softmax(x, -1)
            return x, global_feature, x.data.max(1)[1]
        if self.kind == "trans":                               # translation_ssg.py:44
            return x + mean
        if self.kind == "sign":                                # sign_ssg.py:34-36
            x = torch.sigmoid(x)
            return x, torch.sign(x - 0.5)
        return x                                               # rotation_ssg.py:38


    def _head_fused(self, x, global_feature, mean):
        """Eval mode: fc1/bn1/relu -> fc2/bn2/relu -> fc3 -> epilogue (and the translation heads' mean MLP) in one launch
        (pn2_fc_head_f32) instead of ~14 library launches; dropout is the identity in eval mode."""
        from _pn2_native import HEAD_CLS, HEAD_PLAIN, HEAD_SIGN, HEAD_TRANS
        if self._head_cache is None:
            self._head_cache = self._sa_impl._FoldedHead()
        kind = {"cls": HEAD_CLS, "trans": HEAD_TRANS, "sign": HEAD_SIGN, "reg": HEAD_PLAIN}[self.kind]
        mm = (self.mean_fc1, self.mean_bn1, self.mean_fc2) if (self.kind == "trans" and self.mean_mlp == "True") else None
        out, aux = self._sa_impl.fc_head(x, (self.fc1, self.fc2, self.fc3), (self.bn1, self.bn2), kind, self._head_cache,
                                         mean=mean if self.kind == "trans" else None, mean_mlp=mm)
        if self.kind == "cls":
            return out, global_feature, aux
        if self.kind == "sign":
            return out, aux
        return out


def get_model(name: str, num_category: int = 7, **kw) -> PointNet2Head:
    return PointNet2Head(name, num_category, **kw)


def randomize_bn_(model: nn.Module, seed: int = 7) -> None:
    """Give every BatchNorm non-trivial eval statistics / affine parameters (deterministic, CPU generator) so
    that BN folding is exercised: running_mean~N(0,0.5), running_var~U(0.5,2), gamma~U(0.5,1.5), beta~N(0,0.2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                n = m.num_features
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.5)
                m.running_var.copy_(torch.rand(n, generator=g) * 1.5 + 0.5)
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.2)
