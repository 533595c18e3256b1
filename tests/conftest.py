"""pytest configuration: the `gpu` marker, import paths, and the handles shared by the suites.

  * `-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic (through an oracle-backed fake of
    the native layer that exists ONLY in tests/), C-ABI symbol checks, gloo sharding tests.
  * `-m gpu` runs on a B200: the CUDA path through the C ABI vs the oracle and the golden vectors.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pointnet-like-pose-estimation_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__)), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFERENCE_MODEL_DIR = "/root/reference/model"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir(REFERENCE_MODEL_DIR)
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference checkout not present"))


@pytest.fixture(scope="session")
def oracle():
    import pn2_oracle
    pn2_oracle.build()
    return pn2_oracle


@pytest.fixture()
def fake_native(monkeypatch, oracle):
    """pointnet2_utils with its native layer replaced by the CPU oracle -- host-logic tests only."""
    import pointnet2_utils
    from fake_native import OracleNative
    monkeypatch.setattr(pointnet2_utils, "_native", OracleNative())
    prev = pointnet2_utils.get_precision()
    pointnet2_utils.set_precision("fp32")
    yield pointnet2_utils
    pointnet2_utils.set_precision(prev)
    pointnet2_utils.set_reference_strides(False)
