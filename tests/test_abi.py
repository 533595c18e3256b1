"""CPU suite: the C-ABI shared library loads without a GPU and exports exactly what include/pn2.h declares;
the ctypes binding covers every declared function; a missing library fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

import _pn2_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pn2.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("pn2_fps_f32", "pn2_ball_query_f32", "pn2_index_points_f32", "pn2_square_distance_f32", "pn2_group_f32",
                 "pn2_pack_weights", "pn2_sa_fused", "pn2_last_error_string", "pn2_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_pn2_native.LIB_PATH), "build with `python __graft_entry__.py build`"
    lib = ctypes.CDLL(_pn2_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in pn2.h but not exported"


def test_binding_covers_the_header_and_versions_agree():
    assert sorted(_pn2_native.EXPORTED_SYMBOLS) == declared_functions()
    lib = _pn2_native.load_library()
    version = int(re.search(r"#define PN2_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.pn2_version() == version == _pn2_native.ABI_VERSION


def test_argument_errors_are_reported_not_launched():
    """Host-side validation needs no GPU: bad arguments return < 0 and set the thread-local message."""
    lib = _pn2_native.load_library()
    rc = lib.pn2_fps_f32(None, 0, 0, 0, 1, 8, 3, None, 2, None, None, None, None)
    assert rc < 0 and b"null pointer" in lib.pn2_last_error_string()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.pn2_ball_query_f32(p, 24, 3, 1, p, 24, 3, 1, 1, 2, 8, 3, 0.04, 9, p, None, None, None)
    assert rc < 0 and b"must be in [1, N=8]" in lib.pn2_last_error_string()     # K > N: the reference raises at :89
    rc = lib.pn2_fps_f32(p, 64, 32, 1, 1, 2, 32, p, 1, p, None, None, None)
    assert rc < 0 and b"C=32" in lib.pn2_last_error_string()


def test_missing_library_is_an_import_error(tmp_path, monkeypatch):
    monkeypatch.setattr(_pn2_native, "_lib", None)
    with pytest.raises(ImportError, match="no fallback"):
        _pn2_native.load_library(str(tmp_path / "libpn2_sm100.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pointnet-like-pose-estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pn2_oracle" not in text and "pn2_torch_port" not in text, f


def test_tensor_core_kernels_do_not_spill():
    """A spilled prefetch register in the fused set-abstraction kernels blocks `fence.proxy.async` on the spill store (DESIGN.md,
    prefetch discipline: a ~1000-cycle stall per fence).  Guard it at build time: no stack frame, no local memory."""
    import re
    import shutil
    import subprocess
    import _pn2_native
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", _pn2_native.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    seen = 0
    for fn, usage in re.findall(r"Function (\S+):\n\s*(REG:.*)", out):
        if re.search(r"sa_tc\d?_kernel", fn):
            seen += 1
            stack, local = int(re.search(r"STACK:(\d+)", usage).group(1)), int(re.search(r"LOCAL:(\d+)", usage).group(1))
            assert stack == 0 and local == 0, (fn, usage)
    assert seen >= 4, "the tensor-core kernels were not found in the library"
