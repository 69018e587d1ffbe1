"""The C-ABI library builds, loads and exports every symbol include/nwk.h declares (no GPU)."""
import ctypes
import os
import re

from neighborhoodwatch_b200 import _build, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "nwk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nwk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    declared = header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nwk.h but not exported by libnwk.so"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert lib.nwk_abi_version() == 1


def test_library_is_in_tree_and_native_only():
    path = _native.lib_path()
    assert os.path.dirname(path) == os.path.join(ROOT, "neighborhoodwatch_b200")
    cu = [f for f in os.listdir(_build.CSRC) if f.endswith(".cu")]
    assert sorted(cu) == sorted(_build.SOURCES)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = _native.load()
    buf = (ctypes.c_float * 64)()
    rc = lib.nwk_synth_fill(ctypes.addressof(buf), 4, 8, 1, 0, 1, None)
    assert rc == 4, "compute entry points must fail with NWK_ERR_DEVICE when there is no B200"
    assert b"no CPU fallback" in lib.nwk_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neighborhoodwatch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_workspace_query_is_pure():
    lib = _native.load()
    a = lib.nwk_knn_workspace_bytes(10 ** 6, 10 ** 4, 768, 100, 1, 0)
    b = lib.nwk_knn_workspace_bytes(10 ** 6, 10 ** 4, 768, 100, 1, 0)
    assert a == b and a > 10 ** 4 * 4096 * 8
    assert lib.nwk_merge_workspace_bytes(8, 10 ** 4, 100, 100) >= 256
