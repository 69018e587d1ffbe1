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
    assert lib.nwk_abi_version() == _native.ABI_VERSION == 2


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


def _slab_plan(lib, nb, nq, d, k, slab_rows):
    from ctypes import byref, c_int64
    sr, ns, last = c_int64(), c_int64(), c_int64()
    rc = lib.nwk_host_slab_plan(nb, nq, d, k, slab_rows, byref(sr), byref(ns), byref(last))
    return rc, sr.value, ns.value, last.value


def test_host_slab_plan_every_slab_holds_k_rows():
    """nb=206,k=51,slab_rows=15 and nb=1005,k=10,slab_rows=10 used to give a short (or
    negative) last slab.  Fuzz (nb, k, slab_rows): the slabs tile [0, nb) exactly, each holds >= k rows,
    and the device ring (the longest slab) stays below slab_rows + k."""
    import random
    lib = _native.load()
    cases = [(206, 51, 15), (1005, 10, 10), (1000, 100, 100), (100, 100, 7), (101, 100, 100), (10 ** 6, 100, 0)]
    rnd = random.Random(7)
    for _ in range(20000):
        nb = rnd.randint(1, 5000)
        k = rnd.randint(1, min(nb, 300))
        cases.append((nb, k, rnd.choice([0, rnd.randint(1, 2 * nb), rnd.randint(1, 64)])))
    for nb, k, slab_rows in cases:
        rc, sr, ns, last = _slab_plan(lib, nb, 7, 8, k, slab_rows)
        assert rc == 0, (nb, k, slab_rows)
        assert ns >= 1 and sr >= k and last >= k, (nb, k, slab_rows, sr, ns, last)
        assert (ns - 1) * sr + last == nb, (nb, k, slab_rows, sr, ns, last)
        assert last < sr + k, (nb, k, slab_rows, sr, ns, last)
        if slab_rows >= k:
            assert sr == min(slab_rows, nb)
    # bad arguments are refused, not planned
    assert _slab_plan(lib, 10, 1, 8, 11, 0)[0] == 1
    assert _slab_plan(lib, 0, 1, 8, 1, 0)[0] == 1
