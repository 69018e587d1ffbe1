"""Build libnwk.so (hand-written sm_100a CUDA behind the C ABI of include/nwk.h) in-tree.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting ``neighborhoodwatch_b200/libnwk.so`` travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libnwk.so")
SOURCES = ["nwk_api.cu", "nwk_scan.cu", "nwk_tc.cu", "nwk_synth.cu", "nwk_layout.cu", "nwk_xchg.cu"]
HEADERS = ["nwk_common.cuh", "nwk_ptx.cuh", "nwk_maxsim.cuh", os.path.join("..", "..", "include", "nwk.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libnwk cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(p) > lib_m for p in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out_path: str = None) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path.

    ``defines`` / ``out_path`` build an experiment variant (extra -D macros) next to the
    product library; the product build uses neither.
    """
    global_out = out_path or LIB_PATH
    if not force and not defines and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, "build" if not defines else "build_" + os.path.basename(global_out))
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        log.append(f"== {src} ==\n{out}")
        failed |= proc.returncode != 0
    with open(os.path.join(obj_dir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see output above")
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", global_out + ".tmp", *objs,
            "-cudart", "static"]
    subprocess.run(link, check=True)
    os.replace(global_out + ".tmp", global_out)
    return global_out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
