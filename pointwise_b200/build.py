"""Builds libconv3p_b200.so in-tree with nvcc for sm_100a (no torch involvement, plain C ABI)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libconv3p_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "conv3p_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compiles every .cu under csrc/ for sm_100a and links libconv3p_b200.so.  `defines` / `out` build an experiment
    variant (compile-time switches such as -DC3P_W2_LOOK2=1) into another file, loaded with CONV3P_LIB=<path>
    (tools/build_variants.py): A/B timing of whole libraries in one GPU call."""
    lib_path = out or LIB_PATH
    if not force and not out and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = LIB_DIR if not out else os.path.join(LIB_DIR, "obj_" + os.path.basename(out))
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *ARCH_FLAGS, "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", *[f"-D{d}" for d in defines],
               "-Xptxas", "-v" if verbose else "-warn-spills", "-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(" ".join(cmd))
            print(log)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [NVCC, *ARCH_FLAGS, "-shared", "-o", lib_path, *objs]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    if out:      # a variant's objects are not reused: keep the tree (and what travels to the GPU box) small
        import shutil
        shutil.rmtree(obj_dir, ignore_errors=True)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
