"""Build libpmc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.environ.get("PMC_B200_LIBPATH") or os.path.join(LIBDIR, "libpmc_b200.so")     # override: diagnostic builds (tests/nsf_truth.py)
SOURCES = ["lib.cu", "flow_sweep.cu", "mcmc_ops.cu", "smc_ops.cu", "train_ops.cu", "flow_tc.cu", "flow_train.cu", "flow_tri.cu", "geom_ops.cu", "flow_train_lw.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "shared"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libpmc_b200.so")
    return exe


def needs_build():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pmc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), tag=""):
    """Build the library (in-tree path, or LIBPATH's override).  `defines` / `tag`: a diagnostic variant with extra -D flags,
    objects under build/obj<tag>/ (e.g. -DPMC_TRI_RQS_REFERENCE_HEAD: the spline head in the reference's operation order)."""
    if not force and not needs_build():
        return LIBPATH
    os.makedirs(os.path.dirname(LIBPATH), exist_ok=True)
    objdir = os.path.join(HERE, "..", "build", "obj" + tag)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", LIBPATH, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
