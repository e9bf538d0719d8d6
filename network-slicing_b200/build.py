"""In-tree build of libranslice_b200.so (nvcc, sm_100a).  Called by __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libranslice_b200.so")
SOURCES = ["ranslice_cabi.cu", "embb_step.cu", "embb_fast.cu", "embb_smem.cu", "embb_warp.cu", "mmtc_step.cu", "kbrl.cu", "wrappers.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false",           # fp64 decision arithmetic must not be contracted (parity with NumPy)
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    extra = os.environ.get("RS_NVCC_EXTRA", "").split()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "ranslice_b200.h"))
    headers.append(os.path.join(os.path.dirname(PKG), "include", "kbrl_b200.h"))
    headers.append(os.path.join(os.path.dirname(PKG), "include", "wrapper_b200.h"))
    tag = os.environ.get("RS_BUILD_TAG", "")                # experiment builds: separate objects, lib<tag>.so next to the real one
    out = OUT if not tag else OUT.replace(".so", "_%s.so" % tag)
    objdir = os.path.join(PKG, "build" + ("_" + tag if tag else ""))
    os.makedirs(objdir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in srcs]
    logs = {}

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, [os.path.join(CSRC, src)] + headers):
            return
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))

    with ThreadPoolExecutor(4) as ex:
        list(ex.map(compile_one, zip(srcs, objs)))
    if force or _stale(out, objs):
        cmd = [_nvcc(), "-shared", "-o", out] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for s, l in logs.items():
            sys.stderr.write("== %s\n%s\n" % (s, l))
    with open(os.path.join(objdir, "ptxas.log"), "a") as f:
        for s, l in logs.items():
            f.write("== %s\n%s\n" % (s, l))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
