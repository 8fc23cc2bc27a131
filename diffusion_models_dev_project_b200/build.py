"""Build libscd_b200.so in-tree with nvcc for sm_100a (no torch cpp_extension: the ABI is plain C).

    python -m diffusion_models_dev_project_b200.build [--force] [--verbose] [--debug]

``--debug`` builds a second library, ``_lib/libscd_b200_dbg.so``, with ``-DSCD_DEBUG_STAMPS``: the per-CTA
time stamps of tools/timeline.py and ``scd_debug_set_stamps`` (include/scd_b200_debug.h).  The default
library contains neither.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libscd_b200.so")
DEBUG_LIB_PATH = os.path.join(LIB_DIR, "libscd_b200_dbg.so")
SOURCES = ["geometry.cu", "fp_march.cu", "bp_tile.cu", "vec_ops.cu", "il_ops.cu", "loss_ops.cu", "adapt_ops.cu", "fbp_filter.cu", "cg_solver.cu", "peer_reduce.cu"]
HEADERS = [os.path.join(CSRC, "scd_internal.cuh"), os.path.join(ROOT, "include", "scd_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libscd_b200.so cannot be built (there is no CPU fallback)")
    return cand


def needs_build(path=LIB_PATH):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=False):
    out = DEBUG_LIB_PATH if debug else LIB_PATH
    if not force and not needs_build(out):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", *ARCH,
           "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if debug:
        cmd += ["-DSCD_DEBUG_STAMPS"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug="--debug" in sys.argv)
    print(p)
