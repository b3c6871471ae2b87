"""Builds libdcl_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m doubly_contrastive_semseg_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdcl_b200.so")
SOURCES = ["dcl_api.cu", "dcl_contrast.cu", "dcl_contrast_small.cu", "dcl_sampler.cu", "dcl_plan.cu", "dcl_focal.cu", "dcl_p2p.cu", "dcl_step.cu", "dcl_host_rng.cpp", "dcl_comm.cpp", "dcl_clocks.cpp"]
HEADERS = ["dcl_ptx.cuh", "dcl_common.cuh", "dcl_plan.h", os.path.join("..", "..", "include", "dcl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3", "--use_fast_math", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the library if it is missing or older than its sources. Returns the .so path.
    `extra_flags`/`out` build an experimental variant next to the default library (tools only)."""
    if out is not None:
        return _build_to(out, list(extra_flags), verbose)
    if not force and not _stale():
        return LIB
    return _build_to(LIB, [], verbose)


def _build_to(LIB, extra, verbose):
    objdir = os.path.join(HERE, "build") if not extra else os.path.join(HERE, "build", os.path.basename(LIB) + ".d")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        objs.append(o)
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
