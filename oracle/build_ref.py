"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, taken from where it lies.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/dcl_oracle.py).  The reference's path is one self-contained Python file
(`/root/reference/utils/loss.py`; it imports torch only), so "building" it is placing that file, unmodified, at
`oracle/_ref/loss.py`.  `oracle/_ref/` is git-ignored (reference sources never enter the history) but travels to the
GPU box with the snapshot, where `bench.py --impl reference` and `cpu_baseline` load it by path and time it on the
host cores (`cpu_baseline.kind == "reference"`).  Run in the build container only:

    python oracle/build_ref.py
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/utils/loss.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "loss.py")


def build(verbose=False):
    """-> path of oracle/_ref/loss.py, or None when neither the reference nor a previous build is present."""
    if os.path.exists(SRC):
        os.makedirs(DST_DIR, exist_ok=True)
        if not os.path.exists(DST) or open(SRC, "rb").read() != open(DST, "rb").read():
            shutil.copyfile(SRC, DST)
        with open(os.path.join(DST_DIR, "SOURCE.txt"), "w") as f:
            f.write("%s sha256 %s\n" % (SRC, hashlib.sha256(open(DST, "rb").read()).hexdigest()))
        if verbose:
            print("oracle/_ref/loss.py <-", SRC)
        return DST
    return DST if os.path.exists(DST) else None


def load():
    """The reference module, loaded by file path (its package __init__ needs matplotlib); None if absent."""
    import importlib.util
    path = build()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("dcl_ref_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose=True)
    sys.exit(0 if p else 1)
