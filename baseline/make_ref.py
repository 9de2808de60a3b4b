#!/usr/bin/env python
"""Installs the UNMODIFIED reference (SafeRL-Lab/Massive-MARL-Benchmark, /root/reference) into the git-ignored
`baseline/_ref/` so that `bench.py --impl reference` and the drop-in tests can run the reference's own code on the GPU
box (where /root/reference does not exist; `baseline/_ref/` is git-ignored but travels with the gpurun snapshot).

    python baseline/make_ref.py [--reference /root/reference]

It is the install the task statement names:
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of the reference>
run from a copy under the system temp dir because /root/reference is read-only and setuptools writes build/ and
*.egg-info into the source tree; `--no-deps` because gym / matplotlib / ipdb are not in the offline wheelhouse (they and
`isaacgym` are stubbed at run time by oracle/refshim).  Only the `agents` package is installed
(setup.py: find_packages); cfg/*.yaml is copied next to it for the drop-in tests.  Nothing is patched here: the two
files that cannot run as shipped on current torch (bool-minus-int in one_ant.py:505 / ten_ant.py:1074..1164, four
prints inside a TorchScript function) are patched into a temp dir at import time by oracle/refshim, as in the tests.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def make(reference="/root/reference", quiet=True):
    """Returns DEST, or None when the reference tree is not available (e.g. on the GPU box)."""
    if not os.path.isdir(os.path.join(reference, "agents")):
        return DEST if os.path.isdir(os.path.join(DEST, "agents")) else None
    tmp = tempfile.mkdtemp(prefix="mmb_ref_")
    try:
        src = os.path.join(tmp, "src")
        os.makedirs(src)
        shutil.copytree(os.path.join(reference, "agents"), os.path.join(src, "agents"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for f in ("setup.py", "README.md", "LICENSE"):
            if os.path.exists(os.path.join(reference, f)):
                shutil.copy(os.path.join(reference, f), src)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout + res.stderr)
        if not quiet:
            print(res.stdout[-400:])
        # setup.py's find_packages() skips directories without an __init__.py (agents/algorithms/utils, .../marl/utils, ...),
        # which the reference imports as namespace packages when run from its checkout: complete the tree from the source
        shutil.copytree(os.path.join(reference, "agents"), os.path.join(DEST, "agents"), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if os.path.isdir(os.path.join(reference, "cfg")):
            shutil.copytree(os.path.join(reference, "cfg"), os.path.join(DEST, "cfg"), dirs_exist_ok=True)
        for root, dirs, _ in os.walk(DEST):
            for d in list(dirs):
                if d == "__pycache__":
                    shutil.rmtree(os.path.join(root, d))
                    dirs.remove(d)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    out = make(a.reference, quiet=False)
    print("reference installed in", out if out else "(reference tree not found; nothing done)")
