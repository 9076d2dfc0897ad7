#!/usr/bin/env python
"""Install the UNMODIFIED reference python sources into the git-ignored baseline/_ref/ (bench.py's reference arms).

    python baseline/install_reference.py            # build container only: needs /root/reference

The reference (ms-dot-k/Visual-Context-Attentional-GAN) ships no setup.py / pyproject, so `pip install --target` has
nothing to build; its "install" is a verbatim copy of the importable tree (src/, the four drivers) -- byte-identical
files, checked by sha256 against the source tree.  baseline/_ref/ is listed in .gitignore (reference sources never enter
the history) but NOT in .gpurunignore, so it travels to the GPU box with the snapshot, where /root/reference does not
exist.  Nothing on the product path imports it: only bench.py's `--impl reference`, `--impl reference-gpu` and the
`gpu_eager_baseline` leg do.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VCA_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
WHAT = ["src", "train.py", "train_LRS.py", "test.py", "test_LRS.py", "LICENSE"]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def install(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "src", "models"))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for item in WHAT:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(s):
            shutil.copyfile(s, d)
    for root, _, files in os.walk(DST):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), DST)
                assert _sha(os.path.join(DST, rel)) == _sha(os.path.join(SRC, rel)), rel
                n += 1
    if verbose:
        print(f"installed {n} unmodified reference python files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
