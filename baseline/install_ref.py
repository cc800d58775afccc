"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).

The reference (jcfszxc/jcfszxc-UNet) is a directory of plain Python modules: it has no setup.py / pyproject.toml, so
`pip install /root/reference` has nothing to build (recorded in DESIGN.md §8).  The install is therefore what pip would
do for a pure-Python package: a byte-for-byte copy of the importable modules on the hot path
(UNetFamily/*.py, UNetFamily/utils/unet_parts.py, utils/dice_score.py) into baseline/_ref/, from where
`bench.py --impl reference` (host cores) and `bench.py --impl torch_cuda` (cuDNN on the same B200) import them.
Nothing under baseline/_ref is tracked by git and nothing of the product imports it.

    python baseline/install_ref.py          # runs only where /root/reference exists (not on the GPU box)
"""
from __future__ import annotations

import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["UNetFamily", "utils/dice_score.py"]


def install(verbose: bool = False) -> str | None:
    if not os.path.isdir(REF):
        return DST if os.path.isdir(DST) else None
    os.makedirs(DST, exist_ok=True)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
        if verbose:
            print("installed", rel, file=sys.stderr)
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write(REF + "\n")
    return DST


if __name__ == "__main__":
    print(install(verbose=True))
