#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for `bench.py --impl reference` (and the cpu_baseline leg).

    python baseline/stage_reference.py [/root/reference]

The reference is pure Python whose setup.py selects no packages (`packages=[p for p in find_packages() if
p.startswith('node')]`, setup.py:32-33), so `pip install --target baseline/_ref /root/reference` installs an
empty distribution; what "installing" it means is having `atacom/` importable.  This copies the package
directory as it is — no file is edited — into baseline/_ref/atacom.  baseline/_ref/ is git-ignored (the
reference's sources never enter this repository's history) but NOT gpurun-ignored, so it travels to the GPU box
with the snapshot; `__graft_entry__.build()` runs this whenever /root/reference is present.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def stage(src_root="/root/reference"):
    src = os.path.join(src_root, "atacom")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DEST, "atacom")
    same = os.path.isdir(dst) and not filecmp.dircmp(src, dst, ignore=["__pycache__", "urdf"]).diff_files \
        and os.path.exists(os.path.join(dst, "atacom.py"))
    if not same:
        shutil.rmtree(dst, ignore_errors=True)
        # the URDF meshes (26 STL files) are not on the path; the URDF itself is small and kept
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.stl", "*.STL", "*.dae", "*.obj"))
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as f:
        f.write("%s (unmodified copy of the package directory; see baseline/stage_reference.py)\n" % src)
    return dst


if __name__ == "__main__":
    out = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged:" if out else "reference tree not present; nothing staged", out or "")
