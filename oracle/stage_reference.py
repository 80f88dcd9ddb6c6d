#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for the CPU reference arm of bench.py (TEST/BENCH INFRASTRUCTURE).

    python oracle/stage_reference.py [--ref /root/reference]

Copies gym_pybullet_drones/**/*.py and assets/*.urdf from the read-only reference tree into the git-ignored
oracle/_ref/ (never into history: .gitignore lists oracle/_ref/; it still travels to the GPU box with the gpurun snapshot,
which has no /root/reference) and writes oracle/_ref/MANIFEST.json with the sha256 of every staged file, so that a reader
can check the staged copy is byte-identical to the reference.  __graft_entry__.build() runs this whenever the reference
tree is present.  Nothing under oracle/_ref/ is imported by the product; only `bench.py --impl reference`, the bench's
cpu_baseline leg and tests/ execute it (under the pybullet/gymnasium stand-ins of oracle/refshim)."""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))


def stage(ref_root="/root/reference", out=os.path.join(HERE, "_ref")):
    src = os.path.join(ref_root, "gym_pybullet_drones")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(out, "gym_pybullet_drones")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    manifest = {}
    for root, _dirs, files in os.walk(src):
        rel = os.path.relpath(root, src)
        for f in files:
            keep = f.endswith(".py") or (rel == "assets" and f.endswith(".urdf"))
            if not keep:
                continue
            os.makedirs(os.path.join(dst, rel), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
            with open(os.path.join(root, f), "rb") as fh:
                manifest[os.path.normpath(os.path.join("gym_pybullet_drones", rel, f))] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(out, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    d = stage(a.ref)
    print("staged" if d else "reference tree not found", d or a.ref)
