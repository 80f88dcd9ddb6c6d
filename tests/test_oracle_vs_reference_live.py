"""The oracle against the unmodified reference Python, live (not through fixtures): random configurations replayed in both.
Runs where the reference tree exists (the build container) or where its staged copy oracle/_ref does (the GPU box; the
same check is repeated there under the gpu marker by tests/test_gpu_round2.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference tree (build container) or its staged, unmodified copy (oracle/_ref: written by __graft_entry__.build(), git-ignored,
# travels to the GPU box)
REF = next((r for r in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")) if os.path.isdir(os.path.join(r, "gym_pybullet_drones"))),
           "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gym_pybullet_drones")), reason="reference tree not present")
def test_oracle_matches_live_reference_on_random_configs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_vs_reference.py"), "--ref", REF, "--seeds", "60"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-2000:] + out.stdout[-2000:]
    d = json.loads(lines[-1])
    assert d["cases"] == 60 and d["failures"] == []
    assert d["open_loop_worst"] <= 1e-10 and d["closed_loop_worst"] <= 1e-6
