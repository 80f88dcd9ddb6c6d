"""bench.py contract on the CPU-runnable arm: `--impl reference` prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--envs", "2048"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "drone-substeps/sec" and d["steps"] == 3
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    staged = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "gym_pybullet_drones"))
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")     # the reference's own Python when it is staged
    assert set(d["config"]) >= {"workload", "envs_per_gpu", "substeps_per_step", "parallelism"}
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                          "--warmup", "1", "--envs", "1024"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_port_kind():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-kind", "port", "--steps", "3",
                          "--warmup", "1", "--envs", "2048"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip().startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 1e6
