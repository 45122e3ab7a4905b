"""bench.py contract, CPU side: the reference arm (the oracle timed on host cores) prints ONE JSON line with the keys the
driver reads, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "cpu_baseline", "e2e"}


def run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C3", "--scale", "0.05",
                          "--steps", "1", "--warmup", "1", "--gpus", env_extra.get("WORLD_SIZE", "1")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_line():
    lines = run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "ba_observations_per_s_linearize_schur" and d["unit"] == "obs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["workload"] == "C3"


def test_reference_arm_other_ranks_exit_quietly():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
