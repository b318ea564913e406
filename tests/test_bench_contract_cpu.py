"""bench.py contract checks that need no GPU: the reference arm (the UNMODIFIED reference from baseline/_ref timed on
the host cores) prints exactly ONE JSON line on stdout with the keys the driver reads, and its `config` is the repo
arm's, byte for byte."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, DOD_BENCH_BATCH="64")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["n_gpus"] == 1
    assert j["unit"] == "images/s" and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench._config(1)              # same config object as the repo arm prints
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in j["config"] and j["vs_baseline"] is None


def test_ours_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
