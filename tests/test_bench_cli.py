"""bench.py's command-line contract, the parts that run without a GPU: the reference arm (the CPU restatement timed on
the host cores, rank 0 only) and the loud failure of the GPU arm when no device exists."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--workload", "ml-100k", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE json line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rm2_users_scored_per_sec" and d["unit"] == "users/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert "ml-100k" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["unit"] == "users/s" and cb["cores"] >= 1 and cb["sample"]
    assert cb["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_0_only():
    p = _run(["--impl", "reference", "--workload", "ml-100k", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""      # the other ranks exit 0 without work


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run(["--workload", "tiny", "--steps", "1", "--no-cpu-baseline", "--no-secondary", "--no-e2e"], timeout=300)
    assert p.returncode != 0                                 # no CPU fallback, no JSON line
    assert not any(l.lstrip().startswith("{") for l in p.stdout.splitlines())
