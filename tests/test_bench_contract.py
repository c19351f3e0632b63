"""The parts of bench.py's contract that need no GPU: the reference arm (the reference's own query code on the host
cores) prints one JSON line with the agreed keys, and under a multi-rank launch only rank 0 runs it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kmer_lookups_per_s" and d["unit"] == "lookups/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("tiny") and d["config"]["k"] == 31
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "reads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    r = run_bench("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1", "--gpus", "2",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = run_bench("--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
