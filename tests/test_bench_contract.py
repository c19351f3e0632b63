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


import pytest


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line_with_parity_and_legs():
    """A small run of our arm on the GPU: the contract keys, the roofline / parity / e2e objects, and an attached workload."""
    r = run_bench("--workload", "tiny", "--legs", "tiny,c2q", "--reads", "100000", "--steps", "3", "--warmup", "3", "--quick-cpu", "--no-sharded")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity", "workloads"):
        assert key in d, key
    assert d["metric"] == "kmer_lookups_per_s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] > 0 and d["value"] > 0
    assert d["clocks"]["samples"] >= 1 and d["clocks"]["sm_mhz"] > 0 and d["clocks"]["sm_max_mhz"] >= d["clocks"]["sm_mhz"]
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "l2-latency") and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["frac_of_level_ceiling"] > 0 and "frac_with_io" not in rf
    p = d["parity"]
    assert p["checker"] == "sbwt_ref" and p["reads"] == 100000 and all(p[k_] for k_ in ("lookups_match", "hits_match", "checksum_match", "weighted_checksum_match"))
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["hits_only"]["value"] > 0 and e["bitmap_only"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["cpu_model"]
    assert d["cli_e2e"]["plain"]["lookups"] == d["lookups_per_step_per_gpu"]
    assert d["cli_e2e"]["gzip_input"]["lookups"] == d["cli_e2e"]["bgzf_input"]["lookups"] > 0
    w = d["workloads"]["c2q"]
    assert w["value"] > 0 and w["parity"]["weighted_checksum_match"] and w["roofline"]["kernel_ms"] > 0
