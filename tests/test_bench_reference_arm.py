"""The reference arm of bench.py (`--impl reference`: the oracle timed on the host cores, no GPU) must run on a CPU-only
box and print one JSON line with the contract's keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--scale", "0.01"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["metric"] == "batch_knn_cosine_top10_queries_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 2 and line["warmup"] == 1 and line["vs_baseline"] is None


def test_reference_arm_other_ranks_print_nothing():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--scale", "0.01"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
