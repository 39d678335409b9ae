"""CPU: the `bench.py --impl reference` arm (the reference's CPU data flow timed on the host cores) prints ONE JSON line with
the keys the driver's contract names; the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "1"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MedMamba-T train images/sec @224" and d["unit"] == "images/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["config"]["cpu_batch"] == 1          # --cpu-batch is honoured (round 1 picked the batch from a timing heuristic)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and abs(e["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0", "--no-cpu-baseline"], cwd=ROOT, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode != 0                     # no silent CPU fallback
    assert not [l for l in out.stdout.splitlines() if l.startswith("{") and '"value"' in l]


def test_reference_arm_medssd():
    """BASELINE.json configs[2]: the same arm for the SSD family (eager CPU tree + the from-definition SSD oracle)."""
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--model", "medssd", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1"], cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["metric"] == "MedSSD train images/sec @224" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["config"]["cpu_batch"] == 1
