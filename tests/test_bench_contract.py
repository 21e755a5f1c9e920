"""bench.py's reference arm runs on the host cores only (the reference's own sources under
oracle/_ref, else the oracle port), so its JSON line can be held to the driver's contract on CPU:
one line, the keys the driver reads, no product library mapped."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "tiny", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GoldRush-Path Gbp/s hashed+queried"
    assert d["unit"] == "Gbp/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["same_config"] is True and cb["sample"].startswith("full")
    assert d["e2e"] == {"value": d["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_leaves_the_other_ranks_idle():
    """Under torchrun only rank 0 measures; the others exit 0 without printing."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "tiny", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
