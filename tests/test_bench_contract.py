"""bench.py's contract on a machine without a GPU: the reference arm (the CPU oracle port on the host
cores -- the Rust reference cannot be built in this image) prints one JSON line with the agreed keys,
and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    return p.returncode, lines, p.stderr


def test_reference_arm_prints_one_json_line():
    rc, lines, err = _run("--impl", "reference", "--workload", "gencode_small", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1", "--host-threads", "2", "--cache-dir", "/tmp")
    assert rc == 0, err[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "reads/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("reads/sec pseudoaligned") and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    rc, lines, _ = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert rc == 2 and len(lines) == 1 and "no CUDA device" in json.loads(lines[0])["error"]
