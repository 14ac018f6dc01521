"""bench.py contract: one JSON line with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the CPU restatement of the reference path on a bounded sample (runs without a GPU)."""
    d = _run(["--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "npdr_attr_x_pair_glm_evals_per_s" and d["unit"] == "evals/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "small" and d["vs_baseline"] is None and d["dtype"] == "f64"
    # same static config object as the GPU arm prints (N, p, sample-independent), all host cores asked for explicitly
    assert d["config"]["N"] == 1024 and d["config"]["p"] == 2048 and "sample" not in d["config"]
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_reference_arm_ignores_omp_num_threads_1():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use the host's cores."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--workload", "small", "--steps", "2", "--warmup", "3"])
    assert BASE_KEYS | {"roofline", "clocks", "stages"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["higher_is_better"] is True
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1.0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 1024 * 2048 * 8 and e["d2h_bytes_per_step"] == 2048 * 64
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    assert e["host_memory"] == "pageable" and e["api"].startswith("npdrb_npdr")
    assert d["config"]["workload"] == "small" and d["pairs_M"] > 0 and d["target"] is None   # target only rides on cfg4
