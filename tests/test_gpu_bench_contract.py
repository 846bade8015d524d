"""bench.py end to end on the B200 with a small lattice: the one JSON line carries every key of the contract."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--L", "16", "--n-beta", "8", "--steps", "4",
                          "--warmup", "3", "--ref-sweeps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == "spin_flip_attempts_per_s" and line["value"] > 0 and line["n_gpus"] == 1
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and "workload" in line["config"]
    # a roofline fraction is a fraction; the strong-scaling line carries the sustained leg and the time-to-target rows
    assert 0 < line["roofline"]["frac"] <= 1.2 and line["roofline"]["bound"] == "hbm"
    assert line["scaling"] == "strong" and line["sustained"]["seconds"] >= 4.0 and line["sustained"]["clocks"] is not None
    ttt = line["time_to_target"]
    assert len(ttt) == 3 and all(r["gpu_median_s"] > 0 and r["cpu_median_s"] > 0 for r in ttt)
    assert any("Wishart" in r["instance"] for r in ttt)
    assert "NPT(J, h, mode='production').run" in line["e2e"]["api"] and line["e2e_host_buffers"]["value"] > 0
