"""bench.py must give the same per-step time whatever step count the driver asks for: the timed region
replays graphs that were rehearsed beforehand and starts behind a device-side gate, so a short region
(--steps 20 --warmup 5, the driver's command) holds no graph upload or idle-GPU launch latency."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu-baseline", *flags],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout  # the contract: ONE JSON line on stdout
    return json.loads(lines[0])


def test_short_and_long_runs_agree(cuda_device):
    short = run_bench("--steps", "20", "--warmup", "5")
    long_ = run_bench("--steps", "4000", "--warmup", "50")
    assert short["steps"] == 20 and long_["steps"] == 4000
    a, b = short["ms_per_step"], long_["ms_per_step"]
    assert abs(a - b) <= 0.10 * b, (a, b, short["region_ms"], long_["region_ms"])
    for line in (short, long_):
        assert line["gpu_launches"] > 0 and line["roofline"]["frac"] > 0
        assert line["region_ms"]["max"] <= 1.5 * line["region_ms"]["min"], line["region_ms"]
