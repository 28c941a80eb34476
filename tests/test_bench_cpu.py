"""bench.py contract on the CPU side: the reference arm prints exactly ONE JSON line on stdout (library chatter goes to
stderr) with the keys the driver reads, for the image config and for the Gen1-style event-frame config."""
import json
import os
import subprocess
import sys

from util import ROOT


def _run(*extra):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "resnet10", "--img", "96",
                        "--steps", "1", "--warmup", "1", *extra], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run()
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "images/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_event_frames():
    d = _run("--events", "--T", "5")
    assert "event frames" in d["config"]["workload"] and d["value"] > 0
