"""CPU: bench.py's reference arm prints ONE JSON line with the keys the driver's contract names (the GPU arm needs a
B200; its line is checked on the box by the driver). Also pins the FLOP count the roofline figures rest on."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--img-size", "224"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 0 and "workload" in d["config"]


def test_flop_count_per_image_matches_the_survey():
    sys.path.insert(0, REPO)
    import bench
    assert abs(bench.flop_per_image(518) / 1e9 - 1041.63) < 0.01
    assert abs(bench.flop_per_image(336) / 1e9 - 393.72) < 0.01
