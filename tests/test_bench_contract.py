"""bench.py's reference arm (the unmodified reference's CPU Transport on the bench workload, `--impl reference`) runs
without a GPU: check the JSON contract on a tiny sample. The B200 arm needs a GPU; its line carries the same keys plus
roofline / clocks / gpu_launches (checked by the driver at round end)."""
import json
import os
import subprocess
import sys

import pytest

import support as T


@pytest.mark.skipif(not T.have_reference(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(T.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-histories", "3600"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "photon histories/s" and line["unit"] == "histories/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["config"]["voxels"] == 512 * 512 * 400 and line["config"]["exposures"] == 3600
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "reference" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and "3600 exposures" in cpu["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
