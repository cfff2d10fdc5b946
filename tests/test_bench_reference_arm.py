"""bench.py --impl reference on a small case (CPU only: the reference's own classes from oracle/_ref, no GPU, no product
code on the path): the line keeps the driver's contract -- impl, metric / unit of the b200 arm's leg, a cpu_baseline
describing the run, an e2e object with zero copies, a positive value and ms_per_step."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload, metric, unit", [("q6", "q6_particles_per_sec", "particles/s"),
                                                    ("rdf", "rdf_frames_per_sec", "frames/s")])
def test_reference_arm_line(workload, metric, unit):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                           "--n", "4000", "--steps", "1", "--warmup", "0"], cwd=ROOT, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == metric and line["unit"] == unit
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    assert line["gpu_launches"] == 0 and line["data"] == "synthetic" and "workload" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "reference" and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
