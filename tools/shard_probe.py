#!/usr/bin/env python
"""One GPU plays each of W ranks of the sharded 4 M-point RDF (BASELINE.json configs[3]) in turn and reports the
per-kernel device times of its step -- the Amdahl terms of the multi-GPU curve without needing W GPUs.
usage: shard_probe.py [W=8] [N=4000000] [steps=10]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import _capi, data  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = _capi.Context(0)
box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=0, tilt=(0.3, 0.2, 0.1))
dp = _capi.DevicePoints(ctx, box, pts)
rdf = _capi.DeviceRDF(ctx, 500, 5.0)
names = ("cell_prep", "cell_assign", "scan", "cell_scatter", "search_rdf")
out = {}
for world in (1, W):
    for shard in range(world):
        dp.set_shard(shard, world)
        for _ in range(3):
            rdf.reset(); dp.build_cells(5.0); rdf.accumulate(dp, None, 1, 5.0, 0.0, True)
        ctx.synchronize()
        ctx.profile(True)
        ctx.kernel_time(reset=True)
        for _ in range(steps):
            rdf.reset(); dp.build_cells(5.0); rdf.accumulate(dp, None, 1, 5.0, 0.0, True)
        t = {k: round(ctx.kernel_time(k)[0] / steps * 1e3, 1) for k in names}
        t["sum_us"] = round(sum(t.values()), 1)
        ctx.profile(False)
        ctx.kernel_time(reset=True)
        out[f"{shard}/{world}"] = t
        print(f"shard {shard}/{world}: {t}", flush=True)
dp.set_shard(0, 1)
one = out["0/1"]["sum_us"]
worst = max(v["sum_us"] for k, v in out.items() if k.endswith(f"/{W}"))
print(json.dumps({"kernels_us_1gpu": one, f"kernels_us_worst_of_{W}": worst, "kernel_speedup_bound": round(one / worst, 2)}))
