#!/usr/bin/env python
"""One small frame through every kernel family of the hot path, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(SURVEY.md section 5).  Each result is also checked against the oracle, so a sanitizer run doubles as a parity run.

    compute-sanitizer --tool racecheck python tools/sanitizer_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import _capi, data  # noqa: E402
from oracle import port  # noqa: E402

WRAP, IMAGE, GHOST = 0, 1, 2
ctx = _capi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=3, tilt=(0.2, 0.1, -0.1))
dp = _capi.DevicePoints(ctx, box, pts)
for mapping in (0, 1):  # tile walk, lanes over queries
    ctx.set_tuning("lanes_over_queries", mapping)
    for flavour, pf in ((WRAP, port.WRAP), (IMAGE, port.IMAGE), (GHOST, port.GHOST)):
        got = dp.ball_query(None, flavour, 3.0, 0.0, True).to_host()
        want = port.ball_nlist(pf, box, False, pts, pts, 3.0, 0.0, True)
        assert np.array_equal(got["neighbors"], want.neighbors) and np.array_equal(
            got["vectors"].view(np.uint32), want.vectors.view(np.uint32)), (mapping, flavour)
ctx.set_tuning("lanes_over_queries", -1)
q = pts[: n // 3] + np.float32(0.1)
got = dp.ball_query(q, IMAGE, 3.0, 0.5, False, sort_by_distance=True).to_host()  # query sort path
want = port.ball_nlist(port.IMAGE, box, False, pts, q, 3.0, 0.5, False, True)
assert np.array_equal(got["neighbors"], want.neighbors)
for flavour, pf in ((IMAGE, port.IMAGE), (WRAP, port.WRAP)):
    rdf = _capi.DeviceRDF(ctx, 100, 3.0)
    rdf.accumulate(dp, None, flavour, 3.0, 0.0, True)  # IMAGE: symmetric walk
    assert np.array_equal(rdf.read(), port.rdf_accumulate(pf, box, False, pts, pts, 100, 3.0, 0.0, True))
total = np.zeros(100, np.uint64)
rdf = _capi.DeviceRDF(ctx, 100, 3.0)
for shard in range(3):  # slab build: cp.async.bulk ring, two-piece scan, scatter with atomic ranks
    dp.set_shard(shard, 3)
    rdf.reset()
    rdf.accumulate(dp, None, IMAGE, 3.0, 0.0, True)
    total += rdf.read()
dp.set_shard(0, 1)
assert np.array_equal(total.astype(np.uint32), port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 100, 3.0, 0.0, True))
fbox, fpts = data.make_fcc_system(8, sigma_noise=0.05, seed=2)
dq = _capi.DevicePoints(ctx, fbox, fpts)
nl = dq.knn_query(None, 12, exclude_ii=True)
pnl = port.knn_nlist(fbox, False, fpts, fpts, 12, exclude_ii=True)
assert np.array_equal(nl.to_host()["neighbors"], pnl.neighbors)
st = dq.steinhardt(nl, [4, 6], average=True, wl=True)
want = port.steinhardt(fbox, False, fpts, pnl, [4, 6])
ql_plain = dq.steinhardt(nl, [4, 6])["ql"]
assert np.allclose(ql_plain, want["ql"], rtol=1e-5, atol=1e-6)
ctx.synchronize()
print(f"sanitizer probe ok: {ctx.launch_count} kernel launches")
