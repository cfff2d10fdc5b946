"""Developer probe: wall-clock (synchronised) timings of the C-ABI entry points at the BASELINE.json sizes.

Not a benchmark (bench.py is); used under ncu to get launch lists and to see where a change moved time.
usage: python tools/probe.py [nl|rdf|q6|all] [N] [reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import _capi, data  # noqa: E402


def timed(ctx, fn, reps):
    fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ctx.synchronize()
        ts.append(time.perf_counter() - t0)
        del out
    return min(ts) * 1e3, float(np.median(ts)) * 1e3


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = _capi.Context(0)
    print(_capi.lib().fgpu_version().decode())
    L = (n / 0.08) ** (1.0 / 3.0)
    box, pts = data.make_random_system(L, n, seed=0)
    dp = _capi.DevicePoints(ctx, box, pts)
    if what in ("nl", "all"):
        for flavour, name in ((0, "wrap"), (1, "image")):
            print(f"build_cells r=3: {timed(ctx, lambda: dp.build_cells(3.0), reps)} ms")
            best, med = timed(ctx, lambda: dp.ball_query(None, flavour, 3.0, 0.0, True), reps)
            nl = dp.ball_query(None, flavour, 3.0, 0.0, True)
            print(f"NL {name} r=3 N={n}: bonds={nl.num_bonds} best {best:.3f} ms median {med:.3f} ms "
                  f"({nl.num_bonds / best / 1e3:.1f} M bonds/s)")
            t0 = time.perf_counter()
            h = nl.to_host()
            print(f"   D2H of {sum(v.nbytes for v in h.values()) / 1e6:.0f} MB: {(time.perf_counter() - t0) * 1e3:.1f} ms")
            del nl, h
    if what in ("rdf", "all"):
        for flavour, name in ((1, "image"), (0, "wrap")):
            rdf = _capi.DeviceRDF(ctx, 100, 5.0)
            best, med = timed(ctx, lambda: rdf.accumulate(dp, None, flavour, 5.0, 0.0, True), reps)
            print(f"RDF {name} r=5 bins=100 N={n}: best {best:.3f} ms median {med:.3f} ms -> {1e3 / best:.1f} frames/s; "
                  f"sum={int(rdf.read().astype(np.uint64).sum())}")
    if what in ("q6", "all"):
        m = max(2, round((n / 4) ** (1.0 / 3.0)))
        box, pts = data.make_fcc_system(m, sigma_noise=0.05, seed=0)
        dq = _capi.DevicePoints(ctx, box, pts)
        best, med = timed(ctx, lambda: dq.knn_query(None, 12, exclude_ii=True), reps)
        print(f"kNN k=12 N={len(pts)}: best {best:.3f} ms median {med:.3f} ms")
        nl = dq.knn_query(None, 12, exclude_ii=True)
        best, med = timed(ctx, lambda: dq.steinhardt(nl, [6], want_qlm=False), reps)
        out = dq.steinhardt(nl, [6], want_qlm=False)
        print(f"Steinhardt Q6 N={len(pts)}: best {best:.3f} ms median {med:.3f} ms -> {len(pts) / best / 1e3:.1f} M particles/s; "
              f"mean ql={out['ql'].mean():.6f} order={out['order'][0]:.6f}")
    print("launches:", ctx.launch_count)


if __name__ == "__main__":
    main()
