"""Per-step wall times of a bench workload's device-resident and end-to-end steps (debugging aid).
usage: python tools/e2e_probe.py <hist-client workload> [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from freud_b200 import _capi  # noqa: E402


def main():
    ctx = _capi.Context(0)
    w = bench.workload_hist_client(ctx, 0, int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000, sys.argv[1])
    for name in ("step_dev", "step_e2e", "step_e2e"):
        ts = []
        for _ in range(16):
            t0 = time.perf_counter()
            w[name]()
            ts.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        print(name, " ".join(f"{t:.1f}" for t in ts), "host-binned", w["hist"].host_binned_bonds, flush=True)
    del w
    ctx.close()


main()
