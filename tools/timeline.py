"""Developer probe: device timeline of one step of a workload (kernel begin/end from the library's event pairs), to see
the idle time between dependent launches.  usage: python tools/timeline.py [nl|q6] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import _capi, data  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "nl"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    ctx = _capi.Context(0)
    if what == "nl":
        box, pts = data.make_random_system((n / 0.08) ** (1.0 / 3.0), n, seed=0)
        dp = _capi.DevicePoints(ctx, box, pts)
        def step():
            dp.build_cells(3.0)  # forced rebuild, as bench.py's step
            return dp.ball_query(None, _capi.FLAVOUR_WRAP, 3.0, 0.0, True)
    else:
        m = max(2, round((n / 4) ** (1.0 / 3.0)))
        box, pts = data.make_fcc_system(m, sigma_noise=0.05, seed=0)
        dp = _capi.DevicePoints(ctx, box, pts)
        step = lambda: dp.steinhardt_knn(12, [6])  # noqa: E731
    for _ in range(3):
        out = step()
        del out
    ctx.synchronize()
    ctx.profile(True)
    ctx.kernel_time("", reset=True)
    for rep in range(3):
        out = step()
        del out
        tl = ctx.kernel_timeline()
        ctx.kernel_time("", reset=True)
        prev_end = 0.0
        print(f"--- step {rep}")
        for name, b, e in tl:
            print(f"{name:16s} begin {b:9.1f}  end {e:9.1f}  dur {e - b:7.1f}  gap {b - prev_end:6.1f}")
            prev_end = e
        print(f"kernels {sum(e - b for _, b, e in tl):.1f} us of {tl[-1][2]:.1f} us")


if __name__ == "__main__":
    main()
