#!/usr/bin/env python
"""Benchmark of the neighbour-query + pair-accumulation hot path (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload nl|rdf|q6|rdf4m|traj2d]

Default workload ("nl", BASELINE.json configs[1]): LinkCell neighbour-list query, r_max = 3, exclude_ii, on 1 M
uniform random points in a cubic periodic box at density 0.08; one step = cell-list build + 27-cell pair search +
sorted CSR NeighborList (all five arrays) for one frame.  Metric: neighbour pair evaluations per second, where a
pair evaluation is one execution of the reference's distance arithmetic on a (query, candidate) pair of the
27-cell scheme (the same count for the reference's own LinkCell at cell_width = r_max).
N > 1 (torchrun, one rank per GPU): every rank processes its own frame (seed = rank), no data-path collective,
"scaling": "weak".  The other workloads are the remaining BASELINE.json configs and print the same line shape.

Only the cpu_baseline / --impl reference legs touch oracle/ (as the thing being timed beside the GPU path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO = 0.08
WRAP, IMAGE = 0, 1
TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this workload "
                  "(profiles/ncu_r*_summary.md; the newest capture of the kernel named)")


# ---------------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = max(mx, float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


def pinned_empty(shape, dtype):
    """numpy view of pinned host memory (torch is plumbing here: it owns the page-locked allocation)."""
    import torch

    tdt = {np.float32: torch.float32, np.uint32: torch.int32, np.complex128: torch.complex128}[dtype]
    t = torch.empty(tuple(int(s) for s in np.atleast_1d(shape)), dtype=tdt, pin_memory=True)
    return t.numpy().view(dtype), t  # the caller keeps t alive


# ---------------------------------------------------------------------------------------------------------
# workloads: each returns dict(step_dev, step_e2e, units_per_step, unit, metric, config, h2d, d2h, algo_bytes)
def workload_nl(ctx, rank, n, flavour=WRAP, r_max=3.0):
    from freud_b200 import _capi, data

    L = (n / RHO) ** (1.0 / 3.0)
    box, pts = data.make_random_system(L, n, seed=rank)
    dp = _capi.DevicePoints(ctx, box, pts)
    # pair evaluations of one step, counted on the device in an untimed pass
    ctx.count_pair_evals(True)
    ctx.pair_evals(reset=True)
    nl = dp.ball_query(None, flavour, r_max, 0.0, True)
    evals = ctx.pair_evals(reset=True)
    ctx.count_pair_evals(False)
    n_bonds = nl.num_bonds
    del nl
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts
    out, keep = {}, [keep0]
    for key, shape, dt in (("neighbors", (n_bonds, 2), np.uint32), ("distances", (n_bonds,), np.float32),
                           ("weights", (n_bonds,), np.float32), ("vectors", (n_bonds, 3), np.float32),
                           ("segments", (n,), np.uint32), ("counts", (n,), np.uint32)):
        out[key], t = pinned_empty(shape, dt)
        keep.append(t)

    def step_dev():
        dp.build_cells(r_max)  # forced rebuild: the cell list is part of every frame
        return dp.ball_query(None, flavour, r_max, 0.0, True)

    def step_e2e():
        # the call sequence behind LinkCell(box, points).query(points, dict(r_max=3, exclude_ii=True)).toNeighborList()
        d = _capi.DevicePoints(ctx, box, pin_pts)
        lst = d.ball_query(None, flavour, r_max, 0.0, True)
        lst.to_host(into=out)
        return out["neighbors"][-1, 1]

    n_cells = int(np.prod(dp.build_cells(r_max)))
    algo = {
        # per launch, bytes that must move (DESIGN.md "Kernels"): fp32 positions, 16 B float4 on device
        "cell_assign": 12 * n + 8 * n + 4 * n_cells,
        "cell_scatter": 20 * n + 16 * n,
        "search_nl": 16 * (n + n) + 4 * n_cells + 16 * n_bonds + 8 * n,  # single pass: positions in, 16 B/hit bag out
        "search_count": 16 * (n + n) + 4 * n_cells + 4 * n,
        "search_fill": 16 * (n + n) + 4 * n_cells + 4 * n + 16 * n_bonds,
        "emit": 16 * n_bonds + 16 * n + 12 * n + 4 * n + 28 * n_bonds,
        "pipeline": 16 * (n + n) + 28 * n_bonds + 8 * n,  # SURVEY.md section 8d "NL emit"
    }
    # measured DRAM bytes per launch (ncu, profiles/ncu_r2_summary.md section 3): only for the configuration it was taken on
    traffic = {"search_nl": 24105216 + 105016832, "emit": 230652928 + 208475136} if n == 1_000_000 and r_max == 3.0 else {}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=evals, unit="pair_evals/s", traffic=traffic,
                metric="neighbour_pair_evals_per_sec",
                config={"workload": f"LinkCell NeighborList r_max={r_max:g} exclude_ii N={n} cubic L={L:.4f} rho=0.08 "
                                    f"flavour={'wrap' if flavour == WRAP else 'image'}",
                        "bonds_per_step": n_bonds, "pair_evals_per_step": evals, "n_cells": n_cells},
                # 24 B per bond cross PCIe (indices, distance, vector); the unit weights are written by the host
                h2d=12 * n, d2h=24 * n_bonds + 8 * n, algo=algo, keep=keep, box=box, pts=pts, r_max=r_max, dp=dp,
                traffic_source=TRAFFIC_SOURCE, secondary={"bonds": n_bonds})


def workload_rdf(ctx, rank, n, bins=100, r_max=5.0, flavour=IMAGE, tilt=None, is2d=False, density=RHO):
    from freud_b200 import _capi, data

    L = (n / density) ** (0.5 if is2d else 1.0 / 3.0)
    box, pts = data.make_random_system(L, n, is2D=is2d, seed=rank, tilt=tilt)
    dp = _capi.DevicePoints(ctx, box, pts)
    rdf = _capi.DeviceRDF(ctx, bins, r_max)
    ctx.count_pair_evals(True)
    ctx.pair_evals(reset=True)
    rdf.accumulate(dp, None, flavour, r_max, 0.0, True)
    evals = ctx.pair_evals(reset=True)
    ctx.count_pair_evals(False)
    n_bonds = int(rdf.read().astype(np.uint64).sum())
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    def step_dev():
        dp.build_cells(r_max)
        rdf.accumulate(dp, None, flavour, r_max, 0.0, True)

    def step_e2e():
        # RDF(bins, r_max).compute((box, points), reset=False) then .bin_counts
        d = _capi.DevicePoints(ctx, box, pin_pts)
        rdf.accumulate(d, None, flavour, r_max, 0.0, True)
        return rdf.read()

    n_cells = int(np.prod(dp.build_cells(r_max)))
    algo = {"search_rdf": 16 * (n + n) + 4 * n_cells + 4 * bins, "pipeline": 16 * (n + n) + 4 * bins,
            "cell_assign": 20 * n + 4 * n_cells, "cell_scatter": 36 * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=1, unit="frames/s", metric="rdf_frames_per_sec",
                config={"workload": f"RDF bins={bins} r_max={r_max:g} N={n} L={L:.4f} "
                                    f"{'2D ' if is2d else ''}{'triclinic ' if tilt else ''}"
                                    f"flavour={'wrap' if flavour == WRAP else 'image'} fused (no NeighborList)",
                        "bonds_per_step": n_bonds, "pair_evals_per_step": evals},
                h2d=12 * n, d2h=4 * bins, algo=algo, keep=[keep0], box=box, pts=pts, r_max=r_max, rdf=rdf, dp=dp,
                bins=bins, flavour=flavour, traffic_source=TRAFFIC_SOURCE,
                # measured DRAM bytes of one launch (ncu, profiles/ncu_r1_v6_summary.md "full_rdf"): config 1 M / r=5 / image
                traffic={"search_rdf": 16412672} if (n, r_max, flavour, bins, tilt, is2d) == (1_000_000, 5.0, IMAGE, 100, None, False)
                else {},
                secondary={"pair_evals_per_sec_factor": evals})


def workload_q6(ctx, rank, n):
    from freud_b200 import _capi, data

    m = max(2, round((n / 4) ** (1.0 / 3.0)))
    box, pts = data.make_fcc_system(m, sigma_noise=0.05, seed=rank)
    n = len(pts)
    dp = _capi.DevicePoints(ctx, box, pts)
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    # the window radius fgpu_knn_query starts from (sphere expected to hold 1.5 (k + 1) points), so that the forced
    # rebuild below produces the grid the query uses and the step holds exactly one cell-list build
    r_window = float(np.cbrt(3.0 * 1.5 * 13.0 / (4.0 * np.pi * (n / float(box.volume)))))

    def step_dev():
        # what Steinhardt(6).compute(system, neighbors={"num_neighbors": 12}) runs: cell list, window search, then the
        # k nearest of every row and their Y_lm sums in one kernel (no NeighborList); q_l (4 MB) lands in page-locked memory
        dp.build_cells(r_window)
        return dp.steinhardt_knn(12, [6], exclude_ii=True, out={"ql": pin_ql})

    pin_ql, keep1 = pinned_empty((n, 1), np.float32)
    pin_qlm, keep2 = pinned_empty((n * 13 * 2,), np.float32)

    def step_e2e():
        # Steinhardt(6).compute((box, points), neighbors=dict(num_neighbors=12)) then .particle_order; the
        # per-particle q_lm come back too, as the reference's compute() materialises them on the host
        d = _capi.DevicePoints(ctx, box, pin_pts)
        return d.steinhardt_knn(12, [6], exclude_ii=True, want_qlm=True, out={"ql": pin_ql, "qlm": pin_qlm})["ql"]

    algo = {"steinhardt": 588 * n, "knn": 16 * (n + n) + 8 * 12 * n, "search_nl": 16 * (n + n) + 16 * 26 * n + 8 * n,
            "knn_select": (16 * 26 + 12 + 28 * 12) * n, "knn_ylm": (16 * 20 + 8 + 4) * n, "pipeline": 124 * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s", metric="q6_particles_per_sec",
                config={"workload": f"Steinhardt Q6 num_neighbors=12 FCC {m}^3x4={n} sigma=0.05"},
                h2d=12 * n, d2h=4 * n + 104 * n, algo=algo, keep=[keep0, keep1, keep2], box=box, pts=pts, secondary={},
                dp=dp, traffic_source=TRAFFIC_SOURCE,
                # measured DRAM bytes of one launch (ncu, profiles/ncu_r2_summary.md section 3)
                traffic={"search_nl": 19511808 + 218465280, "knn_ylm": 407950592 + 10454016} if n == 1_000_188 else {})


def workload_q6_sharded(ctx, rank, world, n, comm):
    """configs[2] with the ROWS dealt to the ranks (north_star: "per-GPU ... Ql partials are reduced with a single NCCL
    allreduce"): every rank holds the frame, queries the 12 nearest neighbours of its block of particles
    (q_index_offset = first index of the block), accumulates their q_lm, and the fp64 system sums are reduced with one
    ncclAllReduce inside fgpu_steinhardt_compute(comm).  Strong scaling: the value is particles of the ONE frame per
    second."""
    from freud_b200 import _capi, data, parallel

    m = max(2, round((n / 4) ** (1.0 / 3.0)))
    box, pts = data.make_fcc_system(m, sigma_noise=0.05, seed=0)  # the same frame on every rank
    n = len(pts)
    lo, hi = parallel.shard_bounds(n, rank, world)
    dp = _capi.DevicePoints(ctx, box, pts)
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts
    pin_q, keep1 = pinned_empty((hi - lo, 3), np.float32)
    pin_q[:] = pts[lo:hi]
    pin_ql, keep2 = pinned_empty((hi - lo, 1), np.float32)
    r_window = float(np.cbrt(3.0 * 1.5 * 13.0 / (4.0 * np.pi * (n / float(box.volume)))))

    def rows(d):
        nl = d.knn_query(pin_q, 12, exclude_ii=True, q_index_offset=lo)
        return d.steinhardt(nl, [6], want_qlm=False, comm=comm, n_total=n, out={"ql": pin_ql})

    def step_dev():
        dp.build_cells(r_window)
        return rows(dp)

    def step_e2e():
        return rows(_capi.DevicePoints(ctx, box, pin_pts))["ql"]

    block = hi - lo
    algo = {"search_nl": 16 * (n + block) + 16 * 26 * block + 8 * block, "knn_select": (16 * 26 + 12 + 28 * 12) * block,
            "steinhardt": 588 * block, "pipeline": 124 * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s", metric="q6_particles_per_sec",
                config={"workload": f"Steinhardt Q6 num_neighbors=12 FCC {m}^3x4={n} sigma=0.05, rows (query particles) "
                                    f"sharded over {world} GPU(s), points replicated, fp64 system q_lm summed with one "
                                    f"ncclAllReduce"},
                h2d=12 * n, d2h=4 * block, algo=algo, keep=[keep0, keep1, keep2], box=box, pts=pts, secondary={}, dp=dp,
                rows=rows, block=(lo, hi))


def workload_local_density(ctx, rank, n, r_max=2.5, diameter=1.0):
    """SURVEY.md section 8f rank 3, first client: freud.density.LocalDensity(r_max, diameter).compute((box, points)) on
    the C2 system -- ball query of r_max + diameter / 2 = 3 (IMAGE arithmetic), then the fractional count per row."""
    from freud_b200 import _capi, data

    L = (n / RHO) ** (1.0 / 3.0)
    box, pts = data.make_random_system(L, n, seed=rank)
    dp = _capi.DevicePoints(ctx, box, pts)
    rq = r_max + 0.5 * diameter
    n_bonds = dp.ball_query(None, IMAGE, rq, 0.0, True).num_bonds
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    pin_num, keep1 = pinned_empty((n,), np.float32)
    pin_den, keep2 = pinned_empty((n,), np.float32)

    def step_dev():
        # the C ABI hands the two result arrays (8 MB) to the host: that copy is part of every call
        dp.build_cells(rq)
        # query + count in one call: the bonds are summed where the search left them, no NeighborList
        return dp.local_density(None, IMAGE, rq, r_max, diameter, exclude_ii=True, out=(pin_num, pin_den))

    def step_e2e():
        d = _capi.DevicePoints(ctx, box, pin_pts)
        return d.local_density(None, IMAGE, rq, r_max, diameter, exclude_ii=True, out=(pin_num, pin_den))[1]

    n_cells = int(np.prod(dp.build_cells(rq)))
    algo = {"search_nl": 16 * (n + n) + 4 * n_cells + 16 * n_bonds + 8 * n,
            "emit": 16 * n_bonds + 16 * n + 12 * n + 4 * n + 28 * n_bonds,
            "local_density_rows": 16 * n_bonds + 8 * n + 8 * n,
            "pipeline": 16 * (n + n) + 8 * n}  # fused search + count would read the positions and write two floats
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s",
                metric="local_density_particles_per_sec",
                config={"workload": f"LocalDensity r_max={r_max:g} diameter={diameter:g} (ball query r={rq:g}, image "
                                    f"flavour) N={n} cubic L={L:.4f} rho=0.08", "bonds_per_step": n_bonds},
                h2d=12 * n, d2h=8 * n, algo=algo, keep=[keep0, keep1, keep2], box=box, pts=pts, r_max=r_max,
                diameter=diameter,
                secondary={"bonds": n_bonds})


def workload_correlation(ctx, rank, n, bins=100, r_max=3.0):
    """SURVEY.md section 8f rank 3, second client: freud.density.CorrelationFunction(bins, r_max).compute((box, points),
    values) on the C2 system with complex values -- ball query (IMAGE arithmetic), then counts and complex<double>
    sums per distance bin."""
    from freud_b200 import _capi, data

    L = (n / RHO) ** (1.0 / 3.0)
    box, pts = data.make_random_system(L, n, seed=rank)
    rs = np.random.RandomState(rank + 17)
    values, keep1 = pinned_empty((n,), np.complex128)
    values[:] = rs.standard_normal(n) + 1j * rs.standard_normal(n)
    dp = _capi.DevicePoints(ctx, box, pts)
    cf = _capi.DeviceCorrelation(ctx, bins, r_max)
    n_bonds = dp.ball_query(None, IMAGE, r_max, 0.0, True).num_bonds
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    def step_dev():
        # the values (16 MB of complex128, page-locked) cross PCIe in every call: the C ABI takes them from the host
        cf.reset()
        dp.build_cells(r_max)
        cf.accumulate(dp, None, IMAGE, r_max, values, values, exclude_ii=True)  # query + accumulation, no NeighborList
        return cf.read()

    def step_e2e():
        cf.reset()
        d = _capi.DevicePoints(ctx, box, pin_pts)
        cf.accumulate(d, None, IMAGE, r_max, values, values, exclude_ii=True)
        return cf.read()[1]

    n_cells = int(np.prod(dp.build_cells(r_max)))
    algo = {"search_nl": 16 * (n + n) + 4 * n_cells + 16 * n_bonds + 8 * n,
            "emit": 16 * n_bonds + 16 * n + 12 * n + 4 * n + 28 * n_bonds,
            "correlation_rows": 16 * n_bonds + 16 * n_bonds + 16 * n + 20 * bins,  # bag record + gathered complex128 per bond
            "pipeline": 16 * (n + n) + 32 * n + 20 * bins}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s",
                metric="correlation_function_particles_per_sec",
                config={"workload": f"CorrelationFunction bins={bins} r_max={r_max:g} complex values (image flavour) "
                                    f"N={n} cubic L={L:.4f} rho=0.08", "bonds_per_step": n_bonds},
                h2d=12 * n + 16 * n, d2h=20 * bins, algo=algo, keep=[keep0, keep1], box=box, pts=pts, r_max=r_max,
                bins=bins, values=values, secondary={"bonds": n_bonds})


def workload_pmftxy(ctx, rank, n, x_max=4.0, y_max=3.0, bins=(100, 100)):
    """SURVEY.md section 8f rank 3, third client: freud.pmft.PMFTXY(x_max, y_max, bins).compute((box, points),
    orientations) on the 2-D system of config 5 (areal density 0.5): ball query of sqrt(x_max^2 + y_max^2) = 5 (IMAGE
    arithmetic), then the rotated bond vectors binned in 2-D."""
    from freud_b200 import _capi, data

    L = (n / 0.5) ** 0.5
    box, pts = data.make_random_system(L, n, is2D=True, seed=rank)
    rs = np.random.RandomState(rank + 29)
    angles, keep1 = pinned_empty((n,), np.float32)
    angles[:] = rs.random_sample(n) * 2 * np.pi - np.pi
    r_max = float(np.sqrt(x_max ** 2 + y_max ** 2))
    dp = _capi.DevicePoints(ctx, box, pts)
    pm = _capi.DevicePMFTXY(ctx, x_max, y_max, bins[0], bins[1])
    n_bonds = dp.ball_query(None, IMAGE, r_max, 0.0, True).num_bonds
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    def step_dev():
        # the orientations (4 MB) are taken from the host in every call
        pm.reset()
        dp.build_cells(r_max)
        pm.accumulate(dp, None, IMAGE, r_max, angles, exclude_ii=True)  # query + histogram, no NeighborList
        return pm.read()

    def step_e2e():
        pm.reset()
        d = _capi.DevicePoints(ctx, box, pin_pts)
        pm.accumulate(d, None, IMAGE, r_max, angles, exclude_ii=True)
        return pm.read()

    n_cells = int(np.prod(dp.build_cells(r_max)))
    nb = bins[0] * bins[1]
    algo = {"search_nl": 16 * (n + n) + 4 * n_cells + 16 * n_bonds + 8 * n,
            "emit": 16 * n_bonds + 16 * n + 12 * n + 4 * n + 28 * n_bonds,
            "pmft3_rows": 16 * n_bonds + 12 * n + 4 * nb,  # 16 B bag record per bond; offset, count and angle per query
            "pipeline": 16 * (n + n) + 8 * n + 4 * nb}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s", metric="pmftxy_particles_per_sec",
                config={"workload": f"PMFTXY x_max={x_max:g} y_max={y_max:g} bins={bins[0]}x{bins[1]} (ball query "
                                    f"r={r_max:g}, image flavour) N={n} 2-D square L={L:.4f} areal density 0.5",
                        "bonds_per_step": n_bonds},
                h2d=12 * n + 4 * n, d2h=4 * nb, algo=algo, keep=[keep0, keep1], box=box, pts=pts, angles=angles,
                x_max=x_max, y_max=y_max, bins=bins, secondary={"bonds": n_bonds}, algo_per_step=True,
                # measured DRAM bytes of one launch (ncu, profiles/ncu_r1_v9_summary.md)
                traffic={"pmft3_rows": 705748736 + 4108544} if (n, x_max, y_max) == (1_000_000, 4.0, 3.0) else {})


def cpu_reference_pmftxy(box, pts, angles, x_max, y_max, bins, budget_s=12.0, threads=None):
    """The reference's PMFTXY (AABBQuery engine, all host threads) on a bounded sample of query points."""
    from oracle import ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        return {"value": None, "unit": "particles/s", "cores": threads, "kind": "port", "sample": "unavailable"}
    q = ref.Query("aabb", box, pts, is2d=True)
    probe = 20000
    t0 = time.perf_counter()
    ref.pmftxy(q, angles[:probe], pts[:probe], x_max, y_max, bins[0], bins[1], exclude_ii=True)
    dt = time.perf_counter() - t0
    m = int(min(len(pts), max(probe, probe * budget_s / max(dt, 1e-6) * 0.8)))
    t0 = time.perf_counter()
    ref.pmftxy(q, angles[:m], pts[:m], x_max, y_max, bins[0], bins[1], exclude_ii=True)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "particles/s", "cores": threads, "kind": "reference",
            "sample": f"reference PMFTXY({x_max:g}, {y_max:g}, {bins}).compute via AABBQuery: first {m} of {len(pts)} "
                      f"query points in {dt:.2f} s"}


HIST_CLIENTS = ("pmftxyz", "pmftxyt", "pmftr12", "bond_order")
FCC_SIGMA = 0.05  # noise on the FCC lattice of the bond_order workload (--sigma 0: the perfect lattice, every bond on a bin edge)


def hist_client_inputs(name, n, seed):
    """Synthetic input of the four histogram clients of SURVEY.md section 8f rank 3 that came last: the 2-D system of
    config 5 with random angles (PMFTXYT, PMFTR12), config 2's cubic system with random unit quaternions (PMFTXYZ), and
    config 3's noisy FCC lattice (BondOrder, mode bod, 12 nearest neighbours).  Shared by both bench arms."""
    from freud_b200 import data

    rs = np.random.RandomState(seed + 31)
    if name in ("pmftxyt", "pmftr12"):
        L = (n / 0.5) ** 0.5
        box, pts = data.make_random_system(L, n, is2D=True, seed=seed)
        orient = (rs.random_sample(n) * 2 * np.pi - np.pi).astype(np.float32)
        if name == "pmftxyt":
            spec = dict(kind=1, maxes=(4.0, 3.0), bins=(50, 50, 36), r_max=5.0,
                        label=f"PMFTXYT x_max=4 y_max=3 bins=50x50x36 (ball query r=5, image flavour) N={n} 2-D square "
                              f"L={L:.4f} areal density 0.5")
        else:
            spec = dict(kind=2, maxes=(5.0,), bins=(50, 36, 36), r_max=5.0,
                        label=f"PMFTR12 r_max=5 bins=50x36x36 (ball query r=5, image flavour) N={n} 2-D square L={L:.4f} "
                              f"areal density 0.5")
    elif name == "pmftxyz":
        L = (n / RHO) ** (1.0 / 3.0)
        box, pts = data.make_random_system(L, n, seed=seed)
        q = rs.normal(size=(n, 4))
        orient = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
        spec = dict(kind=0, maxes=(2.0, 2.0, 2.0), bins=(40, 40, 40), r_max=float(np.sqrt(12.0)),
                    label=f"PMFTXYZ x_max=y_max=z_max=2 bins=40^3, identity equivalent orientation (ball query r=3.464, "
                          f"image flavour) N={n} cubic L={L:.4f} rho=0.08")
    else:
        m = max(2, round((n / 4) ** (1.0 / 3.0)))
        box, pts = data.make_fcc_system(m, sigma_noise=FCC_SIGMA, seed=seed)
        orient = None
        spec = dict(kind=None, bins=(72, 36), k=12,
                    label=f"BondOrder bins=72x36 mode=bod num_neighbors=12 FCC {m}^3x4={len(pts)} sigma={FCC_SIGMA:g}")
    return box, pts, orient, spec


def workload_hist_client(ctx, rank, n, name):
    """One frame of PMFTXYZ / PMFTXYT / PMFTR12 / BondOrder through the C ABI: cell list, neighbour search into a device
    NeighborList, the histogram kernel over its bonds, bin counts back."""
    from freud_b200 import _capi

    box, pts, orient, spec = hist_client_inputs(name, n, rank)
    n = len(pts)
    dp = _capi.DevicePoints(ctx, box, pts)
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts
    keep = [keep0]
    if orient is not None:
        pin_o, keep1 = pinned_empty(orient.shape, np.float32)
        pin_o[:] = orient
        keep.append(keep1)
    if name == "bond_order":
        hist = _capi.DeviceBondOrder(ctx, spec["bins"][0], spec["bins"][1], "bod")
        query = lambda d: d.knn_query(None, spec["k"], exclude_ii=True)
        accumulate = lambda nl: hist.accumulate_nlist(nl)
        kernel, r_build = "bond_order", None
    else:
        hist = _capi.DevicePMFT(ctx, spec["kind"], spec["maxes"], spec["bins"])
        equiv = np.float32([[1, 0, 0, 0]])
        # the call behind PMFT*.compute(system, orientations) without a NeighborList handed in: query and histogram
        # in one call, the bonds go from the search's bag straight into the bins
        query = lambda d: d
        accumulate = lambda d: hist.accumulate(d, None, IMAGE, spec["r_max"], pin_o, pin_o, equiv, exclude_ii=True)
        kernel, r_build = "pmft3_rows", spec["r_max"]
    n_bonds = (dp.knn_query(None, spec["k"], exclude_ii=True) if name == "bond_order"
               else dp.ball_query(None, IMAGE, spec["r_max"], 0.0, True)).num_bonds

    def step_dev():
        hist.reset()
        if r_build is not None:
            dp.build_cells(r_build)
        accumulate(query(dp))
        return hist.read()

    def step_e2e():
        hist.reset()
        accumulate(query(_capi.DevicePoints(ctx, box, pin_pts)))
        return hist.read()

    step_dev()
    host_binned = int(hist.host_binned_bonds)  # bonds next to a bin edge, binned by the host's libm
    nb = int(np.prod(spec["bins"]))
    o_bytes = 0 if orient is None else orient.nbytes
    algo = {kernel: 20 * n_bonds + 2 * o_bytes + 4 * nb, "pipeline": 16 * (n + n) + 2 * o_bytes + 4 * nb,
            "search_nl": 16 * (n + n) + 16 * n_bonds + 8 * n, "emit": 16 * n_bonds + 16 * n + 12 * n + 4 * n + 28 * n_bonds,
            "knn_select": (16 * 26 + 12 + 28 * 12) * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=n, unit="particles/s", metric=f"{name}_particles_per_sec",
                config={"workload": spec["label"], "bonds_per_step": n_bonds, "host_binned_bonds_per_step": host_binned},
                h2d=12 * n + o_bytes, d2h=4 * nb, algo=algo,
                keep=keep, box=box, pts=pts, orient=orient, spec=spec, hist=hist, secondary={"bonds": n_bonds},
                algo_per_step=True,
                # measured DRAM bytes of one launch of the client's kernel (ncu, profiles/ncu_r1_v8_summary.md)
                traffic={"bond_order": {"bond_order": 240091648 + 5901312},
                         # ... and profiles/ncu_r1_v9_summary.md for the bag-reading kernels
                         "pmftr12": {"pmft3_rows": 869713664 + 4264960}, "pmftxyz": {"pmft3_rows": 301006592 + 4494592}}
                .get(name, {}) if n in (1_000_000, 1_000_188) else {})


def cpu_reference_hist_client(name, box, pts, orient, spec, budget_s=12.0, threads=None):
    """The reference's PMFTXYZ / PMFTXYT / PMFTR12 / BondOrder (AABBQuery engine, all host threads) on a bounded sample of
    query points."""
    from oracle import ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        return {"value": None, "unit": "particles/s", "cores": threads, "kind": "port", "sample": "unavailable"}
    q = ref.Query("aabb", box, pts, is2d=box.is2D)
    if name == "bond_order":
        ident = np.tile(np.float32([1, 0, 0, 0]), (len(pts), 1))
        run = lambda m: ref.bond_order("bod", q, ident, pts[:m], ident[:m], spec["bins"], mode="nearest",
                                       num_neighbors=spec["k"], exclude_ii=True)
    else:
        equiv = np.float32([[1, 0, 0, 0]]) if spec["kind"] == 0 else None
        run = lambda m: ref.pmft3(spec["kind"], q, None if spec["kind"] == 0 else orient, orient[:m], pts[:m],
                                  spec["maxes"], spec["bins"], equiv=equiv, r_max=spec["r_max"], exclude_ii=True)
    probe = 20000
    t0 = time.perf_counter()
    run(probe)
    dt = time.perf_counter() - t0
    m = int(min(len(pts), max(probe, probe * budget_s / max(dt, 1e-6) * 0.8)))
    t0 = time.perf_counter()
    run(m)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "particles/s", "cores": threads, "kind": "reference",
            "sample": f"reference {spec['label'].split(' N=')[0].split(' FCC')[0]} via AABBQuery: first {m} of {len(pts)} "
                      f"query points in {dt:.2f} s"}


def cpu_reference_correlation(box, pts, values, bins, r_max, budget_s=12.0, threads=None):
    """The reference's CorrelationFunction (AABBQuery engine, all host threads) on a bounded sample of query points."""
    from oracle import ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        return {"value": None, "unit": "particles/s", "cores": threads, "kind": "port", "sample": "unavailable"}
    q = ref.Query("aabb", box, pts)
    probe = 20000
    t0 = time.perf_counter()
    ref.correlation_function(q, values, pts[:probe], values[:probe], bins, r_max, exclude_ii=True)
    dt = time.perf_counter() - t0
    m = int(min(len(pts), max(probe, probe * budget_s / max(dt, 1e-6) * 0.8)))
    t0 = time.perf_counter()
    ref.correlation_function(q, values, pts[:m], values[:m], bins, r_max, exclude_ii=True)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "particles/s", "cores": threads, "kind": "reference",
            "sample": f"reference CorrelationFunction({bins}, {r_max:g}).compute via AABBQuery: first {m} of {len(pts)} "
                      f"query points in {dt:.2f} s"}


def cpu_reference_local_density(box, pts, r_max, diameter, budget_s=12.0, threads=None):
    """The reference's LocalDensity::compute (AABBQuery engine, all host threads) on a bounded sample of query points."""
    from oracle import ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        return {"value": None, "unit": "particles/s", "cores": threads, "kind": "port", "sample": "unavailable"}
    q = ref.Query("aabb", box, pts)
    probe = 20000
    t0 = time.perf_counter()
    ref.local_density(q, pts[:probe], r_max, diameter, exclude_ii=True)
    dt = time.perf_counter() - t0
    m = int(min(len(pts), max(probe, probe * budget_s / max(dt, 1e-6) * 0.8)))
    t0 = time.perf_counter()
    ref.local_density(q, pts[:m], r_max, diameter, exclude_ii=True)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "particles/s", "cores": threads, "kind": "reference",
            "sample": f"reference LocalDensity({r_max:g}, {diameter:g}).compute via AABBQuery: first {m} of {len(pts)} "
                      f"query points in {dt:.2f} s"}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_nl(box, pts, r_max, budget_s=12.0, threads=None):
    """The reference's NeighborList query on this box's host cores, bounded samples of the query points.

    `value` is its practical engine, AABBQuery (what `(box, points)` systems get upstream): pair evaluations of the
    27-cell scheme that the sampled query points stand for, per second, tree build excluded.  The engine BASELINE names,
    LinkCell, is timed beside it as `linkcell_quadratic`: unmodified, it deep-copies the whole cell list per visited
    cell, so it is a footnote, not a comparator."""
    from oracle import port, ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    out = {"cores": threads, "kind": "reference" if ref.available() else "port"}
    if not ref.available():
        t0 = time.perf_counter()
        nl = port.ball_nlist(port.WRAP, box, box.is2D, pts, pts, r_max, 0.0, True)
        dt = time.perf_counter() - t0
        ev = port.count_candidates(box, box.is2D, pts, pts, r_max)
        out.update(value=ev / dt, unit="pair_evals/s", sample=f"oracle port, all {len(pts)} query points, {dt:.2f} s",
                   bonds_per_sec=len(nl) / dt)
        return out
    # practical engine: AABBQuery on a sample sized to the budget
    t0 = time.perf_counter()
    qa = ref.Query("aabb", box, pts)
    build_a = time.perf_counter() - t0
    ma = min(len(pts), 20000)
    t0 = time.perf_counter()
    qa.nlist(pts[:ma], mode="ball", r_max=r_max, exclude_ii=True)
    probe = time.perf_counter() - t0
    ma2 = int(min(len(pts), max(ma, ma * 0.7 * budget_s / max(probe, 1e-6))))
    t0 = time.perf_counter()
    nla = qa.nlist(pts[:ma2], mode="ball", r_max=r_max, exclude_ii=True)
    dta = time.perf_counter() - t0
    eva = port.count_candidates(box, box.is2D, pts, pts[:ma2], r_max)
    out.update(value=eva / dta, unit="pair_evals/s", bonds_per_sec=len(nla) / dta, engine="AABBQuery",
               sample=f"reference AABBQuery.query(ball r_max={r_max:g}).toNeighborList() N={len(pts)}: first {ma2} query "
                      f"points in {dta:.2f} s (+{build_a:.2f} s tree build, not counted); pair evals = the 27-cell "
                      f"candidates those query points have")
    del qa
    # the engine the config names, a few thousand query points of it
    t0 = time.perf_counter()
    q = ref.Query("linkcell", box, pts, cell_width=r_max)
    build_s = time.perf_counter() - t0
    m = min(len(pts), 8 * threads)
    t0 = time.perf_counter()
    q.nlist(pts[:m], mode="ball", r_max=r_max, exclude_ii=True)
    probe = time.perf_counter() - t0
    m2 = int(min(len(pts), max(m, m * 0.3 * budget_s / max(probe, 1e-6))))
    t0 = time.perf_counter()
    nl = q.nlist(pts[:m2], mode="ball", r_max=r_max, exclude_ii=True)
    dt = time.perf_counter() - t0
    ev = port.count_candidates(box, box.is2D, pts, pts[:m2], r_max)
    out["linkcell_quadratic"] = {"value": ev / dt, "unit": "pair_evals/s", "bonds_per_sec": len(nl) / dt,
                                 "sample": f"reference LinkCell(cell_width={r_max:g}) N={len(pts)}: first {m2} query "
                                           f"points in {dt:.2f} s (+{build_s:.2f} s serial build, not counted)"}
    return out


def cpu_reference_rdf(box, pts, bins, r_max, budget_s=12.0, threads=None):
    from oracle import port, ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        t0 = time.perf_counter()
        port.rdf_accumulate(port.IMAGE, box, box.is2D, pts, pts, bins, r_max, 0.0, True)
        dt = time.perf_counter() - t0
        return {"value": 1.0 / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                "sample": f"oracle port full frame {dt:.2f} s"}
    q = ref.Query("raw", box, pts, is2d=box.is2D)
    m = min(len(pts), 20000)
    R = ref.RDF(bins, r_max)
    t0 = time.perf_counter()
    R.accumulate(q, pts[:m], mode="ball", r_max=r_max, exclude_ii=True)
    probe = time.perf_counter() - t0
    m2 = int(min(len(pts), max(m, m * budget_s / max(probe, 1e-6))))
    R = ref.RDF(bins, r_max)
    t0 = time.perf_counter()
    R.accumulate(q, pts[:m2], mode="ball", r_max=r_max, exclude_ii=True)
    R.results()
    dt = time.perf_counter() - t0
    return {"value": (m2 / len(pts)) / dt, "unit": "frames/s", "cores": threads, "kind": "reference",
            "sample": f"reference RDF.accumulate via RawPoints/AABBQuery: first {m2} of {len(pts)} query points in "
                      f"{dt:.2f} s, scaled to a full frame"}


def cpu_reference_q6(box, pts, budget_s=12.0, threads=None):
    from oracle import ref

    threads = threads or os.cpu_count()
    ref.set_num_threads(threads)
    if not ref.available():
        return {"value": None, "unit": "particles/s", "cores": threads, "kind": "port", "sample": "unavailable"}
    q = ref.Query("raw", box, pts)
    S = ref.Steinhardt(6)
    t0 = time.perf_counter()
    S.compute(q, num_neighbors=12, exclude_ii=True)
    dt = time.perf_counter() - t0
    return {"value": len(pts) / dt, "unit": "particles/s", "cores": threads, "kind": "reference",
            "sample": f"reference Steinhardt(6).compute k=12 on all {len(pts)} particles in {dt:.2f} s"}


# ---------------------------------------------------------------------------------------------------------
# parity: the leg's result at its full size against the oracle, outside the timed regions (rank 0 only)
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def oracle_threads(n=None):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the oracle's OpenMP loops (oracle/port.c) are the checker here,
    not the thing measured, and get the host's cores back for the duration of a check."""
    import ctypes

    n = n or os.cpu_count() or 1
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def parity_nl(w, flavour):
    """All five NeighborList arrays + segments / counts of the benchmark's own frame, bit for bit against the oracle's
    restatement of LinkCell / AABBQuery (oracle/port.c: grid + exact per-pair arithmetic; pinned to the compiled
    reference in tests/test_oracle_port.py)."""
    from oracle import port

    oracle_threads()
    t0 = time.perf_counter()
    got = w["dp"].ball_query(None, flavour, w["r_max"], 0.0, True).to_host()
    want = port.ball_nlist(port.WRAP if flavour == WRAP else port.IMAGE, w["box"], w["box"].is2D, w["pts"], w["pts"],
                           w["r_max"], 0.0, True)
    same = {k: bool(np.array_equal(_bits(got[k]), _bits(getattr(want, k))))
            for k in ("neighbors", "distances", "weights", "vectors", "segments", "counts")}
    return {"oracle": "port (oracle/port.c fport_ball_nlist)", "n_bonds": int(len(want)), "n_bonds_gpu": int(len(got["distances"])),
            "arrays": same, "bitwise_equal": all(same.values()), "seconds": round(time.perf_counter() - t0, 2)}


def parity_rdf_counts(got, w, bins, frames=None):
    """Raw bin counts against the oracle (u32, bit-exact).  frames: list of (box, points) accumulated (default: the
    leg's own frame)."""
    from oracle import port

    oracle_threads()
    t0 = time.perf_counter()
    want = np.zeros(bins, np.uint32)
    for box, pts in (frames or [(w["box"], w["pts"])]):
        want = port.rdf_accumulate(port.IMAGE if w.get("flavour", IMAGE) == IMAGE else port.WRAP, box, box.is2D, pts, pts,
                                   bins, w["r_max"], 0.0, True, counts=want)
    return {"oracle": "port (oracle/port.c fport_rdf_accumulate)", "n_bonds": int(want.astype(np.uint64).sum()),
            "n_bonds_gpu": int(np.asarray(got).astype(np.uint64).sum()),
            "bitwise_equal": bool(np.array_equal(np.asarray(got, dtype=np.uint32), want)),
            "seconds": round(time.perf_counter() - t0, 2)}


def parity_q6(w):
    """q_l of every particle against the reference's own Steinhardt on its own AABB kNN list (oracle/_ref), tolerance
    1e-5 relative as north_star states; the kNN NeighborList itself bit for bit."""
    from oracle import ref

    if not ref.available():
        return {"oracle": "unavailable (oracle/_ref missing)", "bitwise_equal": None}
    t0 = time.perf_counter()
    ref.set_num_threads(os.cpu_count())
    dp = w["dp"]
    nl = dp.knn_query(None, 12, exclude_ii=True)
    got = dp.steinhardt_knn(12, [6], exclude_ii=True)["ql"][:, 0]  # the fused route the timed step takes
    got_list = dp.steinhardt(nl, [6], want_qlm=False)["ql"][:, 0]   # ... and the two-kernel route over the list
    got_nl = nl.to_host()
    q = ref.Query("aabb", w["box"], w["pts"])
    want_nl = q.nlist(w["pts"], mode="nearest", num_neighbors=12, exclude_ii=True)
    want = ref.Steinhardt(6).compute(q, nlist=want_nl)["ql"][:, 0]
    same = {k: bool(np.array_equal(_bits(got_nl[k]), _bits(getattr(want_nl, k))))
            for k in ("neighbors", "distances", "vectors", "segments", "counts")}
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    rel_list = np.abs(got_list - want) / np.maximum(np.abs(want), 1e-30)
    return {"oracle": "reference (oracle/_ref: AABBQuery kNN + Steinhardt::compute)", "n_particles": int(len(want)),
            "ql_max_rel_diff": float(rel.max()), "ql_max_rel_diff_list_route": float(rel_list.max()),
            "ql_tolerance_rel": 1e-5, "ql_within_tolerance": bool(rel.max() <= 1e-5 and rel_list.max() <= 1e-5),
            "knn_nlist_arrays": same, "bitwise_equal": all(same.values()),
            "seconds": round(time.perf_counter() - t0, 2)}


def parity_q6_rows(w, got):
    """Sharded Q6: rank 0's rows of q_l and the all-reduced system order parameter against the reference's Steinhardt
    over the whole frame (oracle/_ref), tolerance 1e-5 relative."""
    from oracle import ref

    if not ref.available():
        return {"oracle": "unavailable (oracle/_ref missing)", "bitwise_equal": None}
    t0 = time.perf_counter()
    ref.set_num_threads(os.cpu_count())
    q = ref.Query("aabb", w["box"], w["pts"])
    want_nl = q.nlist(w["pts"], mode="nearest", num_neighbors=12, exclude_ii=True)
    res = ref.Steinhardt(6).compute(q, nlist=want_nl)
    lo, hi = w["block"]
    want = res["ql"][lo:hi, 0]
    rel = np.abs(got["ql"][:, 0] - want) / np.maximum(np.abs(want), 1e-30)
    order_rel = abs(float(got["order"][0]) - float(res["order"][0])) / max(abs(float(res["order"][0])), 1e-30)
    ok = bool(rel.max() <= 1e-5 and order_rel <= 1e-5)
    return {"oracle": "reference (oracle/_ref: AABBQuery kNN + Steinhardt::compute over the whole frame)",
            "rows_checked": [int(lo), int(hi)], "ql_max_rel_diff": float(rel.max()),
            "system_order_rel_diff_after_allreduce": order_rel, "ql_tolerance_rel": 1e-5, "ql_within_tolerance": ok,
            "bitwise_equal": None, "seconds": round(time.perf_counter() - t0, 2)}


# ---------------------------------------------------------------------------------------------------------
# e2e through the drop-in classes (freud_b200.locality / density / order): the call a freud user makes
def api_step(name, w):
    import freud_b200 as fr

    box, pts = w["box"], w["pts"]
    if name in ("nl", "nl_image"):
        cls = fr.locality.LinkCell if name == "nl" else fr.locality.AABBQuery

        def step():
            nq = cls(box, pts, w["r_max"]) if name == "nl" else cls(box, pts)
            nl = nq.query(pts, dict(r_max=w["r_max"], exclude_ii=True)).toNeighborList()
            # touch every array the user can ask for (each is a D2H on first access)
            return (nl.query_point_indices[-1], nl.point_indices[-1], nl.distances[-1], nl.weights[-1],
                    nl.vectors[-1, 2], nl.segments[-1], nl.neighbor_counts[-1])
        return step
    if name in ("rdf", "rdf_wrap"):
        rdf = fr.density.RDF(w["bins"], w["r_max"])

        def step():
            system = (box, pts) if name == "rdf" else fr.locality.LinkCell(box, pts, w["r_max"])
            return rdf.compute(system).bin_counts[-1]
        return step
    if name == "q6":
        st = fr.order.Steinhardt(6)

        def step():
            return st.compute((box, pts), neighbors=dict(num_neighbors=12)).particle_order[-1]
        return step
    return None


def time_api(step, steps):
    if step is None:
        return None
    for _ in range(2):
        step()
    marks = [time.perf_counter()]
    for _ in range(steps):
        step()
        marks.append(time.perf_counter())
    d = np.diff(marks) * 1e3
    return {"ms_per_step": float(d.mean()), "ms_per_step_min_median_max": [round(float(x), 3) for x in (d.min(), np.median(d), d.max())]}


# ---------------------------------------------------------------------------------------------------------
WORKLOADS = ["nl", "nl_image", "rdf", "rdf_wrap", "q6", "q6s", "rdf4m", "traj2d", "local_density", "correlation", "pmftxy",
             "pmftxyz", "pmftxyt", "pmftr12", "bond_order"]
N_DEFAULT = {"nl": 1_000_000, "nl_image": 1_000_000, "rdf": 1_000_000, "rdf_wrap": 1_000_000, "q6": 1_000_188, "q6s": 1_000_188,
             "rdf4m": 4_000_000, "traj2d": 1_000_000, "local_density": 1_000_000, "correlation": 1_000_000,
             "pmftxy": 1_000_000, "pmftxyz": 1_000_000, "pmftxyt": 1_000_000, "pmftr12": 1_000_000,
             "bond_order": 1_000_188}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=WORKLOADS,
                    help="one workload only; default: the three legs of BASELINE.json's metric (nl, rdf, q6) on one GPU, "
                         "the sharded 4 M-point RDF (+ 2-D trajectory, + NeighborList replicas) on several")
    ap.add_argument("--n", type=int, default=None, help="override the particle count (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="rdf4m: launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--sigma", type=float, default=None,
                    help="bond_order: noise on the FCC lattice (0 = the perfect lattice: every bond sits on a bin edge and is "
                         "binned by the host -- the cliff DESIGN.md section 8 names)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="experiment hook: fgpu_ctx_set_tuning(key, value), e.g. span=2 or lanes_over_queries=1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.sigma is not None:
        global FCC_SIGMA
        FCC_SIGMA = args.sigma

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is not None:
        legs = [args.workload]
    elif world == 1:
        # BASELINE.json metric: pair evals/s + RDF frames/s @1M r_max=5 + Q6 particles/s; and configs[3] on this one GPU,
        # the N = 1 point of the strong-scaling curve whose N > 1 points are the headline of the multi-GPU runs
        legs = ["nl", "rdf", "q6", "rdf4m"]
    else:
        # configs[3] (strong scaling, the headline), configs[4], NeighborList replicas, configs[2] with its rows sharded
        legs = ["rdf4m", "traj2d", "nl", "q6s"]

    if args.impl == "reference":
        return run_reference_arm(args, rank, world, legs)

    import torch
    import torch.distributed as dist

    from freud_b200 import _capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    ctx = _capi.Context(local_rank)
    for kv in args.tune:
        key, val = kv.split("=")
        ctx.set_tuning(key, int(val))
    env = dict(torch=torch, dist=dist, ctx=ctx, rank=rank, world=world, local_rank=local_rank,
               stream=torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank)),
               flush=torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda"), comm=None)
    if world > 1:
        from freud_b200 import parallel

        env["comm"] = parallel.make_communicator(ctx)  # NCCL on the library's stream; id travels over torch.distributed

    lines = []
    for k, name in enumerate(legs):
        steps = args.steps if k == 0 or world == 1 else min(args.steps, 5)  # extras of a multi-GPU run stay short
        lines.append(run_leg(name, args, env, steps, args.n or N_DEFAULT[name]))
        ctx.trim()
    if rank == 0:
        line = lines[0]
        if len(lines) > 1:
            line["legs"] = {name: ln for name, ln in zip(legs[1:], lines[1:])}
            for name, ln in zip(legs[1:], lines[1:]):
                key = ln["metric"] if ln["metric"] != line["metric"] and ln["metric"] not in line else f"{name}_{ln['metric']}"
                line[key] = ln["value"]
            line["parity_all_legs"] = all((ln.get("parity") or {}).get("bitwise_equal") is not False
                                          and (ln.get("parity") or {}).get("ql_within_tolerance") is not False
                                          for ln in lines)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_leg(name, args, env, steps, n):
    torch, dist, ctx = env["torch"], env["dist"], env["ctx"]
    rank, world, local_rank, stream, flush, comm = (env[k] for k in ("rank", "world", "local_rank", "stream", "flush", "comm"))
    if name in ("nl", "nl_image"):
        w = workload_nl(ctx, rank, n, WRAP if name == "nl" else IMAGE)
        scaling = "weak"
    elif name in ("rdf", "rdf_wrap"):
        w = workload_rdf(ctx, rank, n, flavour=IMAGE if name == "rdf" else WRAP)
        scaling = "weak"
    elif name == "q6":
        w = workload_q6(ctx, rank, n)
        scaling = "weak"
    elif name == "q6s":
        w = workload_q6_sharded(ctx, rank, world, n, comm)
        scaling = "strong"
    elif name == "local_density":
        w = workload_local_density(ctx, rank, n)
        scaling = "weak"
    elif name == "correlation":
        w = workload_correlation(ctx, rank, n)
        scaling = "weak"
    elif name == "pmftxy":
        w = workload_pmftxy(ctx, rank, n)
        scaling = "weak"
    elif name in HIST_CLIENTS:
        w = workload_hist_client(ctx, rank, n, name)
        scaling = "weak"
    elif name == "rdf4m":
        w = workload_rdf4m(ctx, rank, world, n, comm)
        scaling = "strong"
    else:
        w = workload_traj2d(ctx, rank, world, n, comm)
        scaling = "weak"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ("value") ----------------------------------------------------------------
    # The strong-scaling legs (configs[3], configs[4]) are timed WITHOUT the per-kernel event pairs: at 8 GPUs a step is
    # ~0.17 ms of eight kernels, and two extra event records per kernel are a measurable share of that; their kernel
    # breakdown comes from a separate short pass right after (same inputs, same stream).  Every other leg keeps the
    # event pairs inside the timed region.
    events_in_timed_region = name not in ("rdf4m", "traj2d")
    for _ in range(args.warmup):
        w["step_dev"]()
    barrier()
    # configs[3]: a step is seven short kernels; its build -> search -> reduce sequence is captured once into a CUDA graph
    # on the library's stream and replayed (same kernels, same arguments -- the reduction's epoch lives in device
    # memory for exactly this reason), which takes the host's launch path and most inter-kernel gaps out of a 0.2 ms step
    step_fn, graph_note, launches_per_graph = w["step_dev"], None, None
    if name == "rdf4m" and not args.no_graph:
        try:
            l0 = ctx.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                w["step_dev"]()
            launches_per_graph = ctx.launch_count - l0

            def step_fn():
                with torch.cuda.stream(stream):
                    g.replay()
            for _ in range(2):
                step_fn()
            graph_note = "CUDA graph of one step (captured on the library's stream), replayed"
        except Exception as exc:  # capture is an optimisation of the harness, never a reason to lose the line
            step_fn, launches_per_graph = w["step_dev"], None
            graph_note = f"eager (graph capture failed: {type(exc).__name__}: {exc})"
            torch.cuda.synchronize()
        barrier()
    ctx.profile(events_in_timed_region)
    ctx.kernel_time(reset=True)
    launches0 = ctx.launch_count
    events = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()  # L2 flush between steps (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = step_fn()
            e1.record(stream)
            events.append((e0, e1))
            del out
        barrier()
        t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in events)
    launches = launches_per_graph * steps if launches_per_graph is not None else ctx.launch_count - launches0
    breakdown_steps, breakdown_ms = steps, dev_ms
    if not events_in_timed_region:
        breakdown_steps = min(steps, 5)
        ctx.profile(True)
        ctx.kernel_time(reset=True)
        barrier()
        ev = []
        for _ in range(breakdown_steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            w["step_dev"]()
            e1.record(stream)
            ev.append((e0, e1))
        barrier()
        breakdown_ms = sum(a.elapsed_time(b) for a, b in ev)
    ctx.profile(False)
    # dominant kernel over the timed region
    per_kernel = {}
    names = ("cell_prep", "cell_assign", "cell_scatter", "scan", "search_nl", "search_count", "search_fill", "search_rdf_general",
             "search_rdf", "emit_general", "emit", "segments", "knn_emit", "knn_rows", "knn_select", "knn_ylm", "knn",
             "rdf_distances", "rdf_wait", "rdf_push", "steinhardt", "local_density_rows", "local_density", "correlation_rows", "correlation", "pmftxy", "pmft3_rows", "pmft3", "bond_order", "pmft_add_bins", "pmft_add_hist")
    raw = {nm: ctx.kernel_time(nm) for nm in names}  # prefix match: subtract the longer names
    for nm in names:
        ms, cnt = raw[nm]
        for other in names:
            if other != nm and other.startswith(nm):
                ms, cnt = ms - raw[other][0], cnt - raw[other][1]
        if cnt:
            per_kernel[nm] = (ms, cnt)
    ctx.kernel_time(reset=True)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    ms_per_step = dev_ms_max / steps
    value = w["units"] * world * steps / (dev_ms_max / 1e3)

    # ---- end-to-end leg ("e2e"): host buffers in, host result out, through the C ABI --------------------
    for _ in range(2):
        w["step_e2e"]()
    barrier()
    t0 = time.perf_counter()
    e2e_marks = [t0]
    for _ in range(steps):
        w["step_e2e"]()  # returns once the step's result is on the host
        e2e_marks.append(time.perf_counter())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_steps_ms = np.diff(e2e_marks) * 1e3
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = w["units"] * world * steps / e2e_s

    # ---- the same through the drop-in classes (single GPU legs; pageable numpy in, numpy views out) ------
    api = None
    if world == 1:
        api = time_api(api_step(name, w), max(3, min(steps, 10)))
        if api is not None:
            api["value"] = w["units"] / (api["ms_per_step"] * 1e-3)
            api["unit"] = w["unit"]
            api["vs_capi_e2e"] = round((e2e_s * 1e3 / steps) / api["ms_per_step"], 3)
            api["call"] = {"nl": "LinkCell(box, pts, 3).query(pts, dict(r_max=3, exclude_ii=True)).toNeighborList() + every array",
                           "nl_image": "AABBQuery(box, pts).query(...).toNeighborList() + every array",
                           "rdf": "RDF(100, 5).compute((box, pts)).bin_counts", "rdf_wrap": "RDF(100, 5).compute(LinkCell(box, pts, 5)).bin_counts",
                           "q6": "Steinhardt(6).compute((box, pts), neighbors=dict(num_neighbors=12)).particle_order"}.get(name)

    # ---- parity at the benchmark's own size (outside every timed region) ---------------------------------
    parity = None
    if not args.no_parity:
        try:
            if name in ("nl", "nl_image") and rank == 0:
                parity = parity_nl(w, WRAP if name == "nl" else IMAGE)
            elif name in ("rdf", "rdf_wrap") and rank == 0:
                w["rdf"].reset()
                w["rdf"].accumulate(w["dp"], None, w["flavour"], w["r_max"], 0.0, True)
                parity = parity_rdf_counts(w["rdf"].read(), w, w["bins"])
            elif name == "q6" and rank == 0:
                parity = parity_q6(w)
            elif name == "q6s":
                got = w["rows"](w["dp"])  # collective: every rank takes part in the allreduce
                if rank == 0:
                    parity = parity_q6_rows(w, got)
            elif name == "rdf4m":
                w["step_dev"]()  # every rank takes part in the reduction
                got = w["read_global"]()
                if rank == 0:
                    parity = parity_rdf_counts(got, w, w["bins"])
            elif name == "traj2d":
                # (a) rank 0's own frames against the oracle (the host's cores go to one check, not to `world` of them);
                # (b) the reduced histogram of the step against the sum of every rank's own histogram, all on the GPU
                mine = w["rank_counts"]()  # this rank's own frames, before any reduction
                t = torch.from_numpy(mine.astype(np.int64)).cuda()
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                w["step_dev"]()
                reduced = w["rdf"].read()
                sums_agree = bool(np.array_equal(reduced, (t.cpu().numpy() & 0xFFFFFFFF).astype(np.uint32)))
                if rank == 0:
                    parity = parity_rdf_counts(mine, w, w["bins"], frames=w["frames"])
                    parity["reduced_equals_sum_of_rank_histograms"] = sums_agree
                    parity["bitwise_equal"] = parity["bitwise_equal"] and sums_agree
                    parity["scope"] = (f"rank 0's {len(w['frames'])} frames against the oracle; the step's reduced histogram "
                                       f"against the sum of all {world} ranks' own histograms")
        except Exception as exc:  # a failed check is reported, never hidden
            parity = {"bitwise_equal": False, "error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        return None

    peak, peak_src = peaks()
    roofline = None
    if per_kernel:
        kname = max(per_kernel, key=lambda k: per_kernel[k][0])
        ms, cnt = per_kernel[kname]
        avg_ms = ms / cnt
        algo = w["algo"].get(kname)
        if algo:
            # w["algo"] holds bytes per step; a kernel launched in several chunks per step moves its share per launch
            algo = algo * breakdown_steps / cnt if cnt > breakdown_steps and w.get("algo_per_step") else algo
            achieved = algo / (avg_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": w.get("traffic", {}).get(kname),
                        "traffic_source": w.get("traffic_source") if w.get("traffic", {}).get(kname) else None,
                        "avg_launch_ms": round(avg_ms, 4),
                        "algorithmic_bytes_per_launch": int(algo), "peak_source": peak_src,
                        "share_of_step": round(ms / breakdown_ms, 3),
                        "kernel_ms_per_step": {k: round(v[0] / breakdown_steps, 4) for k, v in per_kernel.items()},
                        "kernel_events": "inside the timed region" if events_in_timed_region
                        else f"separate pass of {breakdown_steps} steps after the timed region (the timed steps carry no "
                             f"per-kernel events)"}
    pipe = w["algo"]["pipeline"] / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(w["config"], l2="256 MB memset between timed steps (outside the event pairs)",
                       timing="CUDA events on the library's stream, max over ranks",
                       **({"launch": graph_note} if graph_note else {})),
        "e2e": {"value": e2e_value, "unit": w["unit"], "h2d_bytes_per_step": int(w["h2d"]),
                "d2h_bytes_per_step": int(w["d2h"]), "ms_per_step": e2e_s * 1e3 / steps,
                "ms_per_step_min_median_max": [round(float(x), 3) for x in (e2e_steps_ms.min(), np.median(e2e_steps_ms),
                                                                            e2e_steps_ms.max())],
                "host_buffers": "pinned", "through": "C ABI (ctypes)"},
        "e2e_api": api,
        "parity": parity,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "roofline_step": {"bound": "hbm", "achieved": round(pipe, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(pipe / peak, 4),
                          "note": "whole step: SURVEY.md 8d algorithmic bytes / ms_per_step"},
        "wall_s": t_wall,
    }
    if "bonds" in w["secondary"]:
        line["bonds_per_sec"] = w["secondary"]["bonds"] * world * steps / (dev_ms_max / 1e3)
    if "pair_evals_per_sec_factor" in w["secondary"]:
        line["pair_evals_per_sec"] = w["secondary"]["pair_evals_per_sec_factor"] * world * steps / (dev_ms_max / 1e3)
    if world == 1 and not args.no_cpu_baseline:
        try:
            if name in ("nl", "nl_image"):
                line["cpu_baseline"] = cpu_reference_nl(w["box"], w["pts"], w["r_max"])
            elif name in ("rdf", "rdf_wrap", "rdf4m", "traj2d"):
                line["cpu_baseline"] = cpu_reference_rdf(w["box"], w["pts"], w["rdf"].bins, w["r_max"])
            elif name == "local_density":
                line["cpu_baseline"] = cpu_reference_local_density(w["box"], w["pts"], w["r_max"], w["diameter"])
            elif name == "correlation":
                line["cpu_baseline"] = cpu_reference_correlation(w["box"], w["pts"], w["values"], w["bins"], w["r_max"])
            elif name == "pmftxy":
                line["cpu_baseline"] = cpu_reference_pmftxy(w["box"], w["pts"], w["angles"], w["x_max"], w["y_max"],
                                                            w["bins"])
            elif name in HIST_CLIENTS:
                line["cpu_baseline"] = cpu_reference_hist_client(name, w["box"], w["pts"], w["orient"], w["spec"])
            else:
                line["cpu_baseline"] = cpu_reference_q6(w["box"], w["pts"])
        except Exception as exc:  # the baseline is a report, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
    return line


def workload_rdf4m(ctx, rank, world, n, comm):
    """BASELINE.json configs[3]: one 4 M-point triclinic frame, query points (home tiles) sharded over the ranks, the
    histograms summed once per frame.  Strong scaling: value = frames/s of the ONE frame."""
    from freud_b200 import _capi, data, parallel

    bins, r_max = 500, 5.0
    L = (n / RHO) ** (1.0 / 3.0)
    box, pts = data.make_random_system(L, n, seed=0, tilt=(0.3, 0.2, 0.1))
    dp = _capi.DevicePoints(ctx, box, pts)
    srdf = parallel.ShardedRDF(ctx, bins, r_max, comm=comm, rank=rank, world=world)
    srdf.keep_shard = True
    rdf = srdf.rdf
    pin_pts, keep0 = pinned_empty((n, 3), np.float32)
    pin_pts[:] = pts

    # the home tiles of the self query are dealt to the ranks; each rank builds the slab of the cell list its
    # tiles see (fgpu_points_set_shard), so neither the search nor the build is replicated work
    shard_arg = None if world == 1 else "tiles"
    if world > 1:
        dp.set_shard(rank, world)

    def step_dev():
        srdf.reset()
        dp.build_cells(r_max)
        # search + exchange: every rank ends the step holding the frame's summed counts on the device
        srdf.accumulate_frame(dp, IMAGE, r_max, 0.0, True, query_shard=shard_arg, reduce=True)

    def step_e2e():
        srdf.reset()
        d = parallel.replicated_points(ctx, box, pin_pts, comm, rank, world)  # each rank uploads 1/world, NVLink does the rest
        srdf.accumulate_frame(d, IMAGE, r_max, 0.0, True, query_shard=shard_arg, reduce=True)
        return srdf.bin_counts()  # D2H of the sum

    step_dev()
    n_bonds = int(srdf.bin_counts().astype(np.uint64).sum())  # the whole frame
    n_cells = int(np.prod(dp.build_cells(r_max)))
    nq = n // world
    algo = {"search_rdf": 16 * (n + nq) + 4 * n_cells + 4 * bins, "pipeline": 16 * (n + nq) + 4 * bins,
            "cell_assign": 20 * n + 4 * n_cells, "cell_scatter": 36 * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=1.0 / world, unit="frames/s",
                metric="rdf_frames_per_sec",
                config={"workload": f"RDF bins=500 r_max=5 N={n} triclinic (xy=.3,xz=.2,yz=.1) L={L:.4f} query points "
                                    f"(home tiles) sharded over {world} GPU(s), points replicated, cell list built "
                                    f"per slab, histograms summed over the ranks every frame ({srdf.reduce_kind()})",
                        "bonds_per_step": n_bonds},
                h2d=12 * n // world, d2h=4 * bins, algo=algo, keep=[keep0], box=box, pts=pts, r_max=r_max,
                rdf=rdf, dp=dp, bins=bins, flavour=IMAGE, read_global=srdf.bin_counts, secondary={})


def workload_traj2d(ctx, rank, world, n, comm, frames_per_rank=8):
    """BASELINE.json configs[4]: 2-D trajectory, reset=False accumulation, frames sharded across ranks."""
    from freud_b200 import _capi, data

    bins, r_max = 100, 5.0
    L = (n / 0.5) ** 0.5
    frames = [data.make_random_system(L, n, is2D=True, seed=rank * frames_per_rank + f) for f in range(frames_per_rank)]
    box = frames[0][0]
    pins, keep = [], []
    for _, p in frames:
        a, t = pinned_empty((n, 3), np.float32)
        a[:] = p
        pins.append(a)
        keep.append(t)
    rdf = _capi.DeviceRDF(ctx, bins, r_max)
    dps = [_capi.DevicePoints(ctx, box, p) for _, p in frames]

    def step_dev():
        rdf.reset()
        for d in dps:
            d.build_cells(r_max)
            rdf.accumulate(d, None, IMAGE, r_max, 0.0, True)
        if comm is not None:
            rdf.allreduce(comm)

    def step_e2e():
        rdf.reset()
        for a in pins:
            d = _capi.DevicePoints(ctx, box, a)
            rdf.accumulate(d, None, IMAGE, r_max, 0.0, True)
        if comm is not None:
            rdf.allreduce(comm)
        return rdf.read()

    def rank_counts():
        rdf.reset()
        for d in dps:
            rdf.accumulate(d, None, IMAGE, r_max, 0.0, True)
        return rdf.read()  # no reduction since the reset: this rank's own frames

    algo = {"search_rdf": 16 * 2 * n + 4 * bins, "pipeline": frames_per_rank * (16 * 2 * n + 4 * bins),
            "cell_assign": 20 * n, "cell_scatter": 36 * n}
    return dict(step_dev=step_dev, step_e2e=step_e2e, units=frames_per_rank, unit="frames/s",
                metric="rdf_frames_per_sec",
                config={"workload": f"trajectory RDF bins=100 r_max=5 reset=False, {frames_per_rank} frames/GPU of N={n} "
                                    f"2-D square L={L:.4f}, frames sharded over {world} GPU(s), one allreduce at the end"},
                h2d=12 * n * frames_per_rank, d2h=4 * bins, algo=algo, keep=keep, box=box, pts=frames[0][1],
                r_max=r_max, rdf=rdf, dp=dps[0], bins=bins, flavour=IMAGE, frames=frames, rank_counts=rank_counts,
                secondary={})


def reference_leg(name, n, steps, budget, threads):
    """One leg of --impl reference: the reference's own CPU implementation (oracle/_ref, every host thread) on a bounded
    sample of the leg's workload; same metric / unit / config as the b200 arm's leg."""
    from freud_b200 import data

    def rep(fn):
        """`steps` bounded samples, each with the wall time it took (the line's ms_per_step)"""
        out = []
        for _ in range(steps):
            t0 = time.perf_counter()
            r = fn()
            r["wall_s"] = time.perf_counter() - t0
            out.append(r)
        return out

    if name in ("nl", "nl_image"):
        L = (n / RHO) ** (1.0 / 3.0)
        box, pts = data.make_random_system(L, n, seed=0)
        runs = rep(lambda: cpu_reference_nl(box, pts, 3.0, budget_s=budget, threads=threads))
        metric, unit = "neighbour_pair_evals_per_sec", "pair_evals/s"
        config = {"workload": f"LinkCell NeighborList r_max=3 exclude_ii N={n} cubic L={L:.4f} rho=0.08 flavour=wrap"}
    elif name in ("rdf", "rdf_wrap", "rdf4m", "traj2d"):
        is2d = name == "traj2d"
        L = (n / (0.5 if is2d else RHO)) ** (0.5 if is2d else 1.0 / 3.0)
        tilt = (0.3, 0.2, 0.1) if name == "rdf4m" else None
        box, pts = data.make_random_system(L, n, is2D=is2d, seed=0, tilt=tilt)
        bins = 500 if name == "rdf4m" else 100
        runs = rep(lambda: cpu_reference_rdf(box, pts, bins, 5.0, budget_s=budget, threads=threads))
        metric, unit = "rdf_frames_per_sec", "frames/s"
        config = {"workload": f"RDF bins={bins} r_max=5 N={n} L={L:.4f}"
                              + (" triclinic (xy=.3,xz=.2,yz=.1)" if tilt else "") + (" 2-D" if is2d else "")}
    elif name == "pmftxy":
        L = (n / 0.5) ** 0.5
        box, pts = data.make_random_system(L, n, is2D=True, seed=0)
        rs = np.random.RandomState(29)
        angles = (rs.random_sample(n) * 2 * np.pi - np.pi).astype(np.float32)
        runs = rep(lambda: cpu_reference_pmftxy(box, pts, angles, 4.0, 3.0, (100, 100), budget_s=budget, threads=threads))
        metric, unit = "pmftxy_particles_per_sec", "particles/s"
        config = {"workload": f"PMFTXY x_max=4 y_max=3 bins=100x100 N={n} 2-D square L={L:.4f} areal density 0.5"}
    elif name in HIST_CLIENTS:
        box, pts, orient, spec = hist_client_inputs(name, n, 0)
        runs = rep(lambda: cpu_reference_hist_client(name, box, pts, orient, spec, budget_s=budget, threads=threads))
        metric, unit = f"{name}_particles_per_sec", "particles/s"
        config = {"workload": spec["label"]}
    elif name == "correlation":
        L = (n / RHO) ** (1.0 / 3.0)
        box, pts = data.make_random_system(L, n, seed=0)
        rs = np.random.RandomState(17)
        values = rs.standard_normal(n) + 1j * rs.standard_normal(n)
        runs = rep(lambda: cpu_reference_correlation(box, pts, values, 100, 3.0, budget_s=budget, threads=threads))
        metric, unit = "correlation_function_particles_per_sec", "particles/s"
        config = {"workload": f"CorrelationFunction bins=100 r_max=3 complex values N={n} cubic L={L:.4f} rho=0.08"}
    elif name == "local_density":
        L = (n / RHO) ** (1.0 / 3.0)
        box, pts = data.make_random_system(L, n, seed=0)
        runs = rep(lambda: cpu_reference_local_density(box, pts, 2.5, 1.0, budget_s=budget, threads=threads))
        metric, unit = "local_density_particles_per_sec", "particles/s"
        config = {"workload": f"LocalDensity r_max=2.5 diameter=1 N={n} cubic L={L:.4f} rho=0.08"}
    else:
        m = max(2, round((n / 4) ** (1.0 / 3.0)))
        box, pts = data.make_fcc_system(m, sigma_noise=0.05, seed=0)
        runs = rep(lambda: cpu_reference_q6(box, pts, threads=threads))
        metric, unit = "q6_particles_per_sec", "particles/s"
        config = {"workload": f"Steinhardt Q6 num_neighbors=12 FCC {m}^3x4={len(pts)} sigma=0.05"}
    vals = [r["value"] for r in runs if r.get("value")]
    value = float(np.median(vals)) if vals else None
    base = dict(runs[-1], value=value)
    return {"impl": "reference", "metric": metric, "value": value, "unit": unit, "steps": steps,
            "warmup": 0, "ms_per_step": float(np.median([r["wall_s"] for r in runs])) * 1e3, "higher_is_better": True,
            "scaling": "strong" if name == "rdf4m" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": base,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_reference_arm(args, rank, world, legs):
    """--impl reference: the reference's own CPU implementation on this box's host cores (rank 0 only), for the same
    legs the b200 arm runs at this world size (the first one is the line's headline)."""
    if rank != 0:
        return
    threads = os.cpu_count()
    steps = max(1, min(args.steps, 3))
    lines = []
    for k, name in enumerate(legs):
        lines.append(reference_leg(name, args.n or N_DEFAULT[name], steps if k == 0 else 1, 8.0 if k == 0 else 6.0, threads))
    line = dict(lines[0], n_gpus=world)
    if len(lines) > 1:
        line["legs"] = {name: ln for name, ln in zip(legs[1:], lines[1:])}
        for name, ln in zip(legs[1:], lines[1:]):
            key = ln["metric"] if ln["metric"] != line["metric"] and ln["metric"] not in line else f"{name}_{ln['metric']}"
            line[key] = ln["value"]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
