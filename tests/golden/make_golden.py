"""Generates the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

The Python package of the reference cannot be imported in the build container (nanobind and oneTBB are absent),
so the vectors come from ``oracle/_ref/libfreud_ref.so``: the unmodified reference C++ sources compiled where they
lie under /root/reference (oracle/Makefile).  Run from the repo root where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Inputs are regenerated from seeds by the tests (freud_b200.data / tests.util), only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from freud_b200 import data  # noqa: E402
from freud_b200.box import Box  # noqa: E402
from oracle import ref  # noqa: E402
from tests.util import random_points  # noqa: E402

CASES = {
    # name: (box, n_points, n_query (0 = self query), r_max, r_min, exclude_ii, seed)
    "tri": (Box(14, 15, 16, 0.3, 0.2, 0.1), 400, 0, 3.0, 0.0, True, 101),
    "tri_q": (Box(14, 15, 16, -0.4, 0.25, 0.15), 350, 120, 3.2, 0.8, False, 102),
    "cubic": (Box.cube(12), 400, 0, 2.5, 0.0, True, 103),
    "sq2d": (Box.square(30), 400, 0, 3.0, 0.0, True, 104),
    "tilt2d": (Box(30, 26, 0, 0.35, 0, 0, is2D=True), 300, 90, 3.0, 0.5, False, 105),
}


def nl_dict(nl, prefix):
    return {f"{prefix}_neighbors": nl.neighbors, f"{prefix}_distances": nl.distances, f"{prefix}_vectors": nl.vectors,
            f"{prefix}_segments": nl.segments, f"{prefix}_counts": nl.counts}


def main():
    ref.set_num_threads(4)
    for name, (box, n, nq, r_max, r_min, excl, seed) in CASES.items():
        pts = random_points(box, n, seed)
        q = pts if nq == 0 else random_points(box, nq, seed + 1000)
        out = {}
        lc = ref.Query("linkcell", box, pts, is2d=box.is2D, cell_width=min(r_max, 0.45 * min(box.Lx, box.Ly)))
        aq = ref.Query("aabb", box, pts, is2d=box.is2D)
        out.update(nl_dict(lc.nlist(q, r_max=r_max, r_min=r_min, exclude_ii=excl), "wrap"))
        out.update(nl_dict(aq.nlist(q, r_max=r_max, r_min=r_min, exclude_ii=excl), "image"))
        out.update(nl_dict(aq.nlist(q, r_max=r_max, r_min=r_min, exclude_ii=excl, sort_by_distance=True), "image_bydist"))
        out.update(nl_dict(aq.nlist(q, num_neighbors=6, exclude_ii=excl), "knn6"))
        for flavour, query in (("wrap", lc), ("image", aq)):
            R = ref.RDF(40, r_max, r_min)
            R.accumulate(query, q, mode="ball", r_max=r_max, exclude_ii=excl)
            res = R.results()
            out.update({f"rdf_{flavour}_{k}": v for k, v in res.items()})
        np.savez_compressed(os.path.join(HERE, f"nl_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.endswith("neighbors")})

    # BASELINE.json configs[0]: RDF bins=100 r_max=5 on make_random_system(box_size=50, num_points=10000)
    box, pts = data.make_random_system(50, 10000, seed=0)
    R = ref.RDF(100, 5.0)
    R.accumulate(ref.Query("raw", box, pts), pts, mode="ball", r_max=5.0, exclude_ii=True)
    res = R.results()
    R2 = ref.RDF(100, 5.0)  # accumulation over two frames, reset=False
    for s in (0, 1):
        b2, p2 = data.make_random_system(50, 10000, seed=s)
        R2.accumulate(ref.Query("raw", b2, p2), p2, mode="ball", r_max=5.0, exclude_ii=True)
    res2 = R2.results()
    np.savez_compressed(os.path.join(HERE, "rdf_config0.npz"), **res, **{f"two_frames_{k}": v for k, v in res2.items()})
    print("rdf_config0", int(res["bin_counts"].sum()), res["bin_counts"][-3:])

    # Steinhardt: noisy FCC, k = 12 (RawPoints -> AABB kNN, per-point iteration in distance order), and ball
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    out = {}
    for ls in ([6], [4, 6], [2, 8], [12]):
        S = ref.Steinhardt(ls)
        r = S.compute(ref.Query("raw", box, pts), num_neighbors=12, exclude_ii=True)
        tag = "_".join(str(l) for l in ls)
        out[f"knn12_ql_{tag}"] = r["ql"]
        out[f"knn12_order_{tag}"] = r["order"]
        for l, qlm in zip(ls, r["qlm"]):
            out[f"knn12_qlm_{tag}_l{l}"] = qlm
    S = ref.Steinhardt([6])
    r = S.compute(ref.Query("raw", box, pts), mode="ball", r_max=1.05, exclude_ii=True)
    out["ball_ql_6"] = r["ql"]
    np.savez_compressed(os.path.join(HERE, "steinhardt_fcc.npz"), **out)
    print("steinhardt", out["knn12_ql_6"][:3, 0])
    steinhardt_options()
    local_density()
    correlation_function()
    pmftxy()
    periodic_buffer()
    pmft3()
    bond_order()


STEINHARDT_OPTIONS = {
    # tag: constructor flags (freud/order/Steinhardt.h:66-76)
    "ave": dict(average=True),
    "wl": dict(wl=True),
    "wln": dict(wl=True, wl_normalize=True),
    "ave_wl": dict(average=True, wl=True),
    "ave_wln": dict(average=True, wl=True, wl_normalize=True),
}


def steinhardt_options():
    """Second-shell average and w_l (Steinhardt.cc:224-289, 329-359) on the noisy FCC system, k = 12."""
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    out = {}
    for tag, flags in STEINHARDT_OPTIONS.items():
        for ls in ([6], [4, 6], [3, 10]):
            r = ref.Steinhardt(ls, **flags).compute(ref.Query("raw", box, pts), num_neighbors=12, exclude_ii=True)
            key = tag + "_" + "_".join(str(l) for l in ls)
            out[f"{key}_particle_order"] = r["particle_order"]
            out[f"{key}_ql"] = r["ql"]
            out[f"{key}_order"] = r["order"]
    np.savez_compressed(os.path.join(HERE, "steinhardt_options.npz"), **out)
    print("steinhardt options", out["ave_wln_6_particle_order"][:3, 0], out["ave_wln_6_order"])


def local_density():
    """LocalDensity(r_max=3, diameter=1) and (2, 0.5) (LocalDensity.cc:38-84): with a NeighborList handed in (bonds summed
    in list order: the bit-exact target) and with the default on-the-fly query (engine order: tolerance)."""
    out = {}
    for name, box, n in (("cube", Box.cube(10), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 123), random_points(box, 500, 124)
        Q = ref.Query("aabb", box, pts, is2d=box.is2D)
        for r_max, diameter in ((3.0, 1.0), (2.0, 0.5)):
            key = f"{name}_{r_max:g}_{diameter:g}"
            nl = Q.nlist(q, mode="ball", r_max=r_max + 0.5 * diameter)
            out[f"{key}_nlist_num"], out[f"{key}_nlist_density"] = ref.local_density(Q, q, r_max, diameter, nlist=nl)
            out[f"{key}_query_num"], out[f"{key}_query_density"] = ref.local_density(Q, q, r_max, diameter)
            out[f"{key}_self_num"], out[f"{key}_self_density"] = ref.local_density(Q, pts, r_max, diameter,
                                                                                   exclude_ii=True)
    np.savez_compressed(os.path.join(HERE, "local_density.npz"), **out)
    print("local density", out["cube_3_1_nlist_num"][:3], out["cube_3_1_nlist_density"][:3])


def correlation_inputs(n, nq, seed):
    """Seeded complex values for the points and the query points (regenerated by the tests)."""
    rs = np.random.RandomState(seed)
    v = rs.standard_normal(n) + 1j * rs.standard_normal(n)
    qv = rs.standard_normal(nq) + 1j * rs.standard_normal(nq)
    return v, qv


def correlation_function():
    """CorrelationFunction(bins=40, r_max=3) (CorrelationFunction.cc:26-95): complex values with separate query points,
    real values of the points against themselves, in a cubic and a tilted 2-D box."""
    out = {}
    for name, box, n in (("cube", Box.cube(12), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 7), random_points(box, 700, 8)
        v, qv = correlation_inputs(n, 700, 3)
        Q = ref.Query("aabb", box, pts, is2d=box.is2D)
        out[f"{name}_complex_corr"], out[f"{name}_complex_counts"] = ref.correlation_function(Q, v, q, qv, 40, 3.0)
        out[f"{name}_real_corr"], out[f"{name}_real_counts"] = ref.correlation_function(Q, v.real, pts, v.real, 40, 3.0,
                                                                                  exclude_ii=True)
    np.savez_compressed(os.path.join(HERE, "correlation_function.npz"), **out)
    print("correlation function", out["cube_complex_corr"][-2:], out["cube_complex_counts"][-2:])


def pmftxy_inputs(n, nq, seed):
    """Seeded orientation angles for the points and for a separate query set (regenerated by the tests)."""
    rs = np.random.RandomState(seed)
    return ((rs.random_sample(n) * 2 * np.pi - np.pi).astype(np.float32),
            (rs.random_sample(nq) * 2 * np.pi).astype(np.float32))


def pmftxy():
    """PMFTXY(3, 2.5, (30, 24)) (PMFTXY.cc:25-87, PMFT.h:73-83) in a square and a tilted 2-D box: separate query points,
    and the points against themselves."""
    out = {}
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        Q = ref.Query("aabb", box, pts, is2d=True)
        out[f"{name}_query_counts"], out[f"{name}_query_pcf"] = ref.pmftxy(Q, th_q, q, 3.0, 2.5, 30, 24)
        out[f"{name}_self_counts"], out[f"{name}_self_pcf"] = ref.pmftxy(Q, th_p, pts, 3.0, 2.5, 30, 24, exclude_ii=True)
    np.savez_compressed(os.path.join(HERE, "pmftxy.npz"), **out)
    print("pmftxy", int(out["sq2d_query_counts"].sum()), out["sq2d_query_pcf"][15, 10:13])


def pmft3_quats(n, seed):
    """Seeded unit quaternions (regenerated by the tests)."""
    q = np.random.RandomState(seed).normal(size=(n, 4))
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


# the 4 proper rotations of a rectangular prism about its axes (w, x, y, z)
PMFT3_EQUIV = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)


def pmft3_lattice():
    """A 12 x 12 square lattice (spacing 1.25) with orientations on multiples of pi / 4: every bond angle sits on a bin
    edge of an 8-bin angular axis, where the bin depends on the last place of libm's atan2f."""
    g = (np.arange(12) - 5.5) * 1.25
    pts = np.stack([np.repeat(g, 12), np.tile(g, 12), np.zeros(144)], axis=1).astype(np.float32)
    return Box.square(15), pts, (np.arange(144) % 8 * (np.pi / 4)).astype(np.float32)


def pmft3():
    """PMFTXYZ (PMFTXYZ.cc:24-147) in a triclinic box; PMFTXYT (PMFTXYT.cc:28-101) and PMFTR12 (PMFTR12.cc:28-113) in a
    square box, a tilted 2-D box and on a square lattice whose bond angles all sit on bin edges."""
    out = {}
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 1500, 31), random_points(box, 400, 32)
    Q = ref.Query("aabb", box, pts)
    mx, bins = (2.0, 2.5, 3.0), (12, 10, 8)
    out["xyz_query_counts"], out["xyz_query_pcf"] = ref.pmft3(ref.PMFT_XYZ, Q, None, pmft3_quats(400, 6), q, mx, bins,
                                                               equiv=PMFT3_EQUIV)
    out["xyz_self_counts"], out["xyz_self_pcf"] = ref.pmft3(ref.PMFT_XYZ, Q, None, pmft3_quats(1500, 7), pts, mx, bins,
                                                             equiv=PMFT3_EQUIV[:1], exclude_ii=True)
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        Q = ref.Query("aabb", box, pts, is2d=True)
        out[f"{name}_xyt_query_counts"], out[f"{name}_xyt_query_pcf"] = ref.pmft3(ref.PMFT_XYT, Q, th_p, th_q, q,
                                                                                   (3.0, 2.5), (14, 12, 9))
        out[f"{name}_xyt_self_counts"], out[f"{name}_xyt_self_pcf"] = ref.pmft3(ref.PMFT_XYT, Q, th_p, th_p, pts,
                                                                                 (3.0, 2.5), (14, 12, 9), exclude_ii=True)
        out[f"{name}_r12_query_counts"], out[f"{name}_r12_query_pcf"] = ref.pmft3(ref.PMFT_R12, Q, th_p, th_q, q, (4.0,),
                                                                                   (10, 11, 12))
        out[f"{name}_r12_self_counts"], out[f"{name}_r12_self_pcf"] = ref.pmft3(ref.PMFT_R12, Q, th_p, th_p, pts, (4.0,),
                                                                                 (10, 11, 12), exclude_ii=True)
    box, pts, th = pmft3_lattice()
    Q = ref.Query("aabb", box, pts, is2d=True)
    out["lattice_xyt_counts"], out["lattice_xyt_pcf"] = ref.pmft3(ref.PMFT_XYT, Q, th, th, pts, (3.0, 3.0), (6, 6, 8),
                                                                   exclude_ii=True)
    out["lattice_r12_counts"], out["lattice_r12_pcf"] = ref.pmft3(ref.PMFT_R12, Q, th, th, pts, (3.0,), (6, 8, 8),
                                                                   exclude_ii=True)
    np.savez_compressed(os.path.join(HERE, "pmft3.npz"), **out)
    print("pmft3", {k: int(v.sum()) for k, v in out.items() if k.endswith("counts")})


def bond_order():
    """BondOrder (BondOrder.cc:30-153), all four modes over the 8 nearest neighbours of separate query points in a
    triclinic box, and mode bod on a perfect FCC lattice (12 nearest neighbours), whose bond directions all sit on the
    edges of an (8, 4) grid of bins -- there the bin hangs on the last place of libm's atan2f / acosf."""
    out = {}
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 800, 41), random_points(box, 300, 42)
    o, qo = pmft3_quats(800, 1), pmft3_quats(300, 2)
    Q = ref.Query("aabb", box, pts)
    for mode in ("bod", "lbod", "obcd", "oocd"):
        out[f"tri_{mode}_counts"], out[f"tri_{mode}_bo"] = ref.bond_order(mode, Q, o, q, qo, (12, 9), mode="nearest",
                                                                            num_neighbors=8)
    box, pts = data.UnitCell.fcc().generate_system(4)
    ident = np.tile(np.float32([1, 0, 0, 0]), (len(pts), 1))
    Q = ref.Query("aabb", box, pts)
    for bins in ((8, 4), (7, 5)):
        tag = f"fcc_{bins[0]}x{bins[1]}"
        out[f"{tag}_counts"], out[f"{tag}_bo"] = ref.bond_order("bod", Q, ident, pts, ident, bins, mode="nearest",
                                                                  num_neighbors=12, exclude_ii=True)
    np.savez_compressed(os.path.join(HERE, "bond_order.npz"), **out)
    print("bond order", {k: int(v.sum()) for k, v in out.items() if k.endswith("counts")})


PBUFF_BOXES = {"cube": Box.cube(5), "tri": Box(4, 5, 6, 0.3, -0.2, 0.1), "tilt2d": Box(4, 5, 0, 0.25, 0, 0, is2D=True)}
PBUFF_MODES = {
    # tag: (buffer, images, include_input_points) -- PeriodicBuffer.cc:23-117
    "img2_all": (2, True, True),
    "img120": ((1, 2, 0), True, False),
    "dist13": (1.3, False, False),
    "dist_mixed_all": ((0.5, 2.2, 1.0), False, True),
}


def periodic_buffer():
    """PeriodicBuffer in both modes (whole images / a distance on every side) and the systems that
    freud.data.UnitCell.generate_system builds on it (freud/data.py:58-150): here the replication only, which is
    all of the reference's C++ involved; the noise is numpy's."""
    out = {}
    for name, box in PBUFF_BOXES.items():
        pts = random_points(box, 40, 300 + len(name))
        q = ref.Query("raw", box, pts, is2d=box.is2D)
        for tag, (buffer, images, incl) in PBUFF_MODES.items():
            bp, ids, box6 = ref.periodic_buffer(q, buffer, images, incl)
            out[f"{name}_{tag}_points"], out[f"{name}_{tag}_ids"], out[f"{name}_{tag}_box"] = bp, ids, box6
    np.savez_compressed(os.path.join(HERE, "periodic_buffer.npz"), **out)
    print("periodic buffer", {k: v.shape for k, v in out.items() if k.endswith("ids")})


if __name__ == "__main__":
    main()
