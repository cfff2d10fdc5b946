"""TEST INFRASTRUCTURE ONLY -- ctypes front end to ``oracle/libfreud_port.so`` (the plain-C restatement in
``port.c``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module; the product (``freud_b200``) never does.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfreud_port.so")
WRAP, IMAGE, GHOST = 0, 1, 2
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "port.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.fport_ball_nlist.restype = C.c_void_p
        L.fport_ball_nlist.argtypes = [C.c_int, _fp, C.c_int, _fp, C.c_uint32, _fp, C.c_uint32, C.c_float, C.c_float,
                                       C.c_int, C.c_int]
        L.fport_knn_nlist.restype = C.c_void_p
        L.fport_knn_nlist.argtypes = [_fp, C.c_int, _fp, C.c_uint32, _fp, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                      C.c_int, C.c_int]
        L.fport_knn_nlist_wrap.restype = C.c_void_p
        L.fport_knn_nlist_wrap.argtypes = L.fport_knn_nlist.argtypes
        L.fport_nlist_size.restype = C.c_uint64
        L.fport_nlist_size.argtypes = [C.c_void_p]
        L.fport_nlist_copy.argtypes = [C.c_void_p, _up, _fp, _fp, _fp, _up, _up]
        L.fport_nlist_free.argtypes = [C.c_void_p]
        L.fport_rdf_accumulate.argtypes = [C.c_int, _fp, C.c_int, _fp, C.c_uint32, _fp, C.c_uint32, C.c_float,
                                           C.c_float, C.c_int, C.c_uint32, C.c_float, C.c_float, _up]
        L.fport_rdf_accumulate_distances.argtypes = [_fp, C.c_uint64, C.c_uint32, C.c_float, C.c_float, _up]
        L.fport_rdf_reduce.argtypes = [_up, C.c_uint32, C.c_float, C.c_float, _fp, C.c_int, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_int, _fp, _fp, _fp, _fp]
        L.fport_steinhardt.argtypes = [_fp, C.c_int, _fp, C.c_uint32, _up, _fp, _fp, _up, _up, _up, C.c_uint32,
                                       C.c_int, _fp, _fp, _fp, _fp]
        L.fport_steinhardt_options.argtypes = [C.c_uint32, _up, _up, _up, _up, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                               _fp, _fp, _fp, _fp, _fp]
        L.fport_wigner3j.argtypes = [C.c_uint32, _fp]
        L.fport_box_apply.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_uint32, _fp]
        L.fport_box_info.argtypes = [_fp, C.c_int, _fp, _fp]
        L.fport_count_candidates.restype = C.c_uint64
        L.fport_count_candidates.argtypes = [_fp, C.c_int, _fp, C.c_uint32, _fp, C.c_uint32, C.c_float]
        _lib = L
    return _lib


def _f32(a, last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(-1, last) if last else a


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


def box6(box):
    if hasattr(box, "Lx"):
        box = (box.Lx, box.Ly, box.Lz, box.xy, box.xz, box.yz)
    return np.asarray(box, dtype=np.float32).copy()


class NeighborList:
    def __init__(self, handle, n_query):
        L = lib()
        if not handle:
            raise MemoryError("oracle port: allocation failed")
        nb = L.fport_nlist_size(handle)
        self.neighbors = np.zeros((nb, 2), np.uint32)
        self.distances = np.zeros(nb, np.float32)
        self.weights = np.zeros(nb, np.float32)
        self.vectors = np.zeros((nb, 3), np.float32)
        self.segments = np.zeros(n_query, np.uint32)
        self.counts = np.zeros(n_query, np.uint32)
        L.fport_nlist_copy(handle, _p(self.neighbors, _up), _p(self.distances), _p(self.weights), _p(self.vectors),
                           _p(self.segments, _up), _p(self.counts, _up))
        L.fport_nlist_free(handle)

    def __len__(self):
        return len(self.distances)


def ball_nlist(flavour, box, is2d, points, query_points, r_max, r_min=0.0, exclude_ii=False, sort_by_distance=False):
    b, p, q = box6(box), _f32(points, 3), _f32(query_points, 3)
    h = lib().fport_ball_nlist(flavour, _p(b), int(is2d), _p(p), len(p), _p(q), len(q), r_max, r_min, int(exclude_ii),
                               int(sort_by_distance))
    return NeighborList(h, len(q))


def knn_nlist(box, is2d, points, query_points, k, r_max=np.inf, r_min=0.0, exclude_ii=False, sort_by_distance=False,
              flavour=IMAGE):
    """flavour IMAGE: AABBQueryIterator (AABBQuery.cc:152-281); WRAP: LinkCellQueryIterator (LinkCell.cc:575-679)."""
    b, p, q = box6(box), _f32(points, 3), _f32(query_points, 3)
    fn = lib().fport_knn_nlist if flavour == IMAGE else lib().fport_knn_nlist_wrap
    h = fn(_p(b), int(is2d), _p(p), len(p), _p(q), len(q), int(k), r_max, r_min, int(exclude_ii),
           int(sort_by_distance))
    return NeighborList(h, len(q))


def rdf_accumulate(flavour, box, is2d, points, query_points, bins, r_max, r_min=0.0, exclude_ii=False, counts=None,
                   query_r_max=None, query_r_min=None):
    b, p, q = box6(box), _f32(points, 3), _f32(query_points, 3)
    if counts is None:
        counts = np.zeros(bins, np.uint32)
    qmax = r_max if query_r_max is None else query_r_max
    qmin = 0.0 if query_r_min is None else query_r_min
    rc = lib().fport_rdf_accumulate(flavour, _p(b), int(is2d), _p(p), len(p), _p(q), len(q), qmax, qmin,
                                    int(exclude_ii), bins, r_min, r_max, _p(counts, _up))
    if rc:
        raise MemoryError("oracle port: allocation failed")
    return counts


def rdf_accumulate_distances(distances, bins, r_max, r_min=0.0, counts=None):
    d = _f32(distances)
    if counts is None:
        counts = np.zeros(bins, np.uint32)
    lib().fport_rdf_accumulate_distances(_p(d), len(d), bins, r_min, r_max, _p(counts, _up))
    return counts


def rdf_reduce(counts, r_max, r_min, box, is2d, n_points, n_query_points, frames=1, finite_size=False):
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    bins = len(c)
    b = box6(box)
    g = np.zeros(bins, np.float32)
    n = np.zeros(bins, np.float32)
    e = np.zeros(bins + 1, np.float32)
    ce = np.zeros(bins, np.float32)
    lib().fport_rdf_reduce(_p(c, _up), bins, r_min, r_max, _p(b), int(is2d), n_points, n_query_points, frames,
                           int(finite_size), _p(g), _p(n), _p(e), _p(ce))
    return dict(bin_counts=c, rdf=g, n_r=n, bin_edges=e, bin_centers=ce)


def pmftxy(box, n_points, nlist, query_orientations, x_max, y_max, n_x, n_y):
    """(bin_counts u32[n_x, n_y], pcf f32[n_x, n_y]) of PMFTXY over the bonds of an oracle NeighborList (one frame)."""
    ij = np.ascontiguousarray(nlist.neighbors, dtype=np.uint32)
    v = _f32(nlist.vectors, 3)
    t = _f32(query_orientations)
    counts, pcf = np.zeros((n_x, n_y), np.uint32), np.zeros((n_x, n_y), np.float32)
    L = lib()
    L.fport_pmftxy.argtypes = [_up, _fp, C.c_uint64, _fp, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_float,
                               C.c_uint32, C.c_uint32, _up, _fp]
    L.fport_pmftxy(_p(ij, _up), _p(v), len(v), _p(t), float(x_max), float(y_max), int(n_x), int(n_y),
                   float(np.float32(box.volume)), int(n_points), len(t), _p(counts, _up), _p(pcf))
    return counts, pcf


PMFT_XYZ, PMFT_XYT, PMFT_R12 = 0, 1, 2


def pmft3(kind, box, n_points, nlist, orientations, query_orientations, maxes, bins, equiv=None):
    """(bin_counts u32[bins], pcf f32[bins]) of PMFTXYZ / PMFTXYT / PMFTR12 over the bonds of an oracle NeighborList (one
    frame); argument meaning as oracle.ref.pmft3."""
    ij = np.ascontiguousarray(nlist.neighbors, dtype=np.uint32)
    v, d = _f32(nlist.vectors, 3), _f32(nlist.distances)
    width = 4 if kind == PMFT_XYZ else None
    o = _f32(orientations, width) if orientations is not None else np.zeros(1, np.float32)
    qo = _f32(query_orientations, width)
    eq = _f32(equiv, 4) if equiv is not None else np.zeros((1, 4), np.float32)
    mx = list(maxes) + [0.0] * (3 - len(maxes))
    counts, pcf = np.zeros(bins, np.uint32), np.zeros(bins, np.float32)
    L = lib()
    L.fport_pmft3.argtypes = [C.c_int, _up, _fp, _fp, C.c_uint64, _fp, _fp, _fp, C.c_uint32, C.c_float, C.c_float,
                              C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, _up, _fp]
    L.fport_pmft3(int(kind), _p(ij, _up), _p(v), _p(d), len(d), _p(o), _p(qo), _p(eq), len(eq) if equiv is not None else 0,
                  float(mx[0]), float(mx[1]), float(mx[2]), int(bins[0]), int(bins[1]), int(bins[2]),
                  float(np.float32(box.volume)), int(n_points), len(qo), _p(counts, _up), _p(pcf))
    return counts, pcf


BOND_ORDER_MODES = {"bod": 0, "lbod": 1, "obcd": 2, "oocd": 3}


def bond_order(mode, nlist, orientations, query_orientations, bins):
    """(bin_counts u32[n_theta, n_phi], bond_order f32[...]) of BondOrder over the bonds of an oracle NeighborList."""
    ij = np.ascontiguousarray(nlist.neighbors, dtype=np.uint32)
    v = _f32(nlist.vectors, 3)
    o, qo = _f32(orientations, 4), _f32(query_orientations, 4)
    counts, bo = np.zeros(bins, np.uint32), np.zeros(bins, np.float32)
    L = lib()
    L.fport_bond_order.argtypes = [C.c_int, _up, _fp, C.c_uint64, _fp, _fp, C.c_uint32, C.c_uint32, _up, _fp]
    L.fport_bond_order(BOND_ORDER_MODES[mode], _p(ij, _up), _p(v), len(v), _p(o), _p(qo), int(bins[0]), int(bins[1]),
                       _p(counts, _up), _p(bo))
    return counts, bo


def correlation_function(nlist, values, query_values, bins, r_max):
    """(correlation complex128[bins], bin_counts) of CorrelationFunction over the bonds of an oracle NeighborList."""
    ij = np.ascontiguousarray(nlist.neighbors, dtype=np.uint32)
    d = _f32(nlist.distances)
    v = np.ascontiguousarray(values, dtype=np.complex128).ravel()
    qv = np.ascontiguousarray(query_values, dtype=np.complex128).ravel()
    corr, counts = np.zeros(bins, np.complex128), np.zeros(bins, np.uint32)
    L = lib()
    dp = C.POINTER(C.c_double)
    L.fport_correlation.argtypes = [_up, _fp, C.c_uint64, dp, dp, C.c_uint32, C.c_float, dp, _up]
    L.fport_correlation(_p(ij, _up), _p(d), len(d), v.ctypes.data_as(dp), qv.ctypes.data_as(dp), int(bins), float(r_max),
                        corr.ctypes.data_as(dp), _p(counts, _up))
    return corr, counts


def local_density(nlist, r_max, diameter, is2d=False):
    """(num_neighbors, density) of LocalDensity::compute over the rows of an oracle NeighborList."""
    d = _f32(nlist.distances)
    seg = np.ascontiguousarray(nlist.segments, dtype=np.uint32)
    cnt = np.ascontiguousarray(nlist.counts, dtype=np.uint32)
    num, den = np.zeros(len(seg), np.float32), np.zeros(len(seg), np.float32)
    L = lib()
    L.fport_local_density.argtypes = [_fp, _up, _up, C.c_uint32, C.c_float, C.c_float, C.c_int, _fp, _fp]
    L.fport_local_density(_p(d), _p(seg, _up), _p(cnt, _up), len(seg), float(r_max), float(diameter), int(bool(is2d)),
                          _p(num), _p(den))
    return num, den


def wigner3j(l):
    """(l l l; m1 m2 m3) in the order of reduceWigner3j's table (Wigner3j.cc:43-55), as float."""
    out = np.zeros(3 * l * l + 3 * l + 1, np.float32)
    if lib().fport_wigner3j(int(l), _p(out)) != len(out):
        raise ValueError("oracle port: l too large")
    return out


def steinhardt(box, is2d, points, nlist, ls, weighted=False, average=False, wl=False, wl_normalize=False):
    b, p = box6(box), _f32(points, 3)
    ls = np.atleast_1d(np.asarray(ls, dtype=np.uint32)).copy()
    n = len(p)
    j = np.ascontiguousarray(nlist.neighbors[:, 1], dtype=np.uint32)
    d = _f32(nlist.distances)
    w = _f32(nlist.weights)
    seg = np.ascontiguousarray(nlist.segments, dtype=np.uint32)
    cnt = np.ascontiguousarray(nlist.counts, dtype=np.uint32)
    tot_m = int(sum(2 * int(l) + 1 for l in ls))
    ql = np.zeros((n, len(ls)), np.float32)
    qlm = np.zeros(n * tot_m * 2, np.float32)
    sys_qlm = np.zeros(tot_m * 2, np.float32)
    order = np.zeros(len(ls), np.float32)
    rc = lib().fport_steinhardt(_p(b), int(is2d), _p(p), n, _p(j, _up), _p(d), _p(w), _p(seg, _up), _p(cnt, _up),
                                _p(ls, _up), len(ls), int(weighted), _p(ql), _p(qlm), _p(sys_qlm), _p(order))
    if rc:
        raise ValueError("oracle port: l too large")
    out, off = [], 0
    for l in ls:
        nm = 2 * int(l) + 1
        blk = qlm[off:off + n * nm * 2].reshape(n, nm, 2)
        out.append((blk[..., 0] + 1j * blk[..., 1]).astype(np.complex64))
        off += n * nm * 2
    if average or wl:
        ql_out, po = np.zeros_like(ql), np.zeros_like(ql)
        rc = lib().fport_steinhardt_options(n, _p(j, _up), _p(seg, _up), _p(cnt, _up), _p(ls, _up), len(ls), int(average),
                                            int(wl), int(wl_normalize), _p(qlm), _p(ql), _p(ql_out), _p(po), _p(order))
        if rc:
            raise IndexError("Wigner 3j coefficients are implemented for l <= 20.")
        return dict(ql=ql_out, qlm=out, order=order, particle_order=po)
    return dict(ql=ql, qlm=out, order=order, particle_order=ql)


def box_apply(box, is2d, op, vecs):
    v = _f32(vecs, 3)
    out = np.empty_like(v)
    b = box6(box)
    lib().fport_box_apply(_p(b), int(is2d), {"wrap": 0, "fractional": 1, "absolute": 2}[op], _p(v), len(v), _p(out))
    return out


def box_info(box, is2d):
    b = box6(box)
    vol = C.c_float()
    pd = np.zeros(3, np.float32)
    lib().fport_box_info(_p(b), int(is2d), C.byref(vol), _p(pd))
    return vol.value, pd


def count_candidates(box, is2d, points, query_points, r_max):
    b, p, q = box6(box), _f32(points, 3), _f32(query_points, 3)
    return int(lib().fport_count_candidates(_p(b), int(is2d), _p(p), len(p), _p(q), len(q), r_max))
