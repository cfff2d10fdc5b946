"""TEST INFRASTRUCTURE ONLY -- ctypes front end to ``oracle/_ref/libfreud_ref.so``.

That library is the UNMODIFIED reference hot path (freud @ e4272dbe) compiled by ``oracle/Makefile``
from the sources under ``/root/reference`` against the std::thread TBB stand-in; see ``ref_capi.cc``
for the reference classes each entry point drives.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import this module; the product (``freud_b200``) never does.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfreud_ref.so")

ENGINE_LINKCELL, ENGINE_AABB, ENGINE_RAW, ENGINE_CELL = 0, 1, 2, 3
MODE_NONE, MODE_BALL, MODE_NEAREST = 0, 1, 2
DEFAULT_NUM_NEIGHBORS = 0xFFFFFFFF

_lib = None
_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.fref_last_error.restype = C.c_char_p
        L.fref_nq_create.restype = C.c_void_p
        L.fref_nq_create.argtypes = [C.c_int, _fp, C.c_int, _fp, C.c_uint, C.c_float]
        L.fref_nq_destroy.argtypes = [C.c_void_p]
        L.fref_query_nlist.restype = C.c_void_p
        L.fref_query_nlist.argtypes = [C.c_void_p, _fp, C.c_uint, C.c_int, C.c_uint, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_int, C.c_int]
        L.fref_nlist_num_bonds.restype = C.c_uint
        L.fref_nlist_num_bonds.argtypes = [C.c_void_p]
        L.fref_nlist_copy.argtypes = [C.c_void_p, _up, _fp, _fp, _fp]
        L.fref_nlist_segments.argtypes = [C.c_void_p, _up, _up]
        L.fref_nlist_destroy.argtypes = [C.c_void_p]
        L.fref_rdf_create.restype = C.c_void_p
        L.fref_rdf_create.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_int]
        L.fref_rdf_destroy.argtypes = [C.c_void_p]
        L.fref_rdf_reset.argtypes = [C.c_void_p]
        L.fref_rdf_accumulate.argtypes = [C.c_void_p, C.c_void_p, _fp, C.c_uint, C.c_void_p, C.c_int, C.c_uint,
                                          C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.fref_rdf_get.argtypes = [C.c_void_p, _up, _fp, _fp, _fp, _fp]
        L.fref_steinhardt_create.restype = C.c_void_p
        L.fref_steinhardt_create.argtypes = [_up, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
        L.fref_steinhardt_destroy.argtypes = [C.c_void_p]
        L.fref_steinhardt_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_float,
                                              C.c_float, C.c_float, C.c_float, C.c_int]
        L.fref_steinhardt_get.argtypes = [C.c_void_p, _fp, _fp, _fp]
        L.fref_steinhardt_get_qlm.argtypes = [C.c_void_p, C.c_uint, _fp]
        L.fref_box_apply.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_uint, _fp]
        L.fref_box_info.argtypes = [_fp, C.c_int, _fp, _fp]
        L.fref_set_num_threads.argtypes = [C.c_uint]
        _lib = L
    return _lib


_ERRORS = {1: ValueError, 2: ValueError, 3: RuntimeError, 4: IndexError, 5: RuntimeError}


def _raise():
    L = lib()
    kind = L.fref_last_error_kind()
    raise _ERRORS.get(kind, RuntimeError)(L.fref_last_error().decode())


def _f32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None:
        a = a.reshape(-1, shape_last)
    return a


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


def box6(box):
    """(Lx, Ly, Lz, xy, xz, yz) float32 from a 6-sequence or an object with those attributes."""
    if hasattr(box, "Lx"):
        box = (box.Lx, box.Ly, box.Lz, box.xy, box.xz, box.yz)
    return np.asarray(box, dtype=np.float32).copy()


def set_num_threads(n):
    lib().fref_set_num_threads(int(n))


def box_apply(box, is2d, op, vecs):
    v = _f32(vecs, 3)
    out = np.empty_like(v)
    b = box6(box)
    if lib().fref_box_apply(_p(b), int(is2d), {"wrap": 0, "fractional": 1, "absolute": 2}[op], _p(v), len(v),
                            _p(out)):
        _raise()
    return out


def box_info(box, is2d):
    b = box6(box)
    vol = C.c_float()
    pd = np.zeros(3, np.float32)
    if lib().fref_box_info(_p(b), int(is2d), C.byref(vol), _p(pd)):
        _raise()
    return vol.value, pd


class NeighborList:
    def __init__(self, handle, n_query):
        self._h = handle
        L = lib()
        nb = L.fref_nlist_num_bonds(handle)
        self.neighbors = np.zeros((nb, 2), np.uint32)
        self.distances = np.zeros(nb, np.float32)
        self.weights = np.zeros(nb, np.float32)
        self.vectors = np.zeros((nb, 3), np.float32)
        L.fref_nlist_copy(handle, _p(self.neighbors, _up), _p(self.distances), _p(self.weights), _p(self.vectors))
        self.segments = np.zeros(n_query, np.uint32)
        self.counts = np.zeros(n_query, np.uint32)
        L.fref_nlist_segments(handle, _p(self.segments, _up), _p(self.counts, _up))

    def __len__(self):
        return len(self.distances)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fref_nlist_destroy(self._h)
            self._h = None


def _qargs(mode=None, num_neighbors=None, r_max=None, r_min=0.0, r_guess=None, scale=None, exclude_ii=False):
    m = {None: MODE_NONE, "none": MODE_NONE, "ball": MODE_BALL, "nearest": MODE_NEAREST}[mode]
    return (m, DEFAULT_NUM_NEIGHBORS if num_neighbors is None else int(num_neighbors),
            -1.0 if r_max is None else float(r_max), float(r_min), -1.0 if r_guess is None else float(r_guess),
            -1.0 if scale is None else float(scale), int(bool(exclude_ii)))


class Query:
    """LinkCell / AABBQuery / RawPoints / CellQuery of the reference."""

    def __init__(self, engine, box, points, is2d=False, cell_width=0.0):
        self.points = _f32(points, 3)
        self.box = box6(box)
        self.is2d = bool(is2d)
        eng = {"linkcell": ENGINE_LINKCELL, "aabb": ENGINE_AABB, "raw": ENGINE_RAW, "cell": ENGINE_CELL}[engine]
        self._h = lib().fref_nq_create(eng, _p(self.box), int(self.is2d), _p(self.points), len(self.points),
                                       float(cell_width))
        if not self._h:
            _raise()

    def nlist(self, query_points, sort_by_distance=False, **qargs):
        q = _f32(query_points, 3)
        h = lib().fref_query_nlist(self._h, _p(q), len(q), *_qargs(**qargs), int(sort_by_distance))
        if not h:
            _raise()
        return NeighborList(h, len(q))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fref_nq_destroy(self._h)
            self._h = None


class RDF:
    def __init__(self, bins, r_max, r_min=0.0, finite_size=False):
        self.bins = int(bins)
        self._h = lib().fref_rdf_create(self.bins, float(r_max), float(r_min), int(finite_size))
        if not self._h:
            _raise()
        self.r_max = float(r_max)

    def accumulate(self, query, query_points, nlist=None, **qargs):
        q = _f32(query_points, 3)
        if not qargs:
            qargs = dict(mode="ball", r_max=self.r_max)
        if lib().fref_rdf_accumulate(self._h, query._h, _p(q), len(q), nlist._h if nlist is not None else None,
                                     *_qargs(**qargs)):
            _raise()

    def reset(self):
        lib().fref_rdf_reset(self._h)

    def results(self):
        counts = np.zeros(self.bins, np.uint32)
        g = np.zeros(self.bins, np.float32)
        n = np.zeros(self.bins, np.float32)
        edges = np.zeros(self.bins + 1, np.float32)
        centers = np.zeros(self.bins, np.float32)
        if lib().fref_rdf_get(self._h, _p(counts, _up), _p(g), _p(n), _p(edges), _p(centers)):
            _raise()
        return dict(bin_counts=counts, rdf=g, n_r=n, bin_edges=edges, bin_centers=centers)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fref_rdf_destroy(self._h)
            self._h = None


def pmftxy(query, query_orientations, query_points, x_max, y_max, n_x, n_y, nlist=None, exclude_ii=False):
    """PMFTXY(x_max, y_max, (n_x, n_y)).compute(...) of the reference: (bin_counts u32[n_x, n_y], pcf f32[n_x, n_y]);
    without a NeighborList it queries a ball of sqrt(x_max^2 + y_max^2) (freud/pmft.py:358)."""
    q = _f32(query_points, 3)
    t = _f32(query_orientations)
    counts, pcf = np.zeros((n_x, n_y), np.uint32), np.zeros((n_x, n_y), np.float32)
    L = lib()
    L.fref_pmftxy.argtypes = [C.c_void_p, _fp, _fp, C.c_uint, C.c_void_p, C.c_float, C.c_float, C.c_uint, C.c_uint,
                              C.c_float, C.c_int, _up, _fp]
    r_max = float(np.sqrt(x_max ** 2 + y_max ** 2))
    if L.fref_pmftxy(query._h, _p(t), _p(q), len(q), nlist._h if nlist is not None else None, float(x_max), float(y_max),
                     int(n_x), int(n_y), r_max, int(bool(exclude_ii)), _p(counts, _up), _p(pcf)):
        _raise()
    return counts, pcf


PMFT_XYZ, PMFT_XYT, PMFT_R12 = 0, 1, 2


def pmft3(kind, query, orientations, query_orientations, query_points, maxes, bins, equiv=None, nlist=None, r_max=None,
          exclude_ii=False):
    """PMFTXYZ / PMFTXYT / PMFTR12 of the reference, one compute: (bin_counts u32[bins], pcf f32[bins]).  ``maxes`` =
    (x_max, y_max, z_max) | (x_max, y_max) | (r_max,); orientations are (N, 4) quaternions for XYZ, angles otherwise.
    Without a NeighborList it queries a ball of r_max (default: the norm of ``maxes``, freud/pmft.py)."""
    q = _f32(query_points, 3)
    width = 4 if kind == PMFT_XYZ else None
    o = _f32(orientations, width) if orientations is not None else np.zeros(1, np.float32)
    qo = _f32(query_orientations, width)
    eq = _f32(equiv, 4) if equiv is not None else np.zeros((1, 4), np.float32)
    mx = list(maxes) + [0.0] * (3 - len(maxes))
    counts, pcf = np.zeros(bins, np.uint32), np.zeros(bins, np.float32)
    L = lib()
    L.fref_pmft3.argtypes = [C.c_int, C.c_void_p, _fp, _fp, _fp, C.c_uint, _fp, C.c_uint, C.c_void_p, C.c_float, C.c_float,
                             C.c_float, C.c_uint, C.c_uint, C.c_uint, C.c_float, C.c_int, _up, _fp]
    if r_max is None:
        r_max = float(np.sqrt(sum(m * m for m in maxes)))
    if L.fref_pmft3(int(kind), query._h, _p(o), _p(qo), _p(q), len(q), _p(eq), len(eq) if equiv is not None else 0,
                    nlist._h if nlist is not None else None, float(mx[0]), float(mx[1]), float(mx[2]), int(bins[0]),
                    int(bins[1]), int(bins[2]), float(r_max), int(bool(exclude_ii)), _p(counts, _up), _p(pcf)):
        _raise()
    return counts, pcf


BOND_ORDER_MODES = {"bod": 0, "lbod": 1, "obcd": 2, "oocd": 3}


def bond_order(bo_mode, query, orientations, query_points, query_orientations, bins, nlist=None, **qargs):
    """BondOrder(bins, mode).compute(...) of the reference: (bin_counts u32[n_theta, n_phi], bond_order f32[...]);
    orientations are (N, 4) quaternions; the bonds are ``nlist`` or the query ``qargs`` describe."""
    q = _f32(query_points, 3)
    o, qo = _f32(orientations, 4), _f32(query_orientations, 4)
    counts, bo = np.zeros(bins, np.uint32), np.zeros(bins, np.float32)
    L = lib()
    L.fref_bond_order.argtypes = [C.c_int, C.c_void_p, _fp, _fp, _fp, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_int,
                                  C.c_uint, C.c_float, C.c_int, _up, _fp]
    qa = _qargs(**qargs) if qargs else _qargs(mode="ball", r_max=1.0)
    mode_q, num_neighbors, r_max, exclude_ii = qa[0], qa[1], qa[2], qa[-1]
    if L.fref_bond_order(BOND_ORDER_MODES[bo_mode], query._h, _p(o), _p(q), _p(qo), len(q),
                         nlist._h if nlist is not None else None, int(bins[0]), int(bins[1]), mode_q, num_neighbors,
                         r_max, exclude_ii, _p(counts, _up), _p(bo)):
        _raise()
    return counts, bo


def correlation_function(query, values, query_points, query_values, bins, r_max, nlist=None, exclude_ii=False):
    """CorrelationFunction(bins, r_max).compute(...) of the reference: (correlation complex128[bins], bin_counts)."""
    q = _f32(query_points, 3)
    v = np.ascontiguousarray(values, dtype=np.complex128).ravel()
    qv = np.ascontiguousarray(query_values, dtype=np.complex128).ravel()
    corr, counts = np.zeros(bins, np.complex128), np.zeros(bins, np.uint32)
    L = lib()
    dp = C.POINTER(C.c_double)
    L.fref_correlation.argtypes = [C.c_void_p, dp, _fp, dp, C.c_uint, C.c_void_p, C.c_uint, C.c_float, C.c_int, dp, _up]
    if L.fref_correlation(query._h, v.ctypes.data_as(dp), _p(q), qv.ctypes.data_as(dp), len(q),
                          nlist._h if nlist is not None else None, int(bins), float(r_max), int(bool(exclude_ii)),
                          corr.ctypes.data_as(dp), _p(counts, _up)):
        _raise()
    return corr, counts


def local_density(query, query_points, r_max, diameter, nlist=None, q_r_max=None, exclude_ii=False):
    """LocalDensity(r_max, diameter).compute(...) of the reference: (num_neighbors, density).  Without a NeighborList
    it queries a ball of q_r_max (default r_max + diameter / 2, freud/density.py:510-514) on the fly."""
    q = _f32(query_points, 3)
    num, den = np.zeros(len(q), np.float32), np.zeros(len(q), np.float32)
    L = lib()
    L.fref_local_density.argtypes = [C.c_void_p, _fp, C.c_uint, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int,
                                     _fp, _fp]
    qr = r_max + 0.5 * diameter if q_r_max is None else q_r_max
    if L.fref_local_density(query._h, _p(q), len(q), nlist._h if nlist is not None else None, float(r_max),
                            float(diameter), float(qr), int(bool(exclude_ii)), _p(num), _p(den)):
        _raise()
    return num, den


def periodic_buffer(query, buffer, images=False, include_input_points=False):
    """PeriodicBuffer().compute(...) of the reference: (buffer_points, buffer_ids, box6 of the grown box)."""
    L = lib()
    L.fref_pbuff_compute.restype = C.c_void_p
    L.fref_pbuff_compute.argtypes = [C.c_void_p, _fp, C.c_int, C.c_int]
    L.fref_pbuff_size.argtypes = [C.c_void_p]
    L.fref_pbuff_size.restype = C.c_uint
    L.fref_pbuff_copy.argtypes = [C.c_void_p, _fp, _up, _fp]
    L.fref_pbuff_destroy.argtypes = [C.c_void_p]
    buff = np.ascontiguousarray(np.broadcast_to(np.asarray(buffer, np.float32), (3,)))
    h = L.fref_pbuff_compute(query._h, _p(buff), int(bool(images)), int(bool(include_input_points)))
    if not h:
        _raise()
    n = L.fref_pbuff_size(h)
    pts, ids, box6 = np.zeros((n, 3), np.float32), np.zeros(n, np.uint32), np.zeros(6, np.float32)
    L.fref_pbuff_copy(h, _p(pts), _p(ids, _up), _p(box6))
    L.fref_pbuff_destroy(h)
    return pts, ids, box6


class Steinhardt:
    def __init__(self, l, average=False, wl=False, weighted=False, wl_normalize=False):
        self.ls = np.atleast_1d(np.asarray(l, dtype=np.uint32)).copy()
        self._h = lib().fref_steinhardt_create(_p(self.ls, _up), len(self.ls), int(average), int(wl),
                                               int(weighted), int(wl_normalize))
        if not self._h:
            _raise()

    def compute(self, query, nlist=None, **qargs):
        if lib().fref_steinhardt_compute(self._h, query._h, nlist._h if nlist is not None else None,
                                         *_qargs(**qargs)):
            _raise()
        n = len(query.points)
        nl = len(self.ls)
        po = np.zeros((n, nl), np.float32)
        ql = np.zeros((n, nl), np.float32)
        order = np.zeros(nl, np.float32)
        if lib().fref_steinhardt_get(self._h, _p(po), _p(ql), _p(order)):
            _raise()
        qlm = []
        for k, l in enumerate(self.ls):
            buf = np.zeros((n, 2 * int(l) + 1), np.complex64)
            if lib().fref_steinhardt_get_qlm(self._h, k, buf.ctypes.data_as(_fp)):
                _raise()
            qlm.append(buf)
        return dict(particle_order=po, ql=ql, order=order, qlm=qlm)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fref_steinhardt_destroy(self._h)
            self._h = None
