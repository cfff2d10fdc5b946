"""ctypes binding of ``libfreud_b200.so`` (the C ABI declared in ``include/freud_b200.h``).

This is the only way Python reaches the GPU path; there is no fallback.  Importing the module never needs a
GPU, loading the library needs the built ``.so`` (``python -c "import __graft_entry__ as g; g.build()"``), and
creating a :class:`Context` needs a CUDA device -- each failure is raised, never papered over.
"""

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfreud_b200.so")

FLAVOUR_WRAP, FLAVOUR_IMAGE, FLAVOUR_GHOST = 0, 1, 2
ST_WEIGHTED, ST_AVERAGE, ST_WL, ST_WL_NORMALIZE = 1, 2, 4, 8
UNIQUE_ID_BYTES = 128

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# every symbol include/freud_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "fgpu_last_error": (C.c_char_p, []),
    "fgpu_version": (C.c_char_p, []),
    "fgpu_device_count": (C.c_int, []),
    "fgpu_ctx_create": (C.c_int, [C.c_int, _vpp]),
    "fgpu_ctx_destroy": (None, [_vp]),
    "fgpu_ctx_trim": (C.c_int, [_vp]),
    "fgpu_ctx_synchronize": (C.c_int, [_vp]),
    "fgpu_ctx_stream": (_vp, [_vp]),
    "fgpu_ctx_launch_count": (C.c_uint64, [_vp]),
    "fgpu_ctx_count_pair_evals": (C.c_int, [_vp, C.c_int]),
    "fgpu_ctx_pair_evals": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.c_int]),
    "fgpu_ctx_force_general_search": (C.c_int, [_vp, C.c_int]),
    "fgpu_ctx_set_tuning": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "fgpu_ctx_profile": (C.c_int, [_vp, C.c_int]),
    "fgpu_ctx_kernel_time": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "fgpu_ctx_kernel_timeline": (C.c_int, [_vp, C.c_char_p, C.c_uint64]),
    "fgpu_points_create": (C.c_int, [_vp, _fp, C.c_int, _fp, C.c_uint32, _vpp]),
    "fgpu_points_create_dev": (C.c_int, [_vp, _fp, C.c_int, _vp, C.c_uint32, _vpp]),
    "fgpu_points_create_replicated": (C.c_int, [_vp, _vp, _fp, C.c_int, _fp, C.c_uint32, _vpp]),
    "fgpu_points_destroy": (None, [_vp]),
    "fgpu_points_set_shard": (C.c_int, [_vp, C.c_int, C.c_int]),
    "fgpu_shard_plan": (C.c_int, [_up, C.c_uint32, C.c_int, C.c_int, _up]),
    "fgpu_points_build_cells": (C.c_int, [_vp, C.c_float, _up]),
    "fgpu_points_read_cells": (C.c_int, [_vp, _up, _up]),
    "fgpu_ball_query": (C.c_int, [_vp, _fp, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                  _vpp]),
    "fgpu_ball_query_dev": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int,
                                      C.c_int, _vpp]),
    "fgpu_knn_query": (C.c_int, [_vp, _fp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_int,
                                 C.c_int, _vpp]),
    "fgpu_nlist_num_bonds": (C.c_uint64, [_vp]),
    "fgpu_nlist_num_query_points": (C.c_uint32, [_vp]),
    "fgpu_nlist_num_points": (C.c_uint32, [_vp]),
    "fgpu_nlist_copy": (C.c_int, [_vp, _up, _fp, _fp, _fp, _up, _up]),
    "fgpu_nlist_copy_begin": (C.c_int, [_vp, _up, _fp, _fp, _fp]),
    "fgpu_nlist_copy_wait": (C.c_int, [_vp, C.c_uint]),
    "fgpu_nlist_from_host": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_uint32, _up, _fp, _fp, _fp, _vpp]),
    "fgpu_nlist_destroy": (None, [_vp]),
    "fgpu_rdf_create": (C.c_int, [_vp, C.c_uint32, C.c_float, C.c_float, _vpp]),
    "fgpu_rdf_destroy": (None, [_vp]),
    "fgpu_rdf_reset": (C.c_int, [_vp]),
    "fgpu_rdf_accumulate": (C.c_int, [_vp, _vp, _fp, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int]),
    "fgpu_rdf_accumulate_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_float,
                                          C.c_int]),
    "fgpu_rdf_accumulate_nlist": (C.c_int, [_vp, _vp]),
    "fgpu_rdf_read": (C.c_int, [_vp, _up]),
    "fgpu_rdf_allreduce": (C.c_int, [_vp, _vp]),
    "fgpu_rdf_attach_comm": (C.c_int, [_vp, _vp]),
    "fgpu_rdf_reduce_transport": (C.c_int, [_vp]),
    "fgpu_rdf_accumulate_reduce": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int]),
    "fgpu_pmftxy_create": (C.c_int, [_vp, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _vpp]),
    "fgpu_pmftxy_destroy": (None, [_vp]),
    "fgpu_pmftxy_reset": (C.c_int, [_vp]),
    "fgpu_pmftxy_accumulate_nlist": (C.c_int, [_vp, _vp, _fp]),
    "fgpu_pmftxy_accumulate": (C.c_int, [_vp, _vp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, _fp]),
    "fgpu_pmftxy_read": (C.c_int, [_vp, _up]),
    "fgpu_pmft_create": (C.c_int, [_vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, _vpp]),
    "fgpu_pmft_destroy": (None, [_vp]),
    "fgpu_pmft_reset": (C.c_int, [_vp]),
    "fgpu_pmft_accumulate_nlist": (C.c_int, [_vp, _vp, _fp, C.c_uint32, _fp, _fp, C.c_uint32]),
    "fgpu_pmft_accumulate": (C.c_int, [_vp, _vp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, _fp, _fp, _fp,
                                       C.c_uint32]),
    "fgpu_pmft_read": (C.c_int, [_vp, _up]),
    "fgpu_pmft_deferred": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "fgpu_bondorder_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int, _vpp]),
    "fgpu_bondorder_destroy": (None, [_vp]),
    "fgpu_bondorder_reset": (C.c_int, [_vp]),
    "fgpu_bondorder_accumulate_nlist": (C.c_int, [_vp, _vp, _fp, C.c_uint32, _fp]),
    "fgpu_bondorder_accumulate": (C.c_int, [_vp, _vp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, _fp, _fp]),
    "fgpu_bondorder_read": (C.c_int, [_vp, _up]),
    "fgpu_bondorder_deferred": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "fgpu_corr_create": (C.c_int, [_vp, C.c_uint32, C.c_float, _vpp]),
    "fgpu_corr_destroy": (None, [_vp]),
    "fgpu_corr_reset": (C.c_int, [_vp]),
    "fgpu_corr_accumulate_nlist": (C.c_int, [_vp, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fgpu_corr_accumulate": (C.c_int, [_vp, _vp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fgpu_corr_read": (C.c_int, [_vp, _up, C.POINTER(C.c_double)]),
    "fgpu_local_density": (C.c_int, [_vp, C.c_float, C.c_float, C.c_int, _fp, _fp]),
    "fgpu_local_density_query": (C.c_int, [_vp, _fp, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                           C.c_float, _fp, _fp]),
    "fgpu_steinhardt_compute": (C.c_int, [_vp, _vp, _up, C.c_uint32, C.c_int, C.c_uint32, _vp, _fp, _fp, _fp, _fp,
                                         _fp]),
    "fgpu_steinhardt_compute_keep": (C.c_int, [_vp, _vp, _up, C.c_uint32, C.c_int, C.c_uint32, _vp, _fp, _fp, _vpp, _fp,
                                              _fp]),
    "fgpu_steinhardt_knn": (C.c_int, [_vp, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_int, _up, C.c_uint32, C.c_int, _fp,
                                     _fp, _vpp, _fp, _fp]),
    "fgpu_buffer_bytes": (C.c_uint64, [_vp]),
    "fgpu_buffer_read": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64]),
    "fgpu_buffer_destroy": (None, [_vp]),
    "fgpu_host_alloc": (C.c_int, [C.c_uint64, _vpp]),
    "fgpu_host_free": (None, [_vp]),
    "fgpu_host_trim": (C.c_int, []),
    "fgpu_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "fgpu_comm_create": (C.c_int, [_vp, C.POINTER(C.c_uint8), C.c_int, C.c_int, _vpp]),
    "fgpu_comm_destroy": (None, [_vp]),
    "fgpu_comm_rank": (C.c_int, [_vp]),
    "fgpu_comm_size": (C.c_int, [_vp]),
    "fgpu_comm_barrier": (C.c_int, [_vp]),
    "fgpu_comm_allreduce_u32": (C.c_int, [_vp, _up, C.c_uint64]),
    "fgpu_comm_allreduce_f64": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_uint64]),
}

_lib = None


class GpuError(RuntimeError):
    pass


_CODE_TO_EXC = {-1: ValueError, -2: ValueError, -3: RuntimeError, -4: GpuError, -5: GpuError, -6: MemoryError,
                -7: IndexError}


def lib():
    """Load libfreud_b200.so and bind every declared symbol; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpuError(f"{LIB_PATH} is missing: build it with `make -C freud_b200/csrc` "
                           "(or __graft_entry__.build()); freud_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise _CODE_TO_EXC.get(rc, RuntimeError)(lib().fgpu_last_error().decode())


def f32(a, last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(-1, last) if last else a


def ptr(a, t=_fp):
    return a.ctypes.data_as(t) if a is not None else None


def box6_of(box):
    if hasattr(box, "as_array6"):
        return box.as_array6(), bool(box.is2D)
    from .box import Box

    b = Box.from_box(box)
    return b.as_array6(), bool(b.is2D)


def shard_plan(dims, n_points, shard, n_shards):
    """Host-only arithmetic of ``fgpu_points_set_shard``: dict of the tickets, cells and slab of one shard."""
    d = np.asarray(dims, dtype=np.uint32).copy()
    out = np.zeros(8, np.uint32)
    check(lib().fgpu_shard_plan(ptr(d, _up), int(n_points), int(shard), int(n_shards), ptr(out, _up)))
    keys = ("ticket_begin", "ticket_end", "n_tickets", "cell_begin", "cell_end", "slab_axis", "slab_lo", "slab_len")
    plan = {k: int(v) for k, v in zip(keys, out)}
    plan["slab_len"] = None if plan["slab_len"] == 0xFFFFFFFF else plan["slab_len"]
    return plan


class Context:
    """One GPU + one stream + scratch memory (``fgpu_ctx``)."""

    _default = {}

    def __init__(self, device=0):
        self._h = _vp()
        check(lib().fgpu_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self._children = weakref.WeakSet()  # device objects that must die before the context does

    @classmethod
    def default(cls, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % max(1, lib().fgpu_device_count())
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def synchronize(self):
        check(lib().fgpu_ctx_synchronize(self._h))

    @property
    def stream(self):
        return lib().fgpu_ctx_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().fgpu_ctx_launch_count(self._h))

    def count_pair_evals(self, enable=True):
        check(lib().fgpu_ctx_count_pair_evals(self._h, int(enable)))

    def pair_evals(self, reset=True):
        out = C.c_uint64()
        check(lib().fgpu_ctx_pair_evals(self._h, C.byref(out), int(reset)))
        return int(out.value)

    def force_general_search(self, enable=True):
        check(lib().fgpu_ctx_force_general_search(self._h, int(enable)))

    def set_tuning(self, key, value):
        """Experiment / test hook (``fgpu_ctx_set_tuning``): "span", "no_symmetry", "lanes_over_queries"."""
        check(lib().fgpu_ctx_set_tuning(self._h, key.encode(), int(value)))

    def profile(self, enable=True):
        check(lib().fgpu_ctx_profile(self._h, int(enable)))

    def kernel_time(self, prefix="", reset=False):
        """(summed ms, launches) of the profiled kernels whose name starts with prefix."""
        ms, n = C.c_double(), C.c_uint64()
        check(lib().fgpu_ctx_kernel_time(self._h, prefix.encode(), C.byref(ms), C.byref(n), int(reset)))
        return ms.value, int(n.value)

    def kernel_timeline(self):
        """[(name, begin_us, end_us)] of the profiled launches since the last reset (``fgpu_ctx_kernel_timeline``)."""
        buf = C.create_string_buffer(1 << 20)
        check(lib().fgpu_ctx_kernel_timeline(self._h, buf, len(buf)))
        rows = [ln.split() for ln in buf.value.decode().splitlines()]
        return [(r[0], float(r[1]), float(r[2])) for r in rows]

    def trim(self):
        """Release the context's grow-only scratch and the arrays kept from the last NeighborList (``fgpu_ctx_trim``)."""
        check(lib().fgpu_ctx_trim(self._h))

    def close(self):
        if self._h:
            for child in list(self._children):
                child._release()
            lib().fgpu_ctx_destroy(self._h)
            self._h = _vp()


class _DeviceObject:
    """Handle owner: destroyed by its own finaliser or, earlier, by ``Context.close()``."""

    _destroy = None  # name of the C destructor

    def _adopt(self, ctx):
        self.ctx = ctx
        ctx._children.add(self)

    def _release(self):
        if getattr(self, "_h", None):
            getattr(lib(), self._destroy)(self._h)
        self._h = None

    def __del__(self):
        self._release()


class DeviceNeighborList(_DeviceObject):
    """Device-resident NeighborList (``fgpu_nlist``); arrays come to the host on demand."""

    _destroy = "fgpu_nlist_destroy"

    def __init__(self, ctx, handle):
        self._adopt(ctx)
        self._h = handle
        L = lib()
        self.num_bonds = int(L.fgpu_nlist_num_bonds(handle))
        self.num_query_points = int(L.fgpu_nlist_num_query_points(handle))
        self.num_points = int(L.fgpu_nlist_num_points(handle))

    def to_host(self, into=None):
        nb, nq = self.num_bonds, self.num_query_points
        out = into or dict(neighbors=np.empty((nb, 2), np.uint32), distances=np.empty(nb, np.float32),
                           weights=np.empty(nb, np.float32), vectors=np.empty((nb, 3), np.float32),
                           segments=np.empty(nq, np.uint32), counts=np.empty(nq, np.uint32))
        check(lib().fgpu_nlist_copy(self._h, ptr(out["neighbors"], _up), ptr(out["distances"]), ptr(out["weights"]),
                                    ptr(out["vectors"]), ptr(out["segments"], _up), ptr(out["counts"], _up)))
        return out

    def local_density(self, r_max, diameter, is2d=False, out=None):
        """(num_neighbors, density) of ``fgpu_local_density`` over this list's rows; ``out``: optional pair of
        preallocated (e.g. page-locked) float32 arrays of length num_query_points."""
        num, den = out if out is not None else (np.empty(self.num_query_points, np.float32),
                                                np.empty(self.num_query_points, np.float32))
        assert num.dtype == np.float32 and den.dtype == np.float32 and num.size == den.size == self.num_query_points
        check(lib().fgpu_local_density(self._h, float(r_max), float(diameter), int(bool(is2d)), ptr(num), ptr(den)))
        return num, den

    @classmethod
    def from_host(cls, ctx, neighbors, distances, weights, vectors, n_query, n_points):
        nbr = np.ascontiguousarray(neighbors, dtype=np.uint32).reshape(-1, 2)
        d = f32(distances)
        w = f32(weights) if weights is not None else None
        v = f32(vectors, 3) if vectors is not None else None
        h = _vp()
        check(lib().fgpu_nlist_from_host(ctx._h, len(d), int(n_query), int(n_points), ptr(nbr, _up), ptr(d), ptr(w),
                                         ptr(v), C.byref(h)))
        return cls(ctx, h)


class DevicePoints(_DeviceObject):
    """Device-resident reference points + box + cell list (``fgpu_points``)."""

    _destroy = "fgpu_points_destroy"

    def __init__(self, ctx, box, points, comm=None):
        """comm: every rank holds the same host array; each uploads its block and NVLink replicates the rest
        (``fgpu_points_create_replicated``, collective)."""
        self._adopt(ctx)
        self.box6, self.is2d = box6_of(box)
        pts = f32(points, 3)
        self.n = len(pts)
        self._h = _vp()
        if comm is None:
            check(lib().fgpu_points_create(ctx._h, ptr(self.box6), int(self.is2d), ptr(pts), self.n, C.byref(self._h)))
        else:
            check(lib().fgpu_points_create_replicated(ctx._h, comm._h, ptr(self.box6), int(self.is2d), ptr(pts), self.n,
                                                      C.byref(self._h)))

    def build_cells(self, r_search):
        dims = np.zeros(3, np.uint32)
        check(lib().fgpu_points_build_cells(self._h, float(r_search), ptr(dims, _up)))
        return dims

    def set_shard(self, shard, n_shards):
        """Deal the home tiles of self-query RDF accumulation to ``n_shards`` ranks; this object is rank ``shard``."""
        check(lib().fgpu_points_set_shard(self._h, int(shard), int(n_shards)))

    def read_cells(self, dims):
        n_cells = int(np.prod(dims))
        cell_start = np.zeros(n_cells + 1, np.uint32)
        order = np.zeros(self.n, np.uint32)
        check(lib().fgpu_points_read_cells(self._h, ptr(cell_start, _up), ptr(order, _up)))
        return cell_start, order

    def ball_query(self, query_points, flavour, r_max, r_min=0.0, exclude_ii=False, sort_by_distance=False,
                   q_index_offset=0):
        q = None if query_points is None else f32(query_points, 3)
        nq = self.n if q is None else len(q)
        h = _vp()
        check(lib().fgpu_ball_query(self._h, ptr(q), nq, int(q_index_offset), int(flavour), float(r_max), float(r_min),
                                    int(bool(exclude_ii)), int(bool(sort_by_distance)), C.byref(h)))
        return DeviceNeighborList(self.ctx, h)

    def knn_query(self, query_points, num_neighbors, r_max=np.inf, r_min=0.0, exclude_ii=False,
                  sort_by_distance=False, q_index_offset=0, flavour=FLAVOUR_IMAGE):
        q = None if query_points is None else f32(query_points, 3)
        nq = self.n if q is None else len(q)
        h = _vp()
        check(lib().fgpu_knn_query(self._h, ptr(q), nq, int(q_index_offset), int(flavour), int(num_neighbors),
                                   float(r_max), float(r_min), int(bool(exclude_ii)), int(bool(sort_by_distance)),
                                   C.byref(h)))
        return DeviceNeighborList(self.ctx, h)

    def local_density(self, query_points, flavour, q_r_max, r_max, diameter, exclude_ii=False, q_r_min=0.0, out=None):
        """(num_neighbors, density) of LocalDensity(r_max, diameter) over a ball query of ``q_r_max`` made for it alone
        (``query_points=None``: the points themselves) -- no NeighborList (``fgpu_local_density_query``)."""
        q = None if query_points is None else f32(query_points, 3)
        nq = self.n if q is None else len(q)
        num, den = out if out is not None else (np.empty(nq, np.float32), np.empty(nq, np.float32))
        assert num.dtype == np.float32 and den.dtype == np.float32 and num.size == den.size == nq
        check(lib().fgpu_local_density_query(self._h, ptr(q), nq, int(flavour), float(q_r_max), float(q_r_min),
                                             int(bool(exclude_ii)), float(r_max), float(diameter), ptr(num), ptr(den)))
        return num, den

    def steinhardt_knn(self, num_neighbors, ls, r_max=np.inf, r_min=0.0, exclude_ii=True, flavour=FLAVOUR_IMAGE,
                       want_qlm=False, out=None):
        """Steinhardt over the k nearest neighbours of every point, query and sums in one call, no NeighborList
        (``fgpu_steinhardt_knn``).  Returns ``ql``, ``sys_qlm``, ``order`` and, with ``want_qlm``, ``qlm``."""
        ls = np.atleast_1d(np.asarray(ls, dtype=np.uint32)).copy()
        n = self.n
        tot_m = int(sum(2 * int(l) + 1 for l in ls))
        out = out or {}
        ql = out["ql"] if "ql" in out else np.empty((n, len(ls)), np.float32)
        sys_qlm = np.empty(tot_m * 2, np.float32)
        order = np.empty(len(ls), np.float32)
        keep = _vp()
        check(lib().fgpu_steinhardt_knn(self._h, int(flavour), int(num_neighbors), float(r_max), float(r_min),
                                        int(bool(exclude_ii)), ptr(ls, _up), len(ls), 0, ptr(ql), None,
                                        C.byref(keep) if want_qlm else None, ptr(sys_qlm), ptr(order)))
        res = {"ql": ql.reshape(n, len(ls)), "sys_qlm": sys_qlm, "order": order}
        if want_qlm:
            flat = out["qlm"] if "qlm" in out else np.empty(n * tot_m * 2, np.float32)
            check(lib().fgpu_buffer_read(keep, flat.ctypes.data_as(_vp), 0, flat.nbytes))
            lib().fgpu_buffer_destroy(keep)
            qlm, off = [], 0
            for l in ls:
                nm = 2 * int(l) + 1
                qlm.append(flat[off:off + n * nm * 2].view(np.complex64).reshape(n, nm))
                off += n * nm * 2
            res["qlm"] = qlm
        return res

    def steinhardt(self, nlist, ls, weighted=False, want_qlm=True, comm=None, n_total=0, out=None, average=False,
                   wl=False, wl_normalize=False):
        """``out``: optional dict of preallocated (e.g. page-locked) float32 arrays ``ql`` (n, len(ls)) and
        ``qlm`` (n * sum(2l+1) * 2,) to receive the per-particle results.  Returns ``ql`` (the averaged q_l when
        ``average``), ``wl`` (when ``wl``), ``qlm`` (always un-averaged), ``sys_qlm`` and ``order``."""
        ls = np.atleast_1d(np.asarray(ls, dtype=np.uint32)).copy()
        n = nlist.num_query_points
        tot_m = int(sum(2 * int(l) + 1 for l in ls))
        out = out or {}
        ql = out["ql"] if "ql" in out else np.empty((n, len(ls)), np.float32)
        qlm = (out["qlm"] if "qlm" in out else np.empty(n * tot_m * 2, np.float32)) if want_qlm else None
        assert ql.dtype == np.float32 and ql.size == n * len(ls) and ql.flags.c_contiguous
        assert qlm is None or (qlm.dtype == np.float32 and qlm.size == n * tot_m * 2 and qlm.flags.c_contiguous)
        wl_arr = np.empty((n, len(ls)), np.float32) if wl else None
        sys_qlm = np.empty(tot_m * 2, np.float32)
        order = np.empty(len(ls), np.float32)
        flags = (ST_WEIGHTED if weighted else 0) | (ST_AVERAGE if average else 0) | (ST_WL if wl else 0) \
            | (ST_WL_NORMALIZE if wl_normalize else 0)
        check(lib().fgpu_steinhardt_compute(self._h, nlist._h, ptr(ls, _up), len(ls), flags, int(n_total),
                                            comm._h if comm is not None else None, ptr(ql), ptr(wl_arr), ptr(qlm),
                                            ptr(sys_qlm), ptr(order)))
        out_qlm, out_sys, off, soff = [], [], 0, 0
        for l in ls:
            nm = 2 * int(l) + 1
            if want_qlm:
                blk = qlm[off:off + n * nm * 2].reshape(n, nm, 2)
                out_qlm.append(blk.view(np.complex64).reshape(n, nm))
                off += n * nm * 2
            out_sys.append(sys_qlm[soff:soff + 2 * nm].view(np.complex64))
            soff += 2 * nm
        return dict(ql=ql, wl=wl_arr, qlm=out_qlm, sys_qlm=out_sys, order=order)


class DeviceRDF(_DeviceObject):
    """Device-resident RDF histogram (``fgpu_rdf``)."""

    _destroy = "fgpu_rdf_destroy"

    def __init__(self, ctx, bins, r_max, r_min=0.0):
        self._adopt(ctx)
        self.bins = int(bins)
        self._h = _vp()
        check(lib().fgpu_rdf_create(ctx._h, self.bins, float(r_max), float(r_min), C.byref(self._h)))

    def reset(self):
        check(lib().fgpu_rdf_reset(self._h))

    def accumulate(self, points, query_points, flavour, r_max, r_min=0.0, exclude_ii=False, q_index_offset=0):
        q = None if query_points is None else f32(query_points, 3)
        nq = points.n if q is None else len(q)
        check(lib().fgpu_rdf_accumulate(self._h, points._h, ptr(q), nq, int(q_index_offset), int(flavour),
                                        float(r_max), float(r_min), int(bool(exclude_ii))))

    def accumulate_nlist(self, nlist):
        check(lib().fgpu_rdf_accumulate_nlist(self._h, nlist._h))

    def read(self):
        counts = np.empty(self.bins, np.uint32)
        check(lib().fgpu_rdf_read(self._h, ptr(counts, _up)))
        return counts

    def allreduce(self, comm):
        check(lib().fgpu_rdf_allreduce(self._h, comm._h))

    def attach_comm(self, comm):
        """Collective: give this RDF a peer-memory mailbox on every rank of ``comm`` (NVLink red.add instead of an NCCL
        launch for the sum).  True if attached, False if the ranks cannot map each other's memory (NCCL stays)."""
        rc = lib().fgpu_rdf_attach_comm(self._h, comm._h)
        if rc == 1:
            return False
        check(rc)
        self._comm_ref = comm  # the communicator must outlive the mailbox
        return True

    @property
    def reduce_transport(self):
        return {1: "nccl", 2: "peer"}[lib().fgpu_rdf_reduce_transport(self._h)]

    def accumulate_reduce(self, points, comm, flavour, r_max, r_min=0.0, exclude_ii=False):
        """Self-query accumulation of (sharded) points and the sum over the ranks in one call."""
        check(lib().fgpu_rdf_accumulate_reduce(self._h, points._h, comm._h, int(flavour), float(r_max), float(r_min),
                                               int(bool(exclude_ii))))


class DevicePMFTXY(_DeviceObject):
    """Device-resident PMFTXY histogram (``fgpu_pmftxy``)."""

    _destroy = "fgpu_pmftxy_destroy"

    def __init__(self, ctx, x_max, y_max, n_x, n_y):
        self._adopt(ctx)
        self.shape = (int(n_x), int(n_y))
        self._h = _vp()
        check(lib().fgpu_pmftxy_create(ctx._h, float(x_max), float(y_max), int(n_x), int(n_y), C.byref(self._h)))

    def reset(self):
        check(lib().fgpu_pmftxy_reset(self._h))

    def accumulate_nlist(self, nlist, query_orientations):
        """query_orientations: angles in radians, one per query point."""
        t = np.ascontiguousarray(query_orientations, dtype=np.float32).ravel()
        assert len(t) == nlist.num_query_points
        check(lib().fgpu_pmftxy_accumulate_nlist(self._h, nlist._h, ptr(t)))

    def accumulate(self, points, query_points, flavour, r_max, query_orientations, r_min=0.0, exclude_ii=False):
        """Ball query of ``points`` (``query_points=None``: against themselves) and the histogram in one call -- no
        NeighborList (``fgpu_pmftxy_accumulate``)."""
        q = None if query_points is None else f32(query_points, 3)
        nq = points.n if q is None else len(q)
        t = np.ascontiguousarray(query_orientations, dtype=np.float32).ravel()
        assert len(t) == nq
        check(lib().fgpu_pmftxy_accumulate(self._h, points._h, ptr(q), nq, int(flavour), float(r_max), float(r_min),
                                           int(bool(exclude_ii)), ptr(t)))

    def read(self):
        counts = np.empty(self.shape, np.uint32)
        check(lib().fgpu_pmftxy_read(self._h, ptr(counts, _up)))
        return counts


PMFT_XYZ, PMFT_XYT, PMFT_R12 = 0, 1, 2


class DevicePMFT(_DeviceObject):
    """Device-resident PMFTXYZ / PMFTXYT / PMFTR12 histogram (``fgpu_pmft``); ``maxes`` and ``bins`` as the reference's
    constructors order them (unused maxima 0)."""

    _destroy = "fgpu_pmft_destroy"

    def __init__(self, ctx, kind, maxes, bins):
        self._adopt(ctx)
        self.kind, self.shape = int(kind), tuple(int(b) for b in bins)
        mx = [float(m) for m in maxes] + [0.0] * (3 - len(maxes))
        self._h = _vp()
        check(lib().fgpu_pmft_create(ctx._h, self.kind, mx[0], mx[1], mx[2], *self.shape, C.byref(self._h)))

    def reset(self):
        check(lib().fgpu_pmft_reset(self._h))

    def accumulate_nlist(self, nlist, orientations, query_orientations, equiv_orientations=None):
        """Angles (XYT, R12) or (N, 4) quaternions (XYZ: ``orientations`` is ignored, ``equiv_orientations`` required)."""
        qo = np.ascontiguousarray(query_orientations, dtype=np.float32)
        if self.kind == PMFT_XYZ:
            eq = np.ascontiguousarray(equiv_orientations, dtype=np.float32).reshape(-1, 4)
            assert qo.shape == (nlist.num_query_points, 4)
            check(lib().fgpu_pmft_accumulate_nlist(self._h, nlist._h, None, nlist.num_points, ptr(qo), ptr(eq), len(eq)))
            return
        o = np.ascontiguousarray(orientations, dtype=np.float32).ravel()
        assert len(o) == nlist.num_points and qo.size == nlist.num_query_points
        check(lib().fgpu_pmft_accumulate_nlist(self._h, nlist._h, ptr(o), len(o), ptr(qo), None, 0))

    def accumulate(self, points, query_points, flavour, r_max, orientations, query_orientations, equiv_orientations=None,
                   r_min=0.0, exclude_ii=False):
        """Ball query of ``points`` (``query_points=None``: against themselves) and the histogram in one call -- no
        NeighborList (``fgpu_pmft_accumulate``); orientation arguments as ``accumulate_nlist``."""
        q = None if query_points is None else f32(query_points, 3)
        nq = points.n if q is None else len(q)
        qo = np.ascontiguousarray(query_orientations, dtype=np.float32)
        o = eq = None
        if self.kind == PMFT_XYZ:
            eq = np.ascontiguousarray(equiv_orientations, dtype=np.float32).reshape(-1, 4)
            assert qo.shape == (nq, 4)
        else:
            o = np.ascontiguousarray(orientations, dtype=np.float32).ravel()
            assert len(o) == points.n and qo.size == nq
        check(lib().fgpu_pmft_accumulate(self._h, points._h, ptr(q), nq, int(flavour), float(r_max), float(r_min),
                                         int(bool(exclude_ii)), ptr(o) if o is not None else None, ptr(qo),
                                         ptr(eq) if eq is not None else None, 0 if eq is None else len(eq)))

    def read(self):
        counts = np.empty(self.shape, np.uint32)
        check(lib().fgpu_pmft_read(self._h, ptr(counts, _up)))
        return counts

    @property
    def host_binned_bonds(self):
        n = C.c_uint64(0)
        check(lib().fgpu_pmft_deferred(self._h, C.byref(n)))
        return n.value


BOND_ORDER_MODES = {"bod": 0, "lbod": 1, "obcd": 2, "oocd": 3}


class DeviceBondOrder(_DeviceObject):
    """Device-resident BondOrder histogram (``fgpu_bondorder``)."""

    _destroy = "fgpu_bondorder_destroy"

    def __init__(self, ctx, n_theta, n_phi, mode="bod"):
        self._adopt(ctx)
        self.shape = (int(n_theta), int(n_phi))
        self._h = _vp()
        check(lib().fgpu_bondorder_create(ctx._h, self.shape[0], self.shape[1], BOND_ORDER_MODES.get(mode, -1),
                                          C.byref(self._h)))

    def reset(self):
        check(lib().fgpu_bondorder_reset(self._h))

    def accumulate_nlist(self, nlist, orientations=None, query_orientations=None):
        """(N, 4) quaternions of the points and of the query points; mode 'bod' needs neither."""
        o = qo = None
        if orientations is not None:
            o = np.ascontiguousarray(orientations, dtype=np.float32).reshape(-1, 4)
            qo = np.ascontiguousarray(query_orientations, dtype=np.float32).reshape(-1, 4)
            assert len(o) == nlist.num_points and len(qo) == nlist.num_query_points
        check(lib().fgpu_bondorder_accumulate_nlist(self._h, nlist._h, ptr(o) if o is not None else None,
                                                    nlist.num_points, ptr(qo) if qo is not None else None))

    def accumulate(self, points, query_points, flavour, r_max, orientations=None, query_orientations=None, r_min=0.0,
                   exclude_ii=False):
        """Ball query of ``points`` (``query_points=None``: against themselves) and the histogram in one call -- no
        NeighborList (``fgpu_bondorder_accumulate``)."""
        q = None if query_points is None else f32(query_points, 3)
        nq = points.n if q is None else len(q)
        o = qo = None
        if orientations is not None:
            o = np.ascontiguousarray(orientations, dtype=np.float32).reshape(-1, 4)
            qo = np.ascontiguousarray(query_orientations, dtype=np.float32).reshape(-1, 4)
            assert len(o) == points.n and len(qo) == nq
        check(lib().fgpu_bondorder_accumulate(self._h, points._h, ptr(q), nq, int(flavour), float(r_max), float(r_min),
                                              int(bool(exclude_ii)), ptr(o) if o is not None else None,
                                              ptr(qo) if qo is not None else None))

    def read(self):
        counts = np.empty(self.shape, np.uint32)
        check(lib().fgpu_bondorder_read(self._h, ptr(counts, _up)))
        return counts

    @property
    def host_binned_bonds(self):
        n = C.c_uint64(0)
        check(lib().fgpu_bondorder_deferred(self._h, C.byref(n)))
        return n.value


class DeviceCorrelation(_DeviceObject):
    """Device-resident CorrelationFunction accumulators (``fgpu_corr``)."""

    _destroy = "fgpu_corr_destroy"

    def __init__(self, ctx, bins, r_max):
        self._adopt(ctx)
        self.bins = int(bins)
        self._h = _vp()
        check(lib().fgpu_corr_create(ctx._h, self.bins, float(r_max), C.byref(self._h)))

    def reset(self):
        check(lib().fgpu_corr_reset(self._h))

    def accumulate_nlist(self, nlist, values, query_values):
        v = np.ascontiguousarray(values, dtype=np.complex128).ravel()
        q = v if query_values is values else np.ascontiguousarray(query_values, dtype=np.complex128).ravel()
        assert len(v) == nlist.num_points and len(q) == nlist.num_query_points
        dp = C.POINTER(C.c_double)
        check(lib().fgpu_corr_accumulate_nlist(self._h, nlist._h, v.ctypes.data_as(dp), q.ctypes.data_as(dp)))

    def accumulate(self, points, query_points, flavour, r_max, values, query_values, r_min=0.0, exclude_ii=False):
        """Ball query of ``points`` (``query_points=None``: against themselves) and the accumulation in one call -- no
        NeighborList (``fgpu_corr_accumulate``)."""
        qp = None if query_points is None else f32(query_points, 3)
        nq = points.n if qp is None else len(qp)
        v = np.ascontiguousarray(values, dtype=np.complex128).ravel()
        q = v if query_values is values else np.ascontiguousarray(query_values, dtype=np.complex128).ravel()
        assert len(v) == points.n and len(q) == nq
        dp = C.POINTER(C.c_double)
        check(lib().fgpu_corr_accumulate(self._h, points._h, ptr(qp), nq, int(flavour), float(r_max), float(r_min),
                                         int(bool(exclude_ii)), v.ctypes.data_as(dp), q.ctypes.data_as(dp)))

    def read(self):
        counts = np.empty(self.bins, np.uint32)
        sums = np.empty(self.bins, np.complex128)
        check(lib().fgpu_corr_read(self._h, ptr(counts, _up), sums.ctypes.data_as(C.POINTER(C.c_double))))
        return counts, sums


class Communicator(_DeviceObject):
    """NCCL communicator, one rank per process (``fgpu_comm``)."""

    _destroy = "fgpu_comm_destroy"

    def __init__(self, ctx, unique_id, rank, size):
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert len(uid) == UNIQUE_ID_BYTES
        self._adopt(ctx)
        self._h = _vp()
        check(lib().fgpu_comm_create(ctx._h, uid.ctypes.data_as(C.POINTER(C.c_uint8)), int(rank), int(size),
                                     C.byref(self._h)))
        self.rank, self.size = int(rank), int(size)

    @staticmethod
    def unique_id():
        uid = np.zeros(UNIQUE_ID_BYTES, np.uint8)
        check(lib().fgpu_comm_unique_id(uid.ctypes.data_as(C.POINTER(C.c_uint8))))
        return uid.tobytes()

    def barrier(self):
        check(lib().fgpu_comm_barrier(self._h))

    def allreduce_u32(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.uint32)
        check(lib().fgpu_comm_allreduce_u32(self._h, ptr(a, _up), a.size))
        return a

    def allreduce_f64(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        check(lib().fgpu_comm_allreduce_f64(self._h, a.ctypes.data_as(C.POINTER(C.c_double)), a.size))
        return a

    def close(self):
        self._release()
