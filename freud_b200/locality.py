"""``freud.locality`` surface of the neighbour-query path, on the C++ host classes (``_freud_b200._locality``).

Mirrors the reference's Python layer for this path only: ``NeighborQuery.from_system`` / ``query`` /
``NeighborQueryResult.toNeighborList`` (``freud/locality.py:201-229, 268-414``), ``AABBQuery`` / ``LinkCell``
/ ``_RawPoints`` (``:839-921``), ``NeighborList`` (``:417-836``), query-argument dictionaries (``:39-175``) and
the ``_PairCompute`` argument resolution (``:924-1016``).  The objects held in ``_cpp_obj`` expose the same
C++ methods the reference's nanobind modules do, so this file reads like upstream's; everything numerical
happens in ``libfreud_b200.so`` on the GPU -- there is no CPU fallback.
"""

import numpy as np

from .box import Box

_VALID_QUERY_KEYS = ("mode", "r_min", "r_max", "r_guess", "num_neighbors", "exclude_ii", "scale")


def _ext():
    try:
        from . import _freud_b200
    except ImportError as exc:  # loud, never a fallback
        raise ImportError("freud_b200._freud_b200 is not built: run `python -c \"import __graft_entry__ as g; "
                          "g.build()\"` (make -C freud_b200/csrc && make -C freud_b200/host)") from exc
    return _freud_b200


def _cpp_box(box):
    b = Box.from_box(box)
    return _ext()._box.Box(b.Lx, b.Ly, b.Lz, b.xy, b.xz, b.yz, b.is2D)


def _points(a, name="points"):
    """freud.util._convert_array(shape=(None, 3), dtype=float32) semantics (freud/util.py:74-137)."""
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must have shape (N, 3), got {a.shape}")
    return np.ascontiguousarray(a, dtype=np.float32)


def _private_points(a):
    """The engine's own copy of the points (freud/locality.py:867-868), in page-locked memory from the library's host
    cache so that the upload behind it runs at the link's rate."""
    return _ext().private_points(_points(a))


def _query_args(d):
    """dict -> C++ QueryArgs (freud/locality.py:39-175: unknown keys are an error, mode is a string)."""
    L = _ext()._locality
    qa = L.QueryArgs()
    for key, val in dict(d).items():
        if key not in _VALID_QUERY_KEYS:
            raise ValueError(f"You have passed an invalid query argument: {key}")
        if key == "mode":
            if val is None or val == "none":
                qa.mode = L.QueryType.none
            elif val == "ball":
                qa.mode = L.QueryType.ball
            elif val == "nearest":
                qa.mode = L.QueryType.nearest
            else:
                raise ValueError("You have passed an invalid mode.")
        elif key == "num_neighbors":
            qa.num_neighbors = int(val)
        elif key == "exclude_ii":
            qa.exclude_ii = bool(val)
        else:
            setattr(qa, key, float(val))
    return qa


def _class_paths(obj):
    return {f"{c.__module__}.{c.__name__}" for c in type(obj).__mro__}


def _system_pair(system):
    """(box-like, positions) of a system-like object; the cases of freud/locality.py:313-372."""
    paths = _class_paths(system)
    if paths & {"MDAnalysis.coordinates.base.Timestep", "MDAnalysis.coordinates.timestep.Timestep"}:
        return system.triclinic_dimensions, system.positions
    if paths & {"gsd.hoomd.Frame", "gsd.hoomd.Snapshot", "hoomd.snapshot.Snapshot"}:
        # HOOMD writes Lz = 1 for 2-D boxes: the configuration's dimensions decide, not Lz
        box = np.array(system.configuration.box, dtype=np.float64)
        if system.configuration.dimensions == 2:
            box[[2, 4, 5]] = 0
        return box, system.particles.position
    if "garnett.trajectory.Frame" in paths:
        return system.box, system.position if hasattr(system, "position") else system.positions
    if paths & {"ovito.data.DataCollection", "ovito.plugins.PyScript.DataCollection", "PyScript.DataCollection"}:
        cell = system.cell
        return Box.from_box(np.asarray(cell.matrix)[:, :3], dimensions=2 if cell.is2D else 3), system.particles.positions
    if hasattr(system, "box") and hasattr(system, "particles") and hasattr(system.particles, "position"):
        box = system.box  # HOOMD-blue 2 snapshot
        if getattr(box, "dimensions", 3) == 2:
            box = Box(box.Lx, box.Ly, xy=getattr(box, "xy", 0), is2D=True)
        return box, system.particles.position
    if hasattr(system, "box") and hasattr(system, "points"):
        return system.box, system.points
    try:
        box, points = system
    except (TypeError, ValueError) as exc:
        raise ValueError("Cannot interpret the system: expected a NeighborQuery, a (box, points) pair, an object "
                         "with box and points, or a frame of a supported reader.") from exc
    return box, points


class NeighborQueryResult:
    """Lazy result of ``NeighborQuery.query`` (freud/locality.py:178-229)."""

    def __init__(self, nq, query_points, query_args):
        self._nq = nq
        self._query_points = query_points
        self._query_args = query_args

    def _iterator(self):
        return self._nq._cpp_obj.query(self._query_points, self._query_args)

    def __iter__(self):
        it = self._iterator()
        term = _ext()._locality.get_iterator_terminator()
        while True:
            bond = it.next()
            if bond[:2] == term[:2]:
                return
            yield (bond[0], bond[1], bond[2])

    def toNeighborList(self, sort_by_distance=False):
        return NeighborList._from_cpp(self._iterator().toNeighborList(bool(sort_by_distance)))


class NeighborQuery:
    """Base of the query engines (freud/locality.py:232-414)."""

    def __init__(self):
        raise RuntimeError("The NeighborQuery class is abstract, and should not be instantiated directly.")

    @classmethod
    def from_system(cls, system, dimensions=None):
        """Anything system-like becomes a NeighborQuery (freud/locality.py:268-383): a NeighborQuery is returned as it
        is; otherwise a box and an (N, 3) position array are taken from, in this order, an MDAnalysis ``Timestep``,
        a GSD / HOOMD-blue 3 frame or snapshot, a garnett ``Frame``, an OVITO ``DataCollection``, a HOOMD-blue 2
        snapshot, any object with ``box`` and ``points``, or a ``(box, points)`` pair.  Foreign types are recognised
        by class path, so none of those packages is imported here."""
        if isinstance(system, cls):
            return system
        if isinstance(system, NeighborQuery) and cls is not NeighborQuery:
            system = (system.box, system.points)  # another engine over the same data
        else:
            system = _system_pair(system)
        box, points = system
        if dimensions is not None:
            box = Box.from_box(box, dimensions)
        return _RawPoints(box, points) if cls is NeighborQuery else cls(box, points)

    @property
    def box(self):
        return self._box

    @property
    def points(self):
        """Read-only view of the engine's private copy: the device copy is uploaded once, so an in-place edit through
        this property would silently desynchronise the two (the reference reads the host array on every query)."""
        ro = getattr(self, "_points_ro", None)
        if ro is None:
            ro = self._points.view()
            ro.setflags(write=False)
            self._points_ro = ro
        return ro

    def query(self, query_points, query_args):
        return NeighborQueryResult(self, _points(query_points, "query_points"), _query_args(query_args))


class AABBQuery(NeighborQuery):
    """freud/locality.py:852-869."""

    def __init__(self, box, points):
        self._box = Box.from_box(box)
        self._points = _private_points(points)  # private copy, as upstream (:867-868)
        self._cpp_obj = _ext()._locality.AABBQuery(_cpp_box(self._box), self._points)


class CellQuery(NeighborQuery):
    """freud/locality.py:900-921: ball queries only, in CellQuery's own (ghost-shift) float32 arithmetic."""

    def __init__(self, box, points):
        self._box = Box.from_box(box)
        self._points = _private_points(points)
        self._cpp_obj = _ext()._locality.CellQuery(_cpp_box(self._box), self._points)


class LinkCell(NeighborQuery):
    """freud/locality.py:872-921."""

    def __init__(self, box, points, cell_width=0):
        self._box = Box.from_box(box)
        self._points = _private_points(points)
        self._cpp_obj = _ext()._locality.LinkCell(_cpp_box(self._box), self._points, float(cell_width))

    @property
    def cell_width(self):
        return self._cpp_obj.getCellWidth()


class _RawPoints(NeighborQuery):
    """freud/locality.py:839-849: what a ``(box, points)`` tuple becomes."""

    def __init__(self, box, points):
        self._box = Box.from_box(box)
        self._points = _private_points(points)
        self._cpp_obj = _ext()._locality.RawPoints(_cpp_box(self._box), self._points)


class NeighborList:
    """freud/locality.py:417-836 (container surface)."""

    def __init__(self):
        self._cpp_obj = _ext()._locality.NeighborList()

    @classmethod
    def _from_cpp(cls, obj):
        self = cls.__new__(cls)
        self._cpp_obj = obj
        return self

    @classmethod
    def from_arrays(cls, num_query_points, num_points, query_point_indices, point_indices, vectors, weights=None):
        qi = np.ascontiguousarray(query_point_indices, dtype=np.uint32)
        pi = np.ascontiguousarray(point_indices, dtype=np.uint32)
        v = _points(vectors, "vectors")
        if not (len(qi) == len(pi) == len(v)):
            raise ValueError("query_point_indices, point_indices and vectors must have the same length.")
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        return cls._from_cpp(_ext()._locality.NeighborList(qi, int(num_query_points), pi, int(num_points), v, w))

    @classmethod
    def all_pairs(cls, system, query_points=None, exclude_ii=True):
        """Every (query point, point) pair as a bond (freud/locality.py:558-600, NeighborList.cc:84-128); O(N^2), for small
        systems.  Vectors are ``box.wrap(query_points[i] - points[j])`` as upstream writes them."""
        nq = NeighborQuery.from_system(system)
        qp = nq.points if query_points is None else _points(query_points, "query_points")
        return cls._from_cpp(_ext()._locality.NeighborList(nq.points, qp, _cpp_box(nq.box), bool(exclude_ii)))

    def __len__(self):
        return self._cpp_obj.getNumBonds()

    def __getitem__(self, key):
        return self._cpp_obj.getNeighbors()[key]

    query_point_indices = property(lambda self: self._cpp_obj.getNeighbors()[:, 0])
    point_indices = property(lambda self: self._cpp_obj.getNeighbors()[:, 1])
    weights = property(lambda self: self._cpp_obj.getWeights())
    distances = property(lambda self: self._cpp_obj.getDistances())
    vectors = property(lambda self: self._cpp_obj.getVectors())
    segments = property(lambda self: self._cpp_obj.getSegments())
    neighbor_counts = property(lambda self: self._cpp_obj.getCounts())
    num_query_points = property(lambda self: self._cpp_obj.getNumQueryPoints())
    num_points = property(lambda self: self._cpp_obj.getNumPoints())

    def copy(self, other=None):
        if other is not None:
            self._cpp_obj.copy(other._cpp_obj)
            return self
        new = NeighborList()
        new._cpp_obj.copy(self._cpp_obj)
        return new

    def find_first_index(self, i):
        return self._cpp_obj.find_first_index(int(i))

    def filter(self, filt):
        self._cpp_obj.filter(np.ascontiguousarray(filt, dtype=bool))
        return self

    def filter_r(self, r_max, r_min=0):
        self._cpp_obj.filter_r(float(r_max), float(r_min))
        return self

    def sort(self, by_distance=False):
        self._cpp_obj.sort(bool(by_distance))
        return self


def _computed(fn):
    """Property that raises AttributeError until compute() ran (``_Compute._computed_property``, freud/util.py:61-80)."""

    def getter(self):
        if not self._called_compute:
            raise AttributeError("Property not computed. Call compute first.")
        return fn(self)

    getter.__doc__ = fn.__doc__
    return property(getter)


class _PairCompute:
    """Argument resolution shared by the computes (freud/locality.py:924-1016); ``compute`` of a subclass sets
    ``_called_compute`` once it ran (freud/util.py:39-59)."""

    _called_compute = False

    def _preprocess_arguments(self, system, query_points=None, neighbors=None):
        nq = NeighborQuery.from_system(system)
        if query_points is None:
            query_points = nq.points
        else:
            query_points = _points(query_points, "query_points")
        nlist, qargs = self._resolve_neighbors(neighbors, query_points is nq.points)
        return nq, nlist, qargs, query_points

    def _resolve_neighbors(self, neighbors, self_query):
        if isinstance(neighbors, NeighborList):
            return neighbors._cpp_obj, _ext()._locality.QueryArgs()
        args = dict(self.default_query_args if neighbors is None else neighbors)
        args.setdefault("exclude_ii", self_query)  # freud/locality.py:982
        return None, _query_args(args)

    @property
    def default_query_args(self):
        raise NotImplementedError(f"The {type(self).__name__} class does not provide default query arguments. "
                                  "You must either provide query arguments or a neighbor list to this compute method.")


class PeriodicBuffer:
    """Replicates points across the periodic boundaries (freud/locality.py:1080-1156, PeriodicBuffer.cc:23-117);
    host-side input preparation, float32 with one rounding per operation in the reference's order.

    ``images=True``: ``buffer`` counts whole images appended on the +x/+y/+z side of the box and every replica is
    wrapped into the grown box; otherwise ``buffer`` is a distance by which the box grows on every side and the
    replicas inside the grown box are kept.  ``buffer_ids`` names the input point of every buffer point."""

    def compute(self, system, buffer, images=False, include_input_points=False):
        nq = NeighborQuery.from_system(system)
        if np.ndim(buffer) == 0:
            buffer = [buffer] * 3
        elif len(buffer) != 3:
            raise ValueError("buffer must be a scalar or have length 3.")
        buff = np.asarray(buffer, dtype=np.float32)
        for axis, b in zip("xyz", buff):
            if b < 0:
                raise ValueError(f"Buffer {axis} distance must be non-negative.")
        box, f32 = nq.box, np.float32
        L = box.L.astype(f32)
        if images:
            reps = np.ceil(buff).astype(np.int64)
            grown = [f32(1 + reps[d]) * L[d] for d in range(3)]
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                reps = np.ceil(buff / L)
            reps = np.where(np.isfinite(reps), reps, 0).astype(np.int64)
            grown = [L[d] + f32(2) * buff[d] for d in range(3)]
        if box.is2D:
            reps[2] = 0
        self._buffer_box = Box(grown[0], grown[1], grown[2], box.xy, box.xz, box.yz, is2D=box.is2D)
        lo = np.zeros(3, dtype=np.int64) if images else -reps
        shifts = np.array([(i, j, k) for i in range(lo[0], reps[0] + 1) for j in range(lo[1], reps[1] + 1)
                           for k in range(lo[2], reps[2] + 1)], dtype=np.int64)
        if not include_input_points:
            shifts = shifts[np.any(shifts != 0, axis=1)]
        a1 = np.array([L[0], 0, 0], dtype=f32)
        a2 = np.array([L[1] * f32(box.xy), L[1], 0], dtype=f32)
        a3 = np.array([L[2] * f32(box.xz), L[2] * f32(box.yz), L[2]], dtype=f32)
        pts = nq.points.astype(f32)
        n, m = len(pts), len(shifts)
        out = np.repeat(pts, m, axis=0)  # point-major, images inner (PeriodicBuffer.cc:69-77)
        sh = np.tile(shifts, (n, 1)).astype(f32)
        out = out + sh[:, 0:1] * a1
        out = out + sh[:, 1:2] * a2
        if not box.is2D:
            out = out + sh[:, 2:3] * a3
        ids = np.repeat(np.arange(n, dtype=np.uint32), m)
        if images:
            out = self._buffer_box.wrap(out) if len(out) else out
        else:
            frac = self._buffer_box.make_fractional(out) if len(out) else np.zeros((0, 3), f32)
            keep = (frac[:, 0] >= 0) & (frac[:, 0] < 1) & (frac[:, 1] >= 0) & (frac[:, 1] < 1)
            if not box.is2D:
                keep &= (frac[:, 2] >= 0) & (frac[:, 2] < 1)
            out, ids = out[keep], ids[keep]
        self._buffer_points, self._buffer_ids = np.ascontiguousarray(out, dtype=f32), ids
        return self

    def _result(self, name):
        if not hasattr(self, name):
            raise AttributeError("PeriodicBuffer: call compute() first.")
        return getattr(self, name)

    buffer_points = property(lambda self: self._result("_buffer_points"))
    buffer_ids = property(lambda self: self._result("_buffer_ids"))
    buffer_box = property(lambda self: self._result("_buffer_box"))

    def __repr__(self):
        return "freud_b200.locality.PeriodicBuffer()"
