"""Shared helpers for the parity tests."""
import numpy as np

from freud_b200.box import Box

BOXES = {
    "cubic": (Box.cube(20), 3000, 2.5),
    "ortho": (Box(18, 25, 31), 3000, 3.0),
    "tri1": (Box(20, 22, 24, 0.3, 0.2, 0.1), 3000, 3.0),
    "tri2": (Box(20, 22, 24, -0.5, 0.4, -0.3), 3000, 2.9),
    "sq2d": (Box.square(40), 2000, 3.0),
    "tilt2d": (Box(40, 35, 0, 0.4, 0, 0, is2D=True), 2000, 3.0),
    "small": (Box.cube(7), 300, 3.0),
}


def random_points(box, n, seed, spill=0.0):
    """Uniform points in the box; spill > 0 pushes a fraction of them outside (un-wrapped inputs)."""
    rs = np.random.RandomState(seed)
    f = rs.random_sample((n, 3))
    if spill:
        f = f * (1 + 2 * spill) - spill
    if box.is2D:
        f[:, 2] = 0
    return box.make_absolute(f)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_nlist_equal(got, want, what=""):
    """Bit-exact comparison of all NeighborList arrays; got is a dict (to_host()), want an oracle list."""
    assert len(got["distances"]) == len(want.distances), f"{what}: bond count {len(got['distances'])} != {len(want.distances)}"
    assert np.array_equal(got["neighbors"], want.neighbors), f"{what}: neighbors differ"
    assert np.array_equal(bits(got["distances"]), bits(want.distances)), f"{what}: distances differ bitwise"
    assert np.array_equal(bits(got["vectors"]), bits(want.vectors)), f"{what}: vectors differ bitwise"
    assert np.array_equal(got["weights"], want.weights), f"{what}: weights differ"
    assert np.array_equal(got["segments"], want.segments), f"{what}: segments differ"
    assert np.array_equal(got["counts"], want.counts), f"{what}: counts differ"
