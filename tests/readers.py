"""TEST INFRASTRUCTURE ONLY -- minimal trajectory readers for the reference's own golden files.

The reference's integration / validation tests read `lj.gsd`, `lj.dcd` and `Test_Configuration.gsd` through the
`gsd` and `MDAnalysis` packages (tests/integration/test_reader_integrations.py:27-60,
tests/validation/test_steinhardt_average.py:14-22); neither is installed in this image, so the two on-disk formats
are parsed here from their published layouts:

* DCD (CHARMM / HOOMD-blue `hoomd.dump.dcd`): Fortran unformatted records.  Header record `CORD` + 20 int32 (frame count
  at icntrl[0], unit-cell flag at icntrl[10]), title record, atom-count record; per frame an optional unit-cell record of
  six float64 `[A, gamma, B, beta, alpha, C]` and three records of N float32 (x, y, z).
* GSD (glotzerlab/gsd file layer 1.0 / 2.0): 256-byte header (magic 0x65DF65DF65DF65DF, index location, namelist
  location, versions), a namelist of chunk names, and an index of 32-byte entries
  `{frame u64, N u64, location i64, M u32, id u16, type u8, flags u8}` that point at raw little-endian arrays.  The
  HOOMD schema stores `configuration/box` (6 x float32: Lx, Ly, Lz, xy, xz, yz), `configuration/dimensions`,
  `particles/N` and `particles/position` (N x 3 float32); a chunk missing from a frame takes frame 0's value.
"""
import struct

import numpy as np

from freud_b200.box import Box


def read_dcd(path):
    """List of (Box, positions float32 (N, 3)) for every frame of a DCD file."""
    with open(path, "rb") as f:
        raw = f.read()
    pos = 0

    def record():
        nonlocal pos
        (n,) = struct.unpack_from("<i", raw, pos)
        body = raw[pos + 4:pos + 4 + n]
        (m,) = struct.unpack_from("<i", raw, pos + 4 + n)
        assert n == m, "corrupt Fortran record"
        pos += 8 + n
        return body

    head = record()
    assert head[:4] == b"CORD", "not a DCD file"
    icntrl = struct.unpack_from("<20i", head, 4)
    n_frames, has_cell = icntrl[0], icntrl[10] != 0
    record()  # titles
    (n_atoms,) = struct.unpack("<i", record())
    frames = []
    while pos < len(raw) and (n_frames == 0 or len(frames) < n_frames):
        box = None
        if has_cell:
            a, gamma, b, beta, alpha, c = struct.unpack("<6d", record())
            # angles are stored as degrees or as cosines depending on the writer; the golden file is orthorhombic
            for ang in (alpha, beta, gamma):
                assert abs(ang - 90.0) < 1e-6 or abs(ang) < 1e-6, "only orthorhombic DCD cells are supported here"
            box = Box(a, b, c)
        xyz = [np.frombuffer(record(), dtype="<f4") for _ in range(3)]
        assert all(len(v) == n_atoms for v in xyz)
        frames.append((box, np.ascontiguousarray(np.stack(xyz, axis=1), dtype=np.float32)))
    return frames


_GSD_TYPES = {1: "<u1", 2: "<u2", 3: "<u4", 4: "<u8", 5: "<i1", 6: "<i2", 7: "<i4", 8: "<i8", 9: "<f4", 10: "<f8"}


class GsdFile:
    """Chunk-level access to a GSD file: `frames`, `chunk(frame, name)`."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self.raw = f.read()
        magic, index_loc, index_alloc, names_loc, names_alloc, schema_version, gsd_version = struct.unpack_from(
            "<QQQQQII", self.raw, 0)
        assert magic == 0x65DF65DF65DF65DF, "not a GSD file"
        self.gsd_version = (gsd_version >> 16, gsd_version & 0xFFFF)
        # namelist: version 1.0 stores fixed 64-byte entries, version 2.0 a run of null-terminated strings
        block = self.raw[names_loc:names_loc + names_alloc * 64]
        if self.gsd_version[0] >= 2:
            self.names = [s.decode() for s in block.split(b"\0") if s]
        else:
            self.names = [block[k * 64:(k + 1) * 64].split(b"\0")[0].decode() for k in range(names_alloc)]
            self.names = [s for s in self.names if s]
        self.index = {}
        self.frames = 0
        for k in range(index_alloc):
            frame, n, loc, m, cid, typ, flags = struct.unpack_from("<QQqIHBB", self.raw, index_loc + 32 * k)
            if loc == 0:
                break
            self.index[(frame, self.names[cid])] = (n, m, loc, typ)
            self.frames = max(self.frames, frame + 1)

    def chunk(self, frame, name, default=None):
        key = (frame, name) if (frame, name) in self.index else (0, name)
        if key not in self.index:
            return default
        n, m, loc, typ = self.index[key]
        dt = np.dtype(_GSD_TYPES[typ])
        a = np.frombuffer(self.raw, dtype=dt, count=n * m, offset=loc)
        return a.reshape(n, m) if m > 1 else a


def read_gsd(path):
    """List of (Box, positions float32 (N, 3)) for every frame of a HOOMD-schema GSD file."""
    g = GsdFile(path)
    frames = []
    for fr in range(g.frames):
        b = np.asarray(g.chunk(fr, "configuration/box", default=np.float32([1, 1, 1, 0, 0, 0])), dtype=np.float64).ravel()
        dims = g.chunk(fr, "configuration/dimensions", default=np.uint8([3]))
        is2d = int(np.asarray(dims).ravel()[0]) == 2
        box = Box(b[0], b[1], 0 if is2d else b[2], b[3], 0 if is2d else b[4], 0 if is2d else b[5], is2D=is2d)
        pos = np.ascontiguousarray(g.chunk(fr, "particles/position"), dtype=np.float32).reshape(-1, 3)
        frames.append((box, pos))
    return frames
