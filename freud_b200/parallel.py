"""Host-side plumbing of the multi-GPU path: one process per GPU, points replicated, query points or frames sharded.

SURVEY.md section 8e: the reference parallelises over query points inside one process
(``freud/locality/NeighborQuery.h:438``, ``freud/locality/NeighborComputeFunctional.h:201-217``) and accumulates
frames with ``compute(..., reset=False)`` (``freud/density.py:632-638``); both axes shard without any data-path
exchange, and the only collective is one sum of ``u32[bins]`` at reduce time.  This module holds the pieces that
are independent of the device: shard arithmetic, merging per-rank NeighborList slices, and the bin-count reduction
(NCCL on the library's stream in production; ``torch.distributed`` host tensors for the gloo tests).
"""

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block of indices ``[lo, hi)`` owned by ``rank``; blocks tile ``range(n)`` in rank order, so the
    per-rank NeighborList slices concatenate into the globally ``(i, j)``-sorted list without a merge."""
    if not (0 <= rank < world):
        raise ValueError("rank must satisfy 0 <= rank < world")
    return (rank * n) // world, ((rank + 1) * n) // world


def frames_of_rank(n_frames, rank, world):
    """Round-robin frame assignment (BASELINE.json configs[4])."""
    if not (0 <= rank < world):
        raise ValueError("rank must satisfy 0 <= rank < world")
    return list(range(rank, n_frames, world))


def merge_nlist_shards(parts, offsets):
    """Concatenate per-rank NeighborList dictionaries (``DeviceNeighborList.to_host()`` layout, local row indices)
    into the list a single GPU would have produced.  ``offsets[k]`` is the first global query index of shard k."""
    out = {}
    nb = 0
    neighbors, segments = [], []
    for part, lo in zip(parts, offsets):
        n = part["neighbors"].copy()
        n[:, 0] += np.uint32(lo)
        neighbors.append(n)
        seg = part["segments"].astype(np.uint32).copy()
        seg[part["counts"] != 0] += np.uint32(nb)  # empty rows keep 0 (NeighborList.cc:199-232)
        segments.append(seg)
        nb += len(part["distances"])
    out["neighbors"] = np.concatenate(neighbors) if neighbors else np.zeros((0, 2), np.uint32)
    out["segments"] = np.concatenate(segments) if segments else np.zeros(0, np.uint32)
    for key in ("distances", "weights", "vectors", "counts"):
        out[key] = np.concatenate([p[key] for p in parts])
    return out


def allreduce_bin_counts(counts, comm=None, group=None):
    """Sum u32 bin counts over all ranks (wrapping like the reference's unsigned counters).

    ``comm``: a ``_capi.Communicator`` -> one ncclAllReduce(u32) staged through device memory.
    otherwise ``torch.distributed`` on host tensors (gloo has no uint32 sum: reduce as int64, wrap back)."""
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    if comm is not None:
        return comm.allreduce_u32(c.copy())
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c.copy()
    t = torch.from_numpy(c.astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return (t.numpy() & 0xFFFFFFFF).astype(np.uint32)


def make_communicator(ctx):
    """NCCL communicator for the library's stream; the unique id travels over the existing torch.distributed
    process group (any backend).  Returns None for a single process."""
    import torch.distributed as dist

    from . import _capi

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [_capi.Communicator.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return _capi.Communicator(ctx, uid[0], rank, world)


def replicated_points(ctx, box, points, comm=None, rank=0, world=1):
    """DevicePoints of a frame every rank holds in host memory: with a communicator each rank uploads 1/world of it and
    the blocks travel over NVLink (``fgpu_points_create_replicated``); a single process uploads all of it."""
    from . import _capi

    return _capi.DevicePoints(ctx, box, points, comm=comm if world > 1 else None)


class ShardedRDF:
    """RDF over query-point shards (config 4) or frame shards (config 5) with one allreduce at read time.

    Mirrors ``RDF.compute(system, reset=False)`` + ``.bin_counts``: ``accumulate_frame`` adds this rank's share of a
    frame, ``bin_counts()`` performs the single exchange."""

    def __init__(self, ctx, bins, r_max, r_min=0.0, comm=None, rank=0, world=1, transport="auto"):
        """transport: "auto" = the peer mailbox over NVLink where the ranks can map each other's memory, else NCCL;
        "nccl" = always ncclAllReduce."""
        from . import _capi

        self.rdf = _capi.DeviceRDF(ctx, bins, r_max, r_min)
        self.ctx, self.comm, self.rank, self.world = ctx, comm, rank, world
        if comm is not None and transport == "auto":
            self.rdf.attach_comm(comm)  # collective; falls back to NCCL where the ranks cannot map each other's memory
        self._reduced = True
        # True: the points stay sharded between calls (a loop over frames of the same DevicePoints keeps the slab's
        # cell list); the caller then restores points.set_shard(0, 1) itself before any other query
        self.keep_shard = False

    def reset(self):
        self.rdf.reset()
        self._reduced = True

    def reduce_kind(self):
        if self.comm is None:
            return "single GPU: no exchange"
        return ("peer mailbox: red.add over NVLink from the search kernel's last block + one-block wait"
                if self.rdf.reduce_transport == "peer" else "ncclAllReduce(u32[bins])")

    def reduce(self):
        """The exchange step, enqueued on the stream: afterwards ``bin_counts()`` is only a D2H."""
        if not self._reduced and self.comm is not None:
            self.rdf.allreduce(self.comm)
        self._reduced = True

    def accumulate_frame(self, points, flavour, r_max, r_min=0.0, exclude_ii=True, query_shard=None, reduce=False):
        """points: DevicePoints (replicated).  reduce=True (``query_shard="tiles"`` or None): the sum over the ranks
        is part of the same call -- with a peer mailbox the search kernel's last block sends the counts itself.
        query_shard:

        * ``None``: all points of the frame are queries (frame sharding, or a single GPU);
        * ``"tiles"``: self query whose home tiles are dealt to the ranks (``fgpu_points_set_shard``): this rank
          searches its share and builds only the slab of the cell list that share can see -- config 4's path;
        * ``(host array, first index)``: an explicit block of query points with ``q_index_offset``."""
        fused = reduce and self.comm is not None and (query_shard is None or isinstance(query_shard, str))

        def run():
            if fused:
                self.rdf.accumulate_reduce(points, self.comm, flavour, r_max, r_min, exclude_ii)
            else:
                self.rdf.accumulate(points, None, flavour, r_max, r_min, exclude_ii)

        if isinstance(query_shard, str) and query_shard == "tiles":
            points.set_shard(self.rank, self.world)
            try:
                run()
            finally:
                if not self.keep_shard:
                    points.set_shard(0, 1)  # the shard is a property of this call, not of the points
        elif query_shard is None:
            run()
        else:
            q, lo = query_shard
            self.rdf.accumulate(points, q, flavour, r_max, r_min, exclude_ii, q_index_offset=lo)
        self._reduced = fused
        if reduce and not fused:
            self.reduce()

    def bin_counts(self):
        """Counts summed over the ranks.  The reduction is out of place (``fgpu_rdf_allreduce``): this rank's own
        histogram keeps only its own frames, so accumulate -> bin_counts -> accumulate -> bin_counts counts every
        frame once."""
        if not self._reduced and self.comm is not None:
            self.rdf.allreduce(self.comm)  # ncclAllReduce(u32[bins]) on the library's stream
        self._reduced = True
        return self.rdf.read()
