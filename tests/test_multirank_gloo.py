"""world_size = 2 on CPU (gloo): the host side of the multi-GPU path -- shard arithmetic, q_index_offset semantics,
NeighborList slice merging and the bin-count reduction -- with the oracle standing in for the per-rank device work.
The same functions drive the NCCL path in bench.py (--gpus N)."""
import os
import socket

import numpy as np
import pytest

from freud_b200 import parallel
from oracle import port
from tests.util import BOXES, random_points


def test_shard_bounds_tile_the_range():
    for n in (0, 1, 7, 1000, 4_000_000):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    assert sorted(sum((parallel.frames_of_rank(64, r, 8) for r in range(8)), [])) == list(range(64))
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box, n, r = BOXES["tri1"]
    pts = random_points(box, n, seed=51)
    lo, hi = parallel.shard_bounds(n, rank, world)
    # per-rank work: this rank's query shard against the replicated points (the oracle plays the GPU)
    part = port.ball_nlist(port.IMAGE, box, False, pts, pts[lo:hi], r, 0.0, False)
    keep = part.neighbors[:, 0].astype(np.int64) + lo != part.neighbors[:, 1]  # exclude_ii with q_index_offset = lo
    counts_local = port.rdf_accumulate_distances(part.distances[keep], 60, r)
    total = parallel.allreduce_bin_counts(counts_local)
    # frame sharding (reset=False accumulation): each rank owns frames rank, rank + world, ...
    frame_counts = np.zeros(60, np.uint32)
    for f in parallel.frames_of_rank(4, rank, world):
        fp = random_points(box, 500, seed=100 + f)
        frame_counts = port.rdf_accumulate(port.IMAGE, box, False, fp, fp, 60, r, 0.0, True, counts=frame_counts)
    frame_total = parallel.allreduce_bin_counts(frame_counts)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), total=total, frame_total=frame_total,
             neighbors=part.neighbors[keep], distances=part.distances[keep], lo=lo)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_result(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    box, n, r = BOXES["tri1"]
    pts = random_points(box, n, seed=51)
    want = port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 60, r, 0.0, True)
    full = port.ball_nlist(port.IMAGE, box, False, pts, pts, r, 0.0, True)
    want_frames = np.zeros(60, np.uint32)
    for f in range(4):
        fp = random_points(box, 500, seed=100 + f)
        want_frames = port.rdf_accumulate(port.IMAGE, box, False, fp, fp, 60, r, 0.0, True, counts=want_frames)
    ranks = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for d in ranks:  # every rank holds the full sum after the exchange
        assert np.array_equal(d["total"], want)
        assert np.array_equal(d["frame_total"], want_frames)
    # contiguous shards concatenate to the globally sorted list (no merge step)
    nbrs = []
    for d in ranks:
        nb = d["neighbors"].copy()
        nb[:, 0] += np.uint32(d["lo"])
        nbrs.append(nb)
    assert np.array_equal(np.concatenate(nbrs), full.neighbors)
    assert np.array_equal(np.concatenate([d["distances"] for d in ranks]).view(np.uint32),
                          full.distances.view(np.uint32))


def test_merge_nlist_shards_segments():
    box, n, r = BOXES["cubic"]
    pts = random_points(box, 600, seed=52)
    full = port.ball_nlist(port.WRAP, box, False, pts, pts, 1.2, 0.0, True)  # sparse: some rows are empty
    parts, offs = [], []
    for rank in range(3):
        lo, hi = parallel.shard_bounds(len(pts), rank, 3)
        p = port.ball_nlist(port.WRAP, box, False, pts, pts[lo:hi], 1.2, 0.0, False)
        keep = p.neighbors[:, 0].astype(np.int64) + lo != p.neighbors[:, 1]
        counts = np.bincount(p.neighbors[keep, 0], minlength=hi - lo).astype(np.uint32)
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
        parts.append(dict(neighbors=p.neighbors[keep], distances=p.distances[keep], weights=p.weights[keep],
                          vectors=p.vectors[keep], counts=counts, segments=np.where(counts != 0, starts, 0).astype(np.uint32)))
        offs.append(lo)
    merged = parallel.merge_nlist_shards(parts, offs)
    assert np.array_equal(merged["neighbors"], full.neighbors)
    assert np.array_equal(merged["segments"], full.segments) and np.array_equal(merged["counts"], full.counts)
    assert (full.counts == 0).any()


def test_home_tile_shards_tile_the_grid():
    """fgpu_shard_plan (host arithmetic of fgpu_points_set_shard, no device): the shards' ticket and cell ranges tile
    the grid in order, and every shard's slab holds the layers of its cells plus one halo layer on each side."""
    from freud_b200 import _capi

    for dims, n_points in (((73, 73, 73), 4_000_000), ((77, 77, 77), 1_000_000), ((282, 282, 1), 1_000_000),
                           ((5, 4, 3), 500), ((9, 7, 1), 300)):
        n_cells = dims[0] * dims[1] * dims[2]
        for shards in (1, 2, 3, 8):
            plans = [_capi.shard_plan(dims, n_points, s, shards) for s in range(shards)]
            assert plans[0]["ticket_begin"] == 0 and plans[-1]["ticket_end"] == plans[0]["n_tickets"]
            assert plans[0]["cell_begin"] == 0 and plans[-1]["cell_end"] == n_cells
            for a, b in zip(plans, plans[1:]):
                assert a["ticket_end"] == b["ticket_begin"] and a["cell_end"] == b["cell_begin"]
            for p in plans:
                if p["cell_end"] == p["cell_begin"]:
                    continue
                axis = p["slab_axis"]
                assert axis == (2 if dims[2] > 1 else 1)
                layers = dims[axis]
                per_layer = dims[0] * dims[1] if axis == 2 else dims[0]
                first = (p["cell_begin"] // per_layer) % layers if axis == 1 else p["cell_begin"] // per_layer
                last = ((p["cell_end"] - 1) // per_layer) % layers if axis == 1 else (p["cell_end"] - 1) // per_layer
                if p["slab_len"] is None or shards == 1:
                    continue  # every layer is kept
                need = {(first - 1) % layers, (last + 1) % layers} | {x % layers for x in range(first, last + 1)}
                have = {(p["slab_lo"] + k) % layers for k in range(p["slab_len"])}
                assert need <= have, (dims, shards, p)
