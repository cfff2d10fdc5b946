"""Multi-GPU parity worker: run under torchrun with one rank per GPU (tests/test_multigpu.py spawns it when the box has
>= 2 GPUs; `gpurun --gpus N -- python -m torch.distributed.run --nproc-per-node N ... tests/multigpu_worker.py`).

Checks, against the oracle and against the single-GPU result:
  * configs[3]'s path at N = 200k: home tiles dealt to the ranks, slab-restricted cell list, the search kernel's last
    block pushing the counts into every rank's peer mailbox (fgpu_rdf_accumulate_reduce) -- bin counts bitwise;
  * the ADVICE round-1 sequence accumulate -> bin_counts -> accumulate -> bin_counts (reset=False over two frames with an
    intermediate read): every frame counted once, on both transports (peer mailbox, NCCL);
  * the frame replicated over NVLink (fgpu_points_create_replicated) gives the same counts as a full upload;
  * multi-GPU Steinhardt: rows sharded over the ranks, q_l of every rank's rows against the port, system q_lm / order
    reduced with the fp64 allreduce (fgpu_steinhardt_compute(comm != NULL)).
Every rank prints "MULTIGPU OK" on success; any failure raises.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from freud_b200 import _capi, data, parallel
    from oracle import port

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ctx = _capi.Context(local)
    comm = parallel.make_communicator(ctx)
    IMAGE = _capi.FLAVOUR_IMAGE

    n, bins, r_max = 200_000, 500, 5.0
    L = (n / 0.08) ** (1 / 3)
    frames = [data.make_random_system(L, n, seed=s, tilt=(0.3, 0.2, 0.1)) for s in (0, 1)]
    box = frames[0][0]
    want = [port.rdf_accumulate(port.IMAGE, box, False, p, p, bins, r_max, 0.0, True) for _, p in frames]

    for transport in ("peer", "nccl"):
        srdf = parallel.ShardedRDF(ctx, bins, r_max, comm=comm, rank=rank, world=world,
                                   transport="auto" if transport == "peer" else "nccl")
        if transport == "peer" and srdf.rdf.reduce_transport != "peer":
            print(f"rank {rank}: peer mailbox unavailable on this box, peer transport not exercised", flush=True)
        # one frame, search + exchange in one call
        dp = parallel.replicated_points(ctx, box, frames[0][1], comm, rank, world)
        srdf.accumulate_frame(dp, IMAGE, r_max, 0.0, True, query_shard="tiles", reduce=True)
        got = srdf.bin_counts()
        assert np.array_equal(got, want[0]), f"{transport}: sharded frame differs from the oracle"
        # several epochs back to back (the bench's loop): reset, frame, reduce
        for _ in range(5):
            srdf.reset()
            srdf.accumulate_frame(dp, IMAGE, r_max, 0.0, True, query_shard="tiles", reduce=True)
        assert np.array_equal(srdf.bin_counts(), want[0]), f"{transport}: repeated epochs"
        # reset=False over two frames with an intermediate read: nothing counted twice
        srdf.reset()
        srdf.accumulate_frame(dp, IMAGE, r_max, 0.0, True, query_shard="tiles")
        first = srdf.bin_counts()
        dp2 = _capi.DevicePoints(ctx, box, frames[1][1])  # plain upload of the whole frame
        srdf.accumulate_frame(dp2, IMAGE, r_max, 0.0, True, query_shard="tiles")
        both = srdf.bin_counts()
        assert np.array_equal(first, want[0]), f"{transport}: first read"
        assert np.array_equal(both, want[0] + want[1]), f"{transport}: accumulate -> read -> accumulate -> read"
        # frames sharded instead (configs[4]'s scheme): rank r takes frame r % 2 whole
        srdf.reset()
        mine = frames[rank % 2][1]
        srdf.accumulate_frame(_capi.DevicePoints(ctx, box, mine), IMAGE, r_max, 0.0, True, query_shard=None, reduce=True)
        n0 = sum(1 for r in range(world) if r % 2 == 0)
        assert np.array_equal(srdf.bin_counts(), (want[0].astype(np.uint64) * n0
                                                  + want[1].astype(np.uint64) * (world - n0)).astype(np.uint32))
        del srdf

    # ---- Steinhardt, rows sharded ------------------------------------------------------------------------
    fbox, fpts = data.make_fcc_system(12, sigma_noise=0.05, seed=2)
    nf = len(fpts)
    lo, hi = parallel.shard_bounds(nf, rank, world)
    dq = _capi.DevicePoints(ctx, fbox, fpts)
    nl = dq.knn_query(fpts[lo:hi], 12, exclude_ii=True, q_index_offset=lo)
    got = dq.steinhardt(nl, [4, 6], comm=comm, n_total=nf)
    full_nl = dq.knn_query(None, 12, exclude_ii=True)
    one = dq.steinhardt(full_nl, [4, 6])
    assert np.allclose(got["ql"], one["ql"][lo:hi], rtol=1e-6, atol=1e-7), "sharded q_l rows differ from the single-GPU rows"
    for a, b in zip(got["sys_qlm"], one["sys_qlm"]):
        assert np.allclose(a, b, atol=1e-6), "system q_lm after the fp64 allreduce"
    assert np.allclose(got["order"], one["order"], rtol=1e-5)
    # ... and the rows against the oracle's Steinhardt over the oracle's own kNN list (brute force, small N)
    pnl = port.knn_nlist(fbox, False, fpts, fpts, 12, exclude_ii=True)
    want_ql = port.steinhardt(fbox, False, fpts, pnl, [4, 6])["ql"]
    assert np.allclose(got["ql"], want_ql[lo:hi], rtol=1e-5, atol=1e-6), "sharded q_l rows differ from the oracle"
    ctx.synchronize()
    dist.barrier()
    print(f"MULTIGPU OK rank {rank}/{world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
