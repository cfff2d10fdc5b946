// Sum of u32 histograms over the ranks of one node through peer-mapped memory (NVLink / NVSwitch), fused into the
// kernel that produces the counts.
//
// The exchange step of the sharded RDF (BASELINE.json configs[3]) is a 2 KB sum.  Through NCCL
// it is a separate launch plus a ring/tree protocol whose latency rivals the eighth of a frame a rank computes; here
// the LAST block of the search kernel adds the rank's finished histogram straight into a mailbox in every rank's
// memory with fire-and-forget red.add over NVLink, fences, and bumps an arrival counter next to it.  A rank's
// mailbox then holds the sum over all ranks once `world` arrivals are in; k_rdf_wait (peer.cu) waits for that on the
// stream, copies the sum out and clears the mailbox for the epoch after next.
//
// Mailbox of one rank (u32 words): hist[2][bins_pad] -- one histogram per epoch parity -- then arrived[2], then the
// rank's own epoch counter (written by its k_rdf_wait only: the kernels read the parity from the device, so a step
// captured in a CUDA graph replays through the epochs without a changing launch argument).
// Why two parities suffice: an epoch ends, on every rank, with the wait for all `world` arrivals.  A peer can only
// start epoch e + 1 after it saw my arrival of epoch e, and my clearing of parity e happens (in stream order) before
// my arrival of epoch e + 1, which every peer waits for before it can touch parity e again in epoch e + 2.
#pragma once
#include <stdint.h>

namespace fgpu {

constexpr int kMaxPeers = 8;

struct PeerBox
{
    uint32_t* box[kMaxPeers]; // mailbox of every rank, mapped into this process (own rank: the local allocation)
    int world;
    int rank;
    uint32_t bins_pad; // words between the two parities' histograms
};

// parity of the epoch this rank is in (its own counter; every rank ends an epoch with exactly one k_rdf_wait)
__device__ __forceinline__ uint32_t peer_parity(const PeerBox& pb)
{
    return *(pb.box[pb.rank] + 2 * (size_t) pb.bins_pad + 2) & 1U;
}

__device__ __forceinline__ void red_add_sys(uint32_t* addr, uint32_t v)
{
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// Called by every thread of ONE block once `hist` (this rank's counts, complete) is visible to it: pushes the
// non-zero counters into every rank's mailbox and announces the arrival.  Ends with a block barrier.
__device__ __forceinline__ void peer_push_block(const PeerBox& pb, const uint32_t* hist, uint32_t bins)
{
    uint32_t const parity = peer_parity(pb);
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
    {
        uint32_t const v = __ldcg(hist + b);
        if (v != 0)
        {
            for (int p = 0; p < pb.world; ++p)
            {
                red_add_sys(pb.box[p] + (size_t) parity * pb.bins_pad + b, v);
            }
        }
    }
    __threadfence_system(); // the adds above are ordered before the arrival below for every observer
    __syncthreads();
    if (threadIdx.x < (unsigned) pb.world)
    {
        red_add_sys(pb.box[threadIdx.x] + 2 * (size_t) pb.bins_pad + parity, 1U);
    }
}

} // namespace fgpu
