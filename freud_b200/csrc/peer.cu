// Stand-alone halves of the peer-memory histogram reduction (see peer.cuh): the push for accumulations whose kernel
// does not carry it, and the stream-ordered wait that ends every epoch.
#include "internal.h"
#include "peer.cuh"

namespace fgpu {

namespace {

__global__ void __launch_bounds__(256) k_rdf_push(PeerBox pb, const uint32_t* __restrict__ hist, uint32_t bins)
{
    peer_push_block(pb, hist, bins);
}

// Waits until every rank's counts of this epoch are in the local mailbox, hands the sum to `reduced` and clears the
// mailbox for the epoch after next.  One block.  *timeout is set (and nothing else is touched) when a peer does not
// arrive within ~10 s -- a dead rank must not hang the stream for ever.
__global__ void __launch_bounds__(256) k_rdf_wait(PeerBox pb, uint32_t bins, uint32_t* __restrict__ reduced,
                                                  int* __restrict__ timeout)
{
    __shared__ int s_ok;
    uint32_t* const mine = pb.box[pb.rank];
    uint32_t const parity = peer_parity(pb);
    uint32_t* const arrived = mine + 2 * (size_t) pb.bins_pad + parity;
    if (threadIdx.x == 0)
    {
        long long const t0 = clock64();
        int ok = 1;
        while (ld_acquire_sys(arrived) < (uint32_t) pb.world)
        {
            __nanosleep(200);
            if (clock64() - t0 > 20000000000LL) // ~10 s at 2 GHz
            {
                ok = 0;
                break;
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    if (s_ok == 0)
    {
        if (threadIdx.x == 0)
        {
            *timeout = 1;
        }
        return;
    }
    uint32_t* const h = mine + (size_t) parity * pb.bins_pad;
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
    {
        reduced[b] = __ldcg(h + b);
        h[b] = 0U;
    }
    __syncthreads(); // every thread has read the parity
    if (threadIdx.x == 0)
    {
        *arrived = 0U;
        mine[2 * (size_t) pb.bins_pad + 2] += 1U; // this rank's next epoch
    }
}

} // namespace

void launch_rdf_push(fgpu_ctx* ctx, const PeerBox& pb, const uint32_t* hist, uint32_t bins)
{
    {
        KernelScope ks(ctx, "rdf_push");
        k_rdf_push<<<1, 256, 0, ctx->stream>>>(pb, hist, bins);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_rdf_wait(fgpu_ctx* ctx, const PeerBox& pb, uint32_t bins, uint32_t* reduced, int* timeout)
{
    {
        KernelScope ks(ctx, "rdf_wait");
        k_rdf_wait<<<1, 256, 0, ctx->stream>>>(pb, bins, reduced, timeout);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
