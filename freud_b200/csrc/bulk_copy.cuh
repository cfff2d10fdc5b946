// cp.async.bulk (the TMA engine's 1-D bulk copy) + mbarrier, as inline PTX for sm_100a.
//
// One thread arms an mbarrier with the byte count it expects and issues the copy; the engine generates the
// addresses, moves the bytes global -> shared without touching a register or an issue slot of the other warps, and
// completes the transaction on the barrier; consumers sleep on try_wait (a hardware wait, not a poll loop).  Source,
// destination and size must be multiples of 16 bytes.
#pragma once
#include <stdint.h>

namespace fgpu {
namespace bulk {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}

// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy writes to it (a
// buffer that was just read is about to be refilled by the copy engine)
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}\n"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while (done == 0);
}

} // namespace bulk
} // namespace fgpu
