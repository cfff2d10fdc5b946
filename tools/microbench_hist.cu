// Micro-benchmark: shared-memory histogram increment strategies on sm_100a (decides the K6 binning design).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench_hist tools/microbench_hist.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
// bin distributed like r^2 dr over [0,1): cube root of a uniform
__device__ __forceinline__ int draw_bin(uint32_t& s, int bins)
{
    float u = (rng(s) >> 8) * (1.0f / 16777216.0f);
    int b = (int) (cbrtf(u) * bins);
    return b < bins ? b : bins - 1;
}

template<int MODE> __global__ void __launch_bounds__(128) k(int bins, int iters, uint32_t* out)
{
    extern __shared__ uint32_t sh[];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const words = MODE == 0 ? bins : MODE == 1 || MODE == 4 ? bins * 4 : MODE == 2 ? bins * 32 * 4 : bins * 16 * 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it)
    {
        int const b = draw_bin(s, bins);
        if (MODE == 0) atomicAdd(&sh[b], 1u);                                  // block-shared ATOMS
        else if (MODE == 1) atomicAdd(&sh[warp * bins + b], 1u);               // warp-private ATOMS
        else if (MODE == 2) { uint32_t* p = &sh[(warp * bins + b) * 32 + lane]; *p = *p + 1; } // lane-private u32
        else if (MODE == 3)
        { // lane-private u16 packed, 2 bins per word
            uint32_t* p = &sh[(warp * (bins / 2) + (b >> 1)) * 32 + lane];
            *p = *p + (1u << ((b & 1) * 16));
        }
        else if (MODE == 4)
        { // warp-private, match_any aggregated
            unsigned const peers = __match_any_sync(0xffffffffu, b);
            if ((__ffs(peers) - 1) == lane) atomicAdd(&sh[warp * bins + b], (uint32_t) __popc(peers));
        }
        acc += b;
    }
    __syncthreads();
    uint32_t t = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) t += sh[i] & 0xffff;
    atomicAdd(out, t + (acc & 1));
}

template<int MODE> void run(const char* name, int bins, size_t smem)
{
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaMemset(d, 0, 4);
    int const iters = 2000, blocks = 148 * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k<MODE><<<blocks, 128, smem>>>(bins, 10, d);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128, smem>>>(bins, iters, d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    double const incs = (double) blocks * 128 * iters;
    printf("%-34s bins=%3d smem=%6zu B  %8.3f ms  %7.1f G incr/s  %s\n", name, bins, smem, ms, incs / ms * 1e-6,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(d);
}

template<int MODE> __global__ void __launch_bounds__(128) k_div(int iters, float L, float rcpL, float* out)
{
    float x = threadIdx.x * 0.37f + 1.0f, acc = 0;
    for (int it = 0; it < iters; ++it)
    {
        float q;
        if (MODE == 0) q = __fdiv_rn(x, L);
        else { float const q0 = __fmul_rn(x, rcpL); float const r = __fmaf_rn(-q0, L, x); q = __fmaf_rn(r, rcpL, q0); }
        if (MODE == 2) q = truncf(q);
        acc += q;
        x += 0.001f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    for (int bins : {100, 500})
    {
        run<0>("block-shared ATOMS", bins, bins * 4);
        run<1>("warp-private ATOMS", bins, bins * 16);
        if (bins * 512 <= 227 * 1024) run<2>("lane-private u32 LDS/STS", bins, (size_t) bins * 512);
        run<3>("lane-private u16 LDS/STS", bins, (size_t) bins * 256);
        run<4>("warp-private match_any + ATOMS", bins, bins * 16);
    }
    float* d;
    cudaMalloc(&d, 148 * 16 * 128 * 4);
    for (int mode = 0; mode < 3; ++mode)
    {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        int const iters = 4000;
        cudaEventRecord(e0);
        if (mode == 0) k_div<0><<<148 * 16, 128>>>(iters, 232.0794f, 1.0f / 232.0794f, d);
        if (mode == 1) k_div<1><<<148 * 16, 128>>>(iters, 232.0794f, 1.0f / 232.0794f, d);
        if (mode == 2) k_div<2><<<148 * 16, 128>>>(iters, 232.0794f, 1.0f / 232.0794f, d);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("div mode %d (%s): %.3f ms, %.1f G/s\n", mode, mode == 0 ? "__fdiv_rn" : mode == 1 ? "markstein" : "markstein+truncf",
               ms, 148.0 * 16 * 128 * iters / ms * 1e-6);
    }
    return 0;
}
