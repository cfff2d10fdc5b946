// k-nearest-neighbour selection on top of the warp-cooperative ball search (the production kNN path).
//
// Replaces AABBQueryIterator::next (freud/locality/AABBQuery.cc:152-281).  By E3 that
// iterator returns, per query point, the k smallest closest-image distances in the IMAGE arithmetic
// r = p_j - (q + image_k) among the points with d >= r_min and r_sq < r_max^2, whatever r_guess and scale
// are.  So the search is one ball query of k_search2<IMAGE, NL> (search2.cu) at a window radius r_win that is
// expected to hold about 2(k + 1) points -- on a grid whose cells are at least r_win thick only one image of
// a point can be inside the window, the one implied by how its cell was reached, and it is the closest one --
// followed by a selection inside every bag row:
//
//   k_knn_rows    counts[q] = min(hits[q], k); a row with fewer than k hits is unresolved unless the window
//                 already is r_max (the host then widens the window for the whole frame and repeats).
//   k_knn_select  one warp per row, lanes over the row's hits: rank every hit by (r_sq, point index) -- the
//                 order in which the reference's std::sort of NeighborBonds resolves the k-th place up to its
//                 unspecified ties -- keep ranks < k, rank the kept hits by point index (or by (d, point
//                 index) for sort_by_distance, NeighborBond.h:80-112) and write the five NeighborList arrays
//                 (NeighborQuery.h:470-478).  Keys are staged in shared memory so that a rank is a loop of
//                 broadcast loads; rows longer than the staging area re-read the bag (L1 hits).
#include "internal.h"

namespace fgpu {

namespace {

constexpr unsigned FULL = 0xffffffffU;
constexpr int kSelWarps = 8;
constexpr uint32_t kStage = 128; // staged keys per warp

__global__ void __launch_bounds__(256) k_knn_rows(KnnRowsArgs a)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t kept = 0;
    bool unresolved = false;
    if (q < a.n_query)
    {
        uint32_t const c = a.hits[q];
        kept = min(c, a.k);
        unresolved = !a.final && c < a.k;
        a.counts[q] = kept;
        a.row_start[q] = kept; // scanned in place afterwards
    }
    unsigned const mu = __ballot_sync(FULL, unresolved);
    unsigned long long sum = kept;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        sum += __shfl_down_sync(FULL, sum, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (mu != 0)
        {
            atomicAdd(a.unresolved, (unsigned long long) __popc(mu));
        }
        if (sum != 0)
        {
            atomicAdd(a.total, sum);
        }
    }
}

__device__ __forceinline__ uint64_t make_key(float f, uint32_t j)
{
    // f >= 0: the bit pattern orders like the value
    return ((uint64_t) __float_as_uint(f) << 32) | (uint64_t) j;
}

template<bool BY_DISTANCE> __global__ void __launch_bounds__(kSelWarps * 32) k_knn_select(KnnSelectArgs a)
{
    __shared__ uint64_t s_key[kSelWarps][kStage];  // selection keys (r_sq, j) of the row
    __shared__ uint64_t s_ord[kSelWarps][kStage];  // ordering keys of the kept hits
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1U << lane) - 1U;
    uint64_t* const sk = s_key[warp];
    uint64_t* const so = s_ord[warp];
    uint32_t const n_warps = gridDim.x * kSelWarps;
    for (uint32_t row = blockIdx.x * kSelWarps + warp; row < a.n_query; row += n_warps)
    {
        uint32_t const n = a.hits[row];
        if (n == 0)
        {
            continue;
        }
        uint32_t const kept = min(n, a.k);
        const float4* __restrict__ const bag = a.bag + a.tmp_start[row];
        uint64_t const out0 = a.row_start[row];
        bool const staged = n <= kStage; // then kept <= kStage as well
        auto key_of = [](const float4& r) { return make_key(dot_exact(r.x, r.y, r.z), __float_as_uint(r.w)); };
        auto ord_of = [](const float4& r) {
            uint32_t const j = __float_as_uint(r.w);
            return BY_DISTANCE ? make_key(__fsqrt_rn(dot_exact(r.x, r.y, r.z)), j) : (uint64_t) j;
        };
        // hits of the row that sort before `key` in the selection order
        auto rank_of = [&](uint64_t key) {
            uint32_t rank = 0;
            if (staged)
            {
#pragma unroll 4
                for (uint32_t i = 0; i < n; ++i)
                {
                    rank += sk[i] < key ? 1U : 0U;
                }
            }
            else
            {
                for (uint32_t i = 0; i < n; ++i)
                {
                    rank += key_of(bag[i]) < key ? 1U : 0U;
                }
            }
            return rank;
        };
        __syncwarp();
        if (staged)
        {
            for (uint32_t h = lane; h < n; h += 32)
            {
                sk[h] = key_of(bag[h]);
            }
        }
        __syncwarp();
        // pass 1: selection rank of every hit (lane l owns hits l, l + 32, ...); kept hits publish their
        // ordering key.  keep_bits remembers the verdicts of the (at most four) rounds of a staged row.
        uint32_t keep_bits = 0, n_seen = 0;
        for (uint32_t h0 = 0, t = 0; h0 < n; h0 += 32, ++t)
        {
            uint32_t const h = h0 + lane;
            bool keep = false;
            float4 r = make_float4(0, 0, 0, 0);
            if (h < n)
            {
                r = bag[h];
                keep = rank_of(key_of(r)) < kept;
            }
            unsigned const mk = __ballot_sync(FULL, keep);
            if (staged)
            {
                if (keep)
                {
                    so[n_seen + __popc(mk & lt_mask)] = ord_of(r);
                    keep_bits |= 1U << t;
                }
                n_seen += __popc(mk);
            }
        }
        __syncwarp();
        // pass 2: position of every kept hit among the kept ones, then the five arrays
        for (uint32_t h0 = 0, t = 0; h0 < n; h0 += 32, ++t)
        {
            uint32_t const h = h0 + lane;
            bool keep = false;
            float4 r = make_float4(0, 0, 0, 0);
            if (h < n)
            {
                r = bag[h];
                keep = staged ? ((keep_bits >> t) & 1U) != 0 : rank_of(key_of(r)) < kept;
            }
            if (__ballot_sync(FULL, keep) == 0)
            {
                continue;
            }
            uint64_t const ord = ord_of(r);
            uint32_t pos = 0;
            if (staged)
            {
#pragma unroll 4
                for (uint32_t i = 0; i < kept; ++i)
                {
                    pos += so[i] < ord ? 1U : 0U;
                }
            }
            else
            {
                // rows longer than the staging area: order against every kept hit of the bag
                for (uint32_t i = 0; i < n; ++i)
                {
                    float4 const u = bag[i];
                    pos += (rank_of(key_of(u)) < kept && ord_of(u) < ord) ? 1U : 0U;
                }
            }
            if (keep)
            {
                uint64_t const out = out0 + pos;
                reinterpret_cast<uint2*>(a.neighbors)[out] = make_uint2(row, __float_as_uint(r.w));
                a.distances[out] = __fsqrt_rn(dot_exact(r.x, r.y, r.z));
                a.weights[out] = 1.0f;
                a.vectors[3 * out] = r.x;
                a.vectors[3 * out + 1] = r.y;
                a.vectors[3 * out + 2] = r.z;
            }
        }
    }
}

} // namespace

void launch_knn_rows(fgpu_ctx* ctx, const KnnRowsArgs& a)
{
    if (a.n_query == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "knn_rows");
        k_knn_rows<<<(a.n_query + 255) / 256, 256, 0, ctx->stream>>>(a);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_knn_select(fgpu_ctx* ctx, int sort_by_distance, const KnnSelectArgs& a)
{
    if (a.n_query == 0)
    {
        return;
    }
    unsigned const want = (a.n_query + kSelWarps - 1) / kSelWarps;
    unsigned const blocks = std::max(1U, std::min(want, (unsigned) ctx->sm_count * 8U));
    {
        KernelScope ks(ctx, "knn_select");
        if (sort_by_distance)
        {
            k_knn_select<true><<<blocks, kSelWarps * 32, 0, ctx->stream>>>(a);
        }
        else
        {
            k_knn_select<false><<<blocks, kSelWarps * 32, 0, ctx->stream>>>(a);
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
