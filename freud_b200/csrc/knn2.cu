// k-nearest-neighbour selection on top of the warp-cooperative ball search (the production kNN path).
//
// Replaces AABBQueryIterator::next (freud/locality/AABBQuery.cc:152-281) and, in the WRAP flavour,
// LinkCellQueryIterator::next (freud/locality/LinkCell.cc:575-679).  By E3 the first
// returns, per query point, the k smallest closest-image distances in the IMAGE arithmetic
// r = p_j - (q + image_k) among the points with d >= r_min and r_sq < r_max^2, whatever r_guess and scale
// are; the second the k smallest wrapped distances with r_min^2 <= r_sq < r_max^2.  So the search is one ball
// query of k_search2<flavour, NL> (search2.cu) at a window radius r_win that is expected to hold about
// 1.5 (k + 1) points -- on a grid whose cells are at least r_win thick only one image of a point can be inside
// the window, the one implied by how its cell was reached, and it is the closest one -- followed by a selection
// inside every bag row:
//
//   k_knn_rows    counts[q] = min(hits[q], k); a row with fewer than k hits is unresolved unless the window
//                 already is r_max.  The host searches the unresolved rows again -- only those, into a second
//                 bag -- with a window 1.5x wider; if many rows are short (or some still are after that) it
//                 widens the window for the whole frame and repeats.
//   k_knn_select  one warp per row, lanes over the row's hits: rank every hit by (r_sq, point index) -- the
//                 order in which the reference's std::sort of NeighborBonds resolves the k-th place up to its
//                 unspecified ties -- keep ranks < k, rank the kept hits by point index (or by (d, point
//                 index) for sort_by_distance, NeighborBond.h:80-112) and write the five NeighborList arrays
//                 (NeighborQuery.h:470-478).  r_sq and point indices are staged in shared memory so that a
//                 rank is a loop of broadcast 16-byte loads and float compares (ties at the k-th place take
//                 a slower exact path); rows longer than the staging area re-read the bag (L1 hits).
#include "internal.h"

namespace fgpu {

namespace {

constexpr unsigned FULL = 0xffffffffU;
constexpr int kSelWarps = 8;
constexpr uint32_t kStage = 128; // staged keys per warp

__global__ void __launch_bounds__(256) k_knn_rows(KnnRowsArgs a)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    int const lane = threadIdx.x & 31;
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
    // one atomic per block on the hot total (one per warp serialises in L2: 31 k adds on one address)
    __shared__ unsigned long long s_sum[256 / 32];
    if (lane == 0)
    {
        s_sum[threadIdx.x >> 5] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long t = 0;
        for (int w = 0; w < 256 / 32; ++w)
        {
            t += s_sum[w];
        }
        if (t != 0)
        {
            atomicAdd(a.total, t);
        }
    }
    unsigned long long base = 0;
    if (lane == 0 && mu != 0)
    {
        base = atomicAdd(a.unresolved, (unsigned long long) __popc(mu));
    }
    base = __shfl_sync(FULL, base, 0);
    if (unresolved)
    {
        a.unresolved_rows[base + __popc(mu & ((1U << lane) - 1U))] = q; // order is irrelevant
    }
}

// positions of a subset of the query points (the rows a wider window has to search again)
__global__ void __launch_bounds__(256) k_gather_points(const float* __restrict__ xyz, const uint32_t* __restrict__ rows,
                                                       uint32_t n_rows, float* __restrict__ out)
{
    uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_rows)
    {
        size_t const r = rows[s];
        out[3 * (size_t) s] = xyz[3 * r];
        out[3 * (size_t) s + 1] = xyz[3 * r + 1];
        out[3 * (size_t) s + 2] = xyz[3 * r + 2];
    }
}

__device__ __forceinline__ uint64_t make_key(float f, uint32_t j)
{
    // f >= 0: the bit pattern orders like the value
    return ((uint64_t) __float_as_uint(f) << 32) | (uint64_t) j;
}

__device__ __forceinline__ uint64_t select_key(const float4& r)
{
    return make_key(dot_exact(r.x, r.y, r.z), __float_as_uint(r.w));
}

template<bool BY_DISTANCE> __device__ __forceinline__ uint64_t order_key(const float4& r)
{
    uint32_t const j = __float_as_uint(r.w);
    return BY_DISTANCE ? make_key(__fsqrt_rn(dot_exact(r.x, r.y, r.z)), j) : (uint64_t) j;
}

__device__ __forceinline__ void write_bond(const KnnSelectArgs& a, uint64_t out, uint32_t row, const float4& r)
{
    reinterpret_cast<uint2*>(a.neighbors)[out] = make_uint2(row, __float_as_uint(r.w));
    a.distances[out] = __fsqrt_rn(dot_exact(r.x, r.y, r.z)); // NeighborBond.h:41-44
    a.weights[out] = 1.0f;
    a.vectors[3 * out] = r.x;
    a.vectors[3 * out + 1] = r.y;
    a.vectors[3 * out + 2] = r.z;
}

// Rows longer than the staging area (very dense spots or a very large k): every rank re-reads the bag.
template<bool BY_DISTANCE>
__device__ __noinline__ void select_long_row(const KnnSelectArgs& a, const float4* __restrict__ bag, uint32_t n,
                                             uint32_t kept, uint32_t row, uint64_t out0, int lane)
{
    auto rank_of = [&](uint64_t key) {
        uint32_t rank = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            rank += select_key(bag[i]) < key ? 1U : 0U;
        }
        return rank;
    };
    for (uint32_t h = lane; h < n; h += 32)
    {
        float4 const r = bag[h];
        if (rank_of(select_key(r)) >= kept)
        {
            continue;
        }
        uint64_t const ord = order_key<BY_DISTANCE>(r);
        uint32_t pos = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            float4 const u = bag[i];
            pos += (rank_of(select_key(u)) < kept && order_key<BY_DISTANCE>(u) < ord) ? 1U : 0U;
        }
        write_bond(a, out0 + pos, row, r);
    }
}

template<bool BY_DISTANCE> __global__ void __launch_bounds__(kSelWarps * 32) k_knn_select(KnnSelectArgs a)
{
    // per warp: r_sq and point index of every hit of the row, then the ordering keys of the kept hits
    __shared__ __align__(16) float s_rsq[kSelWarps][kStage];
    __shared__ __align__(16) uint32_t s_j[kSelWarps][kStage];
    __shared__ __align__(16) uint32_t s_oj[kSelWarps][kStage];
    __shared__ __align__(16) float s_od[kSelWarps][kStage];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1U << lane) - 1U;
    float* const rsq = s_rsq[warp];
    uint32_t* const js = s_j[warp];
    uint32_t* const oj = s_oj[warp];
    float* const od = s_od[warp];
    float const inf = __int_as_float(0x7f800000);
    uint32_t const n_warps = gridDim.x * kSelWarps;
    // What a row needs from memory, fetched one row ahead: the chain length -> bag offset -> record is three
    // dependent loads, and a warp walks its rows one after the other.
    struct RowIn
    {
        uint32_t n;
        const float4* bag;
        uint64_t out0;
        float4 rec; // lane l: hit l of the row (padding beyond the row)
    };
    auto fetch = [&](uint32_t row) {
        RowIn in;
        in.n = a.hits[row];
        uint32_t const ts = a.tmp_start[row];
        // rows searched again with a wider window live in the second bag (flagged in the top bit)
        in.bag = (ts & kSecondBag) != 0 ? a.bag2 + (ts & ~kSecondBag) : a.bag + ts;
        in.out0 = a.row_start[row];
        in.rec = (uint32_t) lane < in.n ? in.bag[lane] : make_float4(inf, 0.0f, 0.0f, __uint_as_float(0x7fffffffU));
        return in;
    };
    uint32_t const first_row = blockIdx.x * kSelWarps + warp;
    RowIn ahead = first_row < a.n_query ? fetch(first_row) : RowIn {0, nullptr, 0, make_float4(0, 0, 0, 0)};
    for (uint32_t row = first_row; row < a.n_query; row += n_warps)
    {
        RowIn const cur = ahead;
        if (row + n_warps < a.n_query && row + n_warps > row)
        {
            ahead = fetch(row + n_warps);
        }
        uint32_t const n = cur.n;
        if (n == 0)
        {
            continue;
        }
        uint32_t const kept = min(n, a.k);
        const float4* __restrict__ const bag = cur.bag;
        uint64_t const out0 = cur.out0;
        if (n > kStage)
        {
            select_long_row<BY_DISTANCE>(a, bag, n, kept, row, out0, lane);
            continue;
        }
        uint32_t const n4 = (n + 3U) & ~3U, rounds = (n + 31U) >> 5;
        __syncwarp();
        if (n <= 32)
        {
            // the common case, straight-line: one hit per lane
            bool const act = (uint32_t) lane < n;
            float4 const r = cur.rec;
            uint32_t const my_j = __float_as_uint(r.w);
            float const my_rsq = act ? dot_exact(r.x, r.y, r.z) : inf;
            uint32_t const mine = __float_as_uint(my_rsq);
            rsq[lane] = my_rsq;
            js[lane] = my_j;
            __syncwarp();
            uint32_t closer = 0;
            for (uint32_t i = 0; i < n4; i += 4)
            {
                uint4 const v = *reinterpret_cast<const uint4*>(rsq + i);
                closer += (v.x - mine) >> 31;
                closer += (v.y - mine) >> 31;
                closer += (v.z - mine) >> 31;
                closer += (v.w - mine) >> 31;
            }
            bool keep = act && closer < kept;
            unsigned mk = __ballot_sync(FULL, keep);
            if ((uint32_t) __popc(mk) != kept)
            {
                // ties at the k-th place: resolve by point index, the order the general kernel and the oracle use
                uint32_t before = 0;
                for (uint32_t i = 0; i < n; ++i)
                {
                    float const v = rsq[i];
                    before += (v < my_rsq || (v == my_rsq && js[i] < my_j)) ? 1U : 0U;
                }
                keep = act && before < kept;
                mk = __ballot_sync(FULL, keep);
            }
            float const my_d = __fsqrt_rn(my_rsq);
            if (keep)
            {
                uint32_t const slot = __popc(mk & lt_mask);
                oj[slot] = my_j;
                if (BY_DISTANCE)
                {
                    od[slot] = my_d;
                }
            }
            if (lane < 4)
            {
                oj[kept + lane] = 0x7fffffffU; // padding sorts after every point index (kept + 3 < kStage)
                od[kept + lane] = inf;
            }
            __syncwarp();
            if (keep)
            {
                uint32_t pos = 0;
                if (!BY_DISTANCE)
                {
                    for (uint32_t i = 0; i < kept; i += 4)
                    {
                        uint4 const v = *reinterpret_cast<const uint4*>(oj + i);
                        pos += (v.x - my_j) >> 31;
                        pos += (v.y - my_j) >> 31;
                        pos += (v.z - my_j) >> 31;
                        pos += (v.w - my_j) >> 31;
                    }
                }
                else
                {
                    for (uint32_t i = 0; i < kept; ++i)
                    {
                        float const d = od[i];
                        pos += (d < my_d || (d == my_d && oj[i] < my_j)) ? 1U : 0U; // NeighborBond.h:96-112
                    }
                }
                uint64_t const out = out0 + pos;
                reinterpret_cast<uint2*>(a.neighbors)[out] = make_uint2(row, my_j);
                a.distances[out] = my_d; // NeighborBond.h:41-44
                a.weights[out] = 1.0f;
                a.vectors[3 * out] = r.x;
                a.vectors[3 * out + 1] = r.y;
                a.vectors[3 * out + 2] = r.z;
            }
            continue;
        }
        for (uint32_t h = lane; h < n4; h += 32)
        {
            float4 const r = h < n ? bag[h] : make_float4(inf, 0.0f, 0.0f, __uint_as_float(0xffffffffU));
            rsq[h] = h < n ? dot_exact(r.x, r.y, r.z) : inf; // padding never sorts before anything
            js[h] = __float_as_uint(r.w);
        }
        __syncwarp();
        // pass 1: hits closer than this one (lane l owns hits l, l + 32, ...).  Hits whose count is below
        // `kept` are the answer unless a group of equal r_sq straddles the k-th place.
        uint32_t keep_bits = 0, n_keep = 0;
        for (uint32_t t = 0; t < rounds; ++t)
        {
            uint32_t const h = t * 32 + lane;
            // r_sq >= +0, so bit patterns order like values and (a - b) >> 31 is [a < b] in two instructions
            uint32_t const mine = __float_as_uint(h < n ? rsq[h] : inf);
            uint32_t closer = 0;
            for (uint32_t i = 0; i < n4; i += 4)
            {
                uint4 const v = *reinterpret_cast<const uint4*>(rsq + i);
                closer += (v.x - mine) >> 31;
                closer += (v.y - mine) >> 31;
                closer += (v.z - mine) >> 31;
                closer += (v.w - mine) >> 31;
            }
            bool const keep = h < n && closer < kept;
            keep_bits |= keep ? 1U << t : 0U;
            n_keep += __popc(__ballot_sync(FULL, keep));
        }
        if (n_keep != kept)
        {
            // ties at the k-th place: resolve by point index, the order the general kernel and the oracle use
            keep_bits = 0;
            for (uint32_t t = 0; t < rounds; ++t)
            {
                uint32_t const h = t * 32 + lane;
                float const mine = h < n ? rsq[h] : inf;
                uint32_t const my_j = h < n ? js[h] : 0xffffffffU;
                uint32_t before = 0;
                for (uint32_t i = 0; i < n; ++i)
                {
                    float const v = rsq[i];
                    before += (v < mine || (v == mine && js[i] < my_j)) ? 1U : 0U;
                }
                keep_bits |= (h < n && before < kept) ? 1U << t : 0U;
            }
        }
        // the kept hits publish their ordering keys (padded to a multiple of four)
        uint32_t n_seen = 0;
        for (uint32_t t = 0; t < rounds; ++t)
        {
            uint32_t const h = t * 32 + lane;
            bool const keep = ((keep_bits >> t) & 1U) != 0;
            unsigned const mk = __ballot_sync(FULL, keep);
            if (keep)
            {
                uint32_t const pos = n_seen + __popc(mk & lt_mask);
                oj[pos] = js[h];
                if (BY_DISTANCE)
                {
                    od[pos] = __fsqrt_rn(rsq[h]);
                }
            }
            n_seen += __popc(mk);
        }
        uint32_t const kept4 = (kept + 3U) & ~3U;
        if (lane < kept4 - kept)
        {
            oj[kept + lane] = 0x7fffffffU; // sorts after every point index
            od[kept + lane] = inf;
        }
        __syncwarp();
        // pass 2: position of every kept hit among the kept ones, then the five arrays
        for (uint32_t t = 0; t < rounds; ++t)
        {
            if (((keep_bits >> t) & 1U) == 0)
            {
                continue;
            }
            uint32_t const h = t * 32 + lane;
            float4 const r = bag[h];
            uint32_t const my_j = __float_as_uint(r.w);
            uint32_t pos = 0;
            if (!BY_DISTANCE)
            {
                // point indices are below 2^31 on this path (checked by the host): same two-instruction compare
                for (uint32_t i = 0; i < kept4; i += 4)
                {
                    uint4 const v = *reinterpret_cast<const uint4*>(oj + i);
                    pos += (v.x - my_j) >> 31;
                    pos += (v.y - my_j) >> 31;
                    pos += (v.z - my_j) >> 31;
                    pos += (v.w - my_j) >> 31;
                }
            }
            else
            {
                float const my_d = __fsqrt_rn(dot_exact(r.x, r.y, r.z));
                for (uint32_t i = 0; i < kept; ++i)
                {
                    float const d = od[i];
                    pos += (d < my_d || (d == my_d && oj[i] < my_j)) ? 1U : 0U; // NeighborBond.h:96-112
                }
            }
            write_bond(a, out0 + pos, row, r);
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

void launch_gather_points(fgpu_ctx* ctx, const float* xyz, const uint32_t* rows, uint32_t n_rows, float* out)
{
    if (n_rows == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "knn_gather");
        k_gather_points<<<(n_rows + 255) / 256, 256, 0, ctx->stream>>>(xyz, rows, n_rows, out);
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
