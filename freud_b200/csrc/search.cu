// 27-cell minimum-image pair search: count / fill (NeighborList) / fused RDF histogram, plus the
// NeighborList emit kernel.
//
// Replaces LinkCellQueryBallIterator::next (freud/locality/LinkCell.cc:496-573, flavour WRAP),
// AABBQueryBallIterator::next (freud/locality/AABBQuery.cc:77-150, flavour IMAGE),
// NeighborQueryIterator::toNeighborList (freud/locality/NeighborQuery.h:434-481) and the on-the-fly branch
// of loopOverNeighbors + RDF's binning lambda (freud/locality/NeighborComputeFunctional.h:195-217,
// freud/density/RDF.cc:101-110).
//
// Version 1 mapping: one thread per query point, queries visited in cell order so that the threads of a
// warp walk the same few cells (L1-resident float4 candidates).  NeighborList build is
//   count (per row) -> exclusive scan -> fill (unsorted 16-byte hit records per row) -> emit
// where emit runs one thread per bond in OUTPUT order: it ranks its hit inside the row by counting smaller
// keys, recomputes the bond with the exact arithmetic and writes all five arrays coalesced.
#include "internal.h"

namespace fgpu {

namespace {

constexpr int kSearchThreads = 128;

__device__ __forceinline__ bool in_window(float r_sq, float r_max_sq, float r_min_sq)
{
    return r_sq < r_max_sq && r_sq >= r_min_sq; // LinkCell.cc:525, AABBQuery.cc:129
}

// Calls on_hit(slot, j, r_sq, rx, ry, rz) for every bond of one query point; returns pair evaluations.
template<int FLAVOUR, typename OnHit>
__device__ __forceinline__ uint32_t visit_hits(const BoxDev& box, const GridDev& g, float qx, float qy, float qz,
                                               uint32_t q_global, int exclude_ii, float r_max_sq, float r_min_sq,
                                               OnHit&& on_hit)
{
    uint32_t evals = 0;
    bool const any_shift = FLAVOUR != FGPU_FLAVOUR_WRAP && *g.any_shift_flag != 0;
    if (FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d)
    {
        qz = 0.0f; // AABBQuery.cc:84-87
    }
    int cx, cy, cz, nqx, nqy, nqz;
    cell_coords(box, g.dx, g.dy, g.dz, qx, qy, qz, cx, cy, cz, nqx, nqy, nqz);
    AxisSlots sx, sy, sz;
    make_slots(g.dx, g.amb_x, cx, sx);
    make_slots(g.dy, g.amb_y, cy, sy);
    make_slots(g.dz, g.amb_z, cz, sz);
    for (int iz = 0; iz < sz.n; ++iz)
    {
        for (int iy = 0; iy < sy.n; ++iy)
        {
            uint32_t const row = ((uint32_t) sz.cell[iz] * g.dy + sy.cell[iy]) * g.dx;
            for (int ix = 0; ix < sx.n; ++ix)
            {
                uint32_t const cell = row + sx.cell[ix];
                uint32_t const beg = __ldg(g.cell_start + cell), end = __ldg(g.cell_start + cell + 1);
                if (beg == end)
                {
                    continue;
                }
                int const wx = sx.w[ix], wy = sy.w[iy], wz = sz.w[iz];
                if (FLAVOUR == FGPU_FLAVOUR_WRAP)
                {
                    for (uint32_t s = beg; s < end; ++s)
                    {
                        float4 const p = __ldg(g.sorted + s);
                        uint32_t const j = __float_as_uint(p.w);
                        if (exclude_ii && j == q_global)
                        {
                            continue; // LinkCell.cc:517-520
                        }
                        ++evals;
                        float rx, ry, rz;
                        wrap_exact(box, __fsub_rn(p.x, qx), __fsub_rn(p.y, qy), __fsub_rn(p.z, qz), rx, ry, rz);
                        float const r_sq = dot_exact(rx, ry, rz);
                        if (in_window(r_sq, r_max_sq, r_min_sq))
                        {
                            on_hit(s, j, r_sq, rx, ry, rz);
                        }
                    }
                }
                else
                {
                    bool const definite = wx != 2 && wy != 2 && wz != 2 && !any_shift;
                    if (definite)
                    {
                        // all points are inside the box: the image is fixed by how the cell was reached
                        int const kx = -nqx - wx, ky = -nqy - wy, kz = -nqz - wz;
                        if (kx < -1 || kx > 1 || ky < -1 || ky > 1 || kz < -1 || kz > 1)
                        {
                            continue;
                        }
                        for (uint32_t s = beg; s < end; ++s)
                        {
                            float4 const p = __ldg(g.sorted + s);
                            uint32_t const j = __float_as_uint(p.w);
                            if (exclude_ii && j == q_global)
                            {
                                continue; // AABBQuery.cc:111-115, CellIterator.h:160-163
                            }
                            ++evals;
                            // AABBQuery.cc:118-122 zeroes z in 2-D boxes, CellQuery does not
                            float const pz = FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d ? 0.0f : p.z;
                            float rx, ry, rz;
                            image_pair<FLAVOUR>(box, p.x, p.y, pz, qx, qy, qz, kx, ky, kz, rx, ry, rz);
                            float const r_sq = dot_exact(rx, ry, rz);
                            if (in_window(r_sq, r_max_sq, r_min_sq))
                            {
                                on_hit(s, j, r_sq, rx, ry, rz);
                            }
                        }
                    }
                    else
                    {
                        // small boxes / points outside the box: try every admissible image per candidate
                        for (uint32_t s = beg; s < end; ++s)
                        {
                            float4 const p = __ldg(g.sorted + s);
                            uint32_t const j = __float_as_uint(p.w);
                            if (exclude_ii && j == q_global)
                            {
                                continue;
                            }
                            int njx = 0, njy = 0, njz = 0;
                            if (any_shift)
                            {
                                unpack_shift(__ldg(g.shift + s), njx, njy, njz);
                            }
                            int const kx0 = wx == 2 ? -1 : njx - nqx - wx, kx1 = wx == 2 ? 1 : kx0;
                            int const ky0 = wy == 2 ? -1 : njy - nqy - wy, ky1 = wy == 2 ? 1 : ky0;
                            int const kz0 = wz == 2 ? -1 : njz - nqz - wz, kz1 = wz == 2 ? 1 : kz0;
                            float const pz = FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d ? 0.0f : p.z;
                            for (int kx = max(kx0, -1); kx <= min(kx1, 1); ++kx)
                            {
                                for (int ky = max(ky0, -1); ky <= min(ky1, 1); ++ky)
                                {
                                    for (int kz = max(kz0, -1); kz <= min(kz1, 1); ++kz)
                                    {
                                        ++evals;
                                        float rx, ry, rz;
                                        image_pair<FLAVOUR>(box, p.x, p.y, pz, qx, qy, qz, kx, ky, kz, rx, ry, rz);
                                        float const r_sq = dot_exact(rx, ry, rz);
                                        if (in_window(r_sq, r_max_sq, r_min_sq))
                                        {
                                            on_hit(s, j, r_sq, rx, ry, rz);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return evals;
}

template<int FLAVOUR, int MODE> __global__ void __launch_bounds__(kSearchThreads) k_search(SearchArgs a)
{
    extern __shared__ uint32_t sh_hist[];
    if (a.only_if != nullptr && *a.only_if == 0)
    {
        return;
    }
    bool const hist_in_smem = MODE == SEARCH_RDF && a.axis.bins * sizeof(uint32_t) <= 48 * 1024;
    if (MODE == SEARCH_RDF && hist_in_smem)
    {
        for (uint32_t b = threadIdx.x; b < a.axis.bins; b += blockDim.x)
        {
            sh_hist[b] = 0;
        }
        __syncthreads();
    }
    uint32_t* const hist = hist_in_smem ? sh_hist : a.hist;
    float const r_max_sq = __fmul_rn(a.r_max, a.r_max); // LinkCell.cc:498, AABBQuery.cc:79
    float const r_min_sq = __fmul_rn(a.r_min, a.r_min);
    unsigned long long my_total = 0, my_evals = 0;

    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_query; t += gridDim.x * blockDim.x)
    {
        float4 const q = a.q_sorted[t];
        uint32_t const qi = __float_as_uint(q.w);
        uint32_t const q_global = qi + a.q_index_offset;
        uint32_t count = 0;
        uint32_t const base = MODE == SEARCH_FILL ? a.row_start[qi] : 0;
        uint32_t const ev = visit_hits<FLAVOUR>(
            a.box, a.grid, q.x, q.y, q.z, q_global, a.exclude_ii, r_max_sq, r_min_sq,
            [&](uint32_t s, uint32_t j, float r_sq, float, float, float) {
                if (MODE == SEARCH_FILL)
                {
                    uint32_t const key_hi = a.sort_by_distance ? __float_as_uint(__fsqrt_rn(r_sq)) : 0U;
                    a.bag[base + count] = make_uint4(key_hi, j, s, qi);
                }
                else if (MODE == SEARCH_RDF)
                {
                    int const bin = axis_bin(a.axis, __fsqrt_rn(r_sq)); // NeighborBond distance = sqrt(r_sq)
                    if (bin >= 0)
                    {
                        atomicAdd(&hist[bin], 1U);
                    }
                }
                ++count;
            });
        if (MODE == SEARCH_COUNT)
        {
            a.row_counts[qi] = count;
            my_total += count;
        }
        my_evals += ev;
    }

    if (MODE == SEARCH_COUNT || a.evals != nullptr)
    {
        // block reduction of the two u64 totals: one atomic each per block
        __shared__ unsigned long long red[2][kSearchThreads / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            my_total += __shfl_down_sync(0xffffffffU, my_total, o);
            my_evals += __shfl_down_sync(0xffffffffU, my_evals, o);
        }
        if ((threadIdx.x & 31) == 0)
        {
            red[0][threadIdx.x >> 5] = my_total;
            red[1][threadIdx.x >> 5] = my_evals;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned long long t0 = 0, t1 = 0;
            for (int w = 0; w < kSearchThreads / 32; ++w)
            {
                t0 += red[0][w];
                t1 += red[1][w];
            }
            if (MODE == SEARCH_COUNT && t0 != 0)
            {
                atomicAdd(a.total, t0);
            }
            if (a.evals != nullptr && t1 != 0)
            {
                atomicAdd(a.evals, t1);
            }
        }
    }
    if (MODE == SEARCH_RDF && hist_in_smem)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < a.axis.bins; b += blockDim.x)
        {
            uint32_t const v = sh_hist[b];
            if (v != 0)
            {
                atomicAdd(&a.hist[b], v); // u32 wraps like the reference's unsigned int counters
            }
        }
    }
}

// ---- emit: one thread per output bond ----------------------------------------------------------------
template<int FLAVOUR> __global__ void __launch_bounds__(256) k_emit(EmitArgs a)
{
    uint64_t const b = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_bonds)
    {
        return;
    }
    uint4 const e = a.bag[b];
    uint32_t const qi = e.w;
    uint32_t const beg = a.row_start[qi], end = a.row_start[qi + 1];
    // rank inside the row: NeighborBond::less_as_tuple / less_as_distance restricted to one row with
    // weight == 1 (NeighborBond.h:80-112); (key_hi, key_lo) = (0, j) or (bits(d), j)
    unsigned long long const my_key = ((unsigned long long) e.x << 32) | e.y;
    uint32_t rank = 0;
    for (uint32_t k = beg; k < end; ++k)
    {
        uint4 const o = a.bag[k];
        unsigned long long const key = ((unsigned long long) o.x << 32) | o.y;
        rank += (key < my_key || (key == my_key && k < b)) ? 1U : 0U;
    }
    uint64_t const out = (uint64_t) beg + rank;

    float4 const p = __ldg(a.sorted + e.z);
    float qx = a.q_xyz[3 * (size_t) qi], qy = a.q_xyz[3 * (size_t) qi + 1], qz = a.q_xyz[3 * (size_t) qi + 2];
    float rx, ry, rz, r_sq;
    if (FLAVOUR == FGPU_FLAVOUR_WRAP)
    {
        wrap_exact(a.box, __fsub_rn(p.x, qx), __fsub_rn(p.y, qy), __fsub_rn(p.z, qz), rx, ry, rz);
        r_sq = dot_exact(rx, ry, rz);
    }
    else
    {
        // the image that produced the hit is the only one inside the window (plane distance > 2 r_max is
        // enforced, NeighborQuery.h:503-510); walk the images in the reference's order and take it
        float const r_max_sq = __fmul_rn(a.r_max, a.r_max), r_min_sq = __fmul_rn(a.r_min, a.r_min);
        float const pz = FLAVOUR == FGPU_FLAVOUR_IMAGE && a.box.is2d ? 0.0f : p.z;
        if (FLAVOUR == FGPU_FLAVOUR_IMAGE && a.box.is2d)
        {
            qz = 0.0f;
        }
        rx = ry = rz = r_sq = 0.0f;
        bool found = false;
        for (int code = 0; code < 27 && !found; ++code)
        {
            // code 0 is the identity image; 1..26 enumerate (i, j, k) in the order of NeighborQuery.h:546-562
            int i = 0, j = 0, k = 0;
            if (code > 0)
            {
                int const c = code - 1 + (code - 1 >= 13 ? 1 : 0); // skip (0,0,0) at position 13
                i = c / 9 - 1;
                j = (c / 3) % 3 - 1;
                k = c % 3 - 1;
            }
            if (a.box.is2d && k != 0)
            {
                continue;
            }
            float tx, ty, tz;
            image_pair<FLAVOUR>(a.box, p.x, p.y, pz, qx, qy, qz, i, j, k, tx, ty, tz);
            float const t_sq = dot_exact(tx, ty, tz);
            if (in_window(t_sq, r_max_sq, r_min_sq))
            {
                rx = tx;
                ry = ty;
                rz = tz;
                r_sq = t_sq;
                found = true;
            }
        }
    }
    a.neighbors[2 * out] = qi;
    a.neighbors[2 * out + 1] = __float_as_uint(p.w);
    a.distances[out] = __fsqrt_rn(r_sq); // NeighborBond(i, j, w, v): distance = sqrt(dot(v, v)), NeighborBond.h:41-44
    a.weights[out] = 1.0f;
    a.vectors[3 * out] = rx;
    a.vectors[3 * out + 1] = ry;
    a.vectors[3 * out + 2] = rz;
}

__global__ void __launch_bounds__(256) k_segments(const uint32_t* __restrict__ row_start,
                                                  const uint32_t* __restrict__ counts, uint32_t* __restrict__ segments,
                                                  uint32_t n_query)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_query)
    {
        segments[i] = counts[i] != 0 ? row_start[i] : 0U; // NeighborList.cc:199-232: untouched rows stay 0
    }
}

__global__ void __launch_bounds__(256) k_rdf_distances(const float* __restrict__ d, uint64_t n, AxisDev axis,
                                                       uint32_t* __restrict__ hist)
{
    extern __shared__ uint32_t sh_hist[];
    bool const in_smem = axis.bins * sizeof(uint32_t) <= 48 * 1024;
    if (in_smem)
    {
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            sh_hist[b] = 0;
        }
        __syncthreads();
    }
    uint32_t* const h = in_smem ? sh_hist : hist;
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t) gridDim.x * blockDim.x)
    {
        int const bin = axis_bin(axis, d[k]);
        if (bin >= 0)
        {
            atomicAdd(&h[bin], 1U);
        }
    }
    if (in_smem)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            if (sh_hist[b] != 0)
            {
                atomicAdd(&hist[b], sh_hist[b]);
            }
        }
    }
}

template<int FLAVOUR> void launch_search_mode(fgpu_ctx* ctx, SearchMode mode, const SearchArgs& a)
{
    if (a.n_query == 0)
    {
        return;
    }
    unsigned blocks = (a.n_query + kSearchThreads - 1) / kSearchThreads;
    size_t smem = 0;
    if (mode == SEARCH_RDF)
    {
        // persistent blocks so that each merges its private histogram once
        blocks = std::min<unsigned>(blocks, (unsigned) ctx->sm_count * 16U);
        smem = a.axis.bins * sizeof(uint32_t) <= 48 * 1024 ? a.axis.bins * sizeof(uint32_t) : 0;
    }
    switch (mode)
    {
    case SEARCH_COUNT:
    {
        KernelScope ks(ctx, "search_count");
        k_search<FLAVOUR, SEARCH_COUNT><<<blocks, kSearchThreads, 0, ctx->stream>>>(a);
        break;
    }
    case SEARCH_FILL:
    {
        KernelScope ks(ctx, "search_fill");
        k_search<FLAVOUR, SEARCH_FILL><<<blocks, kSearchThreads, 0, ctx->stream>>>(a);
        break;
    }
    case SEARCH_RDF:
    {
        KernelScope ks(ctx, "search_rdf_general");
        k_search<FLAVOUR, SEARCH_RDF><<<blocks, kSearchThreads, smem, ctx->stream>>>(a);
        break;
    }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace

void launch_search(fgpu_ctx* ctx, int flavour, SearchMode mode, const SearchArgs& args)
{
    if (flavour == FGPU_FLAVOUR_WRAP)
    {
        launch_search_mode<FGPU_FLAVOUR_WRAP>(ctx, mode, args);
    }
    else
    {
        if (flavour == FGPU_FLAVOUR_GHOST)
        {
            launch_search_mode<FGPU_FLAVOUR_GHOST>(ctx, mode, args);
        }
        else
        {
            launch_search_mode<FGPU_FLAVOUR_IMAGE>(ctx, mode, args);
        }
    }
}

void launch_emit(fgpu_ctx* ctx, int flavour, const EmitArgs& a)
{
    if (a.n_bonds == 0)
    {
        return;
    }
    unsigned const blocks = (unsigned) ((a.n_bonds + 255) / 256);
    {
        KernelScope ks(ctx, "emit_general");
        if (flavour == FGPU_FLAVOUR_WRAP)
        {
            k_emit<FGPU_FLAVOUR_WRAP><<<blocks, 256, 0, ctx->stream>>>(a);
        }
        else
        {
            if (flavour == FGPU_FLAVOUR_GHOST)
            {
                k_emit<FGPU_FLAVOUR_GHOST><<<blocks, 256, 0, ctx->stream>>>(a);
            }
            else
            {
                k_emit<FGPU_FLAVOUR_IMAGE><<<blocks, 256, 0, ctx->stream>>>(a);
            }
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_segments(fgpu_ctx* ctx, const uint32_t* row_start, const uint32_t* counts, uint32_t* segments,
                     uint32_t n_query)
{
    if (n_query == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "segments");
        k_segments<<<(n_query + 255) / 256, 256, 0, ctx->stream>>>(row_start, counts, segments, n_query);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

// CorrelationFunction::accumulate (freud/density/CorrelationFunction.cc:81-95) over a NeighborList: per bond the bin of
// its distance (RegularAxis(bins, 0, r_max)), one count and conj(values[j]) * query_values[i] in complex<double>.
// Block-shared accumulators (u32 + two doubles per bin) merged once per block; bins that do not fit shared memory
// go straight to global atomics.  The order of the double sums is not fixed -- upstream's thread-local sums are not
// either -- so results agree to double rounding, not bit for bit.
__global__ void __launch_bounds__(256) k_correlation(const uint32_t* __restrict__ neighbors, const float* __restrict__ distances,
                                                     uint64_t n_bonds, const double2* __restrict__ values,
                                                     const double2* __restrict__ query_values, AxisDev axis,
                                                     uint32_t* __restrict__ counts, double* __restrict__ sums,
                                                     int use_shared)
{
    extern __shared__ __align__(16) unsigned char corr_smem[];
    double* const s_sum = reinterpret_cast<double*>(corr_smem);                    // 2 * bins
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(s_sum + 2 * (size_t) axis.bins); // bins
    if (use_shared)
    {
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            s_sum[2 * b] = 0.0;
            s_sum[2 * b + 1] = 0.0;
            s_cnt[b] = 0;
        }
        __syncthreads();
    }
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < n_bonds; k += (uint64_t) gridDim.x * blockDim.x)
    {
        int const bin = axis_bin(axis, distances[k]);
        if (bin < 0)
        {
            continue;
        }
        uint2 const ij = reinterpret_cast<const uint2*>(neighbors)[k];
        double2 const x = values[ij.y], y = query_values[ij.x];
        // std::conj(x) * y, CorrelationFunction.cc:69-72
        double const re = __dadd_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y));
        double const im = __dsub_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x));
        if (use_shared)
        {
            atomicAdd(&s_cnt[bin], 1U);
            atomicAdd(&s_sum[2 * bin], re);
            atomicAdd(&s_sum[2 * bin + 1], im);
        }
        else
        {
            atomicAdd(&counts[bin], 1U);
            atomicAdd(&sums[2 * bin], re);
            atomicAdd(&sums[2 * bin + 1], im);
        }
    }
    if (use_shared)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            if (s_cnt[b] != 0)
            {
                atomicAdd(&counts[b], s_cnt[b]);
                atomicAdd(&sums[2 * b], s_sum[2 * b]);
                atomicAdd(&sums[2 * b + 1], s_sum[2 * b + 1]);
            }
        }
    }
}

void launch_correlation(fgpu_ctx* ctx, const uint32_t* neighbors, const float* distances, uint64_t n_bonds,
                        const double* values, const double* query_values, AxisDev axis, uint32_t* counts, double* sums)
{
    if (n_bonds == 0)
    {
        return;
    }
    size_t const smem = (size_t) axis.bins * (2 * sizeof(double) + sizeof(uint32_t));
    int const use_shared = smem <= 40 * 1024 ? 1 : 0;
    // few blocks: every block merges 3 words per occupied bin into the same global accumulators at the end
    unsigned const blocks = (unsigned) std::min<uint64_t>((n_bonds + 255) / 256, (uint64_t) ctx->sm_count * 4U);
    {
        KernelScope ks(ctx, "correlation");
        k_correlation<<<blocks, 256, use_shared ? smem : 0, ctx->stream>>>(
            neighbors, distances, n_bonds, reinterpret_cast<const double2*>(values),
            reinterpret_cast<const double2*>(query_values), axis, counts, sums, use_shared);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

// LocalDensity::compute (freud/density/LocalDensity.cc:38-84) over a NeighborList: one thread per query point walks
// its row in list order and sums, in float and in that order, 1 for a point wholly inside r_max and
// 1 + (r_max - (d + diameter/2)) / diameter for one that straddles it; density = count / area (2-D) or / volume.
// Rows without bonds stay 0 (upstream only ever writes inside the bond loop).
__global__ void __launch_bounds__(256) k_local_density(const uint32_t* __restrict__ row_start, const float* __restrict__ distances,
                                                       uint32_t n_query, float r_max, float diameter, float measure,
                                                       float* __restrict__ num_neighbors, float* __restrict__ density)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_query)
    {
        return;
    }
    uint32_t const beg = row_start[i], end = row_start[i + 1];
    float const half = __fdiv_rn(diameter, 2.0f);
    float const inner = __fsub_rn(r_max, half);
    float num = 0.0f;
    for (uint32_t b = beg; b < end; ++b)
    {
        float const d = distances[b];
        if (d < inner)
        {
            num = __fadd_rn(num, 1.0f);
        }
        else
        {
            float const part = __fdiv_rn(__fsub_rn(r_max, __fadd_rn(d, half)), diameter);
            num = __fadd_rn(num, __fadd_rn(1.0f, part));
        }
    }
    num_neighbors[i] = num;
    density[i] = beg == end ? 0.0f : __fdiv_rn(num, measure);
}

void launch_local_density(fgpu_ctx* ctx, const uint32_t* row_start, const float* distances, uint32_t n_query, float r_max,
                          float diameter, float measure, float* num_neighbors, float* density)
{
    if (n_query == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "local_density");
        k_local_density<<<(n_query + 255) / 256, 256, 0, ctx->stream>>>(row_start, distances, n_query, r_max, diameter,
                                                                        measure, num_neighbors, density);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

// LocalDensity::compute over the rows of the search's bag (the query was made for this compute alone, no
// NeighborList): 8 lanes per query row sum their bonds' contributions -- distance = sqrt(dot(v, v)) as
// NeighborBond.h:41-44 -- and are added up by a butterfly.  The order of the float sum is neither the list's nor
// upstream's iteration order (which is the engine's): results agree to float rounding, as on every on-the-fly query.
__global__ void __launch_bounds__(256) k_local_density_rows(const float4* __restrict__ bag, const uint32_t* __restrict__ row_bag_start,
                                                            const uint32_t* __restrict__ row_counts, uint32_t n_query,
                                                            float r_max, float diameter, float measure,
                                                            float* __restrict__ num_neighbors, float* __restrict__ density)
{
    uint32_t const row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7U;
    bool const live = row < n_query;
    uint32_t const n = live ? row_counts[row] : 0U, start = live ? row_bag_start[row] : 0U;
    float const half = __fdiv_rn(diameter, 2.0f);
    float const inner = __fsub_rn(r_max, half);
    float num = 0.0f;
    for (uint32_t k = sub; k < n; k += 8)
    {
        float4 const r = bag[start + k];
        float const d = __fsqrt_rn(dot_exact(r.x, r.y, r.z));
        num = __fadd_rn(num, d < inner ? 1.0f : __fadd_rn(1.0f, __fdiv_rn(__fsub_rn(r_max, __fadd_rn(d, half)), diameter)));
    }
    num = __fadd_rn(num, __shfl_xor_sync(0xffffffffU, num, 4));
    num = __fadd_rn(num, __shfl_xor_sync(0xffffffffU, num, 2));
    num = __fadd_rn(num, __shfl_xor_sync(0xffffffffU, num, 1));
    if (live && sub == 0)
    {
        num_neighbors[row] = num;
        density[row] = n == 0 ? 0.0f : __fdiv_rn(num, measure);
    }
}

void launch_local_density_rows(fgpu_ctx* ctx, const float4* bag, const uint32_t* row_bag_start, const uint32_t* row_counts,
                               uint32_t n_query, float r_max, float diameter, float measure, float* num_neighbors,
                               float* density)
{
    if (n_query == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "local_density_rows");
        k_local_density_rows<<<(unsigned) (((uint64_t) n_query * 8 + 255) / 256), 256, 0, ctx->stream>>>(
            bag, row_bag_start, row_counts, n_query, r_max, diameter, measure, num_neighbors, density);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

// CorrelationFunction::accumulate over the rows of the search's bag: k_correlation's bond body, 8 lanes per query row
__global__ void __launch_bounds__(256) k_correlation_rows(const float4* __restrict__ bag, const uint32_t* __restrict__ row_bag_start,
                                                          const uint32_t* __restrict__ row_counts, uint32_t n_query,
                                                          const double2* __restrict__ values,
                                                          const double2* __restrict__ query_values, AxisDev axis,
                                                          uint32_t* __restrict__ counts, double* __restrict__ sums,
                                                          int use_shared)
{
    extern __shared__ __align__(16) unsigned char corr_rows_smem[];
    double* const s_sum = reinterpret_cast<double*>(corr_rows_smem);
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(s_sum + 2 * (size_t) axis.bins);
    if (use_shared)
    {
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            s_sum[2 * b] = 0.0;
            s_sum[2 * b + 1] = 0.0;
            s_cnt[b] = 0;
        }
        __syncthreads();
    }
    uint32_t const sub = threadIdx.x & 7U, groups = gridDim.x * blockDim.x / 8U;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < n_query; row += groups)
    {
        uint32_t const n = row_counts[row], start = row_bag_start[row];
        double2 const y = query_values[row];
        for (uint32_t k = sub; k < n; k += 8)
        {
            float4 const r = bag[start + k];
            int const bin = axis_bin(axis, __fsqrt_rn(dot_exact(r.x, r.y, r.z)));
            if (bin < 0)
            {
                continue;
            }
            double2 const x = values[__float_as_uint(r.w)];
            double const re = __dadd_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)); // std::conj(x) * y
            double const im = __dsub_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x));
            if (use_shared)
            {
                atomicAdd(&s_cnt[bin], 1U);
                atomicAdd(&s_sum[2 * bin], re);
                atomicAdd(&s_sum[2 * bin + 1], im);
            }
            else
            {
                atomicAdd(&counts[bin], 1U);
                atomicAdd(&sums[2 * bin], re);
                atomicAdd(&sums[2 * bin + 1], im);
            }
        }
    }
    if (use_shared)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < axis.bins; b += blockDim.x)
        {
            if (s_cnt[b] != 0)
            {
                atomicAdd(&counts[b], s_cnt[b]);
                atomicAdd(&sums[2 * b], s_sum[2 * b]);
                atomicAdd(&sums[2 * b + 1], s_sum[2 * b + 1]);
            }
        }
    }
}

void launch_correlation_rows(fgpu_ctx* ctx, const float4* bag, const uint32_t* row_bag_start, const uint32_t* row_counts,
                             uint32_t n_query, const double* values, const double* query_values, AxisDev axis,
                             uint32_t* counts, double* sums)
{
    if (n_query == 0)
    {
        return;
    }
    size_t const smem = (size_t) axis.bins * (2 * sizeof(double) + sizeof(uint32_t));
    int const use_shared = smem <= 40 * 1024 ? 1 : 0;
    unsigned const blocks
        = (unsigned) std::min<uint64_t>(((uint64_t) n_query * 8 + 255) / 256, (uint64_t) ctx->sm_count * 8U);
    {
        KernelScope ks(ctx, "correlation_rows");
        k_correlation_rows<<<blocks, 256, use_shared ? smem : 0, ctx->stream>>>(
            bag, row_bag_start, row_counts, n_query, reinterpret_cast<const double2*>(values),
            reinterpret_cast<const double2*>(query_values), axis, counts, sums, use_shared);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_rdf_from_distances(fgpu_ctx* ctx, const float* distances, uint64_t n, AxisDev axis, uint32_t* hist)
{
    if (n == 0)
    {
        return;
    }
    unsigned const blocks = (unsigned) std::min<uint64_t>((n + 255) / 256, (uint64_t) ctx->sm_count * 16U);
    size_t const smem = axis.bins * sizeof(uint32_t) <= 48 * 1024 ? axis.bins * sizeof(uint32_t) : 0;
    {
        KernelScope ks(ctx, "rdf_distances");
        k_rdf_distances<<<blocks, 256, smem, ctx->stream>>>(distances, n, axis, hist);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
