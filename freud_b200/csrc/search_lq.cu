// Lanes over queries: the NeighborList search for sparse cells (a handful of points per cell, rows of ~10 bonds).
//
// Same contract as k_search2<FLAVOUR, S2_NL> (search2.cu): LinkCellQueryBallIterator::next
// (freud/locality/LinkCell.cc:496-573, flavour WRAP), AABBQueryBallIterator::next (freud/locality/AABBQuery.cc:77-150,
// flavour IMAGE), CellQuery's ball iterator (freud/locality/CellQuery.cc:96-130, flavour GHOST) and the gather half of
// NeighborQueryIterator::toNeighborList (freud/locality/NeighborQuery.h:434-458): rows of the bag, counts and bag
// offsets per query, total in *cursor.
//
// Why a second mapping.  At configs[1] (1 M points, r_max = 3, 2.2 points per cell) the tile walk spends 224 M warp
// instructions on 59 M pair tests: a home tile holds ~9 queries, so tile setup, candidate flattening, the ballot
// compaction of every filter round and the regrouping of hits by row are paid per handful of queries, and every query
// is tested against the candidates of the whole tile (117 for the 58 of its own 27 cells).  Here a warp takes 32
// consecutive cell-ordered queries, one per lane:
//
//   filter   each lane walks ITS 27 cells as nine runs of the cell-ordered array (x is the fastest cell index, so the
//            three x neighbours of a row are contiguous; the neighbour across the periodic x boundary is a run of its
//            own).  The loop is a plain per-lane loop -- no vote, no shared counter: a survivor of the conservative
//            filter (fused arithmetic, radius r_max + 4E, query pre-shifted by the run's boundary crossings) is
//            pushed on the lane's PRIVATE stack in shared memory (stride 32 words: conflict-free).  15 instructions
//            per candidate trip; lanes in neighbouring cells read neighbouring candidates, which L1 serves.
//   reserve  one warp scan over the stack heights numbers the survivors of the 32 rows consecutively (rows ascending)
//            and one atomicAdd on the bag cursor reserves a record for each of them.
//   decide   the reference's exact un-fused arithmetic on 32 survivors at a time (dense lanes, as the tile walk's
//            stage 2; survivor e belongs to the row whose range of numbers holds e: five shuffles), hits
//            ballot-compacted straight into the reserved bag records (coalesced 16-byte stores) -- still grouped by
//            row, so nothing has to regroup them and no record is staged in shared memory.
//   publish  row counts = stack heights minus the (rare: the filter radius is r_max + 4E) exact rejections; one more
//            scan gives every row its bag offset.  The rejected survivors leave a few unused records at the end of a
//            warp's reservation: *cursor is an upper bound of the bond count (by a few in a thousand), the counts are
//            exact.
//
// A lane whose survivors exceed its stack (sized mean + 7 sigma of an ideal gas) raises fail = 2 and the host repeats
// the frame with the tile walk, which adapts to dense tiles.
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "internal.h"
#include "tile_walk.cuh"

namespace fgpu {

namespace {

constexpr unsigned FULL = 0xffffffffU;
constexpr int kLqWarps = 4;
constexpr int kLqThreads = kLqWarps * 32;
constexpr uint32_t kLqHead = 32 * sizeof(float4) + 32 * sizeof(uint32_t); // queries of the batch, row counts

// per warp: head | private stacks (32 lanes x c words, word i of lane l at [32 i + l])
__host__ __device__ inline size_t lq_warp_bytes(uint32_t c)
{
    return kLqHead + (size_t) 32 * c * sizeof(uint32_t);
}

template<int FLAVOUR, bool TRI> __global__ void __launch_bounds__(kLqThreads, 10) k_search_lq(Search2Args a)
{
    constexpr bool CODED = FLAVOUR != FGPU_FLAVOUR_WRAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1U << lane) - 1U;
    uint32_t const C = a.lq_c;
    unsigned char* const wbase = smem_raw + (size_t) warp * lq_warp_bytes(C);
    float4* __restrict__ const sq = reinterpret_cast<float4*>(wbase);
    uint32_t* __restrict__ const row_cnt = reinterpret_cast<uint32_t*>(wbase + 32 * sizeof(float4));
    uint32_t* __restrict__ const priv = reinterpret_cast<uint32_t*>(wbase + kLqHead);
    const BoxDev& box = a.box;

    if (blockIdx.x == 0 && a.zero_words != nullptr)
    {
        // housekeeping for the row scan that follows this kernel on the stream (as k_search2 does)
        for (uint32_t i = threadIdx.x; i < a.zero_n; i += blockDim.x)
        {
            a.zero_words[i] = 0U;
        }
        if (threadIdx.x == 0)
        {
            *a.zero_tail = 0U;
        }
    }
    // points or queries outside the box: image offsets are not implied by the cell walk -> general kernel
    if (*a.flag_points_outside != 0 || *a.flag_queries_outside != 0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            *a.fail = 1;
        }
        return;
    }

    float const r_max_sq = __fmul_rn(a.r_max, a.r_max); // LinkCell.cc:498, AABBQuery.cc:79
    float const r_min_sq = __fmul_rn(a.r_min, a.r_min);
    float const r_hi_sq = a.r_hi_sq;
    float const knn_r_min = a.knn_r_min;
    int const dx = a.dx, dy = a.dy, dz = a.dz;
    bool const zero_z = FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d; // AABBQuery.cc:84-87, 118-122
    uint32_t const top_max = (uint32_t) lane + 32U * (C - 1U);

    for (;;)
    {
        uint32_t ticket = 0;
        if (lane == 0)
        {
            ticket = atomicAdd(a.work_counter, 1U);
        }
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= a.lq_tickets)
        {
            break;
        }
        uint32_t const t0 = ticket * 32U;
        bool const mine = t0 + (uint32_t) lane < a.n_query;
        __syncwarp(); // the previous batch is done with the stacks and the query slots
        float4 q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        uint32_t qi = 0;
        if (mine)
        {
            q = __ldg(a.q_sorted + t0 + lane);
            qi = __float_as_uint(q.w);
            if (a.q_remap != nullptr)
            {
                qi = __ldg(a.q_remap + qi); // a subset of the rows is searched again (kNN, knn2.cu)
            }
            if (zero_z)
            {
                q.z = 0.0f;
            }
        }
        uint32_t const q_excl = a.exclude_ii ? qi + a.q_index_offset : 0xffffffffU;
        sq[lane] = make_float4(q.x, q.y, q.z, __uint_as_float(q_excl));

        // ---- filter: the lane's nine runs ------------------------------------------------------------------------
        int cx = 0, cy = 0, cz = 0;
        {
            int nx, ny, nz;
            cell_coords(box, dx, dy, dz, q.x, q.y, q.z, cx, cy, cz, nx, ny, nz); // the arithmetic that sorted it there
        }
        int const x0 = max(cx - 1, 0), x1 = min(cx + 1, dx - 1);
        bool const edge = mine && (cx == 0 || cx == dx - 1);
        bool const any_edge = __any_sync(FULL, edge);
        int const r_first = dz == 1 ? 3 : 0, r_last = dz == 1 ? 6 : 9; // 2-D: the three rows of the plane
        uint32_t top = (uint32_t) lane;                                // next free word of the lane's stack
        // run r of the lane: row (oy, oz) = (r % 3 - 1, r / 3 - 1) of its cell, cells [x0, x1]
        auto run_of = [&](int r, uint32_t& b, uint32_t& e, int& wy, int& wz, uint32_t& rowbase) {
            int const oz = r / 3 - 1, oy = r - 3 * (r / 3) - 1;
            int y = cy + oy, z = cz + oz;
            wy = y < 0 ? -1 : (y >= dy ? 1 : 0);
            wz = z < 0 ? -1 : (z >= dz ? 1 : 0);
            y -= wy * dy;
            z -= wz * dz;
            rowbase = ((uint32_t) z * dy + y) * dx;
            b = mine ? __ldg(a.cell_start + rowbase + x0) : 0U;
            e = mine ? __ldg(a.cell_start + rowbase + x1 + 1) : 0U;
        };
        // candidates [b, e) against the query moved to (qx, qy, qz); nothing in the loop is shared between lanes
        auto walk = [&](uint32_t b, uint32_t e, float qx, float qy, float qz, uint32_t tag) {
            for (uint32_t s = b; s < e; ++s)
            {
                float4 const p = __ldg(a.sorted + s);
                float const ddx = p.x - qx, ddy = p.y - qy, ddz = (zero_z ? 0.0f : p.z) - qz;
                float const r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                if (r2 <= r_hi_sq && __float_as_uint(p.w) != q_excl) // LinkCell.cc:517-520, AABBQuery.cc:111-115
                {
                    priv[top] = CODED ? (s | tag) : s;
                    top = min(top + 32U, top_max);
                }
            }
        };
        uint32_t nb = 0, ne = 0, nrow = 0;
        int nwy = 0, nwz = 0;
        run_of(r_first, nb, ne, nwy, nwz, nrow);
        for (int r = r_first; r < r_last; ++r)
        {
            uint32_t const b = nb, e = ne, rowbase = nrow;
            int const wy = nwy, wz = nwz;
            if (r + 1 < r_last)
            {
                run_of(r + 1, nb, ne, nwy, nwz, nrow); // the next row's offsets travel while this row is walked
            }
            // the candidates of a run reached across boundaries (wy, wz) are images shifted by wy b + wz c: move the
            // query the other way instead (fused arithmetic, filter only)
            float const fy = (float) wy, fz = (float) wz;
            float const qx = q.x - (fy * box.bx + fz * box.cx), qy = q.y - (fy * box.by + fz * box.cy),
                        qz = q.z - fz * box.cz;
            uint32_t const code = ((uint32_t) (wy + 1) << 2) | ((uint32_t) (wz + 1) << 4);
            if (any_edge)
            {
                // the x neighbour across the periodic boundary is a run of its own (one cell), walked first so that
                // the lanes that have one are back before the long runs end
                uint32_t b2 = 0, e2 = 0;
                int const wx = cx == 0 ? -1 : 1;
                if (edge)
                {
                    uint32_t const cell = rowbase + (uint32_t) (cx == 0 ? dx - 1 : 0);
                    b2 = __ldg(a.cell_start + cell);
                    e2 = __ldg(a.cell_start + cell + 1);
                }
                walk(b2, e2, qx - (float) wx * box.ax, qy, qz, (code | (uint32_t) (wx + 1)) << 26);
            }
            walk(b, e, qx, qy, qz, (code | 1U) << 26);
        }

        // ---- reserve: number the survivors of the 32 rows consecutively, one bag record each -----------------------
        uint32_t const cnt = (top - (uint32_t) lane) >> 5;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o)
            {
                incl += t;
            }
        }
        uint32_t const S = __shfl_sync(FULL, incl, 31);
        uint32_t const first = incl - cnt; // number of the lane's first survivor
        if (__any_sync(FULL, cnt + 1U >= C))
        {
            // a row beyond its stack (a clustered system): the host repeats the frame with the tile walk
            if (lane == 0)
            {
                *a.fail = 2;
            }
            continue;
        }
        unsigned long long base = 0;
        if (lane == 0 && S != 0)
        {
            base = atomicAdd(a.cursor, (unsigned long long) S);
        }
        base = __shfl_sync(FULL, base, 0);
        bool const fits = base + S <= (unsigned long long) a.temp_cap;
        row_cnt[lane] = cnt;
        __syncwarp();

        // ---- decide: the reference's exact arithmetic, 32 survivors at a time --------------------------------------
        uint32_t n_rec = 0;
        for (uint32_t t = 0; t < S; t += 32)
        {
            uint32_t const e = t + (uint32_t) lane;
            // row of survivor e: the last lane whose first number is <= e (empty rows share their successor's number
            // and lose the comparison to it)
            uint32_t k = 0;
#pragma unroll
            for (uint32_t step = 16; step != 0; step >>= 1)
            {
                uint32_t const f = __shfl_sync(FULL, first, (int) (k + step));
                if (f <= e)
                {
                    k += step;
                }
            }
            uint32_t const first_k = __shfl_sync(FULL, first, (int) k);
            bool hit = false;
            float rx = 0.0f, ry = 0.0f, rz = 0.0f;
            uint32_t j = 0;
            if (e < S)
            {
                uint32_t const ent = priv[k + 32U * (e - first_k)];
                uint32_t const slot = CODED ? ent & 0x3ffffffU : ent;
                uint32_t const cd = CODED ? ent >> 26 : tile::kNoWrap;
                float4 const p = __ldg(a.sorted + slot);
                float4 const qq = sq[k];
                j = __float_as_uint(p.w);
                if (FLAVOUR == FGPU_FLAVOUR_WRAP)
                {
                    wrap_fast<TRI>(box, a.rcp_lx, a.rcp_ly, a.rcp_lz, __fsub_rn(p.x, qq.x), __fsub_rn(p.y, qq.y),
                                   __fsub_rn(p.z, qq.z), rx, ry, rz); // LinkCell.cc:522
                }
                else if (FLAVOUR == FGPU_FLAVOUR_GHOST)
                {
                    // r = (p + shift) - q with the ghost displacement of the crossed boundaries (none: the point as it
                    // is stored), CellQuery.cc:107, 121, CellIterator.h:167
                    float sx, sy, sz;
                    ghost_shift(box, (int) (cd & 3U) - 1, (int) ((cd >> 2) & 3U) - 1, (int) ((cd >> 4) & 3U) - 1, sx, sy,
                                sz);
                    bool const real = cd == tile::kNoWrap;
                    rx = __fsub_rn(real ? p.x : __fadd_rn(p.x, sx), qq.x);
                    ry = __fsub_rn(real ? p.y : __fadd_rn(p.y, sy), qq.y);
                    rz = __fsub_rn(real ? p.z : __fadd_rn(p.z, sz), qq.z);
                }
                else
                {
                    // r = p - (q + image), AABBQuery.cc:93,125; image 0 is +0 and is added like any other
                    float ix, iy, iz;
                    tile::code_image(box, cd, ix, iy, iz);
                    float const pz = box.is2d ? 0.0f : p.z; // AABBQuery.cc:118-122 (q.z was zeroed on load)
                    rx = __fsub_rn(p.x, __fadd_rn(qq.x, ix));
                    ry = __fsub_rn(p.y, __fadd_rn(qq.y, iy));
                    rz = __fsub_rn(pz, __fadd_rn(qq.z, iz));
                }
                float const r_sq = dot_exact(rx, ry, rz);
                hit = in_window2(r_sq, r_max_sq, r_min_sq); // the excluded pair never reaches the stack
                if (FLAVOUR == FGPU_FLAVOUR_IMAGE && knn_r_min > 0.0f)
                {
                    hit = hit && !(__fsqrt_rn(r_sq) < knn_r_min); // kNN filters on the distance, AABBQuery.cc:213
                }
                if (!hit)
                {
                    atomicSub(&row_cnt[k], 1U); // a few in a thousand
                }
            }
            unsigned const mh = __ballot_sync(FULL, hit);
            if (hit && fits)
            {
                a.bag[(uint32_t) base + n_rec + __popc(mh & lt_mask)] = make_float4(rx, ry, rz, __uint_as_float(j));
            }
            n_rec += __popc(mh);
        }
        __syncwarp();

        // ---- publish: counts and bag offsets of the 32 rows ---------------------------------------------------------
        uint32_t const rc = row_cnt[lane];
        uint32_t incl2 = rc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const t = __shfl_up_sync(FULL, incl2, o);
            if (lane >= o)
            {
                incl2 += t;
            }
        }
        if (mine)
        {
            a.counts[qi] = rc;
            if (a.counts_copy != nullptr)
            {
                a.counts_copy[qi] = rc;
            }
            a.tmp_start[qi] = ((uint32_t) base + (incl2 - rc)) | a.tmp_flag;
        }
    }
}

template<int FLAVOUR, bool TRI> void launch_lq(fgpu_ctx* ctx, const Search2Args& a)
{
    size_t const smem = (size_t) kLqWarps * lq_warp_bytes(a.lq_c);
    auto kern = k_search_lq<FLAVOUR, TRI>;
    static bool configured = false; // per instantiation
    if (!configured)
    {
        FGPU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    int per_sm = 0;
    FGPU_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLqThreads, smem));
    if (per_sm < 1)
    {
        throw Error(FGPU_ERUNTIME, "search kernel does not fit the shared memory of this device");
    }
    if (ctx->tune_lq_blocks > 0)
    {
        per_sm = std::min(per_sm, ctx->tune_lq_blocks);
    }
    // shared memory for exactly these blocks; what is left of the SM's 256 KB serves the candidate loads as L1
    static int carved_for = -1;
    if (carved_for != per_sm * 4096 + (int) (smem / 64))
    {
        int const percent = (int) std::min<size_t>(100, ((size_t) per_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        FGPU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, percent));
        carved_for = per_sm * 4096 + (int) (smem / 64);
    }
    unsigned const blocks = (unsigned) std::min<uint64_t>((uint64_t) ctx->sm_count * per_sm,
                                                          ((uint64_t) a.lq_tickets + kLqWarps - 1) / kLqWarps);
    KernelScope ks(ctx, "search_nl");
    kern<<<std::max(blocks, 1U), kLqThreads, smem, ctx->stream>>>(a);
}

} // namespace

// Mapping of the NeighborList search.  Lanes over queries pays per candidate of a query's own 27 cells and needs a
// stack per row sized for the row's bonds; the tile walk pays per candidate of a tile and per tile, and its buffers
// adapt to dense tiles.  Automatic choice: lanes over queries while a row's stack stays within 96 entries (rows of up
// to ~40 bonds).  Measured against the tile walk: configs[1] at 9 bonds per row 272 -> 171 us, the kNN window search
// of configs[2] at ~20 per row 452 -> 278 us, the 2-D ball query of the PMFT benches at 39 per row 553 -> 466 us;
// everything denser stays on the tile walk.
//
// The fused RDF was tried in this mapping too (IMAGE arithmetic, r_sq of a bond on the lane's stack, every lane
// binning its own stack): 314 us against the tile walk's 237 us at 1 M points, r_max = 5 -- the exact test costs the
// same 16 instructions per pair in both, the tile walk runs it on dense lanes (lanes over candidates) where this
// mapping walks to the longest of 32 runs (65 % of the lanes busy): 270 M instructions against 227 M.  Not in the tree.
void search2_choose_mapping(Search2Args& a, int flavour, uint32_t n_query, uint32_t n_points,
                            double expected_hits_per_query, int force)
{
    a.n_query = n_query;
    a.lanes_over_queries = 0;
    a.lq_tickets = (n_query + 31U) / 32U;
    // survivors of the filter: the bonds, a few in a thousand more, and nothing for the excluded pair
    double const mu = std::max(expected_hits_per_query * 1.01, 0.5);
    double const per_lane = mu + 7.0 * std::sqrt(mu) + 4.0;
    a.lq_c = ((uint32_t) std::min(per_lane, 1024.0) + 3U) & ~3U;
    a.lq_f = 0;
    // a stack entry of the IMAGE / GHOST flavours shares its word with the boundary crossings (6 bits)
    bool const fits_word = flavour == FGPU_FLAVOUR_WRAP || n_points <= (1U << 26);
    bool const possible = n_query != 0 && fits_word && lq_warp_bytes(a.lq_c) * kLqWarps <= 160 * 1024;
    bool const automatic = a.lq_c <= 96;
    if (possible && (force > 0 || (force < 0 && automatic)))
    {
        a.lanes_over_queries = 1;
    }
}

bool search2_lq_fallback(Search2Args& a, int fail)
{
    if (a.lanes_over_queries != 0 && fail == 2)
    {
        a.lanes_over_queries = 0;
        return true;
    }
    return false;
}

void launch_search_lq(fgpu_ctx* ctx, int flavour, const Search2Args& a)
{
    bool const tri = a.box.xy != 0.0f || a.box.xz != 0.0f || a.box.yz != 0.0f;
    if (flavour == FGPU_FLAVOUR_WRAP)
    {
        if (tri)
        {
            launch_lq<FGPU_FLAVOUR_WRAP, true>(ctx, a);
        }
        else
        {
            launch_lq<FGPU_FLAVOUR_WRAP, false>(ctx, a);
        }
    }
    else if (flavour == FGPU_FLAVOUR_GHOST)
    {
        launch_lq<FGPU_FLAVOUR_GHOST, false>(ctx, a);
    }
    else
    {
        launch_lq<FGPU_FLAVOUR_IMAGE, false>(ctx, a);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
