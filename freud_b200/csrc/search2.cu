// Warp-cooperative 27-cell minimum-image pair search (the production path for regular grids).
//
// Replaces LinkCellQueryBallIterator::next (freud/locality/LinkCell.cc:496-573, flavour WRAP),
// AABBQueryBallIterator::next (freud/locality/AABBQuery.cc:77-150, flavour IMAGE), the gather half of
// NeighborQueryIterator::toNeighborList (freud/locality/NeighborQuery.h:434-458) and the on-the-fly branch of
// loopOverNeighbors with RDF's binning lambda (freud/locality/NeighborComputeFunctional.h:195-217,
// freud/density/RDF.cc:101-110).
//
// Mapping.  One warp owns one home TILE at a time: a span of consecutive cells of one grid row (tile_walk.cuh;
// the span is chosen so that a tile's candidates fill about four warp rounds), handed out by an atomic ticket.
// The candidate cells of the tile are <= 27 contiguous runs of the cell-ordered float4 array; the runs are
// flattened with a ballot/prefix scheme so that the 32 lanes load 32 consecutive CANDIDATES (coalesced 16-byte
// loads, two rounds held in registers) while the queries of the tile are broadcast from shared memory one
// after the other.  Every warp instruction therefore decides 32 (query, candidate) pairs.
//
// Two stages wherever a decision costs more than a filter (WRAP flavour in both modes, IMAGE flavour in
// NeighborList mode): stage 1 is a conservative filter (fused arithmetic on candidates pre-shifted to the image
// nearest to the home tile, acceptance radius r_max + 4E, E = bound on the rounding of both arithmetics) that
// rejects ~80 % of the pairs in 7 instructions; survivors are ballot-compacted into a per-warp stack of
// {candidate slot, query, boundary crossings} and stage 2 runs the reference's exact, un-fused arithmetic on 32
// stacked pairs at a time (dense lanes), then buffers or bins the hits.  For WRAP, stage 2 may assume what
// stage 1 established -- |fractional displacement| <= 1/3 + eps on every axis -- which makes two exact
// shortcuts legal (wrap_fast below).  For IMAGE the exact test is r = p - (q + image): the image vector of a
// candidate follows from how its cell was reached (all points inside the box, checked on the device; otherwise
// the general kernel in search.cu runs instead).  The fused RDF of the IMAGE flavour has no filter: its exact
// test is 10 instructions, so it decides inline and only stacks r_sq of the hits for the dense sqrt + bin.
//
// NeighborList mode writes each batch of complete rows (hits grouped by row, one 16-byte record per hit) to a
// temporary bag at a position reserved with one atomicAdd per batch, together with the row's count and bag
// offset; k_emit2 then walks the OUTPUT rows, ranks every hit inside its row and writes the five arrays.  RDF mode bins into one
// block-shared histogram with plain shared-memory atomics (measured 0.9 T increments/s, 5.7x faster than
// match_any aggregation: profiles/microbench_r1_hist_div.txt) and merges once per block.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "internal.h"
#include "tile_walk.cuh"

namespace fgpu {

namespace {

using tile::Cand;
using tile::FULL;
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kQueueCap = 96;   // stage-2 stack: < 32 left over + two pushes of <= 32

// Per-warp shared memory.  The stage-2 queue is a stack (push on top, pop the top 32): the order in which
// pairs reach stage 2 is irrelevant, rows are regrouped when a batch is flushed.
struct WarpMemBase
{
    tile::RunScratch runs;                        // non-empty candidate runs of the current home tile
    float4 query[32];                             // current query batch: x, y, z, bits(index to exclude)
    uint2 queue[kQueueCap];                       // stage-2 stack: {candidate slot, query k | crossing code << 8}; IMAGE+RDF: {r_sq, -}
};

struct WarpMemNL : WarpMemBase
{
    uint32_t qid[32];               // original index of the batch's queries
    uint32_t row_cnt[32], row_pos[32];
    // followed by out_cap x {k, j, x, y, z} (dynamic)
};

template<int MODE> struct WarpMemOf
{
    using type = WarpMemBase;
};
template<> struct WarpMemOf<S2_NL>
{
    using type = WarpMemNL;
};

__host__ __device__ inline size_t warp_mem_bytes(int mode, uint32_t out_cap)
{
    return mode == S2_NL ? sizeof(WarpMemNL) + (size_t) out_cap * 5 * sizeof(uint32_t) : sizeof(WarpMemBase);
}

template<int FLAVOUR, int MODE, bool TRI, bool SYM = false>
__global__ void __launch_bounds__(kThreads) k_search2(Search2Args a)
{
    static_assert(!SYM || (FLAVOUR == FGPU_FLAVOUR_IMAGE && MODE == S2_RDF), "symmetric walk: fused IMAGE RDF only");
    using WarpMem = typename WarpMemOf<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1U << lane) - 1U;
    size_t const hist_bytes = MODE == S2_RDF ? ((a.axis.bins * sizeof(uint32_t) + 15) / 16) * 16 : 0;
    uint32_t* const sh_hist = reinterpret_cast<uint32_t*>(smem_raw);
    unsigned char* const wbase = smem_raw + hist_bytes + (size_t) warp * warp_mem_bytes(MODE, a.out_cap);
    WarpMem& wm = *reinterpret_cast<WarpMem*>(wbase);
    float4* __restrict__ const sq = wm.query;
    uint2* __restrict__ const queue = wm.queue;
    // NL only (pointers are never dereferenced otherwise)
    uint32_t* __restrict__ const o_k = reinterpret_cast<uint32_t*>(wbase + sizeof(WarpMemNL));
    uint32_t* __restrict__ const o_j = o_k + a.out_cap;
    float* __restrict__ const o_x = reinterpret_cast<float*>(o_j + a.out_cap);
    float* __restrict__ const o_y = o_x + a.out_cap;
    float* __restrict__ const o_z = o_y + a.out_cap;
    const BoxDev& box = a.box;

    if (MODE == S2_NL && blockIdx.x == 0 && a.zero_words != nullptr)
    {
        // housekeeping for the row scan that follows this kernel on the stream (also when the kernel gives up below:
        // the scan still runs, on counts nobody reads, and must find its scratch clean)
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
        if (MODE == S2_RDF && a.push && blockIdx.x == 0)
        {
            // the frame fails on this rank (and, the points being replicated, on every rank): still arrive, so that
            // no peer's wait has to time out; fgpu_rdf_read raises on the sticky flag
            peer_push_block(a.peer, a.hist, 0);
        }
        return;
    }
    if (MODE == S2_RDF)
    {
        for (uint32_t b = threadIdx.x; b < a.axis.bins; b += blockDim.x)
        {
            sh_hist[b] = 0;
        }
        __syncthreads();
    }
    if (MODE == S2_NL)
    {
        reinterpret_cast<WarpMemNL&>(wm).row_cnt[lane] = 0;
        __syncwarp();
    }

    float const r_max_sq = __fmul_rn(a.r_max, a.r_max); // LinkCell.cc:498, AABBQuery.cc:79
    float const r_min_sq = __fmul_rn(a.r_min, a.r_min);
    float const r_hi_sq = a.r_hi_sq;
    float const knn_r_min = a.knn_r_min;
    // NeighborList mode always filters first and decides on dense lanes (stage 2); the fused RDF of the IMAGE
    // flavour decides inline, its exact test being as cheap as the filter
    constexpr bool FILTERED = FLAVOUR != FGPU_FLAVOUR_IMAGE || MODE == S2_NL;
    int const dx = a.dx, dy = a.dy, dz = a.dz;
    uint32_t q_len = 0;     // stage-2 stack height
    uint32_t o_len = 0;     // NL: buffered hits of the current batch
    uint32_t batch_n = 0;   // NL: queries of the current batch (their hits are buffered until the rows are complete)

    // ---- consumers ------------------------------------------------------------------------------------
    // NL: append the hits of one round (local row k) to the batch buffer
    bool overflow = false; // NL: the batch did not fit the buffer (uniform)
    auto buffer_hits = [&](bool hit, uint32_t k, uint32_t j, float rx, float ry, float rz) {
        WarpMemNL& w = reinterpret_cast<WarpMemNL&>(wm);
        unsigned const mh = __ballot_sync(FULL, hit);
        if (o_len + __popc(mh) > a.out_cap)
        {
            overflow = true;
            return;
        }
        if (hit)
        {
            uint32_t const pos = o_len + __popc(mh & lt_mask);
            o_k[pos] = k;
            o_j[pos] = j;
            o_x[pos] = rx;
            o_y[pos] = ry;
            o_z[pos] = rz;
            atomicAdd(&w.row_cnt[k], 1U);
        }
        o_len += __popc(mh);
    };
    // weight: 1, or 2 for a bond found by the symmetric walk (it stands for (i, j) and (j, i), tile_walk.cuh)
    auto bin_hit = [&](bool hit, float r_sq, uint32_t weight = 1U) {
        if (hit)
        {
            int const bin = axis_bin(a.axis, __fsqrt_rn(r_sq)); // NeighborBond distance = sqrt(dot(v, v))
            if (bin >= 0)
            {
                atomicAdd(&sh_hist[bin], weight);
            }
        }
    };

    // one dense round of stage 2: the top n <= 32 entries of the stack
    auto stage2_round = [&](uint32_t n) {
        __syncwarp();
        bool const act = (uint32_t) lane < n;
        uint32_t const e = q_len - n + lane;
        q_len -= n;
        if (FILTERED)
        {
            // the stack holds pairs that passed the conservative filter: run the reference's exact arithmetic
            bool hit = false;
            float rx = 0, ry = 0, rz = 0, r_sq = 0;
            uint32_t j = 0, k = 0;
            if (act)
            {
                uint2 const ent = queue[e];
                k = ent.y & 0xffU;
                float4 const p = __ldg(a.sorted + ent.x);
                float4 const q = sq[k];
                j = __float_as_uint(p.w);
                if (FLAVOUR == FGPU_FLAVOUR_WRAP)
                {
                    wrap_fast<TRI>(box, a.rcp_lx, a.rcp_ly, a.rcp_lz, __fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y),
                                   __fsub_rn(p.z, q.z), rx, ry, rz); // LinkCell.cc:522
                }
                else if (FLAVOUR == FGPU_FLAVOUR_GHOST)
                {
                    // r = (p + shift) - q with the ghost displacement of the crossed boundaries (none: + 0),
                    // CellQuery.cc:107, CellIterator.h:167
                    uint32_t const cd = ent.y >> 8;
                    float sx, sy, sz;
                    ghost_shift(box, (int) (cd & 3U) - 1, (int) ((cd >> 2) & 3U) - 1, (int) ((cd >> 4) & 3U) - 1, sx, sy,
                                sz);
                    bool const real = cd == tile::kNoWrap; // a real point is stored as it is, CellQuery.cc:121
                    rx = __fsub_rn(real ? p.x : __fadd_rn(p.x, sx), q.x);
                    ry = __fsub_rn(real ? p.y : __fadd_rn(p.y, sy), q.y);
                    rz = __fsub_rn(real ? p.z : __fadd_rn(p.z, sz), q.z);
                }
                else
                {
                    // r = p - (q + image), AABBQuery.cc:93,125; image 0 is +0 and is added like any other
                    float ix, iy, iz;
                    tile::code_image(box, ent.y >> 8, ix, iy, iz);
                    float const pz = box.is2d ? 0.0f : p.z; // AABBQuery.cc:118-122 (q.z was zeroed on load)
                    rx = __fsub_rn(p.x, __fadd_rn(q.x, ix));
                    ry = __fsub_rn(p.y, __fadd_rn(q.y, iy));
                    rz = __fsub_rn(pz, __fadd_rn(q.z, iz));
                }
                r_sq = dot_exact(rx, ry, rz);
                hit = in_window2(r_sq, r_max_sq, r_min_sq) && j != __float_as_uint(q.w);
                if (FLAVOUR == FGPU_FLAVOUR_IMAGE && knn_r_min > 0.0f)
                {
                    hit = hit && !(__fsqrt_rn(r_sq) < knn_r_min); // kNN filters on the distance, AABBQuery.cc:213
                }
            }
            if (MODE == S2_NL)
            {
                buffer_hits(hit, k, j, rx, ry, rz);
            }
            else
            {
                bin_hit(hit, r_sq);
            }
        }
        else
        {
            // IMAGE + RDF: the stack holds r_sq of accepted bonds (sign bit = counts twice, SYM only)
            uint32_t const bits = act ? queue[e].x : 0U;
            bin_hit(act, __uint_as_float(bits & 0x7fffffffU), SYM ? 1U + (bits >> 31) : 1U);
        }
        __syncwarp();
    };

    // NL: the batch is complete -> reserve bag space, write rows grouped, publish counts and offsets
    auto flush_batch = [&]() {
        WarpMemNL& w = reinterpret_cast<WarpMemNL&>(wm);
        __syncwarp();
        uint32_t const cnt = (uint32_t) lane < batch_n ? w.row_cnt[lane] : 0U;
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
        unsigned long long base = 0;
        if (lane == 0 && o_len != 0)
        {
            base = atomicAdd(a.cursor, (unsigned long long) o_len);
        }
        base = __shfl_sync(FULL, base, 0);
        bool const fits = base + o_len <= (unsigned long long) a.temp_cap;
        if ((uint32_t) lane < batch_n)
        {
            uint32_t const qi = w.qid[lane];
            a.counts[qi] = cnt;
            if (a.counts_copy != nullptr)
            {
                a.counts_copy[qi] = cnt;
            }
            a.tmp_start[qi] = ((uint32_t) base + (incl - cnt)) | a.tmp_flag;
            w.row_pos[lane] = incl - cnt;
            w.row_cnt[lane] = 0; // ready for the next batch
        }
        __syncwarp();
        if (fits)
        {
            for (uint32_t i = lane; i < o_len; i += 32)
            {
                uint32_t const k = o_k[i];
                uint32_t const rel = atomicAdd(&w.row_pos[k], 1U);
                uint32_t const dst = (uint32_t) base + rel;
                a.bag[dst] = make_float4(o_x[i], o_y[i], o_z[i], __uint_as_float(o_j[i]));
            }
        }
        o_len = 0;
        __syncwarp();
    };

    // ---- the pair loop: every query of the batch against the (up to) 64 candidates held in registers --------
    // WRAPPED: some candidate run of this home tile crosses a periodic boundary; TWO: c1 holds a second round
    auto pair_loop = [&](auto wrapped_tag, auto two_tag, const Cand& c0, const Cand& c1, uint32_t nqc) {
        constexpr bool WRAPPED = decltype(wrapped_tag)::value;
        constexpr bool TWO = decltype(two_tag)::value;
        for (uint32_t k = 0; k < nqc; ++k)
        {
            float4 const q = sq[k];
            uint32_t const q_excl = __float_as_uint(q.w);
#pragma unroll
            for (int h = 0; h < (TWO ? 2 : 1); ++h)
            {
                const Cand& c = h == 0 ? c0 : c1;
                if (FILTERED)
                {
                    // stage 1: conservative filter on the pre-shifted candidate (fused arithmetic is fine here)
                    float const ddx = c.x - q.x, ddy = c.y - q.y, ddz = c.z - q.z;
                    float const r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                    bool const ok = r2 <= r_hi_sq && c.j != q_excl; // LinkCell.cc:517-520, AABBQuery.cc:111-115
                    unsigned const m = __ballot_sync(FULL, ok);
                    if (ok)
                    {
                        queue[q_len + __popc(m & lt_mask)] = make_uint2(c.slot, WRAPPED ? k | ((c.code & tile::kCodeMask) << 8) : k | (tile::kNoWrap << 8));
                    }
                    q_len += __popc(m);
                }
                else
                {
                    float tx = q.x, ty = q.y, tz = q.z;
                    if (WRAPPED)
                    {
                        tx = __fadd_rn(q.x, c.ix);
                        ty = __fadd_rn(q.y, c.iy);
                        tz = __fadd_rn(q.z, c.iz);
                    }
                    float const rx = __fsub_rn(c.x, tx), ry = __fsub_rn(c.y, ty), rz = __fsub_rn(c.z, tz);
                    float const r_sq = dot_exact(rx, ry, rz);
                    bool const hit = in_window2(r_sq, r_max_sq, r_min_sq) && c.j != q_excl; // AABBQuery.cc:111-115
                    unsigned const m = __ballot_sync(FULL, hit);
                    if (hit)
                    {
                        queue[q_len + __popc(m & lt_mask)].x
                            = SYM ? __float_as_uint(r_sq) | (c.code & tile::kTwice) : __float_as_uint(r_sq);
                    }
                    q_len += __popc(m);
                }
            }
            if (q_len >= 32)
            {
                stage2_round(32);
                if (TWO && q_len >= 32)
                {
                    stage2_round(32);
                }
            }
        }
    };

    // ---- work loop: one home tile (a span of a.span cells of one grid row) per ticket ---------------------
    for (;;)
    {
        uint32_t ticket = 0;
        if (lane == 0)
        {
            ticket = atomicAdd(a.work_counter, 1U);
        }
        ticket = __shfl_sync(FULL, ticket, 0) + a.ticket_begin;
        if (ticket >= a.ticket_end)
        {
            break;
        }
        int cx0, cx1, cy, cz;
        tile::ticket_tile(ticket, a.spans_per_row, a.span, dx, dy, cx0, cx1, cy, cz);
        uint32_t const rowbase = ((uint32_t) cz * dy + cy) * dx;
        uint32_t const qs0 = __ldg(a.q_cell_start + rowbase + cx0), qs1 = __ldg(a.q_cell_start + rowbase + cx1 + 1);
        if (qs0 == qs1)
        {
            continue;
        }
        tile::Runs const runs = tile::setup_runs<SYM>(dx, dy, dz, a.cell_start, cx0, cx1, cy, cz, lane, wm.runs);
        uint32_t const T = runs.T;
        bool const any_wrap = runs.any_wrap;

        // ---- batches of queries ------------------------------------------------------------------------
        // NL: the hits of a batch are buffered until its rows are complete.  The batch size is optimistic
        // (a third of a query's own 27 cells may hit; an ideal gas gives 15.5 %); a batch that overflows the
        // buffer is discarded and redone at half the size, down to one query (whose hits fit if T does).
        uint32_t q_per_batch = 32;
        if (MODE == S2_NL)
        {
            if (T > a.out_cap)
            {
                if (lane == 0)
                {
                    *a.fail = 2; // a single row may exceed the buffer: the host retries with a larger one
                    atomicMax(reinterpret_cast<unsigned int*>(a.fail) + 1, T); // ... sized for the densest tile
                }
                q_per_batch = 0;
            }
            else
            {
                q_per_batch = min(32U, max(1U, (uint32_t) (cx1 - cx0 + 3) * a.out_cap / max(T, 1U)));
            }
        }
        for (uint32_t qb0 = qs0; q_per_batch != 0 && qb0 < qs1;)
        {
            uint32_t const nqc = min(q_per_batch, qs1 - qb0);
            __syncwarp();
            if ((uint32_t) lane < nqc)
            {
                float4 q = __ldg(a.q_sorted + qb0 + lane);
                uint32_t qi = __float_as_uint(q.w);
                if (MODE == S2_NL && a.q_remap != nullptr)
                {
                    qi = __ldg(a.q_remap + qi); // a subset of the rows is searched again (kNN, knn2.cu)
                }
                if (MODE == S2_NL)
                {
                    reinterpret_cast<WarpMemNL&>(wm).qid[lane] = qi;
                }
                if (FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d)
                {
                    q.z = 0.0f; // AABBQuery.cc:84-87
                }
                q.w = __uint_as_float(a.exclude_ii ? qi + a.q_index_offset : 0xffffffffU);
                sq[lane] = q;
            }
            __syncwarp();
            batch_n = nqc;
            for (uint32_t B = 0; B < T; B += 64)
            {
                Cand c0, c1;
                constexpr bool ZERO_Z = FLAVOUR == FGPU_FLAVOUR_IMAGE;
                tile::load_round<FILTERED, ZERO_Z>(runs, box, a.sorted, B, lane, c0);
                if (B + 32 < T)
                {
                    tile::load_round<FILTERED, ZERO_Z>(runs, box, a.sorted, B + 32, lane, c1);
                    if (any_wrap)
                    {
                        pair_loop(std::true_type {}, std::true_type {}, c0, c1, nqc);
                    }
                    else
                    {
                        pair_loop(std::false_type {}, std::true_type {}, c0, c1, nqc);
                    }
                }
                else if (any_wrap)
                {
                    pair_loop(std::true_type {}, std::false_type {}, c0, c0, nqc);
                }
                else
                {
                    pair_loop(std::false_type {}, std::false_type {}, c0, c0, nqc);
                }
            }
            if (FILTERED && q_len != 0)
            {
                // stacked pairs name their query by its slot in this batch (and NeighborList rows must be
                // complete before they are published): drain before the batch changes
                stage2_round(q_len);
            }
            if (MODE == S2_NL)
            {
                if (overflow)
                {
                    __syncwarp();
                    reinterpret_cast<WarpMemNL&>(wm).row_cnt[lane] = 0;
                    o_len = 0;
                    q_len = 0;
                    overflow = false;
                    q_per_batch = max(1U, nqc / 2);
                    continue; // same qb0, smaller batch
                }
                flush_batch();
            }
            qb0 += nqc;
        }
    }
    if (MODE == S2_RDF)
    {
        if (q_len != 0)
        {
            stage2_round(q_len);
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < a.axis.bins; b += blockDim.x)
        {
            uint32_t const v = sh_hist[b];
            if (v != 0)
            {
                atomicAdd(&a.hist[b], v); // u32 wraps like the reference's unsigned int counters
            }
        }
        if (a.push)
        {
            // compute + collective in one kernel: the block that merges last owns the finished histogram and sends it
            // to every rank over NVLink (peer.cuh); nobody waits here, the epoch's k_rdf_wait does
            __shared__ unsigned int s_last;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0)
            {
                s_last = atomicAdd(a.done_counter, 1U) == gridDim.x - 1 ? 1U : 0U;
            }
            __syncthreads();
            if (s_last != 0)
            {
                __threadfence();
                peer_push_block(a.peer, a.hist, a.axis.bins);
            }
        }
    }
}

// Pair evaluations of a query = candidates in its 27 cells, minus the excluded self pair: counted by a
// separate kernel (one thread per cell-ordered query) so that the search loops carry no instrumentation.
__global__ void __launch_bounds__(256) k_count_evals(Search2Args a, uint32_t n_query, const uint32_t* __restrict__ cell_of_point,
                                                     uint32_t n_points)
{
    if (*a.flag_points_outside != 0 || *a.flag_queries_outside != 0)
    {
        return;
    }
    // the cell-ordered queries of the cells this launch covers (all of them unless the tickets are sharded)
    uint32_t const slot0 = a.q_cell_start[a.cell_begin], slot1 = a.q_cell_start[a.cell_end];
    uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x + slot0;
    unsigned long long evals = 0;
    if (t < slot1 && t - slot0 < n_query)
    {
        float4 const q = a.q_sorted[t];
        int cx, cy, cz, nx, ny, nz;
        cell_coords(a.box, a.dx, a.dy, a.dz, q.x, q.y, q.z, cx, cy, cz, nx, ny, nz);
        for (int oz = (a.dz == 1 ? 0 : -1); oz <= (a.dz == 1 ? 0 : 1); ++oz)
        {
            for (int oy = -1; oy <= 1; ++oy)
            {
                int const y = (cy + oy + a.dy) % a.dy, z = (cz + oz + a.dz) % a.dz;
                uint32_t const rowbase = ((uint32_t) z * a.dy + y) * a.dx;
                for (int ox = -1; ox <= 1; ++ox)
                {
                    int const x = (cx + ox + a.dx) % a.dx;
                    evals += a.cell_start[rowbase + x + 1] - a.cell_start[rowbase + x];
                }
            }
        }
        if (a.exclude_ii)
        {
            uint32_t const j = __float_as_uint(q.w) + a.q_index_offset;
            if (j < n_points && cell_of_point == nullptr)
            {
                evals -= 1; // self query: the excluded point is the query itself, in its own cell
            }
            else if (j < n_points)
            {
                uint32_t const c = cell_of_point[j];
                if (c != 0xffffffffU) // else the point is outside this rank's slab: in none of the visited cells
                {
                    int const jx = c % a.dx, jy = (c / a.dx) % a.dy, jz = c / (a.dx * a.dy);
                    auto adjacent = [](int u, int v, int d) {
                        int const diff = (u - v + d) % d;
                        return diff == 0 || diff == 1 || diff == d - 1;
                    };
                    if (adjacent(jx, cx, a.dx) && adjacent(jy, cy, a.dy) && adjacent(jz, cz, a.dz))
                    {
                        evals -= 1;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        evals += __shfl_down_sync(FULL, evals, o);
    }
    if ((threadIdx.x & 31) == 0 && evals != 0)
    {
        atomicAdd(a.evals, evals);
    }
}

// ---- emit: blocks over output rows, threads over the bonds of those rows ------------------------------------
// (A variant that staged a row-id table and the keys of the block's window in shared memory and ranked in a second
// pass was measured slower -- 147 us against 115 us at configs[1]: the second read of the records and two more block
// barriers cost more than the binary search and the L1 re-reads they replaced.  Round 2 measured four more, all with
// identical output and none faster: a warp per chunk of whole rows (row by five shuffles, ranks by one shuffle per
// row member) 120 us; a 4-byte bag of point indices with the vector re-derived from the two points, which halves the
// DRAM reads (231 -> 116 MB) but costs 109 M instructions for 80 M: 136 us; a shared-memory row table instead of the
// binary search: 115 us; the record of the thread's next slot requested one iteration ahead: 125 us; and the offsets
// scan folded into this kernel by look-back over its blocks: 145 us, for the 15 us of the scan kernel it replaces.
// At 107 us under ncu the kernel runs DRAM at 50 %, issue slots at 72 % and L1 at 65 % with 85 % of the warps
// resident -- no single limiter to remove.)
// Ranks every hit inside its row (NeighborBond::less_as_tuple / less_as_distance restricted to one row with
// weight == 1, freud/locality/NeighborBond.h:80-112) and writes the five NeighborList arrays
// (NeighborQuery.h:470-478).  A block owns kEmitRows consecutive OUTPUT rows, so its stores cover one contiguous
// window of every output array (full sectors); the bag rows it gathers are contiguous 16-byte records, and the
// rank loop re-reads them from L1.
constexpr int kEmitRows = 256;

template<bool BY_DISTANCE> __global__ void __launch_bounds__(256) k_emit2(Emit2Args a)
{
    __shared__ uint32_t s_start[kEmitRows + 1];
    __shared__ uint32_t s_tmp[kEmitRows];
    {
        unsigned long long const total = *a.cursor;
        if (*a.fail != 0 || total > a.bag_cap || total > a.out_cap)
        {
            return; // the host repeats the step it has to repeat
        }
    }
    uint32_t const r0 = blockIdx.x * kEmitRows;
    uint32_t const n_rows = min((uint32_t) kEmitRows, a.n_query - r0);
    for (uint32_t i = threadIdx.x; i <= n_rows; i += blockDim.x)
    {
        s_start[i] = a.row_start[r0 + i];
        if (i < n_rows)
        {
            s_tmp[i] = a.tmp_start[r0 + i];
            a.segments[r0 + i] = a.counts[r0 + i] != 0 ? s_start[i] : 0U; // untouched rows stay 0 upstream
        }
    }
    __syncthreads();
    uint32_t const out_end = s_start[n_rows];
    for (uint32_t o = s_start[0] + threadIdx.x; o < out_end; o += blockDim.x)
    {
        // row of output slot o: s_start[row] <= o < s_start[row + 1] (empty rows are skipped by construction)
        uint32_t lo = 0, hi = n_rows;
        while (hi - lo > 1)
        {
            uint32_t const mid = (lo + hi) >> 1;
            if (s_start[mid] <= o)
            {
                lo = mid;
            }
            else
            {
                hi = mid;
            }
        }
        uint32_t const first = s_start[lo], n = s_start[lo + 1] - first;
        const float4* __restrict__ const row = a.bag + s_tmp[lo];
        float4 const h = row[o - first];
        uint32_t const j = __float_as_uint(h.w);
        float const d = __fsqrt_rn(dot_exact(h.x, h.y, h.z)); // NeighborBond.h:41-44
        uint32_t rank = 0;
        if (!BY_DISTANCE)
        {
            for (uint32_t k = 0; k < n; ++k)
            {
                rank += __float_as_uint(row[k].w) < j ? 1U : 0U;
            }
        }
        else
        {
            for (uint32_t k = 0; k < n; ++k)
            {
                float4 const t = row[k];
                float const od = __fsqrt_rn(dot_exact(t.x, t.y, t.z));
                uint32_t const oj = __float_as_uint(t.w);
                rank += (od < d || (od == d && oj < j)) ? 1U : 0U;
            }
        }
        uint64_t const out = (uint64_t) first + rank;
        reinterpret_cast<uint2*>(a.neighbors)[out] = make_uint2(r0 + lo, j);
        a.distances[out] = d;
        a.weights[out] = 1.0f;
        a.vectors[3 * out] = h.x;
        a.vectors[3 * out + 1] = h.y;
        a.vectors[3 * out + 2] = h.z;
    }
}

template<int FLAVOUR, int MODE, bool TRI, bool SYM = false>
void launch_one(fgpu_ctx* ctx, const Search2Args& a, const char* name)
{
    size_t const hist_bytes = MODE == S2_RDF ? ((a.axis.bins * sizeof(uint32_t) + 15) / 16) * 16 : 0;
    size_t const smem = hist_bytes + (size_t) kWarps * warp_mem_bytes(MODE, a.out_cap);
    auto kern = k_search2<FLAVOUR, MODE, TRI, SYM>;
    static bool configured = false; // per instantiation
    if (!configured)
    {
        FGPU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    int per_sm = 0;
    FGPU_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1)
    {
        throw Error(FGPU_ERUNTIME, "search kernel does not fit the shared memory of this device");
    }
    uint64_t const n_tickets = a.ticket_end - a.ticket_begin;
    unsigned const blocks
        = (unsigned) std::min<uint64_t>((uint64_t) ctx->sm_count * per_sm, (n_tickets + kWarps - 1) / kWarps);
    KernelScope ks(ctx, name);
    kern<<<std::max(blocks, 1U), kThreads, smem, ctx->stream>>>(a);
}

} // namespace

void search2_plan(Search2Args& a, uint32_t n_points, int span_override)
{
    // Span of the home tile along x: as many cells as keep the tile's candidates (9 or 3 rows of span + 2 cells)
    // near four warp rounds.  Sparse cells (cell width ~ r_max at liquid densities holds ~2 points) would
    // otherwise spend most of their instructions on per-tile setup rather than on pairs.
    double const per_cell = (double) n_points / (double) std::max(a.n_cells, 1U);
    double const rows = a.dz == 1 ? 3.0 : 9.0;
    int span = 1;
    if (span_override > 0)
    {
        span = span_override; // fgpu_ctx_set_tuning(ctx, "span", n): experiments only
    }
    else
    {
        while (span < 8 && rows * (span + 3) * per_cell <= 136.0)
        {
            ++span;
        }
    }
    span = std::max(1, std::min(span, a.dx));
    a.span = span;
    a.spans_per_row = (uint32_t) ((a.dx + span - 1) / span);
    a.n_tickets = a.spans_per_row * (uint32_t) a.dy * (uint32_t) a.dz;
    a.ticket_begin = 0;
    a.ticket_end = a.n_tickets;
    a.cell_begin = 0;
    a.cell_end = a.n_cells;
}

bool search2_supported(const Search2Args& a, int mode)
{
    bool const grid_ok = a.dx >= 3 && a.dy >= 3 && (a.dz >= 3 || (a.box.is2d && a.dz == 1));
    bool const hist_ok = mode != S2_RDF || a.axis.bins * sizeof(uint32_t) <= 64 * 1024;
    return grid_ok && hist_ok;
}

// Largest hit buffer a block can hold (four warps, 20 B per record, under the 200 KB opted in below); a tile with
// more candidates than this really needs the general kernels.
uint32_t search2_max_out_cap()
{
    return 2048;
}

uint32_t search2_out_cap(double expected_candidates_per_tile)
{
    // A warp buffers the hits of a batch of rows; a single row always fits a buffer that holds every candidate
    // of the tile, and a tile with more candidates than that sends the frame to the general kernel.  Room for
    // the mean + 7 sigma of a Poisson tile (uniform systems never trip it), 20 B per record; the buffer is
    // what limits the resident warps, so it is not rounded up to a power of two.
    double const mu = std::max(expected_candidates_per_tile, 1.0);
    double const want = mu + 7.0 * std::sqrt(mu) + 16.0;
    uint32_t const cap = ((uint32_t) std::min(want, 2048.0) + 31U) & ~31U;
    return std::max(cap, 128U);
}

void launch_count_evals(fgpu_ctx* ctx, const Search2Args& a, uint32_t n_query, const uint32_t* cell_of_point,
                        uint32_t n_points)
{
    if (n_query == 0 || a.evals == nullptr)
    {
        return;
    }
    {
        KernelScope ks(ctx, "count_evals");
        k_count_evals<<<(n_query + 255) / 256, 256, 0, ctx->stream>>>(a, n_query, cell_of_point, n_points);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_search2(fgpu_ctx* ctx, int flavour, int mode, const Search2Args& a)
{
    bool const tri = a.box.xy != 0.0f || a.box.xz != 0.0f || a.box.yz != 0.0f;
    const char* name = mode == S2_NL ? "search_nl" : "search_rdf";
#define FGPU_S2(FL, MD)                                                                                          \
    do                                                                                                           \
    {                                                                                                            \
        if (tri)                                                                                                 \
            launch_one<FL, MD, true>(ctx, a, name);                                                              \
        else                                                                                                     \
            launch_one<FL, MD, false>(ctx, a, name);                                                             \
    } while (0)
    if (mode == S2_NL && a.lanes_over_queries)
    {
        launch_search_lq(ctx, flavour, a); // search_lq.cu
    }
    else if (flavour == FGPU_FLAVOUR_WRAP)
    {
        if (mode == S2_NL)
            FGPU_S2(FGPU_FLAVOUR_WRAP, S2_NL);
        else
            FGPU_S2(FGPU_FLAVOUR_WRAP, S2_RDF);
    }
    else if (flavour == FGPU_FLAVOUR_GHOST)
    {
        if (mode == S2_NL)
            launch_one<FGPU_FLAVOUR_GHOST, S2_NL, false>(ctx, a, name);
        else
            launch_one<FGPU_FLAVOUR_GHOST, S2_RDF, false>(ctx, a, name);
    }
    else
    {
        if (mode == S2_NL)
            FGPU_S2(FGPU_FLAVOUR_IMAGE, S2_NL);
        else if (a.symmetric)
            launch_one<FGPU_FLAVOUR_IMAGE, S2_RDF, false, true>(ctx, a, name); // TRI only matters to wrap_fast
        else
            launch_one<FGPU_FLAVOUR_IMAGE, S2_RDF, false, false>(ctx, a, name);
    }
#undef FGPU_S2
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_emit2(fgpu_ctx* ctx, int sort_by_distance, const Emit2Args& a)
{
    if (a.n_query == 0)
    {
        return;
    }
    unsigned const blocks = (a.n_query + kEmitRows - 1) / kEmitRows;
    {
        KernelScope ks(ctx, "emit");
        if (sort_by_distance)
        {
            k_emit2<true><<<blocks, 256, 0, ctx->stream>>>(a);
        }
        else
        {
            k_emit2<false><<<blocks, 256, 0, ctx->stream>>>(a);
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
