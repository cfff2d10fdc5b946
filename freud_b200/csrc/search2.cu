// Warp-cooperative 27-cell minimum-image pair search (the production path for regular grids).
//
// Replaces LinkCellQueryBallIterator::next (freud/locality/LinkCell.cc:496-573, flavour WRAP),
// AABBQueryBallIterator::next (freud/locality/AABBQuery.cc:77-150, flavour IMAGE), the gather half of
// NeighborQueryIterator::toNeighborList (freud/locality/NeighborQuery.h:434-458) and the on-the-fly branch of
// loopOverNeighbors with RDF's binning lambda (freud/locality/NeighborComputeFunctional.h:195-217,
// freud/density/RDF.cc:101-110).
//
// Mapping.  One warp owns one home cell at a time (work items are runs of kCellChunk consecutive cells, handed
// out by an atomic ticket).  The 27 neighbour cells are <= 18 contiguous runs of the cell-ordered float4
// array (3 x-adjacent cells of a (y, z) row are one run; the periodic x boundary splits a run in two); the
// runs are flattened with a ballot/prefix scheme so that the 32 lanes load 32 consecutive CANDIDATES
// (coalesced 16-byte loads, two rounds held in registers) while the queries of the home cell are broadcast
// from shared memory one after the other.  Every warp instruction therefore decides 32 (query, candidate)
// pairs.
//
// WRAP flavour, two stages: stage 1 is a conservative filter (fused arithmetic on pre-shifted candidates,
// acceptance radius r_max + 4E, E = bound on the rounding of both arithmetics) that rejects ~80 % of the
// pairs in 7 instructions; survivors are ballot-compacted into a per-warp queue and stage 2 runs the
// reference's exact, un-fused arithmetic on 32 queued pairs at a time (dense lanes).  Stage 2 may assume
// what stage 1 established -- |fractional displacement| <= 1/3 + eps on every axis -- which makes two exact
// shortcuts legal (wrap_fast below).
// IMAGE flavour: the exact test r = p - (q + image) is only 10 instructions, so there is no filter; the image
// vector of a candidate follows from how its cell was reached (all points inside the box, checked on the
// device; otherwise the general kernel in search.cu runs instead).
//
// NeighborList mode writes each batch of complete rows (hits grouped by row, 20 B per hit) to a temporary
// bag at a position reserved with one atomicAdd per batch, together with the row's count and bag offset;
// k_emit2 then ranks every hit inside its row and writes the five output arrays.  RDF mode bins into one
// block-shared histogram with plain shared-memory atomics (measured 0.9 T increments/s, 5.7x faster than
// match_any aggregation: profiles/microbench_r1_hist_div.txt) and merges once per block.
#include <type_traits>

#include "internal.h"

namespace fgpu {

namespace {

constexpr unsigned FULL = 0xffffffffU;
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kCellChunk = 8;   // consecutive home cells per work ticket
constexpr int kQueueCap = 96;   // stage-2 stack: < 32 left over + two pushes of <= 32

__device__ __forceinline__ bool in_window2(float r_sq, float r_max_sq, float r_min_sq)
{
    return r_sq < r_max_sq && r_sq >= r_min_sq; // LinkCell.cc:525, AABBQuery.cc:129
}

// Correctly rounded a / L for normal-range operands: q0 = RN(a * y) with y = RN(1 / L) (rounded on the host),
// one fused correction makes the quotient faithful, the second one makes it RN(a / L) (Markstein's theorem;
// the residuals a - L * q are exact in an FMA).  Replaces __fdiv_rn (x86 divss upstream, Box.h:248-250) at
// 5 instead of ~15 issue slots.  Valid because stage 1 bounds |a| / L away from 0 (no underflow).
__device__ __forceinline__ float div_by_const(float a, float L, float y)
{
    float const q0 = __fmul_rn(a, y);
    float const r0 = __fmaf_rn(-q0, L, a);
    float const q1 = __fmaf_rn(r0, y, q0);
    float const r1 = __fmaf_rn(-q1, L, a);
    return __fmaf_rn(r1, y, q1);
}

// util::modulusPositive(f, 1) = fmodf(fmodf(f, 1) + 1, 1) (freud/util/utils.h:29-32) for f in (-1, 2) whose
// intermediate sum stays below 2: truncation is a compare (FSET), not an FRND.
__device__ __forceinline__ float modulus_positive_one_small(float f)
{
    float const t = __fsub_rn(f, f >= 1.0f ? 1.0f : 0.0f);
    float const u = __fadd_rn(t, 1.0f);
    return __fsub_rn(u, u >= 1.0f ? 1.0f : 0.0f);
}

// Box::wrap(v) (freud/box/Box.h:307-329) for displacements whose fractional coordinates are within
// (-1/2 - 0.35, 1/2 + 0.35) + {-1, 0, 1}; bit-identical to wrap_exact on that domain.
template<bool TRI>
__device__ __forceinline__ void wrap_fast(const BoxDev& b, float ylx, float yly, float ylz, float vx, float vy,
                                          float vz, float& rx, float& ry, float& rz)
{
    float dx = __fsub_rn(vx, b.lox);
    float dy = __fsub_rn(vy, b.loy);
    float const dz = __fsub_rn(vz, b.loz);
    if (TRI)
    {
        dx = __fsub_rn(dx, __fadd_rn(__fmul_rn(b.t_xz, vz), __fmul_rn(b.xy, vy)));
        dy = __fsub_rn(dy, __fmul_rn(b.yz, vz));
    }
    float fx = div_by_const(dx, b.Lx, ylx);
    float fy = div_by_const(dy, b.Ly, yly);
    float fz = b.is2d ? 0.0f : div_by_const(dz, b.Lz, ylz);
    fx = modulus_positive_one_small(fx);
    fy = modulus_positive_one_small(fy);
    fz = modulus_positive_one_small(fz);
    float x = __fadd_rn(b.lox, __fmul_rn(fx, b.Lx));
    float y = __fadd_rn(b.loy, __fmul_rn(fy, b.Ly));
    float z = __fadd_rn(b.loz, __fmul_rn(fz, b.Lz));
    if (TRI)
    {
        x = __fadd_rn(x, __fadd_rn(__fmul_rn(b.xy, y), __fmul_rn(b.xz, z)));
        y = __fadd_rn(y, __fmul_rn(b.yz, z));
    }
    if (b.is2d)
    {
        z = 0.0f;
    }
    rx = x;
    ry = y;
    rz = z;
}

// Per-warp shared memory.  The stage-2 queue is a stack (push on top, pop the top 32): the order in which
// pairs reach stage 2 is irrelevant, rows are regrouped when a batch is flushed.
struct WarpMemBase
{
    uint32_t r_excl[32], r_delta[32], r_code[32]; // non-empty candidate runs of the current home cell
    float4 query[32];                             // current query batch: x, y, z, bits(index to exclude)
    uint32_t qa[kQueueCap], qb[kQueueCap];        // stage-2 stack (WRAP: candidate slot, query slot; IMAGE+RDF: r_sq)
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

struct Cand
{
    float x, y, z;    // WRAP: p + lattice shift (approximate); IMAGE: p (z forced to 0 in 2-D)
    float ix, iy, iz; // IMAGE: exact image vector to add to the query
    uint32_t j;       // point index
    uint32_t slot;    // position in the cell-ordered array
};

template<int FLAVOUR, int MODE, bool TRI> __global__ void __launch_bounds__(kThreads) k_search2(Search2Args a)
{
    using WarpMem = typename WarpMemOf<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1U << lane) - 1U, le_mask = (2U << lane) - 1U;
    size_t const hist_bytes = MODE == S2_RDF ? ((a.axis.bins * sizeof(uint32_t) + 15) / 16) * 16 : 0;
    uint32_t* const sh_hist = reinterpret_cast<uint32_t*>(smem_raw);
    unsigned char* const wbase = smem_raw + hist_bytes + (size_t) warp * warp_mem_bytes(MODE, a.out_cap);
    WarpMem& wm = *reinterpret_cast<WarpMem*>(wbase);
    float4* __restrict__ const sq = wm.query;
    uint32_t* __restrict__ const qa = wm.qa;
    uint32_t* __restrict__ const qb = wm.qb;
    // NL only (pointers are never dereferenced otherwise)
    uint32_t* __restrict__ const o_k = reinterpret_cast<uint32_t*>(wbase + sizeof(WarpMemNL));
    uint32_t* __restrict__ const o_j = o_k + a.out_cap;
    float* __restrict__ const o_x = reinterpret_cast<float*>(o_j + a.out_cap);
    float* __restrict__ const o_y = o_x + a.out_cap;
    float* __restrict__ const o_z = o_y + a.out_cap;
    const BoxDev& box = a.box;

    // points or queries outside the box: image offsets are not implied by the cell walk -> general kernel
    if (*a.flag_points_outside != 0 || *a.flag_queries_outside != 0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            *a.fail = 1;
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
    int const dx = a.dx, dy = a.dy, dz = a.dz;
    uint32_t q_len = 0;     // stage-2 stack height
    uint32_t o_len = 0;     // NL: buffered hits of the current batch
    uint32_t batch_base = 0, batch_n = 0; // NL: query slots [batch_base, batch_base + batch_n) form the batch

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
    auto bin_hit = [&](bool hit, float r_sq) {
        if (hit)
        {
            int const bin = axis_bin(a.axis, __fsqrt_rn(r_sq)); // NeighborBond distance = sqrt(dot(v, v))
            if (bin >= 0)
            {
                atomicAdd(&sh_hist[bin], 1U);
            }
        }
    };

    // one dense round of stage 2: the top n <= 32 entries of the stack
    auto stage2_round = [&](uint32_t n) {
        __syncwarp();
        bool const act = (uint32_t) lane < n;
        uint32_t const e = q_len - n + lane;
        q_len -= n;
        if (FLAVOUR == FGPU_FLAVOUR_WRAP)
        {
            bool hit = false;
            float rx = 0, ry = 0, rz = 0, r_sq = 0;
            uint32_t j = 0, qs = 0;
            if (act)
            {
                uint32_t const cs = qa[e];
                qs = qb[e];
                float4 const p = __ldg(a.sorted + cs);
                float4 const q = __ldg(a.q_sorted + qs);
                j = __float_as_uint(p.w);
                wrap_fast<TRI>(box, a.rcp_lx, a.rcp_ly, a.rcp_lz, __fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y),
                               __fsub_rn(p.z, q.z), rx, ry, rz);
                r_sq = dot_exact(rx, ry, rz);
                hit = in_window2(r_sq, r_max_sq, r_min_sq);
            }
            if (MODE == S2_NL)
            {
                buffer_hits(hit, qs - batch_base, j, rx, ry, rz);
            }
            else
            {
                bin_hit(hit, r_sq);
            }
        }
        else
        {
            // IMAGE + RDF: the stack holds r_sq of accepted bonds
            bin_hit(act, act ? __uint_as_float(qa[e]) : 0.0f);
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
            a.tmp_start[qi] = (uint32_t) base + (incl - cnt);
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
                a.tq[dst] = w.qid[k];
                a.tj[dst] = o_j[i];
                a.tv[3 * (size_t) dst] = o_x[i];
                a.tv[3 * (size_t) dst + 1] = o_y[i];
                a.tv[3 * (size_t) dst + 2] = o_z[i];
            }
        }
        o_len = 0;
        __syncwarp();
    };

    // ---- the pair loop: every query of the batch against the (up to) 64 candidates held in registers --------
    // WRAPPED: some candidate run of this home cell crosses a periodic boundary (uniform per cell)
    auto pair_loop = [&](auto wrapped_tag, const Cand& c0, const Cand& c1, bool two, uint32_t nqc) {
        constexpr bool WRAPPED = decltype(wrapped_tag)::value;
        for (uint32_t k = 0; k < nqc; ++k)
        {
            float4 const q = sq[k];
            uint32_t const q_excl = __float_as_uint(q.w);
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const Cand& c = h == 0 ? c0 : c1;
                if (h == 1 && !two)
                {
                    break;
                }
                if (FLAVOUR == FGPU_FLAVOUR_WRAP)
                {
                    // stage 1: conservative filter on the pre-shifted candidate (fused arithmetic is fine here)
                    float const ddx = c.x - q.x, ddy = c.y - q.y, ddz = c.z - q.z;
                    float const r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                    bool const ok = r2 <= r_hi_sq && c.j != q_excl; // LinkCell.cc:517-520
                    unsigned const m = __ballot_sync(FULL, ok);
                    if (ok)
                    {
                        uint32_t const e = q_len + __popc(m & lt_mask);
                        qa[e] = c.slot;
                        qb[e] = batch_base + k;
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
                    if (MODE == S2_NL)
                    {
                        buffer_hits(hit, k, c.j, rx, ry, rz);
                    }
                    else
                    {
                        unsigned const m = __ballot_sync(FULL, hit);
                        if (hit)
                        {
                            qa[q_len + __popc(m & lt_mask)] = __float_as_uint(r_sq);
                        }
                        q_len += __popc(m);
                    }
                }
            }
            if (FLAVOUR == FGPU_FLAVOUR_WRAP || MODE == S2_RDF)
            {
                if (q_len >= 32)
                {
                    stage2_round(32);
                    if (q_len >= 32)
                    {
                        stage2_round(32);
                    }
                }
            }
        }
    };

    // ---- work loop ------------------------------------------------------------------------------------
    for (;;)
    {
        uint32_t ticket = 0;
        if (lane == 0)
        {
            ticket = atomicAdd(a.work_counter, 1U);
        }
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= a.n_tickets)
        {
            break;
        }
        uint32_t const cell0 = ticket * kCellChunk;
        uint32_t const cell1 = min(cell0 + (uint32_t) kCellChunk, a.n_cells);
        int cz = (int) (cell0 / ((uint32_t) dx * dy));
        uint32_t const rem = cell0 - (uint32_t) cz * dx * dy;
        int cy = (int) (rem / (uint32_t) dx);
        int cx = (int) (rem - (uint32_t) cy * dx);

        for (uint32_t cell = cell0; cell < cell1; ++cell)
        {
            uint32_t const qs0 = __ldg(a.q_cell_start + cell), qs1 = __ldg(a.q_cell_start + cell + 1);
            if (qs0 != qs1)
            {
                // ---- candidate runs: lane k < 18 describes run (row = k / 2, segment = k % 2) -------------
                uint32_t len = 0, start = 0, code = 21U;
                {
                    int const row = lane >> 1, seg = lane & 1;
                    int const oz = row / 3 - 1, oy = row - 3 * (row / 3) - 1;
                    bool valid = lane < 18 && !(dz == 1 && oz != 0);
                    int y = cy + oy, z = cz + oz, wy = 0, wz = 0, wx = 0, x0, x1;
                    if (y < 0)
                    {
                        y += dy;
                        wy = -1;
                    }
                    else if (y >= dy)
                    {
                        y -= dy;
                        wy = 1;
                    }
                    if (z < 0)
                    {
                        z += dz;
                        wz = -1;
                    }
                    else if (z >= dz)
                    {
                        z -= dz;
                        wz = 1;
                    }
                    if (seg == 0)
                    {
                        x0 = max(cx - 1, 0);
                        x1 = min(cx + 1, dx - 1);
                    }
                    else
                    {
                        x0 = x1 = cx == 0 ? dx - 1 : 0;
                        wx = cx == 0 ? -1 : 1;
                        valid = valid && (cx == 0 || cx == dx - 1);
                    }
                    if (valid)
                    {
                        uint32_t const rowbase = ((uint32_t) z * dy + y) * dx;
                        start = __ldg(a.cell_start + rowbase + x0);
                        len = __ldg(a.cell_start + rowbase + x1 + 1) - start;
                        code = (uint32_t) (wx + 1) | ((uint32_t) (wy + 1) << 2) | ((uint32_t) (wz + 1) << 4);
                    }
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    uint32_t const t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o)
                    {
                        incl += t;
                    }
                }
                uint32_t const T = __shfl_sync(FULL, incl, 31);
                unsigned const m_ne = __ballot_sync(FULL, len != 0);
                int const R = __popc(m_ne);
                __syncwarp();
                if (len != 0)
                {
                    int const ck = __popc(m_ne & lt_mask);
                    wm.r_excl[ck] = incl - len;
                    wm.r_delta[ck] = start - (incl - len);
                    wm.r_code[ck] = code;
                }
                __syncwarp();
                uint32_t const my_excl = lane < R ? wm.r_excl[lane] : 0xffffffffU;
                uint32_t const my_delta = lane < R ? wm.r_delta[lane] : 0U;
                uint32_t const my_code = lane < R ? wm.r_code[lane] : 21U;
                bool const any_wrap = __any_sync(FULL, my_code != 21U);

                // loads one round of candidates: flattened index f = B + lane
                auto load_round = [&](uint32_t B, Cand& c) {
                    uint32_t const f = B + lane;
                    bool const in = f < T;
                    uint32_t const rel = my_excl - B; // wraps for runs that start before B
                    unsigned const M = __reduce_or_sync(FULL, rel < 32U ? 1U << rel : 0U);
                    int const before = __popc(__ballot_sync(FULL, my_excl < B));
                    int const r = in ? before + __popc(M & le_mask) - 1 : 0;
                    uint32_t const delta = __shfl_sync(FULL, my_delta, r);
                    c.slot = f + delta;
                    c.ix = c.iy = c.iz = 0.0f;
                    if (in)
                    {
                        float4 const p = __ldg(a.sorted + c.slot);
                        c.x = p.x;
                        c.y = p.y;
                        c.z = p.z;
                        c.j = __float_as_uint(p.w);
                    }
                    else
                    {
                        c.x = c.y = c.z = __int_as_float(0x7f800000); // +inf: fails every window test
                        c.j = 0xffffffffU;
                    }
                    if (FLAVOUR == FGPU_FLAVOUR_IMAGE && box.is2d)
                    {
                        c.z = in ? 0.0f : c.z; // AABBQuery.cc:118-122
                    }
                    if (any_wrap)
                    {
                        uint32_t const cd = __shfl_sync(FULL, my_code, r);
                        int const wx = (int) (cd & 3U) - 1, wy = (int) ((cd >> 2) & 3U) - 1,
                                  wz = (int) ((cd >> 4) & 3U) - 1;
                        if (FLAVOUR == FGPU_FLAVOUR_WRAP)
                        {
                            // nearest image of the candidate (approximate: stage 1 only)
                            float const fx = (float) wx, fy = (float) wy, fz = (float) wz;
                            c.x += fx * box.ax + fy * box.bx + fz * box.cx;
                            c.y += fy * box.by + fz * box.cy;
                            c.z += fz * box.cz;
                        }
                        else
                        {
                            // the candidate's cell was reached by crossing w boundaries: the query image that
                            // sees it is k = -w (all points inside the box), NeighborQuery.h:546-562
                            image_vector(box, -wx, -wy, -wz, c.ix, c.iy, c.iz);
                        }
                    }
                };

                // ---- batches of queries --------------------------------------------------------------------
                // NL: the hits of a batch are buffered until its rows are complete.  The batch size is optimistic
                // (a third of the candidates may hit; an ideal gas gives 15.5 %); a batch that overflows the
                // buffer is discarded and redone at half the size, down to one query (whose hits fit if T does).
                uint32_t q_per_batch = 32;
                if (MODE == S2_NL)
                {
                    if (T > a.out_cap)
                    {
                        if (lane == 0)
                        {
                            *a.fail = 2; // a single row may exceed the buffer: the general kernel takes over
                        }
                        q_per_batch = 0;
                    }
                    else
                    {
                        q_per_batch = min(32U, max(1U, 3U * a.out_cap / max(T, 1U)));
                    }
                }
                for (uint32_t qb0 = qs0; q_per_batch != 0 && qb0 < qs1;)
                {
                    uint32_t const nqc = min(q_per_batch, qs1 - qb0);
                    __syncwarp();
                    if ((uint32_t) lane < nqc)
                    {
                        float4 q = __ldg(a.q_sorted + qb0 + lane);
                        uint32_t const qi = __float_as_uint(q.w);
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
                    batch_base = qb0;
                    batch_n = nqc;
                    for (uint32_t B = 0; B < T; B += 64)
                    {
                        Cand c0, c1;
                        load_round(B, c0);
                        bool const two = B + 32 < T;
                        if (two)
                        {
                            load_round(B + 32, c1);
                        }
                        else
                        {
                            c1 = c0;
                        }
                        if (any_wrap)
                        {
                            pair_loop(std::true_type {}, c0, c1, two, nqc);
                        }
                        else
                        {
                            pair_loop(std::false_type {}, c0, c1, two, nqc);
                        }
                    }
                    if (MODE == S2_NL)
                    {
                        if (FLAVOUR == FGPU_FLAVOUR_WRAP && q_len != 0)
                        {
                            stage2_round(q_len); // rows of the batch must be complete before they are published
                        }
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
            if (++cx == dx)
            {
                cx = 0;
                if (++cy == dy)
                {
                    cy = 0;
                    ++cz;
                }
            }
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
    uint32_t const t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long evals = 0;
    if (t < n_query)
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
            if (j < n_points)
            {
                uint32_t const c = cell_of_point[j];
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

// ---- emit: one thread per bagged hit ------------------------------------------------------------------
// Ranks the hit inside its row (NeighborBond::less_as_tuple / less_as_distance restricted to one row with
// weight == 1, freud/locality/NeighborBond.h:80-112) and writes the five NeighborList arrays
// (NeighborQuery.h:470-478).  The bag rows are contiguous, so the rank loop reads L1-resident words.
template<bool BY_DISTANCE> __global__ void __launch_bounds__(256) k_emit2(Emit2Args a)
{
    uint64_t const t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_bonds)
    {
        return;
    }
    uint32_t const qi = a.tq[t], j = a.tj[t];
    uint32_t const beg = a.tmp_start[qi], n = a.counts[qi];
    float const rx = a.tv[3 * t], ry = a.tv[3 * t + 1], rz = a.tv[3 * t + 2];
    float const d = __fsqrt_rn(dot_exact(rx, ry, rz)); // NeighborBond.h:41-44
    uint32_t rank = 0;
    if (!BY_DISTANCE)
    {
        for (uint32_t k = beg; k < beg + n; ++k)
        {
            rank += a.tj[k] < j ? 1U : 0U;
        }
    }
    else
    {
        for (uint32_t k = beg; k < beg + n; ++k)
        {
            float const ox = a.tv[3 * (size_t) k], oy = a.tv[3 * (size_t) k + 1], oz = a.tv[3 * (size_t) k + 2];
            float const od = __fsqrt_rn(dot_exact(ox, oy, oz));
            uint32_t const oj = a.tj[k];
            rank += (od < d || (od == d && oj < j)) ? 1U : 0U;
        }
    }
    uint64_t const out = (uint64_t) a.row_start[qi] + rank;
    reinterpret_cast<uint2*>(a.neighbors)[out] = make_uint2(qi, j);
    a.distances[out] = d;
    a.weights[out] = 1.0f;
    a.vectors[3 * out] = rx;
    a.vectors[3 * out + 1] = ry;
    a.vectors[3 * out + 2] = rz;
}

template<int FLAVOUR, int MODE, bool TRI> void launch_one(fgpu_ctx* ctx, const Search2Args& a, const char* name)
{
    size_t const hist_bytes = MODE == S2_RDF ? ((a.axis.bins * sizeof(uint32_t) + 15) / 16) * 16 : 0;
    size_t const smem = hist_bytes + (size_t) kWarps * warp_mem_bytes(MODE, a.out_cap);
    auto kern = k_search2<FLAVOUR, MODE, TRI>;
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
    unsigned const blocks
        = (unsigned) std::min<uint64_t>((uint64_t) ctx->sm_count * per_sm, ((uint64_t) a.n_tickets + kWarps - 1) / kWarps);
    KernelScope ks(ctx, name);
    kern<<<std::max(blocks, 1U), kThreads, smem, ctx->stream>>>(a);
}

} // namespace

uint32_t search2_tickets(uint32_t n_cells)
{
    return (n_cells + kCellChunk - 1) / kCellChunk;
}

bool search2_supported(const Search2Args& a, int mode)
{
    bool const grid_ok = a.dx >= 3 && a.dy >= 3 && (a.dz >= 3 || (a.box.is2d && a.dz == 1));
    bool const hist_ok = mode != S2_RDF || a.axis.bins * sizeof(uint32_t) <= 64 * 1024;
    return grid_ok && hist_ok;
}

uint32_t search2_out_cap(double expected_candidates_per_query)
{
    // room for 1.5x the expected candidate count of one query (so that a single row always fits, however
    // dense), 20 B per record, four warps per block
    uint32_t cap = 256;
    while (cap < 2048 && (double) cap < 1.5 * expected_candidates_per_query)
    {
        cap *= 2;
    }
    return cap;
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
    if (flavour == FGPU_FLAVOUR_WRAP)
    {
        if (mode == S2_NL)
            FGPU_S2(FGPU_FLAVOUR_WRAP, S2_NL);
        else
            FGPU_S2(FGPU_FLAVOUR_WRAP, S2_RDF);
    }
    else
    {
        if (mode == S2_NL)
            FGPU_S2(FGPU_FLAVOUR_IMAGE, S2_NL);
        else
            FGPU_S2(FGPU_FLAVOUR_IMAGE, S2_RDF);
    }
#undef FGPU_S2
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_emit2(fgpu_ctx* ctx, int sort_by_distance, const Emit2Args& a)
{
    if (a.n_bonds == 0)
    {
        return;
    }
    unsigned const blocks = (unsigned) ((a.n_bonds + 255) / 256);
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
