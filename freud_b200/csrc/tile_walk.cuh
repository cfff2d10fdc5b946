// Candidate enumeration of one home tile, shared by the warp-cooperative kernels (search2.cu, knn2.cu).
//
// A home tile is a span of consecutive cells [cx0, cx1] of one (cy, cz) row of the grid.  Its queries see the
// cells [cx0 - 1, cx1 + 1] of the 9 (3 in 2-D) neighbouring rows: per row one contiguous run of the cell-ordered
// float4 array (x is the fastest cell index) plus, when the span touches a periodic x boundary, the wrapped
// cell on the other side as a run of its own.  Lane k < 27 describes run (row = k / 3, segment = k % 3); the
// non-empty runs are compacted with a ballot and flattened with a prefix sum so that the 32 lanes of a warp
// load 32 consecutive CANDIDATES at a time (coalesced 16-byte loads).
//
// A run carries the periodic boundary crossings (wx, wy, wz) needed to reach it.  The reference engines never
// see cells: LinkCell wraps every displacement (freud/locality/LinkCell.cc:522), AABBQuery tests all 27 images
// (freud/locality/AABBQuery.cc:88-93); the crossing code only tells which single image can be in range
// (SURVEY.md E1-E4), so candidates that are not in the 27 cells of a particular query of the span simply
// fail the distance test.
#pragma once
#include "internal.h"

namespace fgpu {
namespace tile {

constexpr unsigned FULL = 0xffffffffU;
constexpr uint32_t kNoWrap = 21U; // code of (wx, wy, wz) = (0, 0, 0): (0+1) | (0+1) << 2 | (0+1) << 4
constexpr uint32_t kCodeMask = 63U;
// Symmetric self queries (IMAGE flavour, fused RDF): a run that crosses no boundary is only walked from the
// tile that sees it in a "forward" row (oz > 0, or oz == 0 and oy > 0) and its hits count twice -- without an
// image vector r_ij = p_j - p_i and r_ji = p_i - p_j are exact negatives, so both bonds have the same r_sq
// bit for bit.  Runs that do cross a boundary, and the tile's own row, are walked from both sides as before.
constexpr uint32_t kTwice = 0x80000000U;

struct RunScratch
{
    uint32_t excl[32], delta[32], code[32]; // code: boundary crossings | kTwice
};

struct Runs
{
    uint32_t excl;  // lane r < R: flattened index of the first candidate of run r (0xffffffff otherwise)
    uint32_t delta; // lane r < R: slot of a candidate = flattened index + delta
    uint32_t code;  // lane r < R: boundary crossings of run r (| kTwice: its hits count twice)
    uint32_t T;     // candidates of the tile
    int R;          // non-empty runs
    bool any_wrap;  // some run crosses a periodic boundary
    bool any_twice; // some run counts twice
};

struct Cand
{
    float x, y, z;    // SHIFTED: p + lattice shift (approximate, filter only); else p (z forced to 0 if ZERO_Z)
    float ix, iy, iz; // !SHIFTED: exact image vector to add to the query
    uint32_t j;       // point index
    uint32_t slot;    // position in the cell-ordered array
    uint32_t code;    // boundary crossings of the candidate's run (kNoWrap: none)
};

template<bool SYMMETRIC = false>
__device__ __forceinline__ Runs setup_runs(int dx, int dy, int dz, const uint32_t* __restrict__ cell_start, int cx0,
                                           int cx1, int cy, int cz, int lane, RunScratch& s)
{
    uint32_t const lt_mask = (1U << lane) - 1U;
    uint32_t len = 0, start = 0, code = kNoWrap;
    {
        int const row = lane / 3, seg = lane - 3 * row;
        int const oz = row / 3 - 1, oy = row - 3 * (row / 3) - 1;
        bool valid = lane < 27 && !(dz == 1 && oz != 0);
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
            x0 = max(cx0 - 1, 0);
            x1 = min(cx1 + 1, dx - 1);
        }
        else if (seg == 1)
        {
            x0 = x1 = dx - 1; // left neighbour of cell 0
            wx = -1;
            valid = valid && cx0 == 0;
        }
        else
        {
            x0 = x1 = 0; // right neighbour of cell dx - 1
            wx = 1;
            valid = valid && cx1 == dx - 1;
        }
        if (valid)
        {
            code = (uint32_t) (wx + 1) | ((uint32_t) (wy + 1) << 2) | ((uint32_t) (wz + 1) << 4);
            if (SYMMETRIC && code == kNoWrap && (oz != 0 || oy != 0))
            {
                bool const forward = oz > 0 || (oz == 0 && oy > 0);
                valid = forward; // the tile on the other side walks this pair of rows
                code |= kTwice;
            }
        }
        if (valid)
        {
            uint32_t const rowbase = ((uint32_t) z * dy + y) * dx;
            start = __ldg(cell_start + rowbase + x0);
            len = __ldg(cell_start + rowbase + x1 + 1) - start;
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
    Runs r;
    r.T = __shfl_sync(FULL, incl, 31);
    unsigned const m_ne = __ballot_sync(FULL, len != 0);
    r.R = __popc(m_ne);
    __syncwarp();
    if (len != 0)
    {
        int const ck = __popc(m_ne & lt_mask);
        s.excl[ck] = incl - len;
        s.delta[ck] = start - (incl - len);
        s.code[ck] = code;
    }
    __syncwarp();
    r.excl = lane < r.R ? s.excl[lane] : 0xffffffffU;
    r.delta = lane < r.R ? s.delta[lane] : 0U;
    r.code = lane < r.R ? s.code[lane] : kNoWrap;
    r.any_wrap = __any_sync(FULL, (r.code & kCodeMask) != kNoWrap);
    r.any_twice = SYMMETRIC && __any_sync(FULL, (r.code & kTwice) != 0);
    return r;
}

// The candidate's cell was reached by crossing w boundaries: the query image that sees it is k = -w (all
// points inside the box), NeighborQuery.h:546-562.
__device__ __forceinline__ void code_image(const BoxDev& box, uint32_t cd, float& ix, float& iy, float& iz)
{
    int const wx = (int) (cd & 3U) - 1, wy = (int) ((cd >> 2) & 3U) - 1, wz = (int) ((cd >> 4) & 3U) - 1;
    image_vector(box, -wx, -wy, -wz, ix, iy, iz);
}

// One round of candidates: lane l gets flattened candidate B + l (a candidate that fails every window test if
// B + l >= T).  SHIFTED: the coordinates are moved to the image nearest to the home tile with fused arithmetic
// -- good for a conservative filter only; otherwise they stay exact and (ix, iy, iz) is the image vector the
// reference adds to the query.  ZERO_Z: AABBQuery.cc:118-122 (2-D boxes, IMAGE flavour).
template<bool SHIFTED, bool ZERO_Z>
__device__ __forceinline__ void load_round(const Runs& r, const BoxDev& box, const float4* __restrict__ sorted,
                                           uint32_t B, int lane, Cand& c)
{
    uint32_t const le_mask = (2U << lane) - 1U;
    uint32_t const f = B + lane;
    bool const in = f < r.T;
    uint32_t const rel = r.excl - B; // wraps for runs that start before B
    unsigned const M = __reduce_or_sync(FULL, rel < 32U ? 1U << rel : 0U);
    int const before = __popc(__ballot_sync(FULL, r.excl < B));
    int const run = in ? before + __popc(M & le_mask) - 1 : 0;
    uint32_t const delta = __shfl_sync(FULL, r.delta, run);
    c.slot = f + delta;
    c.ix = c.iy = c.iz = 0.0f;
    c.code = kNoWrap;
    if (in)
    {
        float4 const p = __ldg(sorted + c.slot);
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
    if (ZERO_Z && box.is2d)
    {
        c.z = in ? 0.0f : c.z;
    }
    if (r.any_wrap || r.any_twice)
    {
        uint32_t const cd = __shfl_sync(FULL, r.code, run);
        c.code = cd;
    }
    if (r.any_wrap)
    {
        uint32_t const cd = c.code & kCodeMask;
        int const wx = (int) (cd & 3U) - 1, wy = (int) ((cd >> 2) & 3U) - 1, wz = (int) ((cd >> 4) & 3U) - 1;
        if (SHIFTED)
        {
            float const fx = (float) wx, fy = (float) wy, fz = (float) wz;
            c.x += fx * box.ax + fy * box.bx + fz * box.cx;
            c.y += fy * box.by + fz * box.cy;
            c.z += fz * box.cz;
        }
        else
        {
            code_image(box, cd, c.ix, c.iy, c.iz);
        }
    }
}

// Home tile of a work ticket: row = ticket / spans_per_row, span index = ticket % spans_per_row.
__device__ __forceinline__ void ticket_tile(uint32_t ticket, uint32_t spans_per_row, int span, int dx, int dy,
                                            int& cx0, int& cx1, int& cy, int& cz)
{
    uint32_t const row = ticket / spans_per_row;
    uint32_t const s = ticket - row * spans_per_row;
    cz = (int) (row / (uint32_t) dy);
    cy = (int) (row - (uint32_t) cz * dy);
    cx0 = (int) s * span;
    cx1 = min(cx0 + span, dx) - 1;
}

} // namespace tile
} // namespace fgpu
