// Periodic cell list: cell index + count, exclusive scan, scatter to cell-ordered float4.
//
// Replaces (with a different, GPU-shaped data structure) LinkCell::computeCellList
// freud/locality/LinkCell.cc:316-336 / getCellCoord :356-367 and AABBQuery::buildTree
// freud/locality/AABBQuery.cc:53-69.  The reference's structures only generate candidates; results are
// decided by the per-pair arithmetic, so the grid here is free to use its own cell
// width: the smallest width that is provably conservative for the query radius.
//
// Layout in HBM: cell_start u32[n_cells+1]; sorted float4[n] = {x, y, z, bits(point index)} grouped by
// cell (x fastest, then y, then z), so the 3 x-neighbour cells of a (y, z) row are one contiguous run.
// Algorithmic bytes: 12 (read xyz) + 16 (write float4) = 28 B/point, + 4 B/cell counters.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bulk_copy.cuh"
#include "internal.h"

namespace fgpu {

namespace {

constexpr int kScanThreads = 1024; // 8192 elements per tile: the look-back of the last tile is a few rounds, not fifteen
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Single-pass exclusive scan (chained scan with decoupled look-back): a block takes the next tile from an
// atomic ticket (so every predecessor tile is already running), scans it in registers, publishes its
// aggregate, and the first warp walks back over the predecessors' {flag, value} words -- 32 at a time --
// until it meets an inclusive prefix.  One launch and 8 B/element of traffic instead of three launches per
// level; u32 sums wrap like the counters they come from.
enum : unsigned long long
{
    kFlagAggregate = 1ULL << 32,
    kFlagPrefix = 2ULL << 32
};

// The scanned array may be two physical pieces, `head` (split elements, a multiple of 8) followed by `data`: a slab of
// cell layers that wraps around the periodic boundary is one logical run.
__global__ void __launch_bounds__(kScanThreads) k_scan_chained(uint32_t* __restrict__ data_in, size_t n,
                                                               volatile unsigned long long* state,
                                                               unsigned int* ticket, uint32_t* __restrict__ head,
                                                               size_t split)
{
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
    {
        s_tile = atomicAdd(ticket, 1U);
    }
    __syncthreads();
    uint32_t const tile = s_tile;
    // logical index of this thread's first item -> piece and local index (split is a multiple of kScanItems: a
    // thread's items never straddle the pieces); from here on `n` is the length of that piece
    size_t const logical = (size_t) tile * kScanTile + (size_t) threadIdx.x * kScanItems;
    bool const in_head = logical < split;
    uint32_t* const data = in_head ? head : data_in;
    size_t const base = in_head ? logical : logical - split;
    n = in_head ? split : n - split;
    uint32_t v[kScanItems];
    if (base + kScanItems <= n)
    {
        uint4 const a = *reinterpret_cast<const uint4*>(data + base);
        uint4 const b = *reinterpret_cast<const uint4*>(data + base + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    }
    else
    {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
        {
            v[k] = base + k < n ? data[base + k] : 0U;
        }
    }
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const t = __shfl_up_sync(0xffffffffU, incl, o);
        if (lane >= o)
        {
            incl += t;
        }
    }
    if (lane == 31)
    {
        warp_sums[warp] = incl;
    }
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0U;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1)
        {
            uint32_t const t = __shfl_up_sync(0xffffffffU, w, o);
            if (lane >= o)
            {
                w += t;
            }
        }
        uint32_t const total = __shfl_sync(0xffffffffU, w, kScanThreads / 32 - 1);
        if (lane < kScanThreads / 32)
        {
            warp_sums[lane] = w; // inclusive over warps
        }
        // publish the aggregate, then look back for the exclusive prefix of this tile
        uint32_t prefix = 0;
        if (tile == 0)
        {
            if (lane == 0)
            {
                state[0] = kFlagPrefix | total;
            }
        }
        else
        {
            if (lane == 0)
            {
                state[tile] = kFlagAggregate | total;
            }
            for (long long first = (long long) tile - 1;; first -= 32)
            {
                long long const idx = first - lane;
                unsigned long long st;
                do
                {
                    st = idx >= 0 ? state[idx] : kFlagPrefix; // tiles before the first: prefix 0
                } while (__any_sync(0xffffffffU, (st >> 32) == 0));
                unsigned const m = __ballot_sync(0xffffffffU, (st >> 32) == 2);
                int const stop = m != 0 ? __ffs(m) - 1 : 31; // nearest predecessor holding an inclusive prefix
                uint32_t part = lane <= stop ? (uint32_t) st : 0U;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    part += __shfl_xor_sync(0xffffffffU, part, o);
                }
                prefix += part;
                if (m != 0)
                {
                    break;
                }
            }
            if (lane == 0)
            {
                state[tile] = kFlagPrefix | (uint32_t) (prefix + total);
            }
        }
        if (lane == 0)
        {
            s_prefix = prefix;
        }
    }
    __syncthreads();
    uint32_t excl = s_prefix + incl - sum + (warp > 0 ? warp_sums[warp - 1] : 0U);
    uint32_t o[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        o[k] = excl;
        excl += v[k];
    }
    if (base + kScanItems <= n)
    {
        *reinterpret_cast<uint4*>(data + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(data + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    else
    {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
        {
            if (base + k < n)
            {
                data[base + k] = o[k];
            }
        }
    }
}

// A slab restricts the list to the cell layers [lo, lo + len) (cyclic) of one axis: a rank that only searches
// its share of the home tiles needs nothing else (the build must not stay serial).
struct SlabDev
{
    int axis; // 1: y layers (2-D grids), 2: z layers
    int lo;
    int len;  // < 0: no restriction
};

// Cell of a point, the value cell_coords gives, for a fraction of its cost: the fractional coordinates through
// reciprocals (three multiplies instead of three IEEE divisions) settle the cell whenever frac * dim is farther than
// 1e-3 from an integer and the point is farther than 1e-4 (in box units) from a face; the reciprocal arithmetic is
// off by a few 1e-7 there, so both routes name the same cell.  Everything else -- about 1 % of uniform points, and
// every point outside the box -- takes cell_coords itself.  Returns false if the point lies outside the box.
struct CellRecip
{
    float rx, ry, rz, fdx, fdy, fdz;
};

__device__ __forceinline__ bool cell_of_point(const BoxDev& box, const CellRecip& r, int dx, int dy, int dz, float x, float y,
                                              float z, uint32_t& cell)
{
    float const gz = box.is2d ? 0.0f : (z - box.loz) * r.rz;
    float const gy = ((y - box.loy) - box.yz * z) * r.ry;
    float const gx = ((x - box.lox) - (box.t_xz * z + box.xy * y)) * r.rx;
    float const sx = gx * r.fdx, sy = gy * r.fdy, sz = gz * r.fdz;
    float const fx = sx - floorf(sx), fy = sy - floorf(sy), fz = sz - floorf(sz);
    bool const inside = gx > 1e-4f && gx < 0.9999f && gy > 1e-4f && gy < 0.9999f && (box.is2d || (gz > 1e-4f && gz < 0.9999f));
    bool const clear = fx > 1e-3f && fx < 0.999f && fy > 1e-3f && fy < 0.999f && (box.is2d || (fz > 1e-3f && fz < 0.999f));
    if (inside && clear)
    {
        cell = ((uint32_t) (box.is2d ? 0 : (int) sz) * dy + (uint32_t) (int) sy) * dx + (uint32_t) (int) sx;
        return true;
    }
    int cx, cy, cz, nx, ny, nz;
    cell_coords(box, dx, dy, dz, x, y, z, cx, cy, cz, nx, ny, nz);
    cell = ((uint32_t) cz * dy + cy) * dx + cx;
    return (nx | ny | nz) == 0;
}

// K1: cell index + arrival rank of every point, cell populations -- 12 B read, 8 B written per point, one atomic.
// Persistent blocks stream the points through a two-stage cp.async.bulk ring (bulk_copy.cuh); the cell comes from
// cell_of_point (reciprocals, the IEEE route only next to a face).  45 -> 36 us at 4 M points; what is left is the L2's
// atomic rate (4 M atomics on 389 k counters).  Taking the arrival rank in the scatter instead (RED here, a returning
// atomic on a per-cell cursor there) was measured: assign 36 us, scatter 82 -> 92 us -- the scatter is bound by its
// 16-byte random writes and had no room for the atomic, so the rank stays here.
constexpr int kAssignThreads = 256;
constexpr int kAssignChunk = 1024; // points per chunk: 12288 bytes

__global__ void __launch_bounds__(kAssignThreads) k_cell_assign_stream(BoxDev box, int dx, int dy, int dz,
                                                                       const float* __restrict__ xyz, uint32_t n,
                                                                       uint32_t* __restrict__ cell_of,
                                                                       uint32_t* __restrict__ rank_in,
                                                                       uint32_t* __restrict__ cell_count,
                                                                       int* __restrict__ any_shift)
{
    __shared__ __align__(128) float s_raw[2][kAssignChunk * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    uint32_t const n_chunks = (n + kAssignChunk - 1) / kAssignChunk;
    auto chunk_points = [&](uint32_t c) { return min((uint32_t) kAssignChunk, n - c * kAssignChunk); };
    bool const aligned = (reinterpret_cast<uintptr_t>(xyz) & 15U) == 0; // caller-owned arrays may sit anywhere
    auto bulk_ok = [&](uint32_t c) { return aligned && (chunk_points(c) * 12U) % 16U == 0; };
    auto issue = [&](uint32_t c, int stage) {
        if (bulk_ok(c))
        {
            uint32_t const bytes = chunk_points(c) * 12U;
            bulk::mbar_arrive_expect_tx(&s_bar[stage], bytes);
            bulk::copy_g2s(s_raw[stage], xyz + 3 * (size_t) c * kAssignChunk, bytes, &s_bar[stage]);
        }
    };
    if (threadIdx.x == 0)
    {
        bulk::mbar_init(&s_bar[0], 1);
        bulk::mbar_init(&s_bar[1], 1);
        bulk::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < n_chunks)
    {
        issue(blockIdx.x, 0);
    }
    CellRecip const rc {1.0f / box.Lx, 1.0f / box.Ly, box.is2d ? 0.0f : 1.0f / box.Lz, (float) dx, (float) dy, (float) dz};
    uint32_t it = 0, waits[2] = {0, 0};
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it)
    {
        int const stage = (int) (it & 1U);
        uint32_t const here = chunk_points(c);
        if (threadIdx.x == 0)
        {
            uint32_t const next = c + gridDim.x;
            if (next < n_chunks)
            {
                issue(next, stage ^ 1); // that buffer was released by the barrier that ended the previous iteration
            }
        }
        if (bulk_ok(c))
        {
            bulk::mbar_wait(&s_bar[stage], waits[stage] & 1U);
            waits[stage] += 1;
        }
        else
        {
            for (uint32_t e = threadIdx.x; e < 3 * here; e += kAssignThreads)
            {
                s_raw[stage][e] = xyz[3 * (size_t) c * kAssignChunk + e];
            }
            __syncthreads();
        }
        const float* raw = s_raw[stage];
        for (uint32_t k = threadIdx.x; k < here; k += kAssignThreads)
        {
            uint32_t cell;
            if (!cell_of_point(box, rc, dx, dy, dz, raw[3 * k], raw[3 * k + 1], raw[3 * k + 2], cell))
            {
                *any_shift = 1;
            }
            cell_of[(size_t) c * kAssignChunk + k] = cell;
            rank_in[(size_t) c * kAssignChunk + k] = atomicAdd(&cell_count[cell], 1U);
        }
        if (threadIdx.x == 0)
        {
            bulk::fence_proxy_async();
        }
        __syncthreads();
    }
}

// K3: scatter to cell order -- 20 B read, 16 (+4) B written per point
__global__ void __launch_bounds__(256) k_cell_scatter(BoxDev box, int dx, int dy, int dz, const float* __restrict__ xyz,
                                                      uint32_t n, const uint32_t* __restrict__ cell_of,
                                                      const uint32_t* __restrict__ rank_in,
                                                      const uint32_t* __restrict__ cell_start,
                                                      float4* __restrict__ sorted, int* __restrict__ shift,
                                                      const int* __restrict__ any_shift)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
    {
        return;
    }
    uint32_t const cell = cell_of[i];
    float const x = xyz[3 * (size_t) i], y = xyz[3 * (size_t) i + 1], z = xyz[3 * (size_t) i + 2];
    uint32_t const slot = __ldg(cell_start + cell) + rank_in[i];
    sorted[slot] = make_float4(x, y, z, __uint_as_float(i));
    if (shift != nullptr && *any_shift != 0)
    {
        int cx, cy, cz, nx, ny, nz;
        cell_coords(box, dx, dy, dz, x, y, z, cx, cy, cz, nx, ny, nz);
        shift[slot] = pack_shift(nx, ny, nz);
    }
}

// K0: everything the build needs zeroed, in one launch instead of four memsets (launch gaps are a visible share of a
// sharded step): the cell counters (two ranges: a sharded rank only zeroes, scans and reads the cells of its slab),
// the scan's tile states + tickets, the out-of-box flag and the slab counter.
__global__ void __launch_bounds__(256) k_cell_prep(uint32_t* __restrict__ cell_count, uint32_t* __restrict__ cell_fill,
                                                   size_t a0, size_t a1, size_t b0, size_t b1,
                                                   uint32_t* __restrict__ scan_words, uint32_t n_scan,
                                                   int* __restrict__ any_shift, uint32_t* __restrict__ slab_count)
{
    size_t const stride = (size_t) gridDim.x * blockDim.x;
    size_t const t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = a0 + t; i < a1; i += stride)
    {
        cell_count[i] = 0U;
        if (cell_fill != nullptr)
        {
            cell_fill[i] = 0U;
        }
    }
    for (size_t i = b0 + t; i < b1; i += stride)
    {
        cell_count[i] = 0U;
        if (cell_fill != nullptr)
        {
            cell_fill[i] = 0U;
        }
    }
    for (size_t i = t; i < n_scan; i += stride)
    {
        scan_words[i] = 0U;
    }
    if (t == 0)
    {
        *any_shift = 0;
        if (slab_count != nullptr)
        {
            *slab_count = 0U;
        }
    }
}

// K1 for a sharded rank: one streaming pass over ALL points (the input is replicated), of which only the points of
// this rank's slab leave a trace (the build must not stay serial; at 8 ranks the slab holds
// ~1/6 of the points).  Persistent blocks; the points arrive in 12 KB chunks through cp.async.bulk into a two-stage
// shared-memory ring (bulk_copy.cuh): the copy engine keeps a chunk per block in flight while the block works on the
// previous one -- the first two versions of this kernel, load-then-process per block, were latency-bound at 1.5 TB/s
// (32 us for 48 MB; profiles/launches_r2_shard.csv).  Per chunk: a cheap slab test on every point (the z fraction
// -- 2-D: y -- through a reciprocal; anything within 1e-2 of a layer face or 1e-4 of a box face is left undecided),
// survivors to a shared list, then the exact cell arithmetic of cell_coords (three IEEE divisions) on that list only,
// on dense lanes.  Nothing here waits for a global atomic: the cell counters are bumped with RED, the list of
// {position, cell} goes to a fixed window per chunk, and the arrival rank inside the cell is taken by the scatter.
constexpr int kSlabThreads = 256;
constexpr int kSlabChunk = 1024; // points per chunk: 12288 bytes

__global__ void __launch_bounds__(kSlabThreads) k_cell_assign_slab(BoxDev box, int dx, int dy, int dz,
                                                                   const float* __restrict__ xyz, uint32_t n,
                                                                   uint32_t* __restrict__ cell_count,
                                                                   int* __restrict__ any_shift, SlabDev slab,
                                                                   float4* __restrict__ list_pos,
                                                                   uint32_t* __restrict__ list_cell,
                                                                   uint32_t* __restrict__ chunk_count)
{
    __shared__ __align__(128) float s_raw[2][kSlabChunk * 3];
    __shared__ float4 s_pos[kSlabChunk];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_maybe;
    uint32_t const n_chunks = (n + kSlabChunk - 1) / kSlabChunk;
    auto chunk_points = [&](uint32_t c) { return min((uint32_t) kSlabChunk, n - c * kSlabChunk); };
    // only whole 16-byte multiples go through the copy engine (every chunk but possibly the last)
    auto issue = [&](uint32_t c, int stage) {
        uint32_t const bytes = chunk_points(c) * 12U;
        if (bytes % 16U == 0)
        {
            bulk::mbar_arrive_expect_tx(&s_bar[stage], bytes);
            bulk::copy_g2s(s_raw[stage], xyz + 3 * (size_t) c * kSlabChunk, bytes, &s_bar[stage]);
        }
    };
    if (threadIdx.x == 0)
    {
        bulk::mbar_init(&s_bar[0], 1);
        bulk::mbar_init(&s_bar[1], 1);
        bulk::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < n_chunks)
    {
        issue(blockIdx.x, 0);
    }
    int const d = slab.axis == 2 ? dz : dy;
    float const rcp_lx = 1.0f / box.Lx, rcp_ly = 1.0f / box.Ly, rcp_lz = box.is2d ? 0.0f : 1.0f / box.Lz;
    float const fd = (float) d;
    uint32_t it = 0;
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it)
    {
        int const stage = (int) (it & 1U);
        uint32_t const here = chunk_points(c);
        if (threadIdx.x == 0)
        {
            s_maybe = 0;
            uint32_t const next = c + gridDim.x;
            if (next < n_chunks)
            {
                issue(next, stage ^ 1); // that buffer was released by the barrier that ended the previous iteration
            }
        }
        if ((here * 12U) % 16U == 0)
        {
            bulk::mbar_wait(&s_bar[stage], (it >> 1) & 1U);
        }
        else
        {
            for (uint32_t e = threadIdx.x; e < 3 * here; e += kSlabThreads)
            {
                s_raw[stage][e] = xyz[3 * (size_t) c * kSlabChunk + e];
            }
        }
        __syncthreads();
        const float* raw = s_raw[stage];
        for (uint32_t k = threadIdx.x; k < here; k += kSlabThreads)
        {
            float const x = raw[3 * k], y = raw[3 * k + 1], z = raw[3 * k + 2];
            // approximate fractions (the arithmetic of fractional_for_cells with the divisions replaced)
            float const gz = box.is2d ? 0.5f : (z - box.loz) * rcp_lz;
            float const gy = ((y - box.loy) - box.yz * z) * rcp_ly;
            float const gx = ((x - box.lox) - (box.t_xz * z + box.xy * y)) * rcp_lx;
            bool const well_inside = gx > 1e-4f && gx < 0.9999f && gy > 1e-4f && gy < 0.9999f
                && (box.is2d || (gz > 1e-4f && gz < 0.9999f));
            float const gs = (slab.axis == 2 ? gz : gy) * fd; // layer coordinate
            float const frac_layer = gs - floorf(gs);
            bool maybe = true; // near a box face or a layer face: the exact arithmetic decides
            if (well_inside && frac_layer > 1e-2f && frac_layer < 0.99f)
            {
                int rel = (int) gs - slab.lo;
                rel += rel < 0 ? d : 0;
                maybe = rel < slab.len;
            }
            if (maybe)
            {
                s_pos[atomicAdd(&s_maybe, 1U)] = make_float4(x, y, z, __uint_as_float(c * kSlabChunk + k));
            }
        }
        __syncthreads();
        uint32_t const n_maybe = s_maybe;
        size_t const window = (size_t) c * kSlabChunk;
        for (uint32_t k = threadIdx.x; k < n_maybe; k += kSlabThreads)
        {
            float4 const p = s_pos[k];
            int cx, cy, cz, nx, ny, nz;
            cell_coords(box, dx, dy, dz, p.x, p.y, p.z, cx, cy, cz, nx, ny, nz);
            if ((nx | ny | nz) != 0)
            {
                *any_shift = 1;
            }
            int rel = (slab.axis == 2 ? cz : cy) - slab.lo;
            rel += rel < 0 ? d : 0;
            uint32_t cell = 0xffffffffU;
            if (rel < slab.len)
            {
                cell = ((uint32_t) cz * dy + cy) * dx + cx;
                atomicAdd(&cell_count[cell], 1U); // result unused: a RED, nobody waits for it
            }
            list_pos[window + k] = p;
            list_cell[window + k] = cell;
        }
        if (threadIdx.x == 0)
        {
            chunk_count[c] = n_maybe;
            bulk::fence_proxy_async(); // this block's reads of s_raw[stage] come before the engine's next write to it
        }
        __syncthreads();
    }
}

// K3 for a sharded rank: the chunk windows of the list; the arrival rank inside a cell is taken here, where every
// thread is independent and the atomic's latency hides behind the other warps
__global__ void __launch_bounds__(256) k_cell_scatter_slab(const float4* __restrict__ list_pos,
                                                           const uint32_t* __restrict__ list_cell,
                                                           const uint32_t* __restrict__ chunk_count, uint32_t n_chunks,
                                                           const uint32_t* __restrict__ cell_start,
                                                           uint32_t* __restrict__ cell_fill, float4* __restrict__ sorted)
{
    // a block per chunk: its entries are the first chunk_count[c] of the chunk's window, one per thread, so that all
    // of a chunk's atomics are in flight at once (a warp walking its chunk serially was latency-bound: 17 us)
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x)
    {
        uint32_t const cnt = __ldg(chunk_count + c);
        size_t const window = (size_t) c * kSlabChunk;
        for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x)
        {
            uint32_t const cell = list_cell[window + k];
            if (cell != 0xffffffffU)
            {
                uint32_t const slot = __ldg(cell_start + cell) + atomicAdd(&cell_fill[cell], 1U);
                sorted[slot] = list_pos[window + k];
            }
        }
    }
}

// NeighborQuery.h:103-112: a 2-D box takes no point with |z| > 1e-6.  Checked on the device after the upload (a host
// pass over the array costs more than the whole RDF frame it precedes).
__global__ void __launch_bounds__(256) k_check_2d_z(const float* __restrict__ xyz, uint32_t n, int* __restrict__ flag)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && fabsf(xyz[3 * (size_t) i + 2]) > 1e-6f)
    {
        *flag = 1;
    }
}

void choose_dims(const fgpu_points* pts, float r_search, bool force_single_cell, int dim[3], int ambiguous[3])
{
    double const lmax = std::max({(double) pts->box.Lx, (double) pts->box.Ly, (double) pts->box.Lz});
    // cell thickness must exceed r_search by more than the float32 rounding of coordinates of size ~L
    double const w = (double) r_search * (1.0 + 1.0e-4) + 1.0e-5 * lmax;
    for (int d = 0; d < 3; ++d)
    {
        double const q = std::isfinite(w) && w > 0 ? (double) pts->plane_dist[d] / w : 1.0;
        dim[d] = (int) std::max(1.0, std::min(std::floor(q), 1024.0 * 1024.0));
    }
    if (pts->box.is2d)
    {
        dim[2] = 1;
    }
    double const cap = 4.0 * (double) pts->n + 64.0;
    while ((double) dim[0] * dim[1] * dim[2] > cap)
    {
        for (int d = 0; d < 3; ++d)
        {
            if (dim[d] > 1)
            {
                dim[d] = (dim[d] + 1) / 2;
            }
        }
    }
    if (force_single_cell)
    {
        dim[0] = dim[1] = dim[2] = 1;
    }
    for (int d = 0; d < 3; ++d)
    {
        ambiguous[d] = dim[d] < 3 ? 1 : 0;
    }
    if (pts->box.is2d)
    {
        ambiguous[2] = 0; // no z images in 2-D (NeighborQuery.h:519-521)
    }
}

} // namespace

size_t scan_scratch_words(size_t n)
{
    size_t const tiles = (n + kScanTile - 1) / kScanTile;
    return 2 * tiles + 2; // one {flag, value} word per tile + the tile ticket
}

void exclusive_scan_u32(fgpu_ctx* ctx, uint32_t* data, size_t n, bool scratch_is_zero, uint32_t* head, size_t split)
{
    if (n == 0)
    {
        return;
    }
    if ((reinterpret_cast<uintptr_t>(data) & 15U) != 0)
    {
        throw Error(FGPU_ERUNTIME, "exclusive_scan_u32 needs a 16-byte aligned array");
    }
    size_t const tiles = (n + kScanTile - 1) / kScanTile;
    if (!scratch_is_zero)
    {
        ctx->scan_tmp.reserve(scan_scratch_words(n));
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->scan_tmp.ptr, 0, scan_scratch_words(n) * sizeof(uint32_t), ctx->stream));
    }
    if (split % kScanItems != 0 || (split != 0 && (reinterpret_cast<uintptr_t>(head) & 15U) != 0))
    {
        throw Error(FGPU_ERUNTIME, "exclusive_scan_u32: the head piece must be 16-byte aligned and a multiple of 8 long");
    }
    auto* state = reinterpret_cast<unsigned long long*>(ctx->scan_tmp.ptr);
    auto* ticket = reinterpret_cast<unsigned int*>(ctx->scan_tmp.ptr + 2 * tiles);
    {
        KernelScope ks(ctx, "scan");
        k_scan_chained<<<(unsigned) tiles, kScanThreads, 0, ctx->stream>>>(data, n, state, ticket, head, split);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_check_2d_z(fgpu_ctx* ctx, const float* xyz, uint32_t n, int* flag)
{
    if (n == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "check_2d_z");
        k_check_2d_z<<<(n + 255) / 256, 256, 0, ctx->stream>>>(xyz, n, flag);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

GridDev grid_dev(const fgpu_points* pts)
{
    const fgpu_grid& g = pts->grid;
    GridDev d;
    d.dx = g.dim[0];
    d.dy = g.dim[1];
    d.dz = g.dim[2];
    d.amb_x = g.ambiguous[0];
    d.amb_y = g.ambiguous[1];
    d.amb_z = g.ambiguous[2];
    d.any_shift_flag = g.any_shift_flag.ptr;
    d.cell_start = g.cell_start.ptr;
    d.sorted = g.sorted.ptr;
    d.shift = g.shift.ptr;
    return d;
}

ShardPlan shard_plan(const int dim[3], uint32_t n_points, int shard, int n_shards)
{
    // the home tiles of search2_plan, dealt to the shards in contiguous runs of tickets (rows of the grid, z
    // slowest): shard s owns [s T / S, (s + 1) T / S)
    Search2Args a;
    std::memset(&a, 0, sizeof(a));
    a.dx = dim[0];
    a.dy = dim[1];
    a.dz = dim[2];
    a.n_cells = (uint32_t) dim[0] * dim[1] * dim[2];
    search2_plan(a, n_points);
    ShardPlan p;
    // Equal COST, not equal count: the home tiles of the two cell layers at the periodic boundary of the slowest axis
    // see a third of their candidate rows across that boundary, where the symmetric walk cannot halve the pair tests
    // (tile_walk.cuh) -- measured 1.5x the time of an interior tile (one GPU playing ranks 0 and 7 of 8:
    // profiles/ncu_r2_summary.md).  Tickets are ordered layer by layer, so the cumulative cost is piecewise linear.
    {
        int const layers = dim[2] > 1 ? dim[2] : dim[1];
        uint64_t const per_layer = (uint64_t) a.n_tickets / (uint64_t) layers; // tickets of one layer of the slow axis
        double const w_edge = layers >= 3 ? 1.5 : 1.0;
        double const total = (double) a.n_tickets + (w_edge - 1.0) * 2.0 * (double) per_layer;
        auto ticket_at = [&](double cost) {
            double const first = w_edge * (double) per_layer;                       // cost of the first layer
            double const middle = (double) (a.n_tickets - 2 * per_layer);           // ... of the interior layers
            double t;
            if (cost <= first)
            {
                t = cost / w_edge;
            }
            else if (cost <= first + middle)
            {
                t = (double) per_layer + (cost - first);
            }
            else
            {
                t = (double) (a.n_tickets - per_layer) + (cost - first - middle) / w_edge;
            }
            return (uint32_t) std::min<double>(std::max(t, 0.0), (double) a.n_tickets);
        };
        p.ticket_begin = shard == 0 ? 0U : ticket_at(total * (double) shard / (double) n_shards);
        p.ticket_end = shard + 1 == n_shards ? a.n_tickets : ticket_at(total * (double) (shard + 1) / (double) n_shards);
        if (n_shards == 1)
        {
            p.ticket_begin = 0;
            p.ticket_end = a.n_tickets;
        }
    }
    p.slab_axis = dim[2] > 1 ? 2 : 1;
    int const layers = p.slab_axis == 2 ? dim[2] : dim[1];
    p.slab_lo = 0;
    p.slab_len = -1;
    if (n_shards > 1 && p.ticket_end > p.ticket_begin)
    {
        uint32_t const row0 = p.ticket_begin / a.spans_per_row, row1 = (p.ticket_end - 1) / a.spans_per_row;
        int const l0 = p.slab_axis == 2 ? (int) (row0 / (uint32_t) dim[1]) : (int) row0;
        int const l1 = p.slab_axis == 2 ? (int) (row1 / (uint32_t) dim[1]) : (int) row1;
        int const len = l1 - l0 + 3; // one halo layer on each side
        if (len < layers)
        {
            p.slab_lo = (l0 - 1 + layers) % layers;
            p.slab_len = len;
        }
    }
    uint32_t const cells_per_row = (uint32_t) dim[0];
    p.cell_begin = (p.ticket_begin / a.spans_per_row) * cells_per_row
        + std::min((p.ticket_begin % a.spans_per_row) * (uint32_t) a.span, cells_per_row);
    p.cell_end = p.ticket_end == a.n_tickets
        ? a.n_cells
        : (p.ticket_end / a.spans_per_row) * cells_per_row
            + std::min((p.ticket_end % a.spans_per_row) * (uint32_t) a.span, cells_per_row);
    return p;
}

void build_grid(fgpu_points* pts, float r_search, bool force_single_cell)
{
    fgpu_ctx* ctx = pts->ctx;
    fgpu_grid& g = pts->grid;
    int dim[3], amb[3];
    choose_dims(pts, r_search, force_single_cell, dim, amb);
    if (g.r_search >= 0 && dim[0] == g.dim[0] && dim[1] == g.dim[1] && dim[2] == g.dim[2]
        && g.shard == pts->shard && g.n_shards == pts->n_shards)
    {
        g.r_search = std::max(g.r_search, r_search); // same grid already resident
        return;
    }
    SlabDev slab {2, 0, -1};
    if (pts->n_shards > 1)
    {
        ShardPlan const sp = shard_plan(dim, pts->n, pts->shard, pts->n_shards);
        slab = SlabDev {sp.slab_axis, sp.slab_lo, sp.slab_len};
    }
    uint32_t const n = pts->n;
    uint32_t const n_cells = (uint32_t) dim[0] * dim[1] * dim[2];
    bool const slab_only = slab.len >= 0;
    g.cell_start.reserve((size_t) n_cells + 4);
    g.sorted.reserve(n);
    // out-of-box flag: set on the device by K1, read on the device by K3 and the search kernels, so the
    // build needs no host round trip; the shift array is only ever written when the flag is set
    g.shift.reserve(n);
    g.any_shift_flag.reserve(2); // [1]: points in this rank's slab (sharded build)
    int* d_flag = g.any_shift_flag.ptr;
    uint32_t* d_slab_count = reinterpret_cast<uint32_t*>(d_flag + 1);
    // The cells the build zeroes, scans and fills: all of them, or the one or two contiguous index ranges of the slab
    // (layers of the slowest grid axis; a slab that wraps around the periodic boundary is two ranges, the second
    // scan continuing the first).  Every range carries one extra element: cell_start[end] closes its last cell.
    size_t ra0 = 0, ra1 = (size_t) n_cells + 1, rb0 = 0, rb1 = 0;
    bool const ranged = slab_only && slab.axis == (dim[2] > 1 ? 2 : 1);
    if (ranged)
    {
        size_t const per_layer = slab.axis == 2 ? (size_t) dim[0] * dim[1] : (size_t) dim[0];
        int const layers = slab.axis == 2 ? dim[2] : dim[1];
        int const hi = slab.lo + slab.len;
        ra0 = (size_t) slab.lo * per_layer;
        ra1 = (size_t) std::min(hi, layers) * per_layer + 1;
        if (hi > layers)
        {
            rb0 = 0;
            rb1 = (size_t) (hi - layers) * per_layer + 1;
        }
    }
    // 16-byte alignment of the scanned pieces (the scan reads uint4): the high range grows downwards, the low one
    // (scanned first, as the head of one logical run) is padded to a multiple of 8 elements -- the extra cells belong
    // to layers outside the slab, stay empty and are never read
    ra0 &= ~(size_t) 3;
    if (rb1 > rb0)
    {
        rb1 = std::min<size_t>((rb1 + 7) & ~(size_t) 7, ra0);
        if ((rb1 & 7) != 0)
        {
            throw Error(FGPU_ERUNTIME, "sharded cell list: the slab leaves no room between its two cell ranges");
        }
    }
    size_t const scan_n = (ra1 - ra0) + (rb1 - rb0);
    size_t const scan_words = scan_scratch_words(scan_n);
    ctx->scan_tmp.reserve(scan_words);
    {
        KernelScope ks(ctx, "cell_prep");
        unsigned const pblocks = (unsigned) std::min<size_t>((scan_n + 255) / 256, (size_t) ctx->sm_count * 8);
        if (slab_only)
        {
            g.rank_in.reserve((size_t) n_cells + 4); // sharded build: the per-cell fill cursor of the scatter
        }
        k_cell_prep<<<std::max(pblocks, 1U), 256, 0, ctx->stream>>>(g.cell_start.ptr, slab_only ? g.rank_in.ptr : nullptr,
                                                                   ra0, ra1, rb0, rb1, ctx->scan_tmp.ptr,
                                                                   (uint32_t) scan_words, d_flag, d_slab_count);
    }
    unsigned const blocks = (n + 255) / 256;
    if (slab_only)
    {
        // the list has a window of kSlabChunk entries per chunk of points (sized for clustered systems: every point of
        // a chunk may belong to the slab); only the first chunk_count[c] entries of a window are ever touched
        uint32_t const n_chunks = (n + kSlabChunk - 1) / kSlabChunk;
        g.slab_pos.reserve((size_t) n_chunks * kSlabChunk);
        g.cell_of.reserve((size_t) n_chunks * kSlabChunk + n_chunks); // {cell} per list entry, then the chunk counts
        uint32_t* const chunk_count = g.cell_of.ptr + (size_t) n_chunks * kSlabChunk;
        {
            KernelScope ks(ctx, "cell_assign");
            unsigned const ablocks = std::min<unsigned>(n_chunks, (unsigned) ctx->sm_count * 5U);
            k_cell_assign_slab<<<ablocks, kSlabThreads, 0, ctx->stream>>>(pts->box, dim[0], dim[1], dim[2], pts->xyz.ptr, n,
                                                                         g.cell_start.ptr, d_flag, slab, g.slab_pos.ptr,
                                                                         g.cell_of.ptr, chunk_count);
        }
        // one scan over the slab's cells: a wrapped slab is the low range followed by the high one
        exclusive_scan_u32(ctx, g.cell_start.ptr + ra0, scan_n, true, g.cell_start.ptr + rb0, rb1 - rb0);
        {
            KernelScope ks(ctx, "cell_scatter");
            unsigned const sblocks = std::min<unsigned>(n_chunks, (unsigned) ctx->sm_count * 32U);
            k_cell_scatter_slab<<<std::max(sblocks, 1U), 256, 0, ctx->stream>>>(g.slab_pos.ptr, g.cell_of.ptr, chunk_count,
                                                                                n_chunks, g.cell_start.ptr, g.rank_in.ptr,
                                                                                g.sorted.ptr);
        }
        g.cell_of_valid = false;
    }
    else
    {
        g.cell_of.reserve(n);
        g.rank_in.reserve(std::max<size_t>(n, (size_t) n_cells + 4));
        {
            KernelScope ks(ctx, "cell_assign");
            unsigned const n_chunks = (n + kAssignChunk - 1) / kAssignChunk;
            unsigned const ablocks = std::min<unsigned>(n_chunks, (unsigned) ctx->sm_count * 8U);
            k_cell_assign_stream<<<ablocks, kAssignThreads, 0, ctx->stream>>>(pts->box, dim[0], dim[1], dim[2], pts->xyz.ptr,
                                                                             n, g.cell_of.ptr, g.rank_in.ptr,
                                                                             g.cell_start.ptr, d_flag);
        }
        exclusive_scan_u32(ctx, g.cell_start.ptr, (size_t) n_cells + 1, true);
        {
            KernelScope ks(ctx, "cell_scatter");
            k_cell_scatter<<<blocks, 256, 0, ctx->stream>>>(pts->box, dim[0], dim[1], dim[2], pts->xyz.ptr, n, g.cell_of.ptr,
                                                            g.rank_in.ptr, g.cell_start.ptr, g.sorted.ptr, g.shift.ptr,
                                                            d_flag);
        }
        g.cell_of_valid = true;
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
    for (int d = 0; d < 3; ++d)
    {
        g.dim[d] = dim[d];
        g.ambiguous[d] = amb[d];
    }
    g.n_cells = n_cells;
    g.r_search = r_search;
    g.shard = pts->shard;
    g.n_shards = pts->n_shards;
}

void sort_queries(fgpu_points* pts, const float* q_dev, uint32_t n_query)
{
    fgpu_ctx* ctx = pts->ctx;
    const fgpu_grid& g = pts->grid;
    ctx->q_cell.reserve(n_query);
    ctx->q_rank.reserve(n_query);
    ctx->q_cell_start.reserve((size_t) g.n_cells + 4);
    ctx->q_sorted.reserve(n_query);
    ctx->q_outside_flag.reserve(2);
    int* d_flag = ctx->q_outside_flag.ptr; // read on the device by the warp-cooperative search
    size_t const scan_words = scan_scratch_words((size_t) g.n_cells + 1);
    ctx->scan_tmp.reserve(scan_words);
    {
        KernelScope ks(ctx, "cell_prep");
        unsigned const pblocks = (unsigned) std::min<size_t>(((size_t) g.n_cells + 256) / 256, (size_t) ctx->sm_count * 8);
        k_cell_prep<<<std::max(pblocks, 1U), 256, 0, ctx->stream>>>(ctx->q_cell_start.ptr, nullptr, 0,
                                                                   (size_t) g.n_cells + 1, 0, 0, ctx->scan_tmp.ptr,
                                                                   (uint32_t) scan_words, d_flag, nullptr);
    }
    unsigned const blocks = (n_query + 255) / 256;
    {
        KernelScope ks(ctx, "cell_assign");
        unsigned const n_chunks = (n_query + kAssignChunk - 1) / kAssignChunk;
        unsigned const ablocks = std::max(1U, std::min<unsigned>(n_chunks, (unsigned) ctx->sm_count * 8U));
        k_cell_assign_stream<<<ablocks, kAssignThreads, 0, ctx->stream>>>(pts->box, g.dim[0], g.dim[1], g.dim[2], q_dev,
                                                                         n_query, ctx->q_cell.ptr, ctx->q_rank.ptr,
                                                                         ctx->q_cell_start.ptr, d_flag);
    }
    exclusive_scan_u32(ctx, ctx->q_cell_start.ptr, (size_t) g.n_cells + 1, true);
    {
        KernelScope ks(ctx, "cell_scatter");
        k_cell_scatter<<<std::max(blocks, 1U), 256, 0, ctx->stream>>>(pts->box, g.dim[0], g.dim[1], g.dim[2], q_dev, n_query,
                                                                     ctx->q_cell.ptr, ctx->q_rank.ptr, ctx->q_cell_start.ptr,
                                                                     ctx->q_sorted.ptr, nullptr, nullptr);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
