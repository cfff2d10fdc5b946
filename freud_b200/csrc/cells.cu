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

#include "internal.h"

namespace fgpu {

namespace {

constexpr int kScanThreads = 256;
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

__global__ void __launch_bounds__(kScanThreads) k_scan_chained(uint32_t* __restrict__ data, size_t n,
                                                               volatile unsigned long long* state,
                                                               unsigned int* ticket)
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
    size_t const base = (size_t) tile * kScanTile + (size_t) threadIdx.x * kScanItems;
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

// K1: cell index + arrival rank (one atomic per point) -- 12 B read, 8 B written per point
// A slab restricts the list to the cell layers [lo, lo + len) (cyclic) of one axis: a rank that only searches
// its share of the home tiles needs nothing else (the build must not stay serial).
struct SlabDev
{
    int axis; // 1: y layers (2-D grids), 2: z layers
    int lo;
    int len;  // < 0: no restriction
};

__global__ void __launch_bounds__(256) k_cell_assign(BoxDev box, int dx, int dy, int dz, const float* __restrict__ xyz,
                                                     uint32_t n, uint32_t* __restrict__ cell_of,
                                                     uint32_t* __restrict__ rank_in, uint32_t* __restrict__ cell_count,
                                                     int* __restrict__ any_shift, SlabDev slab)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
    {
        return;
    }
    float const x = xyz[3 * (size_t) i], y = xyz[3 * (size_t) i + 1], z = xyz[3 * (size_t) i + 2];
    int cx, cy, cz, nx, ny, nz;
    cell_coords(box, dx, dy, dz, x, y, z, cx, cy, cz, nx, ny, nz);
    if ((nx | ny | nz) != 0)
    {
        *any_shift = 1;
    }
    if (slab.len >= 0)
    {
        int const d = slab.axis == 2 ? dz : dy, c = slab.axis == 2 ? cz : cy;
        int rel = c - slab.lo;
        rel += rel < 0 ? d : 0;
        if (rel >= slab.len)
        {
            cell_of[i] = 0xffffffffU; // not in this rank's slab: left out of the list
            return;
        }
    }
    uint32_t const c = ((uint32_t) cz * dy + cy) * dx + cx;
    cell_of[i] = c;
    rank_in[i] = atomicAdd(&cell_count[c], 1U);
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
    if (cell == 0xffffffffU)
    {
        return; // outside the slab of this rank
    }
    float const x = xyz[3 * (size_t) i], y = xyz[3 * (size_t) i + 1], z = xyz[3 * (size_t) i + 2];
    uint32_t const slot = cell_start[cell] + rank_in[i];
    sorted[slot] = make_float4(x, y, z, __uint_as_float(i));
    if (shift != nullptr && *any_shift != 0)
    {
        int cx, cy, cz, nx, ny, nz;
        cell_coords(box, dx, dy, dz, x, y, z, cx, cy, cz, nx, ny, nz);
        shift[slot] = pack_shift(nx, ny, nz);
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

void exclusive_scan_u32(fgpu_ctx* ctx, uint32_t* data, size_t n)
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
    // scratch: one {flag, value} word per tile + the tile ticket
    ctx->scan_tmp.reserve(2 * tiles + 2);
    FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->scan_tmp.ptr, 0, (2 * tiles + 2) * sizeof(uint32_t), ctx->stream));
    auto* state = reinterpret_cast<unsigned long long*>(ctx->scan_tmp.ptr);
    auto* ticket = reinterpret_cast<unsigned int*>(ctx->scan_tmp.ptr + 2 * tiles);
    {
        KernelScope ks(ctx, "scan");
        k_scan_chained<<<(unsigned) tiles, kScanThreads, 0, ctx->stream>>>(data, n, state, ticket);
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
    p.ticket_begin = (uint32_t) ((uint64_t) a.n_tickets * (uint64_t) shard / (uint64_t) n_shards);
    p.ticket_end = (uint32_t) ((uint64_t) a.n_tickets * (uint64_t) (shard + 1) / (uint64_t) n_shards);
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
    g.cell_of.reserve(n);
    g.rank_in.reserve(n);
    g.cell_start.reserve((size_t) n_cells + 1);
    g.sorted.reserve(n);
    FGPU_CUDA_CHECK(cudaMemsetAsync(g.cell_start.ptr, 0, ((size_t) n_cells + 1) * sizeof(uint32_t), ctx->stream));
    // out-of-box flag: set on the device by K1, read on the device by K3 and the search kernels, so the
    // build needs no host round trip; the shift array is only ever written when the flag is set
    g.shift.reserve(n);
    g.any_shift_flag.reserve(1);
    int* d_flag = g.any_shift_flag.ptr;
    FGPU_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    unsigned const blocks = (n + 255) / 256;
    {
        KernelScope ks(ctx, "cell_assign");
        k_cell_assign<<<blocks, 256, 0, ctx->stream>>>(pts->box, dim[0], dim[1], dim[2], pts->xyz.ptr, n,
                                                       g.cell_of.ptr, g.rank_in.ptr, g.cell_start.ptr, d_flag, slab);
    }
    exclusive_scan_u32(ctx, g.cell_start.ptr, (size_t) n_cells + 1);
    {
        KernelScope ks(ctx, "cell_scatter");
        k_cell_scatter<<<blocks, 256, 0, ctx->stream>>>(pts->box, dim[0], dim[1], dim[2], pts->xyz.ptr, n,
                                                        g.cell_of.ptr, g.rank_in.ptr, g.cell_start.ptr, g.sorted.ptr,
                                                        g.shift.ptr, d_flag);
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
    ctx->q_cell_start.reserve((size_t) g.n_cells + 1);
    ctx->q_sorted.reserve(n_query);
    FGPU_CUDA_CHECK(
        cudaMemsetAsync(ctx->q_cell_start.ptr, 0, ((size_t) g.n_cells + 1) * sizeof(uint32_t), ctx->stream));
    ctx->q_outside_flag.reserve(1);
    int* d_flag = ctx->q_outside_flag.ptr; // read on the device by the warp-cooperative search
    FGPU_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    unsigned const blocks = (n_query + 255) / 256;
    {
        KernelScope ks(ctx, "cell_assign");
        k_cell_assign<<<blocks, 256, 0, ctx->stream>>>(pts->box, g.dim[0], g.dim[1], g.dim[2], q_dev, n_query,
                                                       ctx->q_cell.ptr, ctx->q_rank.ptr, ctx->q_cell_start.ptr,
                                                       d_flag, SlabDev {2, 0, -1});
    }
    exclusive_scan_u32(ctx, ctx->q_cell_start.ptr, (size_t) g.n_cells + 1);
    {
        KernelScope ks(ctx, "cell_scatter");
        k_cell_scatter<<<blocks, 256, 0, ctx->stream>>>(pts->box, g.dim[0], g.dim[1], g.dim[2], q_dev, n_query,
                                                        ctx->q_cell.ptr, ctx->q_rank.ptr, ctx->q_cell_start.ptr,
                                                        ctx->q_sorted.ptr, nullptr, nullptr);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
