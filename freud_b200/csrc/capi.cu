// C ABI of libfreud_b200.so (declared in include/freud_b200.h): argument validation with the reference's
// error behaviour, memory management, and the kernel sequences of each entry point.
#include <dlfcn.h>
#include <nccl.h> // types and enums only; the library is loaded with dlopen at run time

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <cstring>
#include <limits>
#include <nvtx3/nvToolsExt.h>

#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <mutex>
#include <thread>

#include "internal.h"

namespace fgpu {

namespace {
thread_local std::string g_last_error;
}

void set_last_error(const std::string& msg)
{
    g_last_error = msg;
}

cudaStream_t& current_stream()
{
    static thread_local cudaStream_t stream = nullptr;
    return stream;
}

namespace {

// NVTX range around every C-ABI entry point: nvtx3 is header-only and loads the tool's
// injection library on demand, so the ranges cost nothing unless a profiler is attached.
struct NvtxRange
{
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

template<typename F> int guarded_named(const char* name, F&& f)
{
    NvtxRange const range(name);
    try
    {
        f();
        return FGPU_OK;
    }
    catch (const Error& e)
    {
        set_last_error(e.what());
        return e.code;
    }
    catch (const std::bad_alloc&)
    {
        set_last_error("out of host memory");
        return FGPU_ENOMEM;
    }
    catch (const std::exception& e)
    {
        set_last_error(e.what());
        return FGPU_ERUNTIME;
    }
}
// the entry point's own name labels the range
#define guarded(...) guarded_named(__func__, __VA_ARGS__)

void require(bool cond, int code, const char* msg)
{
    if (!cond)
    {
        throw Error(code, msg);
    }
}

// host float arithmetic that must not be contracted (this file is compiled with -ffp-contract=off)
BoxDev make_box(const float* b6, int is2d)
{
    BoxDev b;
    b.is2d = is2d != 0 ? 1 : 0;
    b.Lx = b6[0];
    b.Ly = b6[1];
    b.Lz = b.is2d ? 0.0f : b6[2]; // Box.h:102-106
    b.xy = b6[3];
    b.xz = b6[4];
    b.yz = b6[5];
    volatile float half = 1.0f / 2.0f; // m_hi = m_L / 2.0f multiplies by the reciprocal, VectorMath.h:208-212
    volatile float hx = b.Lx * half, hy = b.Ly * half, hz = b.Lz * half;
    b.lox = -hx;
    b.loy = -hy;
    b.loz = -hz;
    volatile float t = b.yz * b.xy;
    volatile float t_xz = b.xz - t; // Box.h:246
    b.t_xz = t_xz;
    // lattice vectors, Box.h:503-518
    volatile float bx = b.Ly * b.xy, cx = b.Lz * b.xz, cy = b.Lz * b.yz;
    b.ax = b.Lx;
    b.bx = bx;
    b.by = b.Ly;
    b.cx = b.is2d ? 0.0f : cx;
    b.cy = b.is2d ? 0.0f : cy;
    b.cz = b.is2d ? 0.0f : b.Lz;
    return b;
}

// Box::getNearestPlaneDistance, Box.h:489-497
void plane_distances(const BoxDev& b, float out[3])
{
    volatile float t0 = b.xy * b.yz;
    volatile float t = t0 - b.xz;
    volatile float a0 = b.xy * b.xy;
    volatile float a1 = 1.0f + a0;
    volatile float a2 = t * t;
    volatile float a3 = a1 + a2;
    out[0] = b.Lx / std::sqrt((float) a3);
    volatile float c0 = b.yz * b.yz;
    volatile float c1 = 1.0f + c0;
    out[1] = b.Ly / std::sqrt((float) c1);
    out[2] = b.Lz;
}

float box_volume(const BoxDev& b)
{
    volatile float a = b.Lx * b.Ly;
    if (b.is2d)
    {
        return a;
    }
    volatile float v = a * b.Lz;
    return v;
}

void h2d(fgpu_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes != 0)
    {
        FGPU_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
}

void d2h(fgpu_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes != 0)
    {
        FGPU_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
}

void sync(fgpu_ctx* ctx)
{
    FGPU_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void bind_device(fgpu_ctx* ctx)
{
    FGPU_CUDA_CHECK(cudaSetDevice(ctx->device));
    current_stream() = ctx->stream;
}

// for the destroy paths, which must not throw
void bind_quiet(fgpu_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    current_stream() = ctx->stream;
}

struct QueryView
{
    const float4* sorted;        // cell-ordered, w = original index
    const float* xyz;            // original order
    const uint32_t* cell_start;  // n_cells + 1 offsets into sorted
    const int* outside_flag;     // device flag: some query lies outside the box
};

// Stage the query points (if any) and produce their cell-ordered view on the grid of pts.
QueryView prepare_queries(fgpu_points* pts, const float* q_host, const float* q_dev, uint32_t n_query)
{
    fgpu_ctx* ctx = pts->ctx;
    QueryView v;
    if (q_host == nullptr && q_dev == nullptr)
    {
        v.sorted = pts->grid.sorted.ptr;
        v.xyz = pts->xyz.ptr;
        v.cell_start = pts->grid.cell_start.ptr;
        v.outside_flag = pts->grid.any_shift_flag.ptr;
        return v;
    }
    if (q_host != nullptr)
    {
        ctx->q_stage.reserve((size_t) n_query * 3);
        h2d(ctx, ctx->q_stage.ptr, q_host, (size_t) n_query * 3 * sizeof(float));
        q_dev = ctx->q_stage.ptr;
    }
    sort_queries(pts, q_dev, n_query);
    v.sorted = ctx->q_sorted.ptr;
    v.xyz = q_dev;
    v.cell_start = ctx->q_cell_start.ptr;
    v.outside_flag = ctx->q_outside_flag.ptr;
    return v;
}

void validate_ball(const fgpu_points* pts, int flavour, float r_max, float r_min)
{
    require(flavour == FGPU_FLAVOUR_WRAP || flavour == FGPU_FLAVOUR_IMAGE || flavour == FGPU_FLAVOUR_GHOST,
            FGPU_EINVALID, "unknown flavour");
    // NeighborQueryPerPointIterator ctor, NeighborQuery.h:321-328
    require(r_max > 0, FGPU_EINVALID, "NeighborQuery requires r_max to be positive.");
    require(r_max > r_min, FGPU_EINVALID, "NeighborQuery requires that r_max must be greater than r_min.");
    if (flavour != FGPU_FLAVOUR_WRAP)
    {
        // updateImageVectors, NeighborQuery.h:503-510 (all axes periodic); CellQuery::validateQueryArgs,
        // CellQuery.h:186-192
        double const two_r = (double) r_max * 2.0;
        bool const too_large = pts->plane_dist[0] <= two_r || pts->plane_dist[1] <= two_r
            || (!pts->box.is2d && pts->plane_dist[2] <= two_r);
        require(!too_large, FGPU_ERUNTIME,
                flavour == FGPU_FLAVOUR_IMAGE ? "The AABBQuery r_max is too large for this box."
                                              : "The CellQuery r_max is too large for this box.");
    }
}

SearchArgs base_search_args(fgpu_points* pts, const QueryView& qv, uint32_t n_query, uint32_t q_index_offset,
                            float r_max, float r_min, int exclude_ii)
{
    SearchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.box = pts->box;
    a.grid = grid_dev(pts);
    a.q_sorted = qv.sorted;
    a.n_query = n_query;
    a.q_index_offset = q_index_offset;
    a.r_max = r_max;
    a.r_min = r_min;
    a.exclude_ii = exclude_ii != 0;
    a.evals = pts->ctx->count_evals ? pts->ctx->d_evals : nullptr;
    return a;
}

// RN(1 / L) for the Markstein division of search2.cu; long double keeps the double rounding harmless
float rounded_reciprocal(float L)
{
    return L != 0.0f ? (float) (1.0L / (long double) L) : 0.0f;
}

Search2Args base_search2_args(fgpu_points* pts, const QueryView& qv, uint32_t q_index_offset, float r_max,
                              float r_min, int exclude_ii)
{
    fgpu_ctx* ctx = pts->ctx;
    const fgpu_grid& g = pts->grid;
    Search2Args a;
    std::memset(&a, 0, sizeof(a));
    a.box = pts->box;
    a.dx = g.dim[0];
    a.dy = g.dim[1];
    a.dz = g.dim[2];
    a.n_cells = g.n_cells;
    a.cell_start = g.cell_start.ptr;
    a.sorted = g.sorted.ptr;
    a.q_cell_start = qv.cell_start;
    a.q_sorted = qv.sorted;
    a.flag_points_outside = g.any_shift_flag.ptr;
    a.flag_queries_outside = qv.outside_flag;
    a.q_index_offset = q_index_offset;
    a.r_max = r_max;
    a.r_min = r_min;
    a.exclude_ii = exclude_ii != 0;
    a.rcp_lx = rounded_reciprocal(pts->box.Lx);
    a.rcp_ly = rounded_reciprocal(pts->box.Ly);
    a.rcp_lz = rounded_reciprocal(pts->box.Lz);
    // stage-1 acceptance radius: r_max + 4E, E bounding the rounding error of a displacement component in either
    // arithmetic (coordinates of magnitude <= (Lx + Ly + Lz)(1 + |tilts|), a dozen roundings of 2^-24 each)
    double const ext = ((double) pts->box.Lx + pts->box.Ly + pts->box.Lz)
        * (1.0 + std::fabs((double) pts->box.xy) + std::fabs((double) pts->box.xz) + std::fabs((double) pts->box.yz));
    double const E = 16.0 * 5.9604644775390625e-08 * ext;
    double const r_hi = (double) r_max + 4.0 * E;
    a.r_hi_sq = std::nextafter((float) (r_hi * r_hi * (1.0 + 1.0e-6)), INFINITY);
    search2_plan(a, pts->n, ctx->tune_span);
    a.out_cap = search2_out_cap((pts->box.is2d ? 3.0 : 9.0) * (a.span + 2) * (double) pts->n
                                / (double) std::max(g.n_cells, 1U));
    a.fail = reinterpret_cast<int*>(ctx->d_scalars + 4);
    a.cursor = ctx->d_scalars + 5;
    a.work_counter = reinterpret_cast<unsigned int*>(ctx->d_scalars + 6);
    a.evals = ctx->count_evals ? ctx->d_evals : nullptr;
    return a;
}

// bonds a query point expects at the mean density: what sizes the hit buffers and picks the search's mapping
double expected_hits_per_query(const fgpu_points* pts, float r_window)
{
    double const vol = box_volume(pts->box);
    double const shell = pts->box.is2d ? M_PI * (double) r_window * r_window
                                       : 4.0 / 3.0 * M_PI * (double) r_window * r_window * r_window;
    return vol > 0 ? (double) pts->n / vol * shell : 0.0;
}

// First guess of a query's bond count, which sizes the bag and the output arrays (the search reports the exact count
// and the host repeats with it when the guess was short): the previous query's count when this context has one --
// a trajectory's frames resemble each other -- but never more than four times what the mean density predicts for THIS
// query, nor more than every pair: one large query must not make every small query after it allocate gigabytes.
uint64_t bag_capacity(uint64_t hint, uint32_t n_query, uint32_t n_points, double volume, double shell)
{
    double const ideal = (double) n_query * (double) n_points / volume * shell;
    uint64_t const estimate = (uint64_t) (1.25 * ideal) + 4096;
    uint64_t cap = hint != 0 ? hint + hint / 16 + 1024 : estimate;
    cap = std::min<uint64_t>(cap, 4 * estimate);
    cap = std::min<uint64_t>(cap, (uint64_t) n_query * (uint64_t) n_points + 1024);
    return cap;
}

std::unique_ptr<fgpu_nlist> new_nlist(fgpu_ctx* ctx, uint32_t n_query, uint32_t n_points)
{
    std::unique_ptr<fgpu_nlist> nl(new fgpu_nlist());
    nl->ctx = ctx;
    nl->n_query = n_query;
    nl->n_points = n_points;
    nl->swap(ctx->spare_nlist); // the arrays the last destroyed list left behind (empty if none)
    nl->unit_weights = true;    // every query writes weight 1; fgpu_nlist_from_host clears this for given weights
    nl->row_start.reserve((size_t) n_query + 1);
    nl->counts.reserve((size_t) n_query + 1);
    nl->segments.reserve((size_t) n_query + 1);
    return nl;
}

void alloc_bonds(fgpu_nlist* nl, uint64_t n_bonds)
{
    require(n_bonds <= 0xffffffffULL, FGPU_ERUNTIME, "NeighborList would exceed 2^32 - 1 bonds");
    nl->n_bonds = n_bonds;
    nl->neighbors.reserve(n_bonds * 2 + 2);
    nl->distances.reserve(n_bonds + 1);
    nl->weights.reserve(n_bonds + 1);
    nl->vectors.reserve(n_bonds * 3 + 3);
}

// counts (ctx scratch, indexed by original query) -> nl->counts, nl->row_start (exclusive scan); returns total
uint64_t finish_counts(fgpu_ctx* ctx, fgpu_nlist* nl, const uint32_t* counts_dev, unsigned long long* d_total)
{
    uint32_t const nq = nl->n_query;
    FGPU_CUDA_CHECK(cudaMemcpyAsync(nl->counts.ptr, counts_dev, (size_t) nq * sizeof(uint32_t),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyAsync(nl->row_start.ptr, counts_dev, (size_t) nq * sizeof(uint32_t),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemsetAsync(nl->row_start.ptr + nq, 0, sizeof(uint32_t), ctx->stream));
    exclusive_scan_u32(ctx, nl->row_start.ptr, (size_t) nq + 1);
    d2h(ctx, ctx->h_scalars, d_total, sizeof(unsigned long long));
    sync(ctx);
    return ctx->h_scalars[0];
}

void ball_query_impl(fgpu_points* pts, const float* q_host, const float* q_dev, uint32_t n_query,
                     uint32_t q_index_offset, int flavour, float r_max, float r_min, int exclude_ii,
                     int sort_by_distance, fgpu_nlist** out)
{
    require(pts != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
    fgpu_ctx* ctx = pts->ctx;
    bind_device(ctx);
    validate_ball(pts, flavour, r_max, r_min);
    bool const self = q_host == nullptr && q_dev == nullptr;
    require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
    auto nl = new_nlist(ctx, n_query, pts->n);
    nl->q_index_offset = self ? 0 : q_index_offset;
    if (n_query == 0)
    {
        alloc_bonds(nl.get(), 0);
        *out = nl.release();
        return;
    }
    require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
    build_grid(pts, r_max);
    QueryView const qv = prepare_queries(pts, q_host, q_dev, n_query);

    // ---- production path: warp-cooperative single-pass search into a bag, then ranked emit ----------------
    bool fast_counted_evals = false;
    Search2Args s2 = base_search2_args(pts, qv, q_index_offset, r_max, r_min, exclude_ii);
    search2_choose_mapping(s2, flavour, n_query, pts->n, expected_hits_per_query(pts, r_max),
                           ctx->tune_lanes_over_queries);
    if (!ctx->force_general && search2_supported(s2, S2_NL))
    {
        // Capacities of the bag and of the output arrays: the previous query's bond count if there was one, else
        // the ideal-gas expectation.  Search, offsets scan and emit are enqueued back to back -- the emit kernel
        // checks on the device that the search succeeded and that everything fits -- so the frame has one host
        // round trip, at the end, and the GPU never waits for the host in between.
        double const vol = box_volume(pts->box);
        double const shell = pts->box.is2d ? M_PI * (double) r_max * r_max
                                           : 4.0 / 3.0 * M_PI * (double) r_max * r_max * r_max;
        uint64_t cap = bag_capacity(ctx->bag_hint, n_query, pts->n, vol, shell);
        ctx->tmp_start.reserve((size_t) n_query + 1);
        bool done = false, general = false;
        uint64_t n_bonds = 0;
        launch_count_evals(ctx, s2, n_query, pts->grid.cell_of.ptr, pts->n);
        bool const evals_counted = s2.evals != nullptr;
        for (int attempt = 0; attempt < 4 && !done && !general; ++attempt)
        {
            cap = std::min<uint64_t>(cap, 0xffffffffULL);
            ctx->bag4.reserve(cap);
            alloc_bonds(nl.get(), cap); // n_bonds is corrected below
            s2.bag = ctx->bag4.ptr;
            s2.temp_cap = (uint32_t) cap;
            s2.counts = nl->counts.ptr;
            s2.tmp_start = ctx->tmp_start.ptr;
            FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 4, 0, 3 * sizeof(unsigned long long), ctx->stream));
            // the search writes every row's count twice (counts and the array the row scan turns into offsets) and its
            // first block clears the scan's scratch: no D2D copy and no memset between the two kernels
            size_t const scan_words = scan_scratch_words((size_t) n_query + 1);
            ctx->scan_tmp.reserve(scan_words);
            s2.counts_copy = nl->row_start.ptr;
            s2.zero_words = ctx->scan_tmp.ptr;
            s2.zero_n = (uint32_t) scan_words;
            s2.zero_tail = nl->row_start.ptr + n_query;
            launch_search2(ctx, flavour, S2_NL, s2);
            exclusive_scan_u32(ctx, nl->row_start.ptr, (size_t) n_query + 1, true);
            Emit2Args e;
            e.bag = ctx->bag4.ptr;
            e.tmp_start = ctx->tmp_start.ptr;
            e.row_start = nl->row_start.ptr;
            e.counts = nl->counts.ptr;
            e.segments = nl->segments.ptr;
            e.n_query = n_query;
            e.fail = s2.fail;
            e.cursor = s2.cursor;
            e.bag_cap = cap;
            e.out_cap = cap;
            e.neighbors = nl->neighbors.ptr;
            e.distances = nl->distances.ptr;
            e.weights = nl->weights.ptr;
            e.vectors = nl->vectors.ptr;
            launch_emit2(ctx, sort_by_distance, e);
            d2h(ctx, ctx->h_scalars + 4, ctx->d_scalars + 4, 2 * sizeof(unsigned long long));
            // the bond count is the total of the row scan: the bag cursor may run a few records ahead of it (the
            // lanes-over-queries search reserves a record per filter survivor, search_lq.cu)
            ctx->h_scalars[1] = 0;
            d2h(ctx, ctx->h_scalars + 1, nl->row_start.ptr + n_query, sizeof(uint32_t));
            sync(ctx);
            int const fail = (int) (ctx->h_scalars[4] & 0xffffffffULL);
            uint32_t const densest = (uint32_t) (ctx->h_scalars[4] >> 32);
            if (search2_lq_fallback(s2, fail))
            {
                // rows beyond the buffers of the lanes-over-queries search: the tile walk takes the frame
            }
            else if (fail == 2 && densest <= search2_max_out_cap() && s2.out_cap < search2_max_out_cap())
            {
                // a tile denser than the uniform estimate (clustered system): same kernels, a hit buffer that holds
                // it -- fewer resident warps, still far ahead of the thread-per-query family
                s2.out_cap = std::min(search2_max_out_cap(), (densest + densest / 8 + 31U) & ~31U);
            }
            else if (fail != 0)
            {
                general = true; // 1: points outside the box, 2: a tile beyond the largest buffer
                fast_counted_evals = evals_counted && fail != 1; // the count kernel skips case 1 itself
            }
            else if (ctx->h_scalars[5] <= cap)
            {
                n_bonds = ctx->h_scalars[1];
                done = true;
            }
            else
            {
                cap = ctx->h_scalars[5]; // exact size, the search is deterministic
            }
        }
        if (done)
        {
            ctx->bag_hint = n_bonds;
            nl->n_bonds = n_bonds;
            *out = nl.release();
            return;
        }
    }

    // ---- general path (small grids, points outside the box, very long rows): two-pass thread-per-query ------
    ctx->row_counts.reserve((size_t) n_query + 1);
    SearchArgs a = base_search_args(pts, qv, n_query, q_index_offset, r_max, r_min, exclude_ii);
    if (fast_counted_evals)
    {
        a.evals = nullptr;
    }
    a.sort_by_distance = sort_by_distance != 0;
    a.row_counts = ctx->row_counts.ptr;
    a.total = ctx->d_scalars;
    FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(unsigned long long), ctx->stream));
    launch_search(ctx, flavour, SEARCH_COUNT, a);
    uint64_t const n_bonds = finish_counts(ctx, nl.get(), ctx->row_counts.ptr, ctx->d_scalars);
    alloc_bonds(nl.get(), n_bonds);
    if (n_bonds != 0)
    {
        ctx->bag.reserve(n_bonds);
        a.row_start = nl->row_start.ptr;
        a.bag = ctx->bag.ptr;
        a.evals = nullptr; // the second pass repeats the same evaluations; count them once
        launch_search(ctx, flavour, SEARCH_FILL, a);
        EmitArgs e;
        e.box = pts->box;
        e.sorted = pts->grid.sorted.ptr;
        e.q_xyz = qv.xyz;
        e.bag = ctx->bag.ptr;
        e.row_start = nl->row_start.ptr;
        e.n_bonds = n_bonds;
        e.r_max = r_max;
        e.r_min = r_min;
        e.neighbors = nl->neighbors.ptr;
        e.distances = nl->distances.ptr;
        e.weights = nl->weights.ptr;
        e.vectors = nl->vectors.ptr;
        launch_emit(ctx, flavour, e);
    }
    launch_segments(ctx, nl->row_start.ptr, nl->counts.ptr, nl->segments.ptr, n_query);
    // no final sync: every consumer of the list (copy, RDF, Steinhardt, destroy) is ordered on the same stream,
    // and the host query buffer was consumed before the bond total was read back above
    *out = nl.release();
}

// Ball search of n_query rows that leaves its hits in ctx->bag4, grouped by query row: row r owns the records
// [ctx->tmp_start[r], + ctx->row_counts[r]) = {bond vector, bits(point index)}.  For the computes that only bin or
// sum the bonds of a query made for them alone: no ranked emit, no NeighborList.  Returns false when the
// warp-cooperative search does not take the frame (tiny grids, points outside the box, a tile beyond the largest
// buffer, FGPU_SEARCH=general): the caller goes through a NeighborList then.
bool search_to_bag(fgpu_points* pts, const float* q_host, uint32_t n_query, int flavour, float r_max, float r_min,
                   int exclude_ii, uint64_t* n_bonds)
{
    fgpu_ctx* ctx = pts->ctx;
    if (n_query == 0 || ctx->force_general)
    {
        return false;
    }
    build_grid(pts, r_max);
    QueryView const qv = prepare_queries(pts, q_host, nullptr, n_query);
    Search2Args s2 = base_search2_args(pts, qv, 0, r_max, r_min, exclude_ii);
    search2_choose_mapping(s2, flavour, n_query, pts->n, expected_hits_per_query(pts, r_max),
                           ctx->tune_lanes_over_queries);
    if (!search2_supported(s2, S2_NL))
    {
        return false;
    }
    double const vol = box_volume(pts->box);
    double const shell = pts->box.is2d ? M_PI * (double) r_max * r_max : 4.0 / 3.0 * M_PI * (double) r_max * r_max * r_max;
    uint64_t cap = bag_capacity(ctx->bag_hint, n_query, pts->n, vol, shell);
    ctx->tmp_start.reserve((size_t) n_query + 1);
    ctx->row_counts.reserve((size_t) n_query + 1);
    launch_count_evals(ctx, s2, n_query, pts->grid.cell_of.ptr, pts->n);
    s2.evals = nullptr;
    for (int attempt = 0; attempt < 4; ++attempt)
    {
        cap = std::min<uint64_t>(cap, 0xffffffffULL);
        ctx->bag4.reserve(cap);
        s2.bag = ctx->bag4.ptr;
        s2.temp_cap = (uint32_t) cap;
        s2.counts = ctx->row_counts.ptr;
        s2.tmp_start = ctx->tmp_start.ptr;
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 4, 0, 3 * sizeof(unsigned long long), ctx->stream));
        launch_search2(ctx, flavour, S2_NL, s2);
        d2h(ctx, ctx->h_scalars + 4, ctx->d_scalars + 4, 2 * sizeof(unsigned long long));
        sync(ctx);
        int const fail = (int) (ctx->h_scalars[4] & 0xffffffffULL);
        uint32_t const densest = (uint32_t) (ctx->h_scalars[4] >> 32);
        if (search2_lq_fallback(s2, fail))
        {
            // rows beyond the buffers of the lanes-over-queries search: the tile walk takes the frame
        }
        else if (fail == 2 && densest <= search2_max_out_cap() && s2.out_cap < search2_max_out_cap())
        {
            s2.out_cap = std::min(search2_max_out_cap(), (densest + densest / 8 + 31U) & ~31U);
        }
        else if (fail != 0)
        {
            return false;
        }
        else if (ctx->h_scalars[5] <= cap)
        {
            *n_bonds = ctx->h_scalars[5];
            ctx->bag_hint = *n_bonds;
            return true;
        }
        else
        {
            cap = ctx->h_scalars[5]; // exact size, the search is deterministic
        }
    }
    return false;
}

// fuse_push: the caller ends the epoch right after this frame (fgpu_rdf_accumulate_reduce) and the RDF has a peer
// mailbox -- where the frame is one launch of the warp-cooperative kernel with no fallback behind it (the sharded
// path), that launch's last block sends the finished histogram to every rank.  Returns whether it did.
bool rdf_accumulate_impl(fgpu_rdf* rdf, fgpu_points* pts, const float* q_host, const float* q_dev, uint32_t n_query,
                         uint32_t q_index_offset, int flavour, float q_r_max, float q_r_min, int exclude_ii,
                         bool fuse_push = false)
{
    require(rdf != nullptr && pts != nullptr, FGPU_EINVALID, "null argument");
    require(rdf->ctx == pts->ctx, FGPU_EINVALID, "rdf and points belong to different contexts");
    fgpu_ctx* ctx = pts->ctx;
    bind_device(ctx);
    validate_ball(pts, flavour, q_r_max, q_r_min);
    bool const self = q_host == nullptr && q_dev == nullptr;
    require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
    if (n_query == 0)
    {
        return false;
    }
    rdf->reduced_valid = false; // new counts: an earlier sum over the ranks no longer describes this histogram
    bool const sharded = pts->n_shards > 1;
    require(!sharded || self, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
    build_grid(pts, q_r_max);
    QueryView const qv = prepare_queries(pts, q_host, q_dev, n_query);
    Search2Args s2 = base_search2_args(pts, qv, q_index_offset, q_r_max, q_r_min, exclude_ii);
    s2.axis = rdf->axis;
    // Queries are the points themselves and no image vector is involved in most pairs: r_ij and r_ji are exact
    // negatives there, so one test stands for both bonds (IMAGE arithmetic only: Box::wrap is not odd in floating
    // point).  The self pair (i, i) lies in the tile's own row, which is always walked in full.
    s2.symmetric = self && q_index_offset == 0 && flavour == FGPU_FLAVOUR_IMAGE && ctx->tune_no_symmetry == 0;
    bool const fast = !ctx->force_general && search2_supported(s2, S2_RDF);
    if (sharded)
    {
        // This rank bins the pairs of its share of the home tiles; the cell list holds the slab they can see.
        // There is no general-kernel fallback here: it would need the whole list.
        // (FGPU_SEARCH=general does not apply: the sharded path has one kernel family)
        require(search2_supported(s2, S2_RDF), FGPU_ERUNTIME,
                "sharded RDF accumulation needs a grid of at least 3 cells per periodic axis");
        ShardPlan const sp = shard_plan(pts->grid.dim, pts->n, pts->shard, pts->n_shards);
        s2.ticket_begin = sp.ticket_begin;
        s2.ticket_end = sp.ticket_end;
        s2.cell_begin = sp.cell_begin;
        s2.cell_end = sp.cell_end;
        s2.hist = rdf->hist.ptr;
        // a point outside the box makes the kernel give up; the flag is sticky in the word behind the counters
        // and fgpu_rdf_read reports it, so the frame needs no host round trip
        s2.fail = reinterpret_cast<int*>(rdf->hist.ptr + rdf->axis.bins);
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 5, 0, 2 * sizeof(unsigned long long), ctx->stream));
        bool pushed = false;
        if (s2.ticket_end > s2.ticket_begin)
        {
            if (fuse_push && rdf->peer != nullptr)
            {
                s2.push = 1;
                s2.done_counter = reinterpret_cast<unsigned int*>(ctx->d_scalars + 6) + 1; // zeroed just above
                s2.peer = rdf->peer->box;
                pushed = true;
            }
            launch_search2(ctx, flavour, S2_RDF, s2);
            launch_count_evals(ctx, s2, n_query, nullptr, pts->n); // self query: no per-point cell table needed
        }
        return pushed;
    }
    SearchArgs a = base_search_args(pts, qv, n_query, q_index_offset, q_r_max, q_r_min, exclude_ii);
    a.axis = rdf->axis;
    a.hist = rdf->hist.ptr;
    if (fast)
    {
        // The warp-cooperative kernel adds nothing and raises the fail flag when a point lies outside the box;
        // the general kernel is enqueued behind it and returns at once unless that flag is set, so the choice
        // is made on the device and the frame needs no host round trip.
        s2.hist = rdf->hist.ptr;
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 4, 0, 3 * sizeof(unsigned long long), ctx->stream));
        launch_search2(ctx, flavour, S2_RDF, s2);
        launch_count_evals(ctx, s2, n_query, pts->grid.cell_of.ptr, pts->n);
        a.only_if = reinterpret_cast<const int*>(ctx->d_scalars + 4);
        // a.evals stays: the general kernel only runs (and counts) when the fast one and its count bailed out
    }
    launch_search(ctx, flavour, SEARCH_RDF, a);
    if (q_host != nullptr)
    {
        sync(ctx); // the caller may reuse its host buffer / the staging buffer is context scratch
    }
    return false;
}

// ---- NCCL, loaded lazily -----------------------------------------------------------------------------
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
        = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
        {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle != nullptr)
            {
                break;
            }
        }
        if (api.handle == nullptr)
        {
            api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.handle, "ncclBroadcast"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.handle, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.handle, "ncclGroupEnd"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString
            || !api.AllGather || !api.Broadcast || !api.GroupStart || !api.GroupEnd)
        {
            api.why = "libnccl.so.2 lacks a required symbol";
            api.handle = nullptr;
        }
    });
    return api;
}

void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess)
    {
        throw Error(FGPU_ENCCL, std::string(what) + ": " + nccl().GetErrorString(r));
    }
}

NcclApi& nccl_or_throw()
{
    NcclApi& api = nccl();
    if (api.handle == nullptr)
    {
        throw Error(FGPU_ENCCL, api.why);
    }
    return api;
}

} // namespace

int nccl_available(std::string* why)
{
    NcclApi& api = nccl();
    if (api.handle == nullptr && why != nullptr)
    {
        *why = api.why;
    }
    return api.handle != nullptr;
}

} // namespace fgpu

using namespace fgpu;

extern "C" {

const char* fgpu_last_error(void)
{
    return g_last_error.c_str();
}

int fgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* fgpu_version(void)
{
    static std::string text;
    static std::once_flag once;
    std::call_once(once, [] {
        text = "freud_b200 0.1 (sm_100a)";
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0)
        {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, 0) == cudaSuccess)
            {
                char buf[320];
                std::snprintf(buf, sizeof(buf), " on %s, %d SMs, cc %d.%d", prop.name, prop.multiProcessorCount,
                              prop.major, prop.minor);
                text += buf;
            }
        }
        else
        {
            cudaGetLastError();
            text += " (no CUDA device visible)";
        }
    });
    return text.c_str();
}

// ---- context ---------------------------------------------------------------------------------------
int fgpu_ctx_create(int device, fgpu_ctx** out)
{
    return guarded([&] {
        require(out != nullptr, FGPU_EINVALID, "null argument");
        int n = 0;
        cudaError_t const err = cudaGetDeviceCount(&n);
        if (err != cudaSuccess || n == 0)
        {
            cudaGetLastError();
            throw Error(FGPU_ECUDA, "no CUDA device available: freud_b200 has no CPU fallback");
        }
        require(device >= 0 && device < n, FGPU_EINVALID, "device index out of range");
        FGPU_CUDA_CHECK(cudaSetDevice(device));
        std::unique_ptr<fgpu_ctx> ctx(new fgpu_ctx());
        ctx->device = device;
        FGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        current_stream() = ctx->stream;
        // keep freed blocks in the stream-ordered pool instead of handing them back to the driver
        cudaMemPool_t pool = nullptr;
        FGPU_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t threshold = UINT64_MAX;
        FGPU_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        FGPU_CUDA_CHECK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
        FGPU_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ctx->d_scalars), 8 * sizeof(unsigned long long)));
        FGPU_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ctx->d_evals), sizeof(unsigned long long)));
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars, 0, 8 * sizeof(unsigned long long), ctx->stream));
        FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_evals, 0, sizeof(unsigned long long), ctx->stream));
        FGPU_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_scalars), 8 * sizeof(unsigned long long)));
        FGPU_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *out = ctx.release();
    });
}

int fgpu_ctx_trim(fgpu_ctx* ctx)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        bind_device(ctx);
        NlistStorage empty;
        ctx->spare_nlist.swap(empty); // freed when `empty` goes out of scope
        ctx->bag.release();
        ctx->bag4.release();
        ctx->bag4b.release();
        ctx->tmp_start.release();
        ctx->knn_hits.release();
        ctx->knn_unresolved.release();
        ctx->knn_subset.release();
        ctx->knn_d.release();
        ctx->knn_s.release();
        ctx->q_stage.release();
        ctx->q_sorted.release();
        ctx->row_counts.release();
        ctx->row_start.release();
        sync(ctx);
        cudaMemPool_t pool = nullptr;
        FGPU_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        FGPU_CUDA_CHECK(cudaMemPoolTrimTo(pool, 0));
    });
}

void fgpu_ctx_destroy(fgpu_ctx* ctx)
{
    if (ctx == nullptr)
    {
        return;
    }
    bind_quiet(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->timers)
    {
        cudaEventDestroy(t.begin);
        cudaEventDestroy(t.end);
    }
    for (cudaEvent_t e : ctx->event_pool)
    {
        cudaEventDestroy(e);
    }
    cudaFree(ctx->d_scalars);
    cudaFree(ctx->d_evals);
    cudaFreeHost(ctx->h_scalars);
    if (ctx->pinned_stage != nullptr)
    {
        cudaFreeHost(ctx->pinned_stage);
    }
    cudaStream_t const s = ctx->stream;
    delete ctx; // frees the scratch buffers
    cudaStreamDestroy(s);
}

int fgpu_ctx_synchronize(fgpu_ctx* ctx)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        bind_device(ctx);
        sync(ctx);
    });
}

void* fgpu_ctx_stream(fgpu_ctx* ctx)
{
    return ctx != nullptr ? static_cast<void*>(ctx->stream) : nullptr;
}

uint64_t fgpu_ctx_launch_count(fgpu_ctx* ctx)
{
    return ctx != nullptr ? ctx->launches : 0;
}

int fgpu_ctx_count_pair_evals(fgpu_ctx* ctx, int enable)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        ctx->count_evals = enable != 0;
    });
}

int fgpu_ctx_pair_evals(fgpu_ctx* ctx, uint64_t* out, int reset)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        bind_device(ctx);
        d2h(ctx, ctx->h_scalars + 1, ctx->d_evals, sizeof(unsigned long long));
        if (reset)
        {
            FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_evals, 0, sizeof(unsigned long long), ctx->stream));
        }
        sync(ctx);
        *out = ctx->h_scalars[1];
    });
}

int fgpu_ctx_force_general_search(fgpu_ctx* ctx, int enable)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        ctx->force_general = enable != 0;
    });
}

int fgpu_ctx_set_tuning(fgpu_ctx* ctx, const char* key, int value)
{
    return guarded([&] {
        require(ctx != nullptr && key != nullptr, FGPU_EINVALID, "null argument");
        std::string const k(key);
        if (k == "span")
        {
            ctx->tune_span = value;
        }
        else if (k == "no_symmetry")
        {
            ctx->tune_no_symmetry = value;
        }
        else if (k == "lanes_over_queries")
        {
            ctx->tune_lanes_over_queries = value;
        }
        else if (k == "lq_blocks")
        {
            ctx->tune_lq_blocks = value;
        }
        else if (k == "pmft_cluster")
        {
            ctx->tune_pmft_cluster = value;
        }
        else
        {
            throw Error(FGPU_EINVALID, "unknown tuning key: " + k);
        }
    });
}

int fgpu_ctx_kernel_timeline(fgpu_ctx* ctx, char* out, uint64_t cap)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr && cap != 0, FGPU_EINVALID, "null argument");
        bind_device(ctx);
        sync(ctx);
        std::string text;
        for (const auto& t : ctx->timers)
        {
            float b = 0.0f, e = 0.0f;
            FGPU_CUDA_CHECK(cudaEventElapsedTime(&b, ctx->timers.front().begin, t.begin));
            FGPU_CUDA_CHECK(cudaEventElapsedTime(&e, ctx->timers.front().begin, t.end));
            char line[160];
            std::snprintf(line, sizeof(line), "%s %.3f %.3f\n", t.name, (double) b * 1e3, (double) e * 1e3);
            text += line;
        }
        size_t const n = std::min<size_t>(text.size(), (size_t) cap - 1);
        std::memcpy(out, text.data(), n);
        out[n] = '\0';
    });
}

int fgpu_ctx_profile(fgpu_ctx* ctx, int enable)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        ctx->profile = enable != 0;
    });
}

int fgpu_ctx_kernel_time(fgpu_ctx* ctx, const char* prefix, double* ms_out, uint64_t* launches_out, int reset)
{
    return guarded([&] {
        require(ctx != nullptr, FGPU_EINVALID, "null argument");
        bind_device(ctx);
        sync(ctx);
        size_t const plen = prefix != nullptr ? std::strlen(prefix) : 0;
        double ms = 0.0;
        uint64_t launches = 0;
        for (const auto& t : ctx->timers)
        {
            if (plen == 0 || std::strncmp(t.name, prefix, plen) == 0)
            {
                float e = 0.0f;
                FGPU_CUDA_CHECK(cudaEventElapsedTime(&e, t.begin, t.end));
                ms += e;
                launches += 1;
            }
        }
        if (ms_out != nullptr)
        {
            *ms_out = ms;
        }
        if (launches_out != nullptr)
        {
            *launches_out = launches;
        }
        if (reset)
        {
            for (auto& t : ctx->timers)
            {
                ctx->event_pool.push_back(t.begin);
                ctx->event_pool.push_back(t.end);
            }
            ctx->timers.clear();
        }
    });
}

// ---- points ----------------------------------------------------------------------------------------
static void points_create_impl(fgpu_ctx* ctx, const float* box6, int is2d, const float* host, const float* dev,
                               uint32_t n, fgpu_points** out, fgpu_comm* comm = nullptr)
{
    require(ctx != nullptr && box6 != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
    // NeighborQuery ctor, NeighborQuery.h:97-100
    require(n != 0, FGPU_EINVALID, "Cannot create a NeighborQuery with 0 particles.");
    require(host != nullptr || dev != nullptr, FGPU_EINVALID, "null points");
    bind_device(ctx);
    std::unique_ptr<fgpu_points> p(new fgpu_points());
    p->ctx = ctx;
    p->box = make_box(box6, is2d);
    require(p->box.Lx > 0 && p->box.Ly > 0 && (p->box.is2d || p->box.Lz > 0), FGPU_EINVALID,
            "box lengths must be positive");
    plane_distances(p->box, p->plane_dist);
    p->n = n;
    p->xyz.reserve((size_t) n * 3);
    if (host != nullptr && comm != nullptr && comm->size > 1)
    {
        // Every rank holds the same host array: rank r sends only rows [r n / W, (r + 1) n / W) over its own PCIe
        // link and the ranks hand each other their blocks over NVLink (one grouped ncclBroadcast per block), instead
        // of W copies of the whole frame crossing PCIe (the e2e of configs[3] was upload-bound).
        require(comm->ctx == ctx, FGPU_EINVALID, "comm belongs to a different context");
        NcclApi& api = nccl_or_throw();
        uint64_t const W = (uint64_t) comm->size;
        auto lo_of = [&](uint64_t r) { return (uint64_t) n * r / W; };
        uint64_t const lo = lo_of((uint64_t) comm->rank), hi = lo_of((uint64_t) comm->rank + 1);
        h2d(ctx, p->xyz.ptr + 3 * lo, host + 3 * lo, (size_t) (hi - lo) * 3 * sizeof(float));
        nccl_check(api.GroupStart(), "ncclGroupStart");
        for (uint64_t r = 0; r < W; ++r)
        {
            uint64_t const a = lo_of(r), b = lo_of(r + 1);
            if (b > a)
            {
                nccl_check(api.Broadcast(p->xyz.ptr + 3 * a, p->xyz.ptr + 3 * a, (size_t) (b - a) * 3, ncclFloat32, (int) r,
                                         static_cast<ncclComm_t>(comm->nccl_comm), ctx->stream),
                           "ncclBroadcast(points block)");
            }
        }
        nccl_check(api.GroupEnd(), "ncclGroupEnd");
    }
    else if (host != nullptr)
    {
        h2d(ctx, p->xyz.ptr, host, (size_t) n * 3 * sizeof(float));
    }
    else
    {
        FGPU_CUDA_CHECK(cudaMemcpyAsync(p->xyz.ptr, dev, (size_t) n * 3 * sizeof(float), cudaMemcpyDeviceToDevice,
                                        ctx->stream));
    }
    if (p->box.is2d)
    {
        // NeighborQuery.h:103-112, on the device behind the upload
        int* flag = reinterpret_cast<int*>(ctx->d_scalars + 7);
        FGPU_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(unsigned long long), ctx->stream));
        launch_check_2d_z(ctx, p->xyz.ptr, n, flag);
        d2h(ctx, ctx->h_scalars + 7, ctx->d_scalars + 7, sizeof(unsigned long long));
        sync(ctx);
        require((ctx->h_scalars[7] & 0xffffffffULL) == 0, FGPU_EINVALID,
                "A point with z != 0 was provided in a 2D box.");
    }
    else if (host != nullptr)
    {
        sync(ctx); // the reference copies at construction; the caller may free its buffer right away
    }
    *out = p.release();
}

int fgpu_points_create(fgpu_ctx* ctx, const float* box6, int is2d, const float* points_host, uint32_t n,
                       fgpu_points** out)
{
    return guarded([&] { points_create_impl(ctx, box6, is2d, points_host, nullptr, n, out); });
}

int fgpu_points_create_dev(fgpu_ctx* ctx, const float* box6, int is2d, const float* points_dev, uint32_t n,
                           fgpu_points** out)
{
    return guarded([&] { points_create_impl(ctx, box6, is2d, nullptr, points_dev, n, out); });
}

int fgpu_points_create_replicated(fgpu_ctx* ctx, fgpu_comm* comm, const float* box6, int is2d, const float* points_host,
                                  uint32_t n, fgpu_points** out)
{
    return guarded([&] {
        require(comm != nullptr, FGPU_EINVALID, "null argument");
        points_create_impl(ctx, box6, is2d, points_host, nullptr, n, out, comm);
    });
}

void fgpu_points_destroy(fgpu_points* pts)
{
    if (pts != nullptr)
    {
        bind_quiet(pts->ctx); // frees are stream-ordered behind any pending work
        delete pts;
    }
}

int fgpu_points_build_cells(fgpu_points* pts, float r_search, uint32_t* out_dims)
{
    return guarded([&] {
        require(pts != nullptr, FGPU_EINVALID, "null argument");
        require(r_search > 0, FGPU_EINVALID, "r_search must be positive");
        bind_device(pts->ctx);
        pts->grid.r_search = -1.0f; // force a rebuild: this entry point exists to time/test the build
        build_grid(pts, r_search);
        if (out_dims != nullptr)
        {
            for (int d = 0; d < 3; ++d)
            {
                out_dims[d] = (uint32_t) pts->grid.dim[d];
            }
        }
    });
}

int fgpu_points_set_shard(fgpu_points* pts, int shard, int n_shards)
{
    return guarded([&] {
        require(pts != nullptr, FGPU_EINVALID, "null argument");
        require(n_shards >= 1 && shard >= 0 && shard < n_shards, FGPU_EINVALID, "shard index out of range");
        pts->shard = shard;
        pts->n_shards = n_shards; // the next build_grid sees the change and rebuilds
    });
}

int fgpu_shard_plan(const uint32_t* dims, uint32_t n_points, int shard, int n_shards, uint32_t* out)
{
    return guarded([&] {
        require(dims != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        require(n_shards >= 1 && shard >= 0 && shard < n_shards, FGPU_EINVALID, "shard index out of range");
        require(dims[0] >= 1 && dims[1] >= 1 && dims[2] >= 1, FGPU_EINVALID, "empty grid");
        int const d[3] = {(int) dims[0], (int) dims[1], (int) dims[2]};
        ShardPlan const sp = shard_plan(d, n_points, shard, n_shards);
        ShardPlan const last = shard_plan(d, n_points, n_shards - 1, n_shards);
        out[0] = sp.ticket_begin;
        out[1] = sp.ticket_end;
        out[2] = last.ticket_end;
        out[3] = sp.cell_begin;
        out[4] = sp.cell_end;
        out[5] = (uint32_t) sp.slab_axis;
        out[6] = (uint32_t) sp.slab_lo;
        out[7] = sp.slab_len < 0 ? 0xffffffffU : (uint32_t) sp.slab_len;
    });
}

int fgpu_points_read_cells(fgpu_points* pts, uint32_t* cell_start_host, uint32_t* order_host)
{
    return guarded([&] {
        require(pts != nullptr, FGPU_EINVALID, "null argument");
        require(pts->grid.r_search >= 0, FGPU_ERUNTIME, "cell list not built yet");
        fgpu_ctx* ctx = pts->ctx;
        bind_device(ctx);
        if (cell_start_host != nullptr)
        {
            d2h(ctx, cell_start_host, pts->grid.cell_start.ptr, ((size_t) pts->grid.n_cells + 1) * sizeof(uint32_t));
        }
        std::vector<float4> tmp;
        if (order_host != nullptr)
        {
            tmp.resize(pts->n);
            d2h(ctx, tmp.data(), pts->grid.sorted.ptr, (size_t) pts->n * sizeof(float4));
        }
        sync(ctx);
        if (order_host != nullptr)
        {
            for (uint32_t i = 0; i < pts->n; ++i)
            {
                std::memcpy(&order_host[i], &tmp[i].w, sizeof(uint32_t));
            }
        }
    });
}

// ---- ball query --------------------------------------------------------------------------------------
int fgpu_ball_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, uint32_t q_index_offset,
                    int flavour, float r_max, float r_min, int exclude_ii, int sort_by_distance, fgpu_nlist** out)
{
    return guarded([&] {
        ball_query_impl(pts, query_points_host, nullptr, n_query, q_index_offset, flavour, r_max, r_min, exclude_ii,
                        sort_by_distance, out);
    });
}

int fgpu_ball_query_dev(fgpu_points* pts, const float* query_points_dev, uint32_t n_query, uint32_t q_index_offset,
                        int flavour, float r_max, float r_min, int exclude_ii, int sort_by_distance,
                        fgpu_nlist** out)
{
    return guarded([&] {
        ball_query_impl(pts, nullptr, query_points_dev, n_query, q_index_offset, flavour, r_max, r_min, exclude_ii,
                        sort_by_distance, out);
    });
}

// ---- kNN ---------------------------------------------------------------------------------------------
static constexpr double kKnnWindowFill = 1.5;

// bag_consumer: a compute that only needs the k nearest bond vectors of every row (Steinhardt without a NeighborList)
// takes the window search's bag as it is -- rows grouped, unsorted -- instead of the selected, sorted list; it is
// called once, when the warp-cooperative search has resolved every row, and *out then stays null.  When the frame
// goes through the general kernels instead it is not called and *out is the list, as always.
static void knn_query_body(fgpu_points* pts, const float* query_points_host, uint32_t n_query, uint32_t q_index_offset,
                           int flavour, uint32_t num_neighbors, float r_max, float r_min, int exclude_ii,
                           int sort_by_distance, fgpu_nlist** out,
                           const std::function<void(const KnnSelectArgs&)>* bag_consumer)
{
    {
        require(pts != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        // CellQuery::validateQueryArgs, CellQuery.h:180-184
        require(flavour != FGPU_FLAVOUR_GHOST, FGPU_ERUNTIME,
                "CellQuery only supports ball queries (r_max), not nearest queries (num_neighbors). Use AABBQuery "
                "for nearest neighbor queries.");
        require(flavour == FGPU_FLAVOUR_WRAP || flavour == FGPU_FLAVOUR_IMAGE, FGPU_EINVALID, "unknown flavour");
        bool const wrap = flavour == FGPU_FLAVOUR_WRAP;
        fgpu_ctx* ctx = pts->ctx;
        bind_device(ctx);
        require(r_max > 0, FGPU_EINVALID, "NeighborQuery requires r_max to be positive.");
        require(r_max > r_min, FGPU_EINVALID, "NeighborQuery requires that r_max must be greater than r_min.");
        require(num_neighbors != 0xffffffffU, FGPU_ERUNTIME,
                "You must set num_neighbors in the query arguments when performing number of neighbor queries.");
        bool const self = query_points_host == nullptr;
        require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
        require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
        auto nl = new_nlist(ctx, n_query, pts->n);
        nl->q_index_offset = self ? 0 : q_index_offset;
        uint32_t const k = std::min<uint32_t>(num_neighbors, pts->n);
        if (n_query == 0 || k == 0)
        {
            alloc_bonds(nl.get(), 0);
            if (n_query != 0)
            {
                FGPU_CUDA_CHECK(cudaMemsetAsync(nl->counts.ptr, 0, (size_t) n_query * sizeof(uint32_t), ctx->stream));
                FGPU_CUDA_CHECK(
                    cudaMemsetAsync(nl->row_start.ptr, 0, ((size_t) n_query + 1) * sizeof(uint32_t), ctx->stream));
                FGPU_CUDA_CHECK(cudaMemsetAsync(nl->segments.ptr, 0, (size_t) n_query * sizeof(uint32_t), ctx->stream));
                sync(ctx);
            }
            *out = nl.release();
            return;
        }
        ctx->knn_d.reserve((size_t) n_query * k);
        ctx->knn_s.reserve((size_t) n_query * k);
        ctx->row_counts.reserve((size_t) n_query + 1);

        // initial radius: the sphere expected to hold 1.5 (k + 1) points at the mean density (the reference's own
        // r_guess is the k-point sphere, NeighborQuery.h:251-253).  In an ideal gas ~3 % of the rows then hold
        // fewer than k points; they are searched again alone, which is cheaper than a wider window for everybody.
        double const volume = box_volume(pts->box);
        double const density = (double) pts->n / volume;
        double r_search = pts->box.is2d ? std::sqrt(kKnnWindowFill * (k + 1) / (M_PI * density))
                                        : std::cbrt(3.0 * kKnnWindowFill * (k + 1) / (4.0 * M_PI * density));
        QueryView qv;
        bool evals_counted = false; // by the count kernel of the warp-cooperative path
        float r_grid_done = 0.0f;   // grid radius the last warp-cooperative search ran at
        for (int attempt = 0;; ++attempt)
        {
            require(attempt < 64, FGPU_ERUNTIME, "kNN search did not converge");
            float const r_grid = (float) std::min(r_search, (double) r_max * 1.0001);
            build_grid(pts, r_grid);
            const fgpu_grid& g = pts->grid;
            bool const cover_all = g.dim[0] < 3 && g.dim[1] < 3 && (pts->box.is2d || g.dim[2] < 3);
            qv = prepare_queries(pts, query_points_host, nullptr, n_query);

            // ---- production path: warp-cooperative ball search at the window radius, selection per row -----
            // (regular grids with every point inside the box; the thread-per-query kernel below handles the rest)
            float const r_win = r_grid < r_max ? r_grid : r_max;
            // r_min: LinkCell compares squares (LinkCell.cc:619), AABBQuery the distance itself (AABBQuery.cc:213)
            Search2Args s2 = base_search2_args(pts, qv, q_index_offset, r_win, wrap ? r_min : 0.0f, exclude_ii);
            s2.knn_r_min = !wrap && r_min > 0.0f ? r_min : 0.0f;
            search2_choose_mapping(s2, flavour, n_query, pts->n, expected_hits_per_query(pts, r_win),
                                   ctx->tune_lanes_over_queries);
            if (!ctx->force_general && !cover_all && search2_supported(s2, S2_NL) && pts->n < 0x7fffffffU)
            {
                ctx->tmp_start.reserve((size_t) n_query + 1);
                ctx->knn_hits.reserve((size_t) n_query + 1);
                ctx->knn_unresolved.reserve(2 * (size_t) n_query + 2);
                if (attempt == 0)
                {
                    launch_count_evals(ctx, s2, n_query, pts->grid.cell_of.ptr, pts->n);
                }
                enum WindowResult
                {
                    WINDOW_DONE,
                    WINDOW_GENERAL
                };
                // One search of `rows` query points at the window of `args` into `bag`, then the row bookkeeping
                // over all n_query rows, enqueued behind it: one host round trip per search.
                auto run_window = [&](Search2Args& args, DevBuf<float4>& bag, uint64_t rows, float window,
                                      uint32_t* unresolved_rows) {
                    bool const final_window = !(window < r_max);
                    double const shell = pts->box.is2d ? M_PI * (double) window * window
                                                       : 4.0 / 3.0 * M_PI * (double) window * window * window;
                    uint64_t cap = (uint64_t) (1.25 * (double) rows * density * shell) + 4096;
                    for (int pass = 0; pass < 4; ++pass)
                    {
                        cap = std::min<uint64_t>(cap, 0x7fffffffULL); // the top bit of a bag offset names the bag
                        bag.reserve(cap);
                        args.bag = bag.ptr;
                        args.temp_cap = (uint32_t) cap;
                        args.counts = ctx->knn_hits.ptr;
                        args.tmp_start = ctx->tmp_start.ptr;
                        FGPU_CUDA_CHECK(
                            cudaMemsetAsync(ctx->d_scalars + 2, 0, 5 * sizeof(unsigned long long), ctx->stream));
                        launch_search2(ctx, flavour, S2_NL, args);
                        KnnRowsArgs ra;
                        ra.hits = ctx->knn_hits.ptr;
                        ra.n_query = n_query;
                        ra.k = k;
                        ra.final = final_window ? 1 : 0;
                        ra.counts = nl->counts.ptr;
                        ra.row_start = nl->row_start.ptr;
                        ra.unresolved = ctx->d_scalars + 2;
                        ra.total = ctx->d_scalars + 3;
                        ra.unresolved_rows = unresolved_rows;
                        launch_knn_rows(ctx, ra);
                        if (bag_consumer == nullptr)
                        {
                            // offsets of the NeighborList's rows; a compute that reads the bag itself has no use for them
                            FGPU_CUDA_CHECK(cudaMemsetAsync(nl->row_start.ptr + n_query, 0, sizeof(uint32_t), ctx->stream));
                            exclusive_scan_u32(ctx, nl->row_start.ptr, (size_t) n_query + 1);
                        }
                        d2h(ctx, ctx->h_scalars + 2, ctx->d_scalars + 2, 4 * sizeof(unsigned long long));
                        sync(ctx);
                        int const failed = (int) (ctx->h_scalars[4] & 0xffffffffULL);
                        uint32_t const densest = (uint32_t) (ctx->h_scalars[4] >> 32);
                        if (search2_lq_fallback(args, failed))
                        {
                            continue; // rows beyond the buffers of the lanes-over-queries search: the tile walk
                        }
                        if (failed == 2 && densest <= search2_max_out_cap() && args.out_cap < search2_max_out_cap())
                        {
                            args.out_cap = std::min(search2_max_out_cap(), (densest + densest / 8 + 31U) & ~31U);
                            continue; // a denser tile than the uniform estimate: retry with a buffer that holds it
                        }
                        if (failed != 0)
                        {
                            return WINDOW_GENERAL; // 1: points outside the box, 2: a tile beyond the largest buffer
                        }
                        if (ctx->h_scalars[5] <= cap)
                        {
                            return WINDOW_DONE;
                        }
                        cap = ctx->h_scalars[5]; // exact size, the search is deterministic
                    }
                    return WINDOW_GENERAL;
                };
                WindowResult res = run_window(s2, ctx->bag4, n_query, r_win, ctx->knn_unresolved.ptr);
                int const fail = (int) (ctx->h_scalars[4] & 0xffffffffULL);
                evals_counted = evals_counted
                    || (s2.evals != nullptr && attempt == 0 && (res == WINDOW_DONE || fail != 1));
                uint64_t n_short = res == WINDOW_DONE ? ctx->h_scalars[2] : 0;
                bool second_bag = false;
                if (res == WINDOW_DONE && n_short != 0 && n_short <= (uint64_t) n_query / 16 + 1)
                {
                    // A few rows hold fewer than k points: search those again -- only those -- with a window 1.5x
                    // wider on a coarser grid, into a second bag.  The first bag stays valid: its records carry
                    // the bond vectors, not references into the old grid.
                    float const r_grid2 = (float) std::min((double) r_grid * 1.5, (double) r_max * 1.0001);
                    build_grid(pts, r_grid2);
                    const fgpu_grid& g2 = pts->grid;
                    bool const cover_all2 = g2.dim[0] < 3 && g2.dim[1] < 3 && (pts->box.is2d || g2.dim[2] < 3);
                    ctx->knn_subset.reserve(3 * (size_t) n_short);
                    launch_gather_points(ctx, qv.xyz, ctx->knn_unresolved.ptr, (uint32_t) n_short,
                                         ctx->knn_subset.ptr);
                    sort_queries(pts, ctx->knn_subset.ptr, (uint32_t) n_short);
                    QueryView sub;
                    sub.sorted = ctx->q_sorted.ptr;
                    sub.xyz = ctx->knn_subset.ptr;
                    sub.cell_start = ctx->q_cell_start.ptr;
                    sub.outside_flag = ctx->q_outside_flag.ptr;
                    float const r_win2 = r_grid2 < r_max ? r_grid2 : r_max;
                    Search2Args s2b = base_search2_args(pts, sub, q_index_offset, r_win2, wrap ? r_min : 0.0f, exclude_ii);
                    s2b.knn_r_min = s2.knn_r_min;
                    s2b.q_remap = ctx->knn_unresolved.ptr;
                    s2b.tmp_flag = kSecondBag;
                    search2_choose_mapping(s2b, flavour, (uint32_t) n_short, pts->n, expected_hits_per_query(pts, r_win2),
                                           ctx->tune_lanes_over_queries);
                    if (!cover_all2 && search2_supported(s2b, S2_NL)
                        && run_window(s2b, ctx->bag4b, n_short, r_win2, ctx->knn_unresolved.ptr + n_query)
                            == WINDOW_DONE)
                    {
                        second_bag = true;
                        n_short = ctx->h_scalars[2];
                    }
                    r_grid_done = r_grid2;
                }
                else
                {
                    r_grid_done = r_grid;
                }
                if (res == WINDOW_DONE)
                {
                    if (n_short != 0)
                    {
                        r_search = (double) r_grid_done * 1.5; // many short rows: widen the window of the whole frame
                        continue;
                    }
                    uint64_t const n_bonds = ctx->h_scalars[3];
                    if (bag_consumer != nullptr)
                    {
                        KnnSelectArgs sa;
                        std::memset(&sa, 0, sizeof(sa));
                        sa.bag = ctx->bag4.ptr;
                        sa.bag2 = second_bag ? ctx->bag4b.ptr : nullptr;
                        sa.tmp_start = ctx->tmp_start.ptr;
                        sa.hits = ctx->knn_hits.ptr;
                        sa.n_query = n_query;
                        sa.k = k;
                        (*bag_consumer)(sa);
                        *out = nullptr;
                        return;
                    }
                    alloc_bonds(nl.get(), n_bonds);
                    if (n_bonds != 0)
                    {
                        KnnSelectArgs sa;
                        sa.bag = ctx->bag4.ptr;
                        sa.bag2 = second_bag ? ctx->bag4b.ptr : nullptr;
                        sa.tmp_start = ctx->tmp_start.ptr;
                        sa.hits = ctx->knn_hits.ptr;
                        sa.row_start = nl->row_start.ptr;
                        sa.n_query = n_query;
                        sa.k = k;
                        sa.neighbors = nl->neighbors.ptr;
                        sa.distances = nl->distances.ptr;
                        sa.weights = nl->weights.ptr;
                        sa.vectors = nl->vectors.ptr;
                        launch_knn_select(ctx, sort_by_distance, sa);
                    }
                    launch_segments(ctx, nl->row_start.ptr, nl->counts.ptr, nl->segments.ptr, n_query);
                    // no final sync: consumers are ordered on the same stream; the host query buffer was
                    // consumed before the totals were read back above
                    *out = nl.release();
                    return;
                }
            }
            KnnArgs a;
            std::memset(&a, 0, sizeof(a));
            a.box = pts->box;
            a.grid = grid_dev(pts);
            a.flavour = flavour;
            a.q_sorted = qv.sorted;
            a.n_query = n_query;
            a.q_index_offset = q_index_offset;
            a.k = k;
            a.r_max = r_max;
            a.r_min = r_min;
            a.r_safe = r_grid;
            a.exclude_ii = exclude_ii != 0;
            a.cover_all = cover_all ? 1 : 0;
            a.knn_d = ctx->knn_d.ptr;
            a.knn_s = ctx->knn_s.ptr;
            a.row_counts = ctx->row_counts.ptr;
            a.unresolved = ctx->d_scalars + 2;
            a.total = ctx->d_scalars + 3;
            a.evals = ctx->count_evals && !evals_counted ? ctx->d_evals : nullptr;
            FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 2, 0, 2 * sizeof(unsigned long long), ctx->stream));
            launch_knn(ctx, a);
            d2h(ctx, ctx->h_scalars + 2, ctx->d_scalars + 2, 2 * sizeof(unsigned long long));
            sync(ctx);
            if (ctx->h_scalars[2] == 0 || cover_all)
            {
                break; // the total (d_scalars[3]) is read back by finish_counts below
            }
            r_search = (double) r_grid * 1.5;
        }
        FGPU_CUDA_CHECK(cudaMemcpyAsync(ctx->d_scalars, ctx->d_scalars + 3, sizeof(unsigned long long),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        uint64_t const n_bonds = finish_counts(ctx, nl.get(), ctx->row_counts.ptr, ctx->d_scalars);
        alloc_bonds(nl.get(), n_bonds);
        if (n_bonds != 0)
        {
            KnnEmitArgs e;
            e.box = pts->box;
            e.flavour = flavour;
            e.sorted = pts->grid.sorted.ptr;
            e.q_xyz = qv.xyz;
            e.knn_d = ctx->knn_d.ptr;
            e.knn_s = ctx->knn_s.ptr;
            e.row_start = nl->row_start.ptr;
            e.row_counts = nl->counts.ptr;
            e.n_query = n_query;
            e.k = k;
            e.sort_by_distance = sort_by_distance != 0;
            e.neighbors = nl->neighbors.ptr;
            e.distances = nl->distances.ptr;
            e.weights = nl->weights.ptr;
            e.vectors = nl->vectors.ptr;
            launch_knn_emit(ctx, e);
        }
        launch_segments(ctx, nl->row_start.ptr, nl->counts.ptr, nl->segments.ptr, n_query);
        sync(ctx);
        *out = nl.release();
    }
}

int fgpu_knn_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, uint32_t q_index_offset,
                   int flavour, uint32_t num_neighbors, float r_max, float r_min, int exclude_ii,
                   int sort_by_distance, fgpu_nlist** out)
{
    return guarded([&] {
        knn_query_body(pts, query_points_host, n_query, q_index_offset, flavour, num_neighbors, r_max, r_min, exclude_ii,
                       sort_by_distance, out, nullptr);
    });
}

// ---- NeighborList -------------------------------------------------------------------------------------
uint64_t fgpu_nlist_num_bonds(const fgpu_nlist* nl)
{
    return nl != nullptr ? nl->n_bonds : 0;
}

uint32_t fgpu_nlist_num_query_points(const fgpu_nlist* nl)
{
    return nl != nullptr ? nl->n_query : 0;
}

uint32_t fgpu_nlist_num_points(const fgpu_nlist* nl)
{
    return nl != nullptr ? nl->n_points : 0;
}

namespace {

void fill_unit_weights(float* weights_host, size_t nb)
{
    unsigned const n_threads = nb >= (1U << 20) ? std::max(1U, std::min(4U, std::thread::hardware_concurrency())) : 1U;
    auto fill = [&](size_t lo, size_t hi) { std::fill(weights_host + lo, weights_host + hi, 1.0f); };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; ++t)
    {
        pool.emplace_back(fill, nb * t / n_threads, nb * (t + 1) / n_threads);
    }
    fill(0, nb / n_threads);
    for (auto& th : pool)
    {
        th.join();
    }
}

} // namespace

int fgpu_nlist_copy(const fgpu_nlist* nl, uint32_t* neighbors_host, float* distances_host, float* weights_host,
                    float* vectors_host, uint32_t* segments_host, uint32_t* counts_host)
{
    return guarded([&] {
        require(nl != nullptr, FGPU_EINVALID, "null argument");
        fgpu_ctx* ctx = nl->ctx;
        bind_device(ctx);
        size_t const nb = nl->n_bonds;
        if (neighbors_host != nullptr)
        {
            d2h(ctx, neighbors_host, nl->neighbors.ptr, nb * 2 * sizeof(uint32_t));
        }
        if (distances_host != nullptr)
        {
            d2h(ctx, distances_host, nl->distances.ptr, nb * sizeof(float));
        }
        bool const fill_weights = weights_host != nullptr && nl->unit_weights;
        if (weights_host != nullptr && !fill_weights)
        {
            d2h(ctx, weights_host, nl->weights.ptr, nb * sizeof(float));
        }
        if (vectors_host != nullptr)
        {
            d2h(ctx, vectors_host, nl->vectors.ptr, nb * 3 * sizeof(float));
        }
        if (segments_host != nullptr)
        {
            d2h(ctx, segments_host, nl->segments.ptr, (size_t) nl->n_query * sizeof(uint32_t));
        }
        if (counts_host != nullptr)
        {
            d2h(ctx, counts_host, nl->counts.ptr, (size_t) nl->n_query * sizeof(uint32_t));
        }
        if (fill_weights)
        {
            // A list built by a query carries weight 1 on every bond: write the ones here, on host threads, while
            // the other arrays cross PCIe (the copy of a frame's list is bound by that link) instead of sending a
            // seventh of the bytes for a constant.
            fill_unit_weights(weights_host, nb);
        }
        sync(ctx);
    });
}

int fgpu_nlist_copy_begin(const fgpu_nlist* nl, uint32_t* neighbors_host, float* distances_host, float* weights_host,
                          float* vectors_host)
{
    return guarded([&] {
        require(nl != nullptr, FGPU_EINVALID, "null argument");
        fgpu_ctx* ctx = nl->ctx;
        bind_device(ctx);
        size_t const nb = nl->n_bonds;
        void* const dst[4] = {neighbors_host, distances_host, weights_host, vectors_host};
        const void* const src[4] = {nl->neighbors.ptr, nl->distances.ptr, nl->weights.ptr, nl->vectors.ptr};
        size_t const bytes[4] = {nb * 2 * sizeof(uint32_t), nb * sizeof(float), nb * sizeof(float), nb * 3 * sizeof(float)};
        for (int k = 0; k < 4; ++k)
        {
            if (dst[k] == nullptr)
            {
                continue;
            }
            if (nl->copy_done[k] == nullptr)
            {
                FGPU_CUDA_CHECK(cudaEventCreateWithFlags(&nl->copy_done[k], cudaEventDisableTiming));
            }
            if (k == 2 && nl->unit_weights)
            {
                nl->pending_unit_weights = weights_host; // a constant: written by the host when somebody waits for it
            }
            else
            {
                d2h(ctx, dst[k], src[k], bytes[k]);
            }
            FGPU_CUDA_CHECK(cudaEventRecord(nl->copy_done[k], ctx->stream));
        }
    });
}

int fgpu_nlist_copy_wait(const fgpu_nlist* nl, unsigned which)
{
    return guarded([&] {
        require(nl != nullptr, FGPU_EINVALID, "null argument");
        bind_device(nl->ctx);
        for (int k = 0; k < 4; ++k)
        {
            if ((which & (1U << k)) == 0 || nl->copy_done[k] == nullptr)
            {
                continue;
            }
            if (k == 2 && nl->pending_unit_weights != nullptr)
            {
                fill_unit_weights(nl->pending_unit_weights, nl->n_bonds);
                nl->pending_unit_weights = nullptr;
            }
            FGPU_CUDA_CHECK(cudaEventSynchronize(nl->copy_done[k]));
        }
    });
}

int fgpu_nlist_from_host(fgpu_ctx* ctx, uint64_t n_bonds, uint32_t n_query, uint32_t n_points,
                         const uint32_t* neighbors_host, const float* distances_host, const float* weights_host,
                         const float* vectors_host, fgpu_nlist** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        require(n_bonds == 0 || (neighbors_host != nullptr && distances_host != nullptr), FGPU_EINVALID,
                "neighbors and distances are required");
        bind_device(ctx);
        auto nl = new_nlist(ctx, n_query, n_points);
        alloc_bonds(nl.get(), n_bonds);
        // row structure on the host (the list is sorted by query index, NeighborList::validate)
        std::vector<uint32_t> counts((size_t) n_query + 1, 0), row_start((size_t) n_query + 1, 0),
            segments((size_t) n_query + 1, 0);
        uint32_t last = 0;
        for (uint64_t b = 0; b < n_bonds; ++b)
        {
            uint32_t const i = neighbors_host[2 * b];
            require(i < n_query && neighbors_host[2 * b + 1] < n_points, FGPU_EINVALID,
                    "NeighborList index out of range");
            require(i >= last, FGPU_EINVALID, "NeighborList must be sorted by query point index");
            last = i;
            counts[i] += 1;
        }
        uint32_t acc = 0;
        for (uint32_t i = 0; i < n_query; ++i)
        {
            row_start[i] = acc;
            segments[i] = counts[i] != 0 ? acc : 0;
            acc += counts[i];
        }
        row_start[n_query] = acc;
        std::vector<float> ones;
        h2d(ctx, nl->neighbors.ptr, neighbors_host, n_bonds * 2 * sizeof(uint32_t));
        h2d(ctx, nl->distances.ptr, distances_host, n_bonds * sizeof(float));
        nl->unit_weights = weights_host == nullptr;
        if (weights_host == nullptr)
        {
            ones.assign(n_bonds, 1.0f);
            weights_host = ones.data();
        }
        h2d(ctx, nl->weights.ptr, weights_host, n_bonds * sizeof(float));
        if (vectors_host != nullptr)
        {
            h2d(ctx, nl->vectors.ptr, vectors_host, n_bonds * 3 * sizeof(float));
        }
        else if (n_bonds != 0)
        {
            FGPU_CUDA_CHECK(cudaMemsetAsync(nl->vectors.ptr, 0, n_bonds * 3 * sizeof(float), ctx->stream));
        }
        h2d(ctx, nl->counts.ptr, counts.data(), (size_t) n_query * sizeof(uint32_t));
        h2d(ctx, nl->row_start.ptr, row_start.data(), ((size_t) n_query + 1) * sizeof(uint32_t));
        h2d(ctx, nl->segments.ptr, segments.data(), (size_t) n_query * sizeof(uint32_t));
        sync(ctx);
        *out = nl.release();
    });
}

void fgpu_nlist_destroy(fgpu_nlist* nl)
{
    if (nl != nullptr)
    {
        bind_quiet(nl->ctx);
        for (cudaEvent_t& e : nl->copy_done)
        {
            if (e != nullptr)
            {
                cudaEventSynchronize(e); // a copy begun by fgpu_nlist_copy_begin still owns its host destination
                cudaEventDestroy(e);
                e = nullptr;
            }
        }
        if (nl->bytes() > nl->ctx->spare_nlist.bytes())
        {
            nl->swap(nl->ctx->spare_nlist); // keep the larger set for the next list, free the smaller one
        }
        delete nl;
    }
}

// ---- RDF ----------------------------------------------------------------------------------------------
int fgpu_rdf_create(fgpu_ctx* ctx, uint32_t bins, float r_max, float r_min, fgpu_rdf** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        // RDF::RDF, RDF.cc:27-42
        require(bins != 0, FGPU_EINVALID, "RDF requires a nonzero number of bins.");
        require(r_max > 0, FGPU_EINVALID, "RDF requires r_max to be positive.");
        require(r_min >= 0, FGPU_EINVALID, "RDF requires r_min to be non-negative.");
        require(r_max > r_min, FGPU_EINVALID, "RDF requires that r_max must be greater than r_min.");
        bind_device(ctx);
        std::unique_ptr<fgpu_rdf> r(new fgpu_rdf());
        r->ctx = ctx;
        // RegularAxis ctor, Histogram.h:126-138
        volatile float span = r_max - r_min;
        volatile float width = span / (float) bins;
        volatile float inv = 1.0f / width;
        r->axis.r_min = r_min;
        r->axis.r_max = r_max;
        r->axis.inv_width = inv;
        r->axis.bins = bins;
        // one word behind the counters: sticky "a sharded accumulation could not run" flag, checked by read()
        r->hist.reserve((size_t) bins + 1);
        FGPU_CUDA_CHECK(cudaMemsetAsync(r->hist.ptr, 0, ((size_t) bins + 1) * sizeof(uint32_t), ctx->stream));
        sync(ctx);
        *out = r.release();
    });
}

void fgpu_rdf_destroy(fgpu_rdf* rdf)
{
    if (rdf != nullptr)
    {
        bind_quiet(rdf->ctx);
        if (rdf->peer != nullptr)
        {
            cudaStreamSynchronize(rdf->ctx->stream);
            for (void* p : rdf->peer->opened)
            {
                if (p != nullptr)
                {
                    cudaIpcCloseMemHandle(p);
                }
            }
            cudaFree(rdf->peer->mailbox);
            delete rdf->peer;
        }
        delete rdf;
    }
}

int fgpu_rdf_reset(fgpu_rdf* rdf)
{
    return guarded([&] {
        require(rdf != nullptr, FGPU_EINVALID, "null argument");
        bind_device(rdf->ctx);
        rdf->reduced_valid = false;
        FGPU_CUDA_CHECK(
            cudaMemsetAsync(rdf->hist.ptr, 0, ((size_t) rdf->axis.bins + 1) * sizeof(uint32_t), rdf->ctx->stream));
    });
}

int fgpu_rdf_accumulate(fgpu_rdf* rdf, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                        uint32_t q_index_offset, int flavour, float q_r_max, float q_r_min, int exclude_ii)
{
    return guarded([&] {
        rdf_accumulate_impl(rdf, pts, query_points_host, nullptr, n_query, q_index_offset, flavour, q_r_max, q_r_min,
                            exclude_ii);
    });
}

int fgpu_rdf_accumulate_dev(fgpu_rdf* rdf, fgpu_points* pts, const float* query_points_dev, uint32_t n_query,
                            uint32_t q_index_offset, int flavour, float q_r_max, float q_r_min, int exclude_ii)
{
    return guarded([&] {
        rdf_accumulate_impl(rdf, pts, nullptr, query_points_dev, n_query, q_index_offset, flavour, q_r_max, q_r_min,
                            exclude_ii);
    });
}

int fgpu_rdf_accumulate_nlist(fgpu_rdf* rdf, const fgpu_nlist* nl)
{
    return guarded([&] {
        require(rdf != nullptr && nl != nullptr, FGPU_EINVALID, "null argument");
        require(rdf->ctx == nl->ctx, FGPU_EINVALID, "rdf and nlist belong to different contexts");
        bind_device(rdf->ctx);
        rdf->reduced_valid = false;
        launch_rdf_from_distances(rdf->ctx, nl->distances.ptr, nl->n_bonds, rdf->axis, rdf->hist.ptr);
    });
}

int fgpu_rdf_read(fgpu_rdf* rdf, uint32_t* counts_host)
{
    return guarded([&] {
        require(rdf != nullptr && counts_host != nullptr, FGPU_EINVALID, "null argument");
        bind_device(rdf->ctx);
        fgpu_ctx* ctx = rdf->ctx;
        // after fgpu_rdf_allreduce: the sum over the ranks; the rank's own counts stay in hist, so accumulating
        // further frames and reducing again never counts a frame twice
        d2h(ctx, counts_host, rdf->reduced_valid ? rdf->reduced.ptr : rdf->hist.ptr,
            (size_t) rdf->axis.bins * sizeof(uint32_t));
        d2h(ctx, ctx->h_scalars + 7, rdf->hist.ptr + rdf->axis.bins, sizeof(uint32_t));
        bool const peer_sum = rdf->reduced_valid && rdf->peer != nullptr;
        if (peer_sum)
        {
            d2h(ctx, ctx->h_scalars + 3, rdf->reduced.ptr + rdf->axis.bins, sizeof(uint32_t));
        }
        sync(ctx);
        require(!peer_sum || (ctx->h_scalars[3] & 0xffffffffULL) == 0, FGPU_ENCCL,
                "histogram reduction over peer memory: a rank did not arrive within 10 s");
        // raised on the device by a sharded accumulation (fgpu_points_set_shard), which has no general-kernel
        // fallback and no host round trip of its own
        require((ctx->h_scalars[7] & 0xffffffffULL) == 0, FGPU_ERUNTIME,
                "sharded RDF accumulation needs every point inside the box (wrap the points first)");
    });
}

namespace {

// Ends the current epoch of a peer-attached RDF: sends this rank's counts unless a fused kernel already did, waits
// (on the stream) for every rank's, leaves the sum in rdf->reduced.
void peer_reduce(fgpu_rdf* rdf, bool already_pushed)
{
    fgpu_ctx* ctx = rdf->ctx;
    fgpu_rdf_peer* pr = rdf->peer;
    PeerBox const pb = pr->box; // the epoch's parity lives in the mailbox: nothing here changes from call to call
    if (!already_pushed)
    {
        launch_rdf_push(ctx, pb, rdf->hist.ptr, rdf->axis.bins);
    }
    launch_rdf_wait(ctx, pb, rdf->axis.bins, rdf->reduced.ptr, reinterpret_cast<int*>(rdf->reduced.ptr + rdf->axis.bins));
    pr->epoch += 1;
    rdf->reduced_valid = true;
}

void nccl_reduce(fgpu_rdf* rdf, fgpu_comm* comm)
{
    NcclApi& api = nccl_or_throw();
    // out of place: hist keeps this rank's counts (compute(..., reset=False) goes on adding to them)
    rdf->reduced.reserve((size_t) rdf->axis.bins + 1);
    nccl_check(api.AllReduce(rdf->hist.ptr, rdf->reduced.ptr, rdf->axis.bins, ncclUint32, ncclSum,
                             static_cast<ncclComm_t>(comm->nccl_comm), rdf->ctx->stream),
               "ncclAllReduce(u32 histogram)");
    rdf->reduced_valid = true;
}

} // namespace

int fgpu_rdf_allreduce(fgpu_rdf* rdf, fgpu_comm* comm)
{
    return guarded([&] {
        require(rdf != nullptr && comm != nullptr, FGPU_EINVALID, "null argument");
        require(rdf->ctx == comm->ctx, FGPU_EINVALID, "rdf and comm belong to different contexts");
        bind_device(rdf->ctx);
        if (rdf->peer != nullptr && rdf->peer->comm == comm)
        {
            peer_reduce(rdf, false);
        }
        else
        {
            nccl_reduce(rdf, comm);
        }
    });
}

int fgpu_rdf_attach_comm(fgpu_rdf* rdf, fgpu_comm* comm)
{
    return guarded([&] {
        require(rdf != nullptr && comm != nullptr, FGPU_EINVALID, "null argument");
        require(rdf->ctx == comm->ctx, FGPU_EINVALID, "rdf and comm belong to different contexts");
        require(rdf->peer == nullptr, FGPU_ERUNTIME, "this RDF already has a peer mailbox");
        fgpu_ctx* ctx = rdf->ctx;
        bind_device(ctx);
        NcclApi& api = nccl_or_throw();
        int const world = comm->size;
        // Collective, so every rank must reach the same verdict: each one reports whether it could allocate and
        // export its mailbox, and later whether it could open every peer's; anything short of "all yes" leaves the
        // RDF on the NCCL route (returns 1: not an error, a different transport).
        uint32_t const bins_pad = (rdf->axis.bins + 31U) & ~31U;
        size_t const words = 2 * (size_t) bins_pad + 32;
        std::unique_ptr<fgpu_rdf_peer> pr(new fgpu_rdf_peer());
        cudaIpcMemHandle_t mine;
        std::memset(&mine, 0, sizeof(mine));
        bool ok = world <= kMaxPeers;
        if (ok)
        {
            ok = cudaMalloc(reinterpret_cast<void**>(&pr->mailbox), words * sizeof(uint32_t)) == cudaSuccess
                && cudaMemset(pr->mailbox, 0, words * sizeof(uint32_t)) == cudaSuccess
                && (world == 1 || cudaIpcGetMemHandle(&mine, pr->mailbox) == cudaSuccess);
            cudaGetLastError();
        }
        // exchange {ok, handle} of every rank
        constexpr size_t kRec = 128;
        static_assert(sizeof(cudaIpcMemHandle_t) + 4 <= kRec, "ipc record");
        std::vector<unsigned char> all((size_t) world * kRec, 0);
        unsigned char rec[kRec] = {};
        rec[0] = ok ? 1 : 0;
        std::memcpy(rec + 4, &mine, sizeof(mine));
        comm->stage.reserve((size_t) (world + 1) * kRec);
        h2d(ctx, comm->stage.ptr + (size_t) world * kRec, rec, kRec);
        nccl_check(api.AllGather(comm->stage.ptr + (size_t) world * kRec, comm->stage.ptr, kRec, ncclUint8,
                                 static_cast<ncclComm_t>(comm->nccl_comm), ctx->stream),
                   "ncclAllGather(ipc handles)");
        d2h(ctx, all.data(), comm->stage.ptr, (size_t) world * kRec);
        sync(ctx);
        for (int r = 0; r < world; ++r)
        {
            ok = ok && all[(size_t) r * kRec] == 1;
        }
        if (ok)
        {
            for (int r = 0; r < world && ok; ++r)
            {
                if (r == comm->rank)
                {
                    pr->box.box[r] = pr->mailbox;
                    continue;
                }
                cudaIpcMemHandle_t h;
                std::memcpy(&h, all.data() + (size_t) r * kRec + 4, sizeof(h));
                void* p = nullptr;
                ok = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                cudaGetLastError();
                pr->opened[r] = ok ? p : nullptr;
                pr->box.box[r] = static_cast<uint32_t*>(p);
            }
        }
        // second verdict: did everybody open everything?
        uint32_t verdict = ok ? 1U : 0U;
        h2d(ctx, comm->stage.ptr, &verdict, sizeof(verdict));
        nccl_check(api.AllReduce(comm->stage.ptr, comm->stage.ptr, 1, ncclUint32, ncclMin,
                                 static_cast<ncclComm_t>(comm->nccl_comm), ctx->stream),
                   "ncclAllReduce(ipc verdict)");
        d2h(ctx, &verdict, comm->stage.ptr, sizeof(verdict));
        sync(ctx);
        if (verdict == 0)
        {
            for (void* p : pr->opened)
            {
                if (p != nullptr)
                {
                    cudaIpcCloseMemHandle(p);
                }
            }
            cudaFree(pr->mailbox);
            set_last_error("peer mailbox unavailable (no CUDA IPC / peer access between the ranks): reductions use NCCL");
            throw Error(1, "peer mailbox unavailable: reductions use NCCL");
        }
        pr->comm = comm;
        pr->box.world = world;
        pr->box.rank = comm->rank;
        pr->box.bins_pad = bins_pad;
        rdf->reduced.reserve((size_t) rdf->axis.bins + 1);
        FGPU_CUDA_CHECK(cudaMemsetAsync(rdf->reduced.ptr, 0, ((size_t) rdf->axis.bins + 1) * sizeof(uint32_t), ctx->stream));
        sync(ctx);
        rdf->peer = pr.release();
    });
}

int fgpu_rdf_reduce_transport(const fgpu_rdf* rdf)
{
    return rdf != nullptr && rdf->peer != nullptr ? 2 : 1;
}

int fgpu_rdf_accumulate_reduce(fgpu_rdf* rdf, fgpu_points* pts, fgpu_comm* comm, int flavour, float q_r_max,
                               float q_r_min, int exclude_ii)
{
    return guarded([&] {
        require(rdf != nullptr && pts != nullptr && comm != nullptr, FGPU_EINVALID, "null argument");
        require(rdf->ctx == comm->ctx, FGPU_EINVALID, "rdf and comm belong to different contexts");
        bool const peer = rdf->peer != nullptr && rdf->peer->comm == comm;
        bool const pushed = rdf_accumulate_impl(rdf, pts, nullptr, nullptr, pts->n, 0, flavour, q_r_max, q_r_min,
                                                exclude_ii, peer);
        if (peer)
        {
            peer_reduce(rdf, pushed);
        }
        else
        {
            nccl_reduce(rdf, comm);
        }
    });
}

// ---- Steinhardt ------------------------------------------------------------------------------------------
static AxisDev regular_axis(uint32_t bins, float lo, float hi)
{
    // RegularAxis ctor, freud/util/Histogram.h:126-138
    volatile float span = hi - lo;
    volatile float width = span / (float) bins;
    volatile float inv = 1.0f / width;
    AxisDev a;
    a.r_min = lo;
    a.r_max = hi;
    a.inv_width = inv;
    a.bins = bins;
    return a;
}

// ---- PMFTXYZ / PMFTXYT / PMFTR12 ----------------------------------------------------------------------------------
namespace {

// RegularAxis::bin (Histogram.h:152-173) on the host; -1 = outside
int host_axis_bin(const AxisDev& a, float value)
{
    if (!(value >= a.r_min) || value >= a.r_max)
    {
        return -1;
    }
    volatile float d = value - a.r_min;
    volatile float val = d * a.inv_width;
    int bin = (int) val;
    return (uint32_t) bin == a.bins ? bin - 1 : bin;
}

float host_mod_two_pi(float a) // util::modulusPositive(a, TWO_PI), utils.h:29-32
{
    float const two_pi = (float) (2.0 * M_PI); // Box.h:24
    volatile float inner = std::fmod(a, two_pi) + two_pi;
    return std::fmod(inner, two_pi);
}

constexpr uint64_t kPmftChunk = 1ULL << 24; // bonds per launch: bounds the list of host-binned bonds (20 B each)

} // namespace

int fgpu_pmft_create(fgpu_ctx* ctx, int kind, float max0, float max1, float max2, uint32_t n0, uint32_t n1, uint32_t n2,
                     fgpu_pmft** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        require(kind == FGPU_PMFT_XYZ || kind == FGPU_PMFT_XYT || kind == FGPU_PMFT_R12 || kind == FGPU_PMFT_XY,
                FGPU_EINVALID, "unknown PMFT kind");
        float const two_pi = (float) (2.0 * M_PI);
        std::unique_ptr<fgpu_pmft> p(new fgpu_pmft());
        p->ctx = ctx;
        p->kind = kind;
        if (kind == FGPU_PMFT_XYZ) // PMFTXYZ.cc:27-50
        {
            require(n0 >= 1, FGPU_EINVALID, "PMFTXYZ requires at least 1 bin in X.");
            require(n1 >= 1, FGPU_EINVALID, "PMFTXYZ requires at least 1 bin in Y.");
            require(n2 >= 1, FGPU_EINVALID, "PMFTXYZ requires at least 1 bin in Z.");
            require(!(max0 < 0), FGPU_EINVALID, "PMFTXYZ requires that x_max must be positive.");
            require(!(max1 < 0), FGPU_EINVALID, "PMFTXYZ requires that y_max must be positive.");
            require(!(max2 < 0), FGPU_EINVALID, "PMFTXYZ requires that z_max must be positive.");
            p->a0 = regular_axis(n0, -max0, max0);
            p->a1 = regular_axis(n1, -max1, max1);
            p->a2 = regular_axis(n2, -max2, max2);
        }
        else if (kind == FGPU_PMFT_XY) // PMFTXY.cc:27-42; the third axis is a single bin that every bond falls into
        {
            require(n0 >= 1, FGPU_EINVALID, "PMFTXY requires at least 1 bin in X.");
            require(n1 >= 1, FGPU_EINVALID, "PMFTXY requires at least 1 bin in Y.");
            require(!(max0 < 0), FGPU_EINVALID, "PMFTXY requires that x_max must be positive.");
            require(!(max1 < 0), FGPU_EINVALID, "PMFTXY requires that y_max must be positive.");
            p->a0 = regular_axis(n0, -max0, max0);
            p->a1 = regular_axis(n1, -max1, max1);
            p->a2 = regular_axis(1, 0.0f, 1.0f);
            n2 = 1;
        }
        else if (kind == FGPU_PMFT_XYT) // PMFTXYT.cc:30-49
        {
            require(n0 >= 1, FGPU_EINVALID, "PMFTXYT requires at least 1 bin in X.");
            require(n1 >= 1, FGPU_EINVALID, "PMFTXYT requires at least 1 bin in Y.");
            require(n2 >= 1, FGPU_EINVALID, "PMFTXYT requires at least 1 bin in T.");
            require(!(max0 < 0), FGPU_EINVALID, "PMFTXYT requires that x_max must be positive.");
            require(!(max1 < 0), FGPU_EINVALID, "PMFTXYT requires that y_max must be positive.");
            p->a0 = regular_axis(n0, -max0, max0);
            p->a1 = regular_axis(n1, -max1, max1);
            p->a2 = regular_axis(n2, 0.0f, two_pi);
        }
        else // PMFTR12.cc:30-45
        {
            require(n0 >= 1, FGPU_EINVALID, "PMFTR12 requires at least 1 bin in R.");
            require(n1 >= 1, FGPU_EINVALID, "PMFTR12 requires at least 1 bin in T1.");
            require(n2 >= 1, FGPU_EINVALID, "PMFTR12 requires at least 1 bin in T2.");
            require(!(max0 < 0), FGPU_EINVALID, "PMFTR12 requires that r_max must be positive.");
            p->a0 = regular_axis(n0, 0.0f, max0);
            p->a1 = regular_axis(n1, 0.0f, two_pi);
            p->a2 = regular_axis(n2, 0.0f, two_pi);
        }
        uint64_t const n_bins = (uint64_t) n0 * n1 * n2;
        require(n_bins < (1ULL << 31), FGPU_EINVALID, "PMFT histogram too large");
        bind_device(ctx);
        p->hist.reserve((size_t) n_bins);
        FGPU_CUDA_CHECK(cudaMemsetAsync(p->hist.ptr, 0, (size_t) n_bins * sizeof(uint32_t), ctx->stream));
        sync(ctx);
        *out = p.release();
    });
}

void fgpu_pmft_destroy(fgpu_pmft* pmft)
{
    if (pmft != nullptr)
    {
        bind_quiet(pmft->ctx);
        delete pmft;
    }
}

int fgpu_pmft_reset(fgpu_pmft* pmft)
{
    return guarded([&] {
        require(pmft != nullptr, FGPU_EINVALID, "null argument");
        bind_device(pmft->ctx);
        size_t const n_bins = (size_t) pmft->a0.bins * pmft->a1.bins * pmft->a2.bins;
        FGPU_CUDA_CHECK(cudaMemsetAsync(pmft->hist.ptr, 0, n_bins * sizeof(uint32_t), pmft->ctx->stream));
        pmft->deferred_total = 0;
    });
}

namespace {

// Page-locked scratch from the library's host cache (fgpu_host_alloc): the bonds left to the host come back at the
// link's rate and without a zero-fill (a perfect lattice leaves all of them: 240 MB per million-particle frame).
struct HostScratch
{
    void* p = nullptr;
    explicit HostScratch(size_t bytes)
    {
        if (fgpu_host_alloc(bytes, &p) != FGPU_OK)
        {
            throw Error(FGPU_ENOMEM, fgpu_last_error());
        }
    }
    HostScratch(const HostScratch&) = delete;
    HostScratch& operator=(const HostScratch&) = delete;
    ~HostScratch()
    {
        fgpu_host_free(p);
    }
};

struct PmftHostInputs
{
    const float* orientations;
    const float* query_orientations;
};

// uploads the orientations of one frame and fills the kernel arguments that do not depend on where the bonds are
Pmft3Args stage_pmft(fgpu_pmft* pmft, uint32_t n_points, uint32_t n_query, const float* orientations_host,
                     const float* query_orientations_host, const float* equiv_orientations_host, uint32_t n_equiv)
{
    fgpu_ctx* ctx = pmft->ctx;
    int const kind = pmft->kind;
    Pmft3Args a {};
    a.a0 = pmft->a0;
    a.a1 = pmft->a1;
    a.a2 = pmft->a2;
    a.hist = pmft->hist.ptr;
    if (kind == FGPU_PMFT_XYZ)
    {
        require(equiv_orientations_host != nullptr && n_equiv >= 1, FGPU_EINVALID,
                "PMFTXYZ needs at least one equivalent orientation");
        pmft->stage_a.reserve(4 * (size_t) n_query + 4);
        pmft->stage_b.reserve(4 * (size_t) n_equiv);
        h2d(ctx, pmft->stage_a.ptr, query_orientations_host, 4 * (size_t) n_query * sizeof(float));
        h2d(ctx, pmft->stage_b.ptr, equiv_orientations_host, 4 * (size_t) n_equiv * sizeof(float));
        a.query_quats = reinterpret_cast<const float4*>(pmft->stage_a.ptr);
        a.equiv_quats = reinterpret_cast<const float4*>(pmft->stage_b.ptr);
        a.n_equiv = n_equiv;
        return a;
    }
    if (kind != FGPU_PMFT_XY)
    {
        require(orientations_host != nullptr, FGPU_EINVALID, "null orientations");
        pmft->stage_b.reserve((size_t) n_points + 1);
        h2d(ctx, pmft->stage_b.ptr, orientations_host, (size_t) n_points * sizeof(float));
        a.orientations = pmft->stage_b.ptr;
    }
    pmft->stage_a.reserve((size_t) n_query + 1);
    h2d(ctx, pmft->stage_a.ptr, query_orientations_host, (size_t) n_query * sizeof(float));
    a.query_orientations = pmft->stage_a.ptr;
    return a;
}

// One launch of the histogram kernel over the bonds `a` describes, counting into a.hist, with room for `cap` bonds
// left to the host; those are binned here with the host's libm and added to a.hist.  Returns false -- nothing
// reliable in a.hist -- if more than `cap` bonds were left over.
bool run_pmft_pass(fgpu_pmft* pmft, Pmft3Args& a, uint64_t cap, const PmftHostInputs& in)
{
    fgpu_ctx* ctx = pmft->ctx;
    int const kind = pmft->kind;
    if (kind == FGPU_PMFT_XYZ)
    {
        launch_pmft3(ctx, kind, a);
        return true;
    }
    pmft->deferred.reserve((size_t) cap);
    if (kind == FGPU_PMFT_R12)
    {
        pmft->deferred_dist.reserve((size_t) cap);
    }
    a.deferred = pmft->deferred.ptr;
    a.deferred_dist = pmft->deferred_dist.ptr;
    a.deferred_cap = (uint32_t) cap;
    a.deferred_count = reinterpret_cast<uint32_t*>(ctx->d_scalars + 7);
    FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 7, 0, sizeof(unsigned long long), ctx->stream));
    launch_pmft3(ctx, kind, a);
    d2h(ctx, ctx->h_scalars + 7, ctx->d_scalars + 7, sizeof(unsigned long long));
    sync(ctx);
    uint32_t const n_def = (uint32_t) (ctx->h_scalars[7] & 0xffffffffULL);
    if (n_def > cap)
    {
        return false;
    }
    if (n_def == 0)
    {
        return true;
    }
    // the bonds whose coordinate or angle sits within a few ulps of a bin edge: the reference's own libm decides
    HostScratch rec_mem((size_t) n_def * sizeof(uint4)), dist_mem(kind == FGPU_PMFT_R12 ? (size_t) n_def * sizeof(float) : 16);
    const uint4* const rec = static_cast<uint4*>(rec_mem.p);
    const float* const rec_dist = static_cast<float*>(dist_mem.p);
    d2h(ctx, rec_mem.p, pmft->deferred.ptr, (size_t) n_def * sizeof(uint4));
    if (kind == FGPU_PMFT_R12)
    {
        d2h(ctx, dist_mem.p, pmft->deferred_dist.ptr, (size_t) n_def * sizeof(float));
    }
    sync(ctx);
    // a few libm calls per bond: tens of thousands of bonds per frame (all of them on a lattice) are worth the host's threads
    HostScratch slot_mem((size_t) n_def * sizeof(uint32_t));
    uint32_t* const slot = static_cast<uint32_t*>(slot_mem.p);
    uint32_t const kNone = 0xffffffffU;
    auto bin_range = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t r = lo; r < hi; ++r)
    {
        uint32_t const i = rec[r].x, j = rec[r].y;
        float vx, vy;
        std::memcpy(&vx, &rec[r].z, sizeof(float));
        std::memcpy(&vy, &rec[r].w, sizeof(float));
        int c0, c1, c2;
        if (kind == FGPU_PMFT_XYT || kind == FGPU_PMFT_XY)
        {
            float const t = -in.query_orientations[i]; // rotmat2::fromAngle, VectorMath.h:912-921
            float const c = std::cos(t), sn = std::sin(t);
            volatile float x1 = c * vx, x2 = -sn * vy, y1 = sn * vx, y2 = c * vy;
            c0 = host_axis_bin(a.a0, x1 + x2);
            c1 = host_axis_bin(a.a1, y1 + y2);
            c2 = 0;
            if (kind == FGPU_PMFT_XYT)
            {
                float const d_theta = std::atan2(-vy, -vx); // PMFTXYT.cc:94
                c2 = host_axis_bin(a.a2, host_mod_two_pi(in.orientations[j] - d_theta));
            }
        }
        else
        {
            c0 = host_axis_bin(a.a0, rec_dist[r]);
            float const d_theta1 = std::atan2(vy, vx), d_theta2 = std::atan2(-vy, -vx); // PMFTR12.cc:103-104
            c1 = host_axis_bin(a.a1, host_mod_two_pi(in.orientations[j] - d_theta1));
            c2 = host_axis_bin(a.a2, host_mod_two_pi(in.query_orientations[i] - d_theta2));
        }
        slot[r] = c0 >= 0 && c1 >= 0 && c2 >= 0 ? ((uint32_t) c0 * a.a1.bins + (uint32_t) c1) * a.a2.bins + (uint32_t) c2
                                                : kNone;
    }
    };
    unsigned const n_threads = n_def >= 4096 ? std::max(1U, std::min(32U, std::thread::hardware_concurrency())) : 1U;
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t)
        {
            pool.emplace_back(bin_range, (uint32_t) ((uint64_t) n_def * t / n_threads),
                              (uint32_t) ((uint64_t) n_def * (t + 1) / n_threads));
        }
        bin_range(0, n_def / n_threads);
        for (auto& th : pool)
        {
            th.join();
        }
    }
    pmft->deferred_total += n_def;
    {
        // every slot goes up; the kernel skips the bonds that fell outside the axes
        pmft->host_bins.reserve(n_def);
        h2d(ctx, pmft->host_bins.ptr, slot, (size_t) n_def * sizeof(uint32_t));
        launch_add_bins(ctx, pmft->host_bins.ptr, n_def, a.hist);
        sync(ctx); // the scratch goes back to the cache
    }
    return true;
}

} // namespace

int fgpu_pmft_accumulate_nlist(fgpu_pmft* pmft, const fgpu_nlist* nl, const float* orientations_host, uint32_t n_points,
                               const float* query_orientations_host, const float* equiv_orientations_host,
                               uint32_t n_equiv)
{
    return guarded([&] {
        require(pmft != nullptr && nl != nullptr && query_orientations_host != nullptr, FGPU_EINVALID, "null argument");
        require(pmft->ctx == nl->ctx, FGPU_EINVALID, "pmft and nlist belong to different contexts");
        fgpu_ctx* ctx = pmft->ctx;
        bind_device(ctx);
        Pmft3Args a = stage_pmft(pmft, n_points, nl->n_query, orientations_host, query_orientations_host,
                                 equiv_orientations_host, n_equiv);
        PmftHostInputs const in {orientations_host, query_orientations_host};
        for (uint64_t b0 = 0; b0 < nl->n_bonds; b0 += kPmftChunk)
        {
            uint64_t const nb = std::min<uint64_t>(kPmftChunk, nl->n_bonds - b0);
            a.neighbors = nl->neighbors.ptr + 2 * b0;
            a.vectors = nl->vectors.ptr + 3 * b0;
            a.distances = nl->distances.ptr + b0;
            a.n_bonds = nb;
            run_pmft_pass(pmft, a, nb, in); // room for every bond of the chunk: cannot fail
        }
        sync(ctx); // the caller's arrays were consumed
    });
}

// The query and the histogram in one call, without a NeighborList: the search leaves its hits in the bag, grouped by
// query row, and the histogram kernel reads them there (k_pmft3<., ROWS>) -- no ranking, no 28 B per bond written and
// read back.  The frame is counted into a histogram of its own first, so that a frame that leaves more bonds to the
// host than the list has room for can simply be repeated with a longer list.
int fgpu_pmft_accumulate(fgpu_pmft* pmft, fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour,
                         float r_max, float r_min, int exclude_ii, const float* orientations_host,
                         const float* query_orientations_host, const float* equiv_orientations_host, uint32_t n_equiv)
{
    return guarded([&] {
        require(pmft != nullptr && pts != nullptr && query_orientations_host != nullptr, FGPU_EINVALID, "null argument");
        require(pmft->ctx == pts->ctx, FGPU_EINVALID, "pmft and points belong to different contexts");
        fgpu_ctx* ctx = pmft->ctx;
        bind_device(ctx);
        validate_ball(pts, flavour, r_max, r_min);
        bool const self = query_points_host == nullptr;
        require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
        require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
        uint64_t n_bonds = 0;
        bool const fused = search_to_bag(pts, query_points_host, n_query, flavour, r_max, r_min, exclude_ii, &n_bonds);
        if (!fused)
        {
            // tiny grids, points outside the box, ...: through a NeighborList
            fgpu_nlist* nl = nullptr;
            ball_query_impl(pts, query_points_host, nullptr, n_query, 0, flavour, r_max, r_min, exclude_ii, 0, &nl);
            std::unique_ptr<fgpu_nlist, void (*)(fgpu_nlist*)> guard(nl, fgpu_nlist_destroy);
            int const rc = fgpu_pmft_accumulate_nlist(pmft, nl, orientations_host, pts->n, query_orientations_host,
                                                      equiv_orientations_host, n_equiv);
            if (rc != FGPU_OK)
            {
                throw Error(rc, fgpu_last_error());
            }
            return;
        }
        Pmft3Args a = stage_pmft(pmft, pts->n, n_query, orientations_host, query_orientations_host,
                                 equiv_orientations_host, n_equiv);
        PmftHostInputs const in {orientations_host, query_orientations_host};
        size_t const n_bins = (size_t) pmft->a0.bins * pmft->a1.bins * pmft->a2.bins;
        pmft->frame_hist.reserve(n_bins);
        a.hist = pmft->frame_hist.ptr;
        a.bag = ctx->bag4.ptr;
        a.row_bag_start = ctx->tmp_start.ptr;
        a.row_counts = ctx->row_counts.ptr;
        a.n_rows = n_query;
        double const per_row = (double) n_bonds / (double) n_query;
        a.group = per_row < 6.0 ? 4U : (per_row <= 96.0 ? 8U : 32U);
        uint64_t room = std::min<uint64_t>(std::max<uint64_t>(n_bonds / 16, 4096), n_bonds);
        uint64_t const before = pmft->deferred_total;
        for (;;)
        {
            FGPU_CUDA_CHECK(cudaMemsetAsync(pmft->frame_hist.ptr, 0, n_bins * sizeof(uint32_t), ctx->stream));
            pmft->deferred_total = before;
            if (run_pmft_pass(pmft, a, room, in))
            {
                break;
            }
            room = n_bonds; // e.g. a lattice whose every bond angle sits on a bin edge
        }
        launch_add_hist(ctx, pmft->frame_hist.ptr, (uint32_t) n_bins, pmft->hist.ptr);
        sync(ctx); // the caller's arrays were consumed
    });
}

int fgpu_pmft_read(fgpu_pmft* pmft, uint32_t* counts_host)
{
    return guarded([&] {
        require(pmft != nullptr && counts_host != nullptr, FGPU_EINVALID, "null argument");
        bind_device(pmft->ctx);
        size_t const n_bins = (size_t) pmft->a0.bins * pmft->a1.bins * pmft->a2.bins;
        d2h(pmft->ctx, counts_host, pmft->hist.ptr, n_bins * sizeof(uint32_t));
        sync(pmft->ctx);
    });
}

int fgpu_pmft_deferred(const fgpu_pmft* pmft, uint64_t* bonds)
{
    return guarded([&] {
        require(pmft != nullptr && bonds != nullptr, FGPU_EINVALID, "null argument");
        *bonds = pmft->deferred_total;
    });
}

// ---- PMFTXY: the two-axis histogram is the XY kind of fgpu_pmft ------------------------------------------------
int fgpu_pmftxy_create(fgpu_ctx* ctx, float x_max, float y_max, uint32_t n_x, uint32_t n_y, fgpu_pmftxy** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        fgpu_pmft* inner = nullptr;
        int const rc = fgpu_pmft_create(ctx, FGPU_PMFT_XY, x_max, y_max, 0.0f, n_x, n_y, 1, &inner);
        if (rc != FGPU_OK)
        {
            throw Error(rc, fgpu_last_error());
        }
        std::unique_ptr<fgpu_pmftxy> p(new fgpu_pmftxy());
        p->inner = inner;
        *out = p.release();
    });
}

void fgpu_pmftxy_destroy(fgpu_pmftxy* pmft)
{
    if (pmft != nullptr)
    {
        fgpu_pmft_destroy(pmft->inner);
        delete pmft;
    }
}

int fgpu_pmftxy_reset(fgpu_pmftxy* pmft)
{
    return pmft == nullptr ? fgpu_pmft_reset(nullptr) : fgpu_pmft_reset(pmft->inner);
}

int fgpu_pmftxy_accumulate_nlist(fgpu_pmftxy* pmft, const fgpu_nlist* nl, const float* query_orientations_host)
{
    return fgpu_pmft_accumulate_nlist(pmft == nullptr ? nullptr : pmft->inner, nl, nullptr, 0, query_orientations_host,
                                      nullptr, 0);
}

int fgpu_pmftxy_accumulate(fgpu_pmftxy* pmft, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                           int flavour, float r_max, float r_min, int exclude_ii, const float* query_orientations_host)
{
    return fgpu_pmft_accumulate(pmft == nullptr ? nullptr : pmft->inner, pts, query_points_host, n_query, flavour, r_max,
                                r_min, exclude_ii, nullptr, query_orientations_host, nullptr, 0);
}

int fgpu_pmftxy_read(fgpu_pmftxy* pmft, uint32_t* counts_host)
{
    return fgpu_pmft_read(pmft == nullptr ? nullptr : pmft->inner, counts_host);
}

// ---- BondOrder ----------------------------------------------------------------------------------------------------
namespace {

// rotate(q, v), VectorMath.h:810-818, in float with one rounding per operation (this file is built without contraction)
void host_quat_rotate(float s, float qx, float qy, float qz, float& x, float& y, float& z)
{
    float const bx = x, by = y, bz = z;
    float const p1 = qx * qx, p2 = qy * qy, p3 = qz * qz;
    float const vv = (p1 + p2) + p3;
    float const ss = s * s;
    float const a = ss - vv;
    float const two_s = 2.0f * s;
    float const c1 = qy * bz, c2 = qz * by, c3 = qz * bx, c4 = qx * bz, c5 = qx * by, c6 = qy * bx;
    float const cx = c1 - c2, cy = c3 - c4, cz = c5 - c6;
    float const d1 = qx * bx, d2 = qy * by, d3 = qz * bz;
    float const vb = (d1 + d2) + d3;
    float const two_vb = 2.0f * vb;
    float const t1x = bx * a, t2x = cx * two_s, t3x = qx * two_vb;
    float const t1y = by * a, t2y = cy * two_s, t3y = qy * two_vb;
    float const t1z = bz * a, t2z = cz * two_s, t3z = qz * two_vb;
    x = (t1x + t2x) + t3x;
    y = (t1y + t2y) + t3y;
    z = (t1z + t2z) + t3z;
}

} // namespace

int fgpu_bondorder_create(fgpu_ctx* ctx, uint32_t n_theta, uint32_t n_phi, int mode, fgpu_bondorder** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        // BondOrder.cc:34-41
        require(n_theta >= 2, FGPU_EINVALID, "BondOrder requires at least 2 bins in theta.");
        require(n_phi >= 2, FGPU_EINVALID, "BondOrder requires at least 2 bins in phi.");
        require(mode >= FGPU_BOND_ORDER_BOD && mode <= FGPU_BOND_ORDER_OOCD, FGPU_EINVALID, "unknown BondOrder mode");
        require((uint64_t) n_theta * n_phi < (1ULL << 31), FGPU_EINVALID, "BondOrder histogram too large");
        bind_device(ctx);
        std::unique_ptr<fgpu_bondorder> b(new fgpu_bondorder());
        b->ctx = ctx;
        b->mode = mode;
        b->at = regular_axis(n_theta, 0.0f, (float) (2.0 * M_PI)); // BondOrder.cc:70-71
        b->ap = regular_axis(n_phi, 0.0f, (float) M_PI);
        b->hist.reserve((size_t) n_theta * n_phi);
        FGPU_CUDA_CHECK(cudaMemsetAsync(b->hist.ptr, 0, (size_t) n_theta * n_phi * sizeof(uint32_t), ctx->stream));
        sync(ctx);
        *out = b.release();
    });
}

void fgpu_bondorder_destroy(fgpu_bondorder* bo)
{
    if (bo != nullptr)
    {
        bind_quiet(bo->ctx);
        delete bo;
    }
}

int fgpu_bondorder_reset(fgpu_bondorder* bo)
{
    return guarded([&] {
        require(bo != nullptr, FGPU_EINVALID, "null argument");
        bind_device(bo->ctx);
        FGPU_CUDA_CHECK(cudaMemsetAsync(bo->hist.ptr, 0, (size_t) bo->at.bins * bo->ap.bins * sizeof(uint32_t),
                                        bo->ctx->stream));
        bo->deferred_total = 0;
    });
}

namespace {

// uploads the orientations of one frame and fills the kernel arguments that do not depend on where the bonds are
BondOrderArgs stage_bond_order(fgpu_bondorder* bo, uint32_t n_points, uint32_t n_query, const float* orientations_host,
                               const float* query_orientations_host)
{
    fgpu_ctx* ctx = bo->ctx;
    require(bo->mode == FGPU_BOND_ORDER_BOD || (orientations_host != nullptr && query_orientations_host != nullptr),
            FGPU_EINVALID, "null orientations");
    BondOrderArgs a {};
    a.at = bo->at;
    a.ap = bo->ap;
    a.mode = bo->mode;
    a.hist = bo->hist.ptr;
    if (bo->mode != FGPU_BOND_ORDER_BOD)
    {
        bo->stage_a.reserve(4 * (size_t) n_points + 4);
        bo->stage_b.reserve(4 * (size_t) n_query + 4);
        h2d(ctx, bo->stage_a.ptr, orientations_host, 4 * (size_t) n_points * sizeof(float));
        h2d(ctx, bo->stage_b.ptr, query_orientations_host, 4 * (size_t) n_query * sizeof(float));
        a.orientations = reinterpret_cast<const float4*>(bo->stage_a.ptr);
        a.query_orientations = reinterpret_cast<const float4*>(bo->stage_b.ptr);
    }
    return a;
}

// One launch of k_bond_order over the bonds `a` describes, counting into a.hist, with room for `cap` bonds left to
// the host (binned here with its libm and added to a.hist).  Returns false if more than `cap` were left over.
bool run_bond_order_pass(fgpu_bondorder* bo, BondOrderArgs& a, uint64_t cap, const float* orientations_host,
                         const float* query_orientations_host)
{
    fgpu_ctx* ctx = bo->ctx;
    int const mode = bo->mode;
    bo->deferred.reserve((size_t) cap);
    bo->deferred_z.reserve((size_t) cap);
    a.deferred = bo->deferred.ptr;
    a.deferred_z = bo->deferred_z.ptr;
    a.deferred_cap = (uint32_t) cap;
    a.deferred_count = reinterpret_cast<uint32_t*>(ctx->d_scalars + 7);
    FGPU_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 7, 0, sizeof(unsigned long long), ctx->stream));
    launch_bond_order(ctx, a);
    d2h(ctx, ctx->h_scalars + 7, ctx->d_scalars + 7, sizeof(unsigned long long));
    sync(ctx);
    uint32_t const n_def = (uint32_t) (ctx->h_scalars[7] & 0xffffffffULL);
    if (n_def > cap)
    {
        return false;
    }
    if (n_def == 0)
    {
        return true;
    }
    HostScratch rec_mem((size_t) n_def * sizeof(uint4)), z_mem((size_t) n_def * sizeof(float));
    const uint4* const rec = static_cast<uint4*>(rec_mem.p);
    const float* const rec_z = static_cast<float*>(z_mem.p);
    d2h(ctx, rec_mem.p, bo->deferred.ptr, (size_t) n_def * sizeof(uint4));
    d2h(ctx, z_mem.p, bo->deferred_z.ptr, (size_t) n_def * sizeof(float));
    sync(ctx);
    // two libm calls per bond: a perfect lattice leaves every bond here (each sits on a bin edge), so the loop runs on
    // every host thread (one thread took 0.5 s for the 12 M bonds of a million-particle FCC frame)
    HostScratch slot_mem((size_t) n_def * sizeof(uint32_t));
    uint32_t* const slot = static_cast<uint32_t*>(slot_mem.p);
    uint32_t const kNone = 0xffffffffU;
    auto bin_range = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; ++r)
        {
            uint32_t const i = rec[r].x, j = rec[r].y;
            float x, y, z = rec_z[r];
            std::memcpy(&x, &rec[r].z, sizeof(float));
            std::memcpy(&y, &rec[r].w, sizeof(float));
            if (mode != FGPU_BOND_ORDER_BOD) // BondOrder.cc:108-134
            {
                const float* rq = orientations_host + 4 * (size_t) j;
                const float* q = query_orientations_host + 4 * (size_t) i;
                if (mode == FGPU_BOND_ORDER_OOCD)
                {
                    x = 0.0f;
                    y = 0.0f;
                    z = 1.0f;
                    host_quat_rotate(q[0], q[1], q[2], q[3], x, y, z);
                }
                host_quat_rotate(rq[0], -rq[1], -rq[2], -rq[3], x, y, z);
                if (mode == FGPU_BOND_ORDER_OBCD)
                {
                    host_quat_rotate(q[0], q[1], q[2], q[3], x, y, z);
                }
            }
            float const theta = host_mod_two_pi(std::atan2(y, x)); // BondOrder.cc:140-141
            float const xx = x * x, yy = y * y, zz = z * z;
            float const dot = (xx + yy) + zz;
            float const arg = z / std::sqrt(dot);
            float const phi = std::acos(arg); // :144
            int const bt = host_axis_bin(a.at, theta), bp = host_axis_bin(a.ap, phi);
            slot[r] = bt >= 0 && bp >= 0 ? (uint32_t) bt * a.ap.bins + (uint32_t) bp : kNone;
        }
    };
    unsigned const n_threads = n_def >= 4096 ? std::max(1U, std::min(32U, std::thread::hardware_concurrency())) : 1U;
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t)
        {
            pool.emplace_back(bin_range, (uint32_t) ((uint64_t) n_def * t / n_threads),
                              (uint32_t) ((uint64_t) n_def * (t + 1) / n_threads));
        }
        bin_range(0, (uint32_t) ((uint64_t) n_def / n_threads));
        for (auto& th : pool)
        {
            th.join();
        }
    }
    bo->deferred_total += n_def;
    bo->host_bins.reserve(n_def);
    h2d(ctx, bo->host_bins.ptr, slot, (size_t) n_def * sizeof(uint32_t)); // the kernel skips the bonds outside the axes
    launch_add_bins(ctx, bo->host_bins.ptr, n_def, a.hist);
    sync(ctx);
    return true;
}

} // namespace

int fgpu_bondorder_accumulate_nlist(fgpu_bondorder* bo, const fgpu_nlist* nl, const float* orientations_host,
                                    uint32_t n_points, const float* query_orientations_host)
{
    return guarded([&] {
        require(bo != nullptr && nl != nullptr, FGPU_EINVALID, "null argument");
        require(bo->ctx == nl->ctx, FGPU_EINVALID, "bond order and nlist belong to different contexts");
        fgpu_ctx* ctx = bo->ctx;
        bind_device(ctx);
        BondOrderArgs a = stage_bond_order(bo, n_points, nl->n_query, orientations_host, query_orientations_host);
        for (uint64_t b0 = 0; b0 < nl->n_bonds; b0 += kPmftChunk)
        {
            uint64_t const nb = std::min<uint64_t>(kPmftChunk, nl->n_bonds - b0);
            a.neighbors = nl->neighbors.ptr + 2 * b0;
            a.vectors = nl->vectors.ptr + 3 * b0;
            a.n_bonds = nb;
            run_bond_order_pass(bo, a, nb, orientations_host, query_orientations_host); // room for every bond
        }
        sync(ctx); // the caller's arrays were consumed
    });
}

// query + histogram in one call, the bonds read from the search's bag (see fgpu_pmft_accumulate)
int fgpu_bondorder_accumulate(fgpu_bondorder* bo, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                              int flavour, float r_max, float r_min, int exclude_ii, const float* orientations_host,
                              const float* query_orientations_host)
{
    return guarded([&] {
        require(bo != nullptr && pts != nullptr, FGPU_EINVALID, "null argument");
        require(bo->ctx == pts->ctx, FGPU_EINVALID, "bond order and points belong to different contexts");
        fgpu_ctx* ctx = bo->ctx;
        bind_device(ctx);
        validate_ball(pts, flavour, r_max, r_min);
        bool const self = query_points_host == nullptr;
        require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
        require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
        uint64_t n_bonds = 0;
        if (!search_to_bag(pts, query_points_host, n_query, flavour, r_max, r_min, exclude_ii, &n_bonds))
        {
            fgpu_nlist* nl = nullptr;
            ball_query_impl(pts, query_points_host, nullptr, n_query, 0, flavour, r_max, r_min, exclude_ii, 0, &nl);
            std::unique_ptr<fgpu_nlist, void (*)(fgpu_nlist*)> guard(nl, fgpu_nlist_destroy);
            int const rc = fgpu_bondorder_accumulate_nlist(bo, nl, orientations_host, pts->n, query_orientations_host);
            if (rc != FGPU_OK)
            {
                throw Error(rc, fgpu_last_error());
            }
            return;
        }
        BondOrderArgs a = stage_bond_order(bo, pts->n, n_query, orientations_host, query_orientations_host);
        size_t const n_bins = (size_t) bo->at.bins * bo->ap.bins;
        bo->frame_hist.reserve(n_bins);
        a.hist = bo->frame_hist.ptr;
        a.bag = ctx->bag4.ptr;
        a.row_bag_start = ctx->tmp_start.ptr;
        a.row_counts = ctx->row_counts.ptr;
        a.n_rows = n_query;
        double const per_row = (double) n_bonds / (double) n_query;
        a.group = per_row < 6.0 ? 4U : (per_row <= 96.0 ? 8U : 32U);
        uint64_t room = std::min<uint64_t>(std::max<uint64_t>(n_bonds / 16, 4096), n_bonds);
        uint64_t const before = bo->deferred_total;
        for (;;)
        {
            FGPU_CUDA_CHECK(cudaMemsetAsync(bo->frame_hist.ptr, 0, n_bins * sizeof(uint32_t), ctx->stream));
            bo->deferred_total = before;
            if (run_bond_order_pass(bo, a, room, orientations_host, query_orientations_host))
            {
                break;
            }
            room = n_bonds; // e.g. a perfect lattice: every bond direction on a bin edge
        }
        launch_add_hist(ctx, bo->frame_hist.ptr, (uint32_t) n_bins, bo->hist.ptr);
        sync(ctx); // the caller's arrays were consumed
    });
}

int fgpu_bondorder_read(fgpu_bondorder* bo, uint32_t* counts_host)
{
    return guarded([&] {
        require(bo != nullptr && counts_host != nullptr, FGPU_EINVALID, "null argument");
        bind_device(bo->ctx);
        d2h(bo->ctx, counts_host, bo->hist.ptr, (size_t) bo->at.bins * bo->ap.bins * sizeof(uint32_t));
        sync(bo->ctx);
    });
}

int fgpu_bondorder_deferred(const fgpu_bondorder* bo, uint64_t* bonds)
{
    return guarded([&] {
        require(bo != nullptr && bonds != nullptr, FGPU_EINVALID, "null argument");
        *bonds = bo->deferred_total;
    });
}

int fgpu_corr_create(fgpu_ctx* ctx, uint32_t bins, float r_max, fgpu_corr** out)
{
    return guarded([&] {
        require(ctx != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        // CorrelationFunction.cc:28-35
        require(bins != 0, FGPU_EINVALID, "CorrelationFunction  requires a nonzero number of bins.");
        require(r_max > 0, FGPU_EINVALID, "CorrelationFunction requires r_max to be positive.");
        bind_device(ctx);
        std::unique_ptr<fgpu_corr> c(new fgpu_corr());
        c->ctx = ctx;
        volatile float span = r_max - 0.0f; // RegularAxis ctor, Histogram.h:126-138
        volatile float width = span / (float) bins;
        volatile float inv = 1.0f / width;
        c->axis.r_min = 0.0f;
        c->axis.r_max = r_max;
        c->axis.inv_width = inv;
        c->axis.bins = bins;
        c->counts.reserve(bins);
        c->sums.reserve(2 * (size_t) bins);
        FGPU_CUDA_CHECK(cudaMemsetAsync(c->counts.ptr, 0, (size_t) bins * sizeof(uint32_t), ctx->stream));
        FGPU_CUDA_CHECK(cudaMemsetAsync(c->sums.ptr, 0, 2 * (size_t) bins * sizeof(double), ctx->stream));
        sync(ctx);
        *out = c.release();
    });
}

void fgpu_corr_destroy(fgpu_corr* corr)
{
    if (corr != nullptr)
    {
        bind_quiet(corr->ctx);
        delete corr;
    }
}

int fgpu_corr_reset(fgpu_corr* corr)
{
    return guarded([&] {
        require(corr != nullptr, FGPU_EINVALID, "null argument");
        bind_device(corr->ctx);
        FGPU_CUDA_CHECK(cudaMemsetAsync(corr->counts.ptr, 0, (size_t) corr->axis.bins * sizeof(uint32_t), corr->ctx->stream));
        FGPU_CUDA_CHECK(cudaMemsetAsync(corr->sums.ptr, 0, 2 * (size_t) corr->axis.bins * sizeof(double), corr->ctx->stream));
    });
}

int fgpu_corr_accumulate_nlist(fgpu_corr* corr, const fgpu_nlist* nl, const double* values_host,
                               const double* query_values_host)
{
    return guarded([&] {
        require(corr != nullptr && nl != nullptr && values_host != nullptr && query_values_host != nullptr, FGPU_EINVALID,
                "null argument");
        require(corr->ctx == nl->ctx, FGPU_EINVALID, "corr and nlist belong to different contexts");
        fgpu_ctx* ctx = corr->ctx;
        bind_device(ctx);
        corr->values.reserve(2 * (size_t) nl->n_points + 2);
        h2d(ctx, corr->values.ptr, values_host, 2 * (size_t) nl->n_points * sizeof(double));
        const double* d_query_values = corr->values.ptr;
        if (query_values_host != values_host || nl->n_query != nl->n_points)
        {
            // (the points against themselves hand in one array for both: uploaded once)
            corr->query_values.reserve(2 * (size_t) nl->n_query + 2);
            h2d(ctx, corr->query_values.ptr, query_values_host, 2 * (size_t) nl->n_query * sizeof(double));
            d_query_values = corr->query_values.ptr;
        }
        launch_correlation(ctx, nl->neighbors.ptr, nl->distances.ptr, nl->n_bonds, corr->values.ptr, d_query_values,
                           corr->axis, corr->counts.ptr, corr->sums.ptr);
        sync(ctx); // the caller's value arrays were consumed
    });
}

int fgpu_corr_accumulate(fgpu_corr* corr, fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour,
                         float r_max, float r_min, int exclude_ii, const double* values_host,
                         const double* query_values_host)
{
    return guarded([&] {
        require(corr != nullptr && pts != nullptr && values_host != nullptr && query_values_host != nullptr, FGPU_EINVALID,
                "null argument");
        require(corr->ctx == pts->ctx, FGPU_EINVALID, "corr and points belong to different contexts");
        fgpu_ctx* ctx = corr->ctx;
        bind_device(ctx);
        validate_ball(pts, flavour, r_max, r_min);
        bool const self = query_points_host == nullptr;
        require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
        require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
        uint64_t n_bonds = 0;
        if (!search_to_bag(pts, query_points_host, n_query, flavour, r_max, r_min, exclude_ii, &n_bonds))
        {
            fgpu_nlist* nl = nullptr;
            ball_query_impl(pts, query_points_host, nullptr, n_query, 0, flavour, r_max, r_min, exclude_ii, 0, &nl);
            std::unique_ptr<fgpu_nlist, void (*)(fgpu_nlist*)> guard(nl, fgpu_nlist_destroy);
            int const rc = fgpu_corr_accumulate_nlist(corr, nl, values_host, query_values_host);
            if (rc != FGPU_OK)
            {
                throw Error(rc, fgpu_last_error());
            }
            return;
        }
        corr->values.reserve(2 * (size_t) pts->n + 2);
        h2d(ctx, corr->values.ptr, values_host, 2 * (size_t) pts->n * sizeof(double));
        const double* d_query_values = corr->values.ptr;
        if (query_values_host != values_host || n_query != pts->n)
        {
            corr->query_values.reserve(2 * (size_t) n_query + 2);
            h2d(ctx, corr->query_values.ptr, query_values_host, 2 * (size_t) n_query * sizeof(double));
            d_query_values = corr->query_values.ptr;
        }
        launch_correlation_rows(ctx, ctx->bag4.ptr, ctx->tmp_start.ptr, ctx->row_counts.ptr, n_query, corr->values.ptr,
                                d_query_values, corr->axis, corr->counts.ptr, corr->sums.ptr);
        sync(ctx); // the caller's value arrays were consumed
    });
}

int fgpu_corr_read(fgpu_corr* corr, uint32_t* counts_host, double* sums_host)
{
    return guarded([&] {
        require(corr != nullptr, FGPU_EINVALID, "null argument");
        fgpu_ctx* ctx = corr->ctx;
        bind_device(ctx);
        if (counts_host != nullptr)
        {
            d2h(ctx, counts_host, corr->counts.ptr, (size_t) corr->axis.bins * sizeof(uint32_t));
        }
        if (sums_host != nullptr)
        {
            d2h(ctx, sums_host, corr->sums.ptr, 2 * (size_t) corr->axis.bins * sizeof(double));
        }
        sync(ctx);
    });
}

namespace {

// LocalDensity.cc:48-49: area = M_PI * r * r (double, rounded once); volume = float(4/3 pi) * r * r * r
float local_density_measure(float r_max, int is2d)
{
    volatile float vol = static_cast<float>(4.0 / 3.0 * M_PI);
    vol = vol * r_max;
    vol = vol * r_max;
    vol = vol * r_max;
    return is2d ? (float) (M_PI * (double) r_max * (double) r_max) : (float) vol;
}

} // namespace

int fgpu_local_density_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour, float q_r_max,
                             float q_r_min, int exclude_ii, float r_max, float diameter, float* num_neighbors_host,
                             float* density_host)
{
    return guarded([&] {
        require(pts != nullptr, FGPU_EINVALID, "null argument");
        require(r_max > 0, FGPU_EINVALID, "LocalDensity requires r_max to be positive.");
        require(!(diameter < 0), FGPU_EINVALID, "LocalDensity requires diameter to be non-negative.");
        fgpu_ctx* ctx = pts->ctx;
        bind_device(ctx);
        validate_ball(pts, flavour, q_r_max, q_r_min);
        bool const self = query_points_host == nullptr;
        require(!self || n_query == pts->n, FGPU_EINVALID, "self query requires n_query == n_points");
        require(pts->n_shards == 1, FGPU_ERUNTIME, "sharded points serve self-query RDF accumulation only");
        uint64_t n_bonds = 0;
        if (!search_to_bag(pts, query_points_host, n_query, flavour, q_r_max, q_r_min, exclude_ii, &n_bonds))
        {
            fgpu_nlist* nl = nullptr;
            ball_query_impl(pts, query_points_host, nullptr, n_query, 0, flavour, q_r_max, q_r_min, exclude_ii, 0, &nl);
            std::unique_ptr<fgpu_nlist, void (*)(fgpu_nlist*)> guard(nl, fgpu_nlist_destroy);
            int const rc = fgpu_local_density(nl, r_max, diameter, pts->box.is2d ? 1 : 0, num_neighbors_host, density_host);
            if (rc != FGPU_OK)
            {
                throw Error(rc, fgpu_last_error());
            }
            return;
        }
        DevBuf<float> d_num, d_den;
        d_num.reserve((size_t) n_query + 1);
        d_den.reserve((size_t) n_query + 1);
        launch_local_density_rows(ctx, ctx->bag4.ptr, ctx->tmp_start.ptr, ctx->row_counts.ptr, n_query, r_max, diameter,
                                  local_density_measure(r_max, pts->box.is2d ? 1 : 0), d_num.ptr, d_den.ptr);
        if (num_neighbors_host != nullptr)
        {
            d2h(ctx, num_neighbors_host, d_num.ptr, (size_t) n_query * sizeof(float));
        }
        if (density_host != nullptr)
        {
            d2h(ctx, density_host, d_den.ptr, (size_t) n_query * sizeof(float));
        }
        sync(ctx);
    });
}

int fgpu_local_density(const fgpu_nlist* nl, float r_max, float diameter, int is2d, float* num_neighbors_host,
                       float* density_host)
{
    return guarded([&] {
        require(nl != nullptr, FGPU_EINVALID, "null argument");
        require(r_max > 0, FGPU_EINVALID, "LocalDensity requires r_max to be positive.");
        require(!(diameter < 0), FGPU_EINVALID, "LocalDensity requires diameter to be non-negative.");
        fgpu_ctx* ctx = nl->ctx;
        bind_device(ctx);
        uint32_t const n = nl->n_query;
        float const measure = local_density_measure(r_max, is2d);
        DevBuf<float> d_num, d_den;
        d_num.reserve((size_t) n + 1);
        d_den.reserve((size_t) n + 1);
        launch_local_density(ctx, nl->row_start.ptr, nl->distances.ptr, n, r_max, diameter, measure, d_num.ptr,
                             d_den.ptr);
        if (num_neighbors_host != nullptr)
        {
            d2h(ctx, num_neighbors_host, d_num.ptr, (size_t) n * sizeof(float));
        }
        if (density_host != nullptr)
        {
            d2h(ctx, density_host, d_den.ptr, (size_t) n * sizeof(float));
        }
        sync(ctx);
    });
}

// fused: the neighbours are not a list but the bag of a k-nearest-neighbour window search over the points themselves
// (nl == nullptr then): q_l / q_lm come out of k_knn_ylm, everything behind the kernels is shared.
static void steinhardt_compute_body(fgpu_points* pts, const fgpu_nlist* nl, const KnnSelectArgs* fused, int fused_flavour,
                                    const uint32_t* ls, uint32_t n_ls, int flags, uint32_t n_total, fgpu_comm* comm,
                                    float* ql_host, float* wl_host, float* qlm_host, fgpu_buffer** qlm_keep,
                                    float* sys_qlm_host, float* order_host, bool* fused_too_long)
{
    {
        require(pts != nullptr && (nl != nullptr || fused != nullptr) && ls != nullptr && n_ls != 0, FGPU_EINVALID,
                "null argument");
        require(nl == nullptr || pts->ctx == nl->ctx, FGPU_EINVALID, "points and nlist belong to different contexts");
        require(nl == nullptr || nl->n_points == pts->n, FGPU_EINVALID,
                "NeighborList was built for a different number of points");
        require(nl == nullptr || (uint64_t) nl->n_query + nl->q_index_offset <= pts->n, FGPU_EINVALID,
                "NeighborList has more rows than there are points");
        bool const weighted = (flags & FGPU_ST_WEIGHTED) != 0, average = (flags & FGPU_ST_AVERAGE) != 0;
        bool const wl = (flags & FGPU_ST_WL) != 0, wl_normalize = wl && (flags & FGPU_ST_WL_NORMALIZE) != 0;
        fgpu_ctx* ctx = pts->ctx;
        bind_device(ctx);
        std::vector<uint32_t> lv(ls, ls + n_ls);
        size_t tot_m = 0;
        for (uint32_t l : lv)
        {
            tot_m += 2 * (size_t) l + 1;
            // getWigner3j, freud/order/Wigner3j.cc: std::out_of_range beyond the tabulated range
            require(!wl || l <= 20, FGPU_ERANGE, "Wigner 3j coefficients are implemented for l <= 20.");
        }
        uint32_t const n = nl != nullptr ? nl->n_query : pts->n; // rows held by this rank (== pts->n on a single GPU)
        require(!average || n == pts->n, FGPU_ERUNTIME,
                "Steinhardt average needs the q_lm of every neighbour: all rows must be on this rank");
        if (n_total == 0)
        {
            n_total = n;
        }
        DevBuf<float> d_ql, d_qlm, d_ql_ave, d_qlm_ave, d_wl, d_w3j;
        DevBuf<uint32_t> d_w3j_off;
        DevBuf<double> d_sys;
        d_ql.reserve((size_t) n * n_ls + 1);
        bool const need_qlm = qlm_host != nullptr || qlm_keep != nullptr || average || wl;
        if (need_qlm)
        {
            d_qlm.reserve((size_t) n * tot_m * 2 + 2);
        }
        d_sys.reserve(tot_m * 2);
        FGPU_CUDA_CHECK(cudaMemsetAsync(d_sys.ptr, 0, tot_m * 2 * sizeof(double), ctx->stream));
        SteinhardtArgs a;
        a.box = pts->box;
        a.rcp_lx = rounded_reciprocal(pts->box.Lx);
        a.rcp_ly = rounded_reciprocal(pts->box.Ly);
        a.rcp_lz = rounded_reciprocal(pts->box.Lz);
        if (!pts->xyz4_ready)
        {
            pts->xyz4.reserve(pts->n);
            launch_pad_positions(ctx, pts->xyz.ptr, pts->n, pts->xyz4.ptr);
            pts->xyz4_ready = true;
        }
        a.xyz = pts->xyz.ptr;
        a.xyz4 = pts->xyz4.ptr;
        a.n = n;
        a.row_offset = nl != nullptr ? nl->q_index_offset : 0;
        a.neighbors = nl != nullptr ? nl->neighbors.ptr : nullptr;
        a.distances = nl != nullptr ? nl->distances.ptr : nullptr;
        a.weights = nl != nullptr ? nl->weights.ptr : nullptr;
        a.row_start = nl != nullptr ? nl->row_start.ptr : nullptr;
        a.weighted = weighted;
        a.n_total = n_total;
        a.ql = d_ql.ptr;
        a.qlm = need_qlm ? d_qlm.ptr : nullptr;
        a.sys_qlm = average ? nullptr : d_sys.ptr; // the system sums come from the averaged q_lm then
        a.sys_partials = nullptr;
        if (fused != nullptr)
        {
            require(!average && knn_ylm_supported(lv, fused->k), FGPU_ERUNTIME, "fused kNN -> Ylm: unsupported options");
            launch_knn_ylm(ctx, a, lv[0], *fused);
        }
        else
        {
            launch_steinhardt(ctx, a, lv);
        }
        if (average)
        {
            d_ql_ave.reserve((size_t) n * n_ls + 1);
            d_qlm_ave.reserve((size_t) n * tot_m * 2 + 2);
            SteinhardtAveArgs av;
            av.n = n;
            av.neighbors = nl->neighbors.ptr;
            av.row_start = nl->row_start.ptr;
            av.qlm = d_qlm.ptr;
            av.qlm_ave = d_qlm_ave.ptr;
            av.ql_ave = d_ql_ave.ptr;
            av.sys_qlm = d_sys.ptr;
            av.sys_partials = nullptr;
            launch_steinhardt_average(ctx, av, (int) n_ls, (uint32_t) tot_m);
        }
        std::vector<std::vector<float>> w3j(n_ls);
        if (wl)
        {
            std::vector<float> flat;
            std::vector<uint32_t> off(n_ls);
            for (uint32_t r = 0; r < n_ls; ++r)
            {
                w3j[r] = wigner3j_table(lv[r]);
                off[r] = (uint32_t) flat.size();
                flat.insert(flat.end(), w3j[r].begin(), w3j[r].end());
            }
            d_w3j.reserve(flat.size());
            d_w3j_off.reserve(n_ls);
            d_wl.reserve((size_t) n * n_ls + 1);
            h2d(ctx, d_w3j.ptr, flat.data(), flat.size() * sizeof(float));
            h2d(ctx, d_w3j_off.ptr, off.data(), n_ls * sizeof(uint32_t));
            SteinhardtWlArgs wa;
            wa.n = n;
            wa.qlm = average ? d_qlm_ave.ptr : d_qlm.ptr;
            wa.ql = average ? d_ql_ave.ptr : d_ql.ptr;
            wa.w3j = d_w3j.ptr;
            wa.w3j_off = d_w3j_off.ptr;
            wa.normalize = wl_normalize ? 1 : 0;
            wa.wl = d_wl.ptr;
            launch_steinhardt_wl(ctx, wa, (int) n_ls);
            sync(ctx); // the pageable host tables above were consumed
        }
        if (comm != nullptr && comm->size > 1)
        {
            NcclApi& api = nccl_or_throw();
            nccl_check(api.AllReduce(d_sys.ptr, d_sys.ptr, tot_m * 2, ncclFloat64, ncclSum,
                                     static_cast<ncclComm_t>(comm->nccl_comm), ctx->stream),
                       "ncclAllReduce(f64 system q_lm)");
        }
        std::vector<double> sys(tot_m * 2);
        d2h(ctx, sys.data(), d_sys.ptr, tot_m * 2 * sizeof(double));
        if (ql_host != nullptr)
        {
            d2h(ctx, ql_host, average ? d_ql_ave.ptr : d_ql.ptr, (size_t) n * n_ls * sizeof(float));
        }
        if (wl_host != nullptr && wl)
        {
            d2h(ctx, wl_host, d_wl.ptr, (size_t) n * n_ls * sizeof(float));
        }
        if (qlm_host != nullptr)
        {
            d2h(ctx, qlm_host, d_qlm.ptr, (size_t) n * tot_m * 2 * sizeof(float));
        }
        sync(ctx);
        if (qlm_keep != nullptr)
        {
            // the per-particle q_lm (104 MB at l = 6, N = 1e6) stay on the device until somebody asks for them
            std::unique_ptr<fgpu_buffer> keep(new fgpu_buffer());
            keep->ctx = ctx;
            keep->bytes = (uint64_t) n * tot_m * 2 * sizeof(float);
            keep->data.swap(d_qlm);
            *qlm_keep = keep.release();
        }
        // system q_lm = sum_i q_lm(i) / N ; derive m < 0 ; normalizeSystem (Steinhardt.cc:291-327)
        size_t off = 0;
        for (uint32_t r = 0; r < n_ls; ++r)
        {
            uint32_t const l = lv[r];
            double norm = 0.0;
            for (uint32_t m = 0; m <= l; ++m)
            {
                sys[2 * (off + m)] /= (double) n_total;
                sys[2 * (off + m) + 1] /= (double) n_total;
            }
            for (uint32_t m = 1; m <= l; ++m)
            {
                double const phase = (m & 1U) ? -1.0 : 1.0;
                sys[2 * (off + l + m)] = phase * sys[2 * (off + m)];
                sys[2 * (off + l + m) + 1] = -phase * sys[2 * (off + m) + 1];
            }
            for (uint32_t k = 0; k < 2 * l + 1; ++k)
            {
                double const re = sys[2 * (off + k)], im = sys[2 * (off + k) + 1];
                norm += re * re + im * im;
                if (sys_qlm_host != nullptr)
                {
                    sys_qlm_host[2 * (off + k)] = (float) re;
                    sys_qlm_host[2 * (off + k) + 1] = (float) im;
                }
            }
            double const nf = 4.0 * M_PI / (2 * l + 1);
            double order = std::sqrt(norm * nf);
            if (wl)
            {
                // reduceWigner3j over the system q_lm (Wigner3j.cc:43-55), then the optional normalisation
                double const ql_system = order;
                double acc = 0.0;
                size_t counter = 0;
                int const li = (int) l;
                auto at = [&](int m) { return off + (size_t) (m < 0 ? li - m : m); };
                for (int m1 = -li; m1 <= li; ++m1)
                {
                    for (int m2 = std::max(-li - m1, -li); m2 <= std::min(li - m1, li); ++m2)
                    {
                        int const m3 = -m1 - m2;
                        std::complex<double> const s1(sys[2 * at(m1)], sys[2 * at(m1) + 1]);
                        std::complex<double> const s2(sys[2 * at(m2)], sys[2 * at(m2) + 1]);
                        std::complex<double> const s3(sys[2 * at(m3)], sys[2 * at(m3) + 1]);
                        acc += ((double) w3j[r][counter] * s1 * s2 * s3).real();
                        ++counter;
                    }
                }
                if (wl_normalize)
                {
                    double const nrm = std::sqrt(nf) / ql_system;
                    acc *= nrm * nrm * nrm;
                }
                order = acc;
            }
            if (order_host != nullptr)
            {
                order_host[r] = (float) order;
            }
            off += 2 * (size_t) l + 1;
        }
    }
}

static int steinhardt_compute_impl(fgpu_points* pts, const fgpu_nlist* nl, const uint32_t* ls, uint32_t n_ls, int flags,
                                   uint32_t n_total, fgpu_comm* comm, float* ql_host, float* wl_host, float* qlm_host,
                                   fgpu_buffer** qlm_keep, float* sys_qlm_host, float* order_host)
{
    return guarded([&] {
        require(nl != nullptr, FGPU_EINVALID, "null argument");
        bool unused = false;
        steinhardt_compute_body(pts, nl, nullptr, 0, ls, n_ls, flags, n_total, comm, ql_host, wl_host, qlm_host, qlm_keep,
                                sys_qlm_host, order_host, &unused);
    });
}

int fgpu_steinhardt_compute(fgpu_points* pts, const fgpu_nlist* nl, const uint32_t* ls, uint32_t n_ls, int flags,
                            uint32_t n_total, fgpu_comm* comm, float* ql_host, float* wl_host, float* qlm_host,
                            float* sys_qlm_host, float* order_host)
{
    return steinhardt_compute_impl(pts, nl, ls, n_ls, flags, n_total, comm, ql_host, wl_host, qlm_host, nullptr,
                                   sys_qlm_host, order_host);
}

int fgpu_steinhardt_compute_keep(fgpu_points* pts, const fgpu_nlist* nl, const uint32_t* ls, uint32_t n_ls, int flags,
                                 uint32_t n_total, fgpu_comm* comm, float* ql_host, float* wl_host,
                                 fgpu_buffer** qlm_dev_out, float* sys_qlm_host, float* order_host)
{
    if (qlm_dev_out == nullptr)
    {
        set_last_error("null argument");
        return FGPU_EINVALID;
    }
    *qlm_dev_out = nullptr;
    return steinhardt_compute_impl(pts, nl, ls, n_ls, flags, n_total, comm, ql_host, wl_host, nullptr, qlm_dev_out,
                                   sys_qlm_host, order_host);
}

int fgpu_steinhardt_knn(fgpu_points* pts, int flavour, uint32_t num_neighbors, float r_max, float r_min, int exclude_ii,
                        const uint32_t* ls, uint32_t n_ls, int flags, float* ql_host, float* wl_host,
                        fgpu_buffer** qlm_dev_out, float* sys_qlm_host, float* order_host)
{
    return guarded([&] {
        require(pts != nullptr && ls != nullptr && n_ls != 0, FGPU_EINVALID, "null argument");
        if (qlm_dev_out != nullptr)
        {
            *qlm_dev_out = nullptr;
        }
        std::vector<uint32_t> const lv(ls, ls + n_ls);
        bool const fusable = (flags & (FGPU_ST_AVERAGE | FGPU_ST_WL)) == 0
            && knn_ylm_supported(lv, std::min<uint32_t>(num_neighbors, pts->n));
        bool done = false, too_long = false;
        std::function<void(const KnnSelectArgs&)> const consumer = [&](const KnnSelectArgs& sa) {
            steinhardt_compute_body(pts, nullptr, &sa, flavour, ls, n_ls, flags, pts->n, nullptr, ql_host, wl_host, nullptr,
                                    qlm_dev_out, sys_qlm_host, order_host, &too_long);
            done = !too_long;
        };
        fgpu_nlist* nl = nullptr;
        knn_query_body(pts, nullptr, pts->n, 0, flavour, num_neighbors, r_max, r_min, exclude_ii, 0, &nl,
                       fusable ? &consumer : nullptr);
        if (done)
        {
            return;
        }
        std::unique_ptr<fgpu_nlist, void (*)(fgpu_nlist*)> guard(nl, fgpu_nlist_destroy);
        if (nl == nullptr)
        {
            // the fused kernel declined the frame (a row beyond its staging): the same query again, as a list
            fgpu_nlist* again = nullptr;
            knn_query_body(pts, nullptr, pts->n, 0, flavour, num_neighbors, r_max, r_min, exclude_ii, 0, &again, nullptr);
            guard.reset(again);
        }
        bool unused = false;
        steinhardt_compute_body(pts, guard.get(), nullptr, 0, ls, n_ls, flags, pts->n, nullptr, ql_host, wl_host, nullptr,
                                qlm_dev_out, sys_qlm_host, order_host, &unused);
    });
}

// ---- device buffers kept for a later read ----------------------------------------------------------------------
uint64_t fgpu_buffer_bytes(const fgpu_buffer* buf)
{
    return buf != nullptr ? buf->bytes : 0;
}

int fgpu_buffer_read(fgpu_buffer* buf, void* host, uint64_t offset_bytes, uint64_t bytes)
{
    return guarded([&] {
        require(buf != nullptr && (host != nullptr || bytes == 0), FGPU_EINVALID, "null argument");
        require(offset_bytes <= buf->bytes && bytes <= buf->bytes - offset_bytes, FGPU_ERANGE, "read beyond the buffer");
        bind_device(buf->ctx);
        d2h(buf->ctx, host, reinterpret_cast<const unsigned char*>(buf->data.ptr) + offset_bytes, bytes);
        sync(buf->ctx);
    });
}

void fgpu_buffer_destroy(fgpu_buffer* buf)
{
    if (buf != nullptr)
    {
        bind_quiet(buf->ctx);
        delete buf;
    }
}

// ---- page-locked host memory, cached ---------------------------------------------------------------------------
// The outputs of this path are large (225 MB of NeighborList arrays, 104 MB of q_lm per 1 M-point frame) and a
// device -> pageable-host copy runs at a fraction of the link's rate, while page-locking memory costs about as much
// as the copy it speeds up.  So the host classes take their arrays from here: blocks are page-locked once, handed
// back on free and reused by the next frame's arrays of the same size class.  Without a CUDA device (host-only unit
// tests of the container classes) the blocks are ordinary aligned memory.
namespace {

struct HostPool
{
    std::mutex mu;
    std::unordered_map<void*, std::pair<size_t, bool>> live;   // block -> {size class, page-locked}
    std::unordered_map<void*, std::pair<size_t, bool>> parked; // same for the cached blocks
    std::multimap<size_t, void*> free_by_size;
    size_t cached_bytes = 0;
    int no_device = -1; // -1: unknown
};

HostPool& host_pool()
{
    static HostPool* p = new HostPool(); // never destroyed: blocks may outlive static destruction order
    return *p;
}

size_t host_size_class(size_t bytes)
{
    size_t const unit = bytes <= (1U << 20) ? 4096 : (2U << 20);
    return std::max<size_t>(unit, (bytes + unit - 1) / unit * unit);
}

constexpr size_t kHostCacheLimit = 6ULL << 30;

} // namespace

int fgpu_host_alloc(uint64_t bytes, void** out)
{
    return guarded([&] {
        require(out != nullptr, FGPU_EINVALID, "null argument");
        HostPool& hp = host_pool();
        size_t const cls = host_size_class((size_t) bytes);
        std::lock_guard<std::mutex> lock(hp.mu);
        auto it = hp.free_by_size.lower_bound(cls);
        if (it != hp.free_by_size.end() && it->first <= cls + cls / 4)
        {
            void* p = it->second;
            hp.free_by_size.erase(it);
            auto const rec = hp.parked[p];
            hp.parked.erase(p);
            hp.cached_bytes -= rec.first;
            hp.live[p] = rec;
            *out = p;
            return;
        }
        void* p = nullptr;
        bool pinned = false;
        if (hp.no_device != 1)
        {
            cudaError_t const err = cudaHostAlloc(&p, cls, cudaHostAllocPortable);
            if (err == cudaSuccess)
            {
                pinned = true;
                hp.no_device = 0;
            }
            else
            {
                cudaGetLastError();
                p = nullptr;
                if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver)
                {
                    hp.no_device = 1;
                }
            }
        }
        if (p == nullptr)
        {
            p = std::aligned_alloc(4096, cls);
        }
        if (p == nullptr)
        {
            throw Error(FGPU_ENOMEM, "out of host memory");
        }
        hp.live[p] = {cls, pinned};
        *out = p;
    });
}

void fgpu_host_free(void* p)
{
    if (p == nullptr)
    {
        return;
    }
    HostPool& hp = host_pool();
    std::lock_guard<std::mutex> lock(hp.mu);
    auto it = hp.live.find(p);
    if (it == hp.live.end())
    {
        return; // not ours
    }
    auto const rec = it->second;
    hp.live.erase(it);
    if (hp.cached_bytes + rec.first <= kHostCacheLimit)
    {
        hp.parked[p] = rec;
        hp.free_by_size.emplace(rec.first, p);
        hp.cached_bytes += rec.first;
        return;
    }
    if (rec.second)
    {
        cudaFreeHost(p);
    }
    else
    {
        std::free(p);
    }
}

int fgpu_host_trim(void)
{
    return guarded([&] {
        HostPool& hp = host_pool();
        std::lock_guard<std::mutex> lock(hp.mu);
        for (auto& kv : hp.parked)
        {
            if (kv.second.second)
            {
                cudaFreeHost(kv.first);
            }
            else
            {
                std::free(kv.first);
            }
        }
        hp.parked.clear();
        hp.free_by_size.clear();
        hp.cached_bytes = 0;
    });
}

// ---- NCCL ------------------------------------------------------------------------------------------------
int fgpu_comm_unique_id(uint8_t* unique_id_out)
{
    return guarded([&] {
        require(unique_id_out != nullptr, FGPU_EINVALID, "null argument");
        static_assert(sizeof(ncclUniqueId) == FGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");
        NcclApi& api = nccl_or_throw();
        ncclUniqueId id;
        nccl_check(api.GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(unique_id_out, &id, sizeof(id));
    });
}

int fgpu_comm_create(fgpu_ctx* ctx, const uint8_t* unique_id, int rank, int n_ranks, fgpu_comm** out)
{
    return guarded([&] {
        require(ctx != nullptr && unique_id != nullptr && out != nullptr, FGPU_EINVALID, "null argument");
        require(n_ranks >= 1 && rank >= 0 && rank < n_ranks, FGPU_EINVALID, "bad rank / size");
        bind_device(ctx);
        NcclApi& api = nccl_or_throw();
        ncclUniqueId id;
        std::memcpy(&id, unique_id, sizeof(id));
        ncclComm_t c = nullptr;
        nccl_check(api.CommInitRank(&c, n_ranks, id, rank), "ncclCommInitRank");
        std::unique_ptr<fgpu_comm> comm(new fgpu_comm());
        comm->ctx = ctx;
        comm->nccl_comm = c;
        comm->rank = rank;
        comm->size = n_ranks;
        *out = comm.release();
    });
}

void fgpu_comm_destroy(fgpu_comm* comm)
{
    if (comm != nullptr)
    {
        bind_quiet(comm->ctx);
        cudaStreamSynchronize(comm->ctx->stream);
        if (comm->nccl_comm != nullptr && nccl().handle != nullptr)
        {
            nccl().CommDestroy(static_cast<ncclComm_t>(comm->nccl_comm));
        }
        delete comm;
    }
}

int fgpu_comm_rank(const fgpu_comm* comm)
{
    return comm != nullptr ? comm->rank : 0;
}

int fgpu_comm_size(const fgpu_comm* comm)
{
    return comm != nullptr ? comm->size : 1;
}

static void allreduce_host(fgpu_comm* comm, void* host, uint64_t count, size_t elem, ncclDataType_t type)
{
    require(comm != nullptr && (host != nullptr || count == 0), FGPU_EINVALID, "null argument");
    if (count == 0)
    {
        return;
    }
    fgpu_ctx* ctx = comm->ctx;
    bind_device(ctx);
    NcclApi& api = nccl_or_throw();
    comm->stage.reserve(count * elem);
    h2d(ctx, comm->stage.ptr, host, count * elem);
    nccl_check(api.AllReduce(comm->stage.ptr, comm->stage.ptr, count, type, ncclSum,
                             static_cast<ncclComm_t>(comm->nccl_comm), ctx->stream),
               "ncclAllReduce");
    d2h(ctx, host, comm->stage.ptr, count * elem);
    sync(ctx);
}

int fgpu_comm_allreduce_u32(fgpu_comm* comm, uint32_t* host_inout, uint64_t count)
{
    return guarded([&] { allreduce_host(comm, host_inout, count, sizeof(uint32_t), ncclUint32); });
}

int fgpu_comm_allreduce_f64(fgpu_comm* comm, double* host_inout, uint64_t count)
{
    return guarded([&] { allreduce_host(comm, host_inout, count, sizeof(double), ncclFloat64); });
}

int fgpu_comm_barrier(fgpu_comm* comm)
{
    return guarded([&] {
        uint32_t token = 1;
        allreduce_host(comm, &token, 1, sizeof(uint32_t), ncclUint32);
    });
}

} // extern "C"
