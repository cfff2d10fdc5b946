// Internal declarations shared by the translation units of libfreud_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/freud_b200.h"
#include "pair_math.cuh"
#include "peer.cuh"

namespace fgpu {

// ---- error plumbing ------------------------------------------------------------------------------
struct Error : std::runtime_error
{
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define FGPU_CUDA_CHECK(expr)                                                                                    \
    do                                                                                                           \
    {                                                                                                            \
        cudaError_t const err__ = (expr);                                                                        \
        if (err__ != cudaSuccess)                                                                                \
        {                                                                                                        \
            throw ::fgpu::Error(FGPU_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));              \
        }                                                                                                        \
    } while (0)

// The stream every allocation / free of the calling thread is ordered on; set by bind_device(ctx) at the top
// of each entry point.  Allocation goes through CUDA's stream-ordered pool (release threshold = never), so a
// NeighborList per frame costs microseconds instead of the milliseconds cudaMalloc/cudaFree take.
cudaStream_t& current_stream();

// ---- device buffer (grow-only, stream-ordered use on the context's single stream) -----------------
template<typename T> struct DevBuf
{
    T* ptr = nullptr;
    size_t cap = 0; // elements
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf()
    {
        release();
    }
    void release()
    {
        if (ptr != nullptr)
        {
            cudaFreeAsync(ptr, current_stream());
        }
        ptr = nullptr;
        cap = 0;
    }
    void swap(DevBuf& other)
    {
        std::swap(ptr, other.ptr);
        std::swap(cap, other.cap);
    }
    // contents are NOT preserved on growth
    void reserve(size_t n)
    {
        if (n <= cap)
        {
            return;
        }
        release();
        size_t const want = n + n / 8 + 64;
        FGPU_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&ptr), want * sizeof(T), current_stream()));
        cap = want;
    }
};

} // namespace fgpu

// The arrays of a NeighborList.  A destroyed list leaves them with its context (one spare set, the largest seen),
// and the next list starts from them: a trajectory loop that drops each frame's list before the next query runs
// without a single allocation per frame -- gigabyte-sized blocks otherwise come and go in the stream-ordered pool,
// which now and then has to map fresh memory for them (tens to hundreds of milliseconds).
struct NlistStorage
{
    fgpu::DevBuf<uint32_t> neighbors; // n_bonds x 2
    fgpu::DevBuf<float> distances;
    fgpu::DevBuf<float> weights;
    fgpu::DevBuf<float> vectors;      // n_bonds x 3
    fgpu::DevBuf<uint32_t> row_start; // n_query + 1 (exclusive scan; internal)
    fgpu::DevBuf<uint32_t> counts;    // n_query
    fgpu::DevBuf<uint32_t> segments;  // n_query (0 for empty rows, as upstream)
    void swap(NlistStorage& o)
    {
        neighbors.swap(o.neighbors);
        distances.swap(o.distances);
        weights.swap(o.weights);
        vectors.swap(o.vectors);
        row_start.swap(o.row_start);
        counts.swap(o.counts);
        segments.swap(o.segments);
    }
    size_t bytes() const
    {
        return 4 * (neighbors.cap + distances.cap + weights.cap + vectors.cap + row_start.cap + counts.cap + segments.cap);
    }
};

// ---- opaque handle definitions ----------------------------------------------------------------------
struct fgpu_ctx
{
    NlistStorage spare_nlist;
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    uint64_t launches = 0;
    bool count_evals = false;
    unsigned long long* d_evals = nullptr; // device counter of pair evaluations
    // scratch
    fgpu::DevBuf<uint32_t> scan_tmp;
    fgpu::DevBuf<float> q_stage;          // staged query points (H2D)
    fgpu::DevBuf<uint32_t> q_cell;        // per query: cell index
    fgpu::DevBuf<uint32_t> q_rank;        // per query: arrival rank inside its cell
    fgpu::DevBuf<uint32_t> q_cell_start;  // per cell: start of the cell-ordered queries
    fgpu::DevBuf<float4> q_sorted;        // cell-ordered query points, w = original query index
    fgpu::DevBuf<uint32_t> row_counts;    // per query: number of bonds
    fgpu::DevBuf<uint32_t> row_start;     // exclusive scan of row_counts (n_query + 1)
    fgpu::DevBuf<uint4> bag;              // unsorted hits {key_hi, key_lo, slot, query}
    fgpu::DevBuf<float4> bag4;            // search2 bag: {bond vector, bits(point index)} per hit
    fgpu::DevBuf<uint32_t> tmp_start;     // per query: offset of its row in the bag
    fgpu::DevBuf<uint32_t> knn_hits;      // kNN: per query, hits inside the search window
    fgpu::DevBuf<uint32_t> knn_unresolved; // kNN: rows with fewer than k hits (searched again, wider)
    fgpu::DevBuf<float> knn_subset;       // kNN: positions of those rows' query points
    fgpu::DevBuf<float4> bag4b;           // kNN: bag of the second, wider search
    fgpu::DevBuf<int> q_outside_flag;     // device flag: a query point lies outside the box
    uint64_t bag_hint = 0;                // bonds of the previous query (sizes the next bag)
    int force_general = 0;                // fgpu_ctx_force_general_search: always run the search.cu kernels (testing)
    // fgpu_ctx_set_tuning: experiment / test hooks, never read from the environment
    int tune_span = 0;                    // > 0: cells per home tile of the tile walk
    int tune_no_symmetry = 0;             // 1: self-query IMAGE RDF without the symmetric walk
    int tune_lanes_over_queries = -1;     // -1 automatic, 0 / 1: force the NeighborList search's mapping
    int tune_pmft_cluster = 0;            // 1: PMFT histograms beyond one block's shared memory count in cluster DSMEM (slower, pmft.cu)
    int tune_lq_blocks = 0;               // > 0: resident blocks per SM of the lanes-over-queries search (the rest of
                                          // the SM's 256 KB stays L1)
    fgpu::DevBuf<double> st_partials;     // Steinhardt: per-block partial sums of the system q_lm
    fgpu::DevBuf<float> knn_d;            // kNN scratch, [k][n_query]
    fgpu::DevBuf<uint32_t> knn_s;         // kNN scratch, [k][n_query] (slot | image code)
    unsigned long long* d_scalars = nullptr; // 8 x u64 device scalars (totals, flags)
    unsigned long long* h_scalars = nullptr; // pinned mirror
    void* pinned_stage = nullptr;            // pinned staging for pageable H2D/D2H
    size_t pinned_bytes = 0;
    // per-kernel timing (fgpu_ctx_profile)
    struct TimerRec
    {
        const char* name;
        cudaEvent_t begin, end;
    };
    bool profile = false;
    std::vector<TimerRec> timers;
    std::vector<cudaEvent_t> event_pool; // events of cleared timers, reused (cudaEventCreate costs more than a launch)
};

namespace fgpu {
// Brackets one kernel launch: counts it and, when profiling, records an event pair on the stream.
struct KernelScope
{
    fgpu_ctx* ctx;
    cudaEvent_t begin = nullptr, end = nullptr;
    const char* name;
    KernelScope(fgpu_ctx* c, const char* n) : ctx(c), name(n)
    {
        ctx->launches += 1;
        if (ctx->profile)
        {
            begin = take(ctx);
            end = take(ctx);
            cudaEventRecord(begin, ctx->stream);
        }
    }
    ~KernelScope()
    {
        if (begin != nullptr)
        {
            cudaEventRecord(end, ctx->stream);
            ctx->timers.push_back({name, begin, end});
        }
    }
    static cudaEvent_t take(fgpu_ctx* c)
    {
        cudaEvent_t e = nullptr;
        if (!c->event_pool.empty())
        {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        }
        else
        {
            cudaEventCreate(&e);
        }
        return e;
    }
};
} // namespace fgpu

struct fgpu_grid
{
    float r_search = -1.0f; // radius the grid is conservative for (< 0: not built)
    int dim[3] = {0, 0, 0};
    uint32_t n_cells = 0;
    int ambiguous[3] = {0, 0, 0}; // dim < 3 on a periodic axis: several images may map to one cell
    fgpu::DevBuf<int> any_shift_flag; // device flag: some point lies outside the box (image offset != 0)
    fgpu::DevBuf<uint32_t> cell_of;    // per point (sharded build: the compact {index, cell} list of the slab instead)
    bool cell_of_valid = false;        // cell_of is indexed by point (k_count_evals reads it)
    fgpu::DevBuf<uint32_t> rank_in;    // per point: arrival rank inside its cell (sharded build: per list entry)
    fgpu::DevBuf<uint32_t> cell_start; // n_cells + 1
    fgpu::DevBuf<float4> sorted;       // cell-ordered positions, w = bit pattern of the point index
    fgpu::DevBuf<float4> slab_pos;     // sharded build: the slab's points {x, y, z, bits(index)} in arrival order
    fgpu::DevBuf<int> shift;           // cell-ordered packed integer image offsets (10 bits per axis, biased)
    int shard = 0, n_shards = 1;       // the slab this list was built for (fgpu_points_set_shard)
};

struct fgpu_points
{
    fgpu_ctx* ctx = nullptr;
    fgpu::BoxDev box;
    float plane_dist[3];
    uint32_t n = 0;
    fgpu::DevBuf<float> xyz; // original order, n x 3
    fgpu::DevBuf<float4> xyz4; // original order, padded; built on first use by Steinhardt (xyz4_ready)
    bool xyz4_ready = false;
    fgpu_grid grid;
    int shard = 0, n_shards = 1; // > 1: this rank searches one share of the home tiles (self-query RDF only)
};

struct fgpu_nlist : NlistStorage
{
    fgpu_ctx* ctx = nullptr;
    uint64_t n_bonds = 0;
    uint32_t n_query = 0;
    uint32_t n_points = 0;
    bool unit_weights = false; // built by a query: every weight is 1 (NeighborQuery.h:470-478)
    uint32_t q_index_offset = 0; // built by a query of a contiguous shard of the points: row k is point k + offset
    // fgpu_nlist_copy_begin / _wait: one event behind each array's copy, and the ones still to be written on the host
    mutable cudaEvent_t copy_done[4] = {nullptr, nullptr, nullptr, nullptr};
    mutable float* pending_unit_weights = nullptr;
};

// Peer-memory reduction state of one RDF (fgpu_rdf_attach_comm, peer.cuh)
struct fgpu_rdf_peer
{
    fgpu_comm* comm = nullptr;
    uint32_t* mailbox = nullptr;           // this rank's mailbox (cudaMalloc: stream-ordered pool memory has no IPC handle)
    void* opened[fgpu::kMaxPeers] = {};    // peers' mailboxes as opened here (nullptr for the own rank)
    fgpu::PeerBox box;                     // parity is set per epoch
    uint64_t epoch = 0;
    bool pushed = false;                   // the counts of the current epoch are already on their way (fused push)
};

struct fgpu_rdf
{
    fgpu_ctx* ctx = nullptr;
    fgpu_rdf_peer* peer = nullptr;
    fgpu::AxisDev axis;
    fgpu::DevBuf<uint32_t> hist;    // this rank's counts (+ one sticky flag word), never touched by a reduction
    fgpu::DevBuf<uint32_t> reduced; // sum over the ranks, valid from fgpu_rdf_allreduce until the next accumulate / reset
    bool reduced_valid = false;
};

struct fgpu_pmft;
struct fgpu_pmftxy
{
    fgpu_pmft* inner = nullptr; // kind FGPU_PMFT_XY
};

struct fgpu_pmft
{
    fgpu_ctx* ctx = nullptr;
    int kind = 0; // FGPU_PMFT_*
    fgpu::AxisDev a0, a1, a2;
    fgpu::DevBuf<uint32_t> hist;      // n0 * n1 * n2, row-major (axis 0 slow)
    fgpu::DevBuf<float> stage_a;      // staged per call: XYT (cos, sin)(-theta_i) | R12 theta_i | XYZ query quaternions
    fgpu::DevBuf<float> stage_b;      // staged per call: theta_j of the points | XYZ equivalent orientations
    fgpu::DevBuf<uint4> deferred;     // bonds whose angle bin is left to the host's libm: (i, j, bits vx, bits vy)
    fgpu::DevBuf<float> deferred_dist;
    fgpu::DevBuf<uint32_t> host_bins; // ... and the bins the host found for them
    fgpu::DevBuf<uint32_t> frame_hist; // fused path: the histogram of the frame being accumulated
    uint64_t deferred_total = 0;      // statistics: bonds the host binned since the last reset
};

struct fgpu_bondorder
{
    fgpu_ctx* ctx = nullptr;
    int mode = 0; // FGPU_BOND_ORDER_*
    fgpu::AxisDev at, ap;
    fgpu::DevBuf<uint32_t> hist;           // n_theta * n_phi, row-major (theta slow)
    fgpu::DevBuf<float> stage_a, stage_b;  // staged per call: orientations of the points | of the query points
    fgpu::DevBuf<uint4> deferred;          // bonds left to the host's libm: (i, j, bits vx, bits vy) ...
    fgpu::DevBuf<float> deferred_z;        // ... and vz
    fgpu::DevBuf<uint32_t> host_bins;
    fgpu::DevBuf<uint32_t> frame_hist;     // one-call route: the histogram of the frame being accumulated
    uint64_t deferred_total = 0;
};

struct fgpu_corr
{
    fgpu_ctx* ctx = nullptr;
    fgpu::AxisDev axis;
    fgpu::DevBuf<uint32_t> counts; // bins
    fgpu::DevBuf<double> sums;     // bins x (re, im)
    fgpu::DevBuf<double> values, query_values; // staged per call
};

struct fgpu_buffer
{
    fgpu_ctx* ctx = nullptr;
    uint64_t bytes = 0;
    fgpu::DevBuf<float> data;
};

struct fgpu_comm
{
    fgpu_ctx* ctx = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0;
    int size = 1;
    fgpu::DevBuf<unsigned char> stage;
};

namespace fgpu {

// ---- launchers (one per kernel family; all enqueue on ctx->stream) ---------------------------------
enum SearchMode
{
    SEARCH_COUNT = 0,
    SEARCH_FILL = 1,
    SEARCH_RDF = 2
};

struct GridDev
{
    int dx, dy, dz;
    int amb_x, amb_y, amb_z;
    const int* any_shift_flag; // device flag, see fgpu_grid
    const uint32_t* cell_start;
    const float4* sorted;
    const int* shift;
};

struct SearchArgs
{
    BoxDev box;
    GridDev grid;
    const float4* q_sorted; // cell-ordered queries (w = original index)
    uint32_t n_query;
    uint32_t q_index_offset;
    float r_max, r_min;
    int exclude_ii;
    int sort_by_distance;
    // COUNT
    uint32_t* row_counts;
    unsigned long long* total; // u64 bond total (overflow detection)
    // FILL
    const uint32_t* row_start;
    uint4* bag;
    // RDF
    AxisDev axis;
    uint32_t* hist;
    // instrumentation
    unsigned long long* evals; // may be nullptr
    // run only if *only_if != 0 (nullptr: always): device-side fallback behind the warp-cooperative kernel
    const int* only_if;
};

// in place, n elements; scratch_is_zero: ctx->scan_tmp (scan_scratch_words(n) words) was zeroed by the caller's kernel
// head / split: the n logical elements are `split` elements at `head` (16-byte aligned, a multiple of 8 long)
// followed by n - split elements at `data`
void exclusive_scan_u32(fgpu_ctx* ctx, uint32_t* data, size_t n, bool scratch_is_zero = false, uint32_t* head = nullptr,
                        size_t split = 0);
size_t scan_scratch_words(size_t n);
void build_grid(fgpu_points* pts, float r_search, bool force_single_cell = false);
void launch_check_2d_z(fgpu_ctx* ctx, const float* xyz, uint32_t n, int* flag); // *flag = 1 if some |z| > 1e-6
GridDev grid_dev(const fgpu_points* pts);
// cell-sorts arbitrary query points with the grid of pts; result in ctx->q_sorted
void sort_queries(fgpu_points* pts, const float* q_dev, uint32_t n_query);
void launch_search(fgpu_ctx* ctx, int flavour, SearchMode mode, const SearchArgs& args);

struct EmitArgs
{
    BoxDev box;
    const float4* sorted;
    const float* q_xyz; // original order n_query x 3
    const uint4* bag;
    const uint32_t* row_start;
    uint64_t n_bonds;
    float r_max, r_min;
    uint32_t* neighbors;
    float* distances;
    float* weights;
    float* vectors;
};
void launch_emit(fgpu_ctx* ctx, int flavour, const EmitArgs& args);
void launch_segments(fgpu_ctx* ctx, const uint32_t* row_start, const uint32_t* counts, uint32_t* segments,
                     uint32_t n_query);

// ---- warp-cooperative search (search2.cu): the production path on regular grids -------------------------
enum Search2Mode
{
    S2_NL = 0,
    S2_RDF = 1
};

struct Search2Args
{
    BoxDev box;
    int dx, dy, dz;
    uint32_t n_cells;
    int lanes_over_queries;         // NeighborList mode: 32 cell-ordered queries per ticket, one per lane (search_lq.cu)
    uint32_t n_query;               // queries in q_sorted
    uint32_t lq_tickets;            // lanes over queries: ceil(n_query / 32)
    uint32_t lq_c, lq_f;            // ... filter survivors a lane / a warp can hold (dynamic shared memory)
    int span;                       // cells per home tile along x (search2_plan)
    uint32_t spans_per_row;
    uint32_t n_tickets;             // work items: one home tile each
    uint32_t ticket_begin;          // this launch handles tickets [ticket_begin, ticket_end) (ticket_end == 0: all)
    uint32_t ticket_end;
    uint32_t cell_begin, cell_end;  // the cells those tickets cover (count_evals)
    const uint32_t* cell_start;     // candidates: cell list of the reference points
    const float4* sorted;
    const uint32_t* q_cell_start;   // queries, cell-sorted on the same grid
    const float4* q_sorted;
    const int* flag_points_outside;  // device flags: some point / query lies outside the box
    const int* flag_queries_outside;
    uint32_t q_index_offset;
    const uint32_t* q_remap;        // not null: the queries are a subset, q_remap[i] is the row of subset query i
    uint32_t tmp_flag;              // or-ed into tmp_start (kSecondBag when the bag is the second one)
    float r_max, r_min;
    int exclude_ii;
    float rcp_lx, rcp_ly, rcp_lz;   // RN(1 / L), rounded on the host
    float r_hi_sq;                  // stage-1 acceptance bound (WRAP)
    int symmetric;                  // IMAGE + RDF, queries == points: walk every boundary-free pair of rows once
                                    // and count its hits twice (tile_walk.cuh)
    float knn_r_min;                // > 0 (IMAGE + NL, kNN only): also reject sqrt(r_sq) < knn_r_min (AABBQuery.cc:213)
    // NeighborList mode
    float4* bag;                    // bag: {vector, bits(point index)} of every hit, rows contiguous
    uint32_t temp_cap;
    uint32_t out_cap;               // hit records a warp can buffer per batch (dynamic shared memory)
    uint32_t* counts;               // per query (original order)
    uint32_t* counts_copy;          // not null: the same values again (the array the row scan runs on: no D2D copy)
    uint32_t* zero_words;           // not null: block 0 clears zero_n words there (the row scan's scratch) and ...
    uint32_t zero_n;
    uint32_t* zero_tail;            // ... *zero_tail (the element closing the row scan), instead of three memsets
    uint32_t* tmp_start;            // per query: offset of its row in the bag
    unsigned long long* cursor;     // bag records reserved so far (== total bonds at the end)
    int* fail;                      // != 0: the result cannot be represented by this path (1: points outside the box,
                                    // 2: a tile has more candidates than out_cap; fail[1] = that count)
    // RDF mode
    AxisDev axis;
    uint32_t* hist;
    // scheduling / instrumentation
    unsigned int* work_counter;
    unsigned long long* evals;      // may be nullptr
    // RDF mode, multi-GPU: the last block to finish pushes the finished histogram into every rank's mailbox (peer.cuh)
    int push;
    unsigned int* done_counter;     // blocks that have merged their histogram (zeroed with work_counter)
    PeerBox peer;
};
void search2_plan(Search2Args& a, uint32_t n_points, int span_override = 0); // sets span, spans_per_row, n_tickets
// NeighborList mode: tile walk or lanes over queries (force: -1 automatic, 0 / 1 fgpu_ctx_set_tuning); sets n_query
void search2_choose_mapping(Search2Args& a, int flavour, uint32_t n_query, uint32_t n_points,
                            double expected_hits_per_query, int force);
// the lanes-over-queries launch gave up (fail == 2: a row or a warp's rows beyond its buffers): back to the tile walk
bool search2_lq_fallback(Search2Args& a, int fail);
void launch_search_lq(fgpu_ctx* ctx, int flavour, const Search2Args& a);

// Share of one rank when the home tiles of a self query are dealt to n_shards ranks: a contiguous run of tickets,
// the cells they cover, and the slab of cell layers (with one halo layer on each side) their candidates live in.
struct ShardPlan
{
    uint32_t ticket_begin, ticket_end;
    uint32_t cell_begin, cell_end;
    int slab_axis, slab_lo, slab_len; // slab_len < 0: every layer
};
ShardPlan shard_plan(const int dim[3], uint32_t n_points, int shard, int n_shards);
bool search2_supported(const Search2Args& a, int mode);
void launch_search2(fgpu_ctx* ctx, int flavour, int mode, const Search2Args& a);
uint32_t search2_out_cap(double expected_candidates_per_query);
uint32_t search2_max_out_cap();
// adds the pair evaluations of the query set to *a.evals (no-op when a.evals == nullptr)
void launch_count_evals(fgpu_ctx* ctx, const Search2Args& a, uint32_t n_query, const uint32_t* cell_of_point,
                        uint32_t n_points);

struct Emit2Args
{
    const float4* bag;
    const uint32_t* tmp_start;
    const uint32_t* row_start; // n_query + 1
    const uint32_t* counts;    // n_query
    uint32_t* segments;        // n_query: counts ? row_start : 0 (NeighborList.cc:199-232), written here too
    uint32_t n_query;
    // launched before the host knows how the search went: the kernel returns at once if the search gave up
    // (*fail), if its bag overflowed (*cursor > bag_cap) or if the output arrays are too small (> out_cap)
    const int* fail;
    const unsigned long long* cursor;
    uint64_t bag_cap, out_cap;
    uint32_t* neighbors;
    float* distances;
    float* weights;
    float* vectors;
};
void launch_emit2(fgpu_ctx* ctx, int sort_by_distance, const Emit2Args& a);

struct KnnArgs
{
    BoxDev box;
    GridDev grid;
    int flavour;
    const float4* q_sorted;
    uint32_t n_query;
    uint32_t q_index_offset;
    uint32_t k;
    float r_max, r_min;
    float r_safe; // every point closer than this is inside the visited cells
    int exclude_ii;
    int cover_all; // grid visits every point with every image: nothing can be unresolved
    float* knn_d;
    uint32_t* knn_s;
    uint32_t* row_counts;
    unsigned long long* unresolved;
    unsigned long long* total;
    unsigned long long* evals;
};
void launch_knn(fgpu_ctx* ctx, const KnnArgs& args);

struct KnnEmitArgs
{
    BoxDev box;
    int flavour;
    const float4* sorted;
    const float* q_xyz;
    const float* knn_d;
    const uint32_t* knn_s;
    const uint32_t* row_start;
    const uint32_t* row_counts;
    uint32_t n_query;
    uint32_t k;
    int sort_by_distance;
    uint32_t* neighbors;
    float* distances;
    float* weights;
    float* vectors;
};
void launch_knn_emit(fgpu_ctx* ctx, const KnnEmitArgs& args);

// ---- kNN on the warp-cooperative search (knn2.cu): ball search into the bag, then per-row selection ---------
struct KnnRowsArgs
{
    const uint32_t* hits;        // per query: hits inside the window (bag row length)
    uint32_t n_query;
    uint32_t k;
    int final;                   // the window already is r_max: short rows are complete
    uint32_t* counts;            // per query: min(hits, k)
    uint32_t* row_start;         // same values, scanned in place by the caller
    unsigned long long* unresolved;
    unsigned long long* total;
    uint32_t* unresolved_rows;   // n_query slots: the rows counted in *unresolved, in any order
};
void launch_knn_rows(fgpu_ctx* ctx, const KnnRowsArgs& a);
void launch_gather_points(fgpu_ctx* ctx, const float* xyz, const uint32_t* rows, uint32_t n_rows, float* out);
constexpr uint32_t kSecondBag = 0x80000000U; // flag in tmp_start: the row lives in the second bag

struct KnnSelectArgs
{
    const float4* bag;
    const float4* bag2;          // rows whose tmp_start carries kSecondBag
    const uint32_t* tmp_start;   // per query: offset of its row in the bag
    const uint32_t* hits;        // per query: bag row length
    const uint32_t* row_start;   // n_query + 1, output offsets
    uint32_t n_query;
    uint32_t k;
    uint32_t* neighbors;
    float* distances;
    float* weights;
    float* vectors;
};
void launch_knn_select(fgpu_ctx* ctx, int sort_by_distance, const KnnSelectArgs& a);

void launch_rdf_from_distances(fgpu_ctx* ctx, const float* distances, uint64_t n, AxisDev axis, uint32_t* hist);
#define FGPU_PMFT_XY 3 // internal: PMFTXY as a three-axis histogram whose last axis has one bin (fgpu_pmftxy_*)
struct Pmft3Args
{
    AxisDev a0, a1, a2;
    const uint32_t* neighbors;
    const float* vectors;
    const float* distances;
    uint64_t n_bonds;
    // ... or the bonds still in the search's bag, grouped by query row (no NeighborList): bag != nullptr
    const float4* bag;
    const uint32_t* row_bag_start;
    const uint32_t* row_counts;
    uint32_t n_rows;
    uint32_t group;                  // lanes per row: 4, 8 or 32
    const float* orientations;       // XYT, R12: per point
    const float* query_orientations; // XY, XYT, R12: per query point
    const float4* query_quats;       // XYZ: per query point, (s, x, y, z)
    const float4* equiv_quats;       // XYZ
    uint32_t n_equiv;
    uint32_t* hist;
    uint4* deferred;
    float* deferred_dist;
    uint32_t deferred_cap;
    uint32_t* deferred_count;
    int use_shared;            // 0: global atomics, 1: the block's shared memory, 2: slices over a cluster (pmft.cu)
    uint32_t slice;            // bins per CTA of the cluster
};
void launch_pmft3(fgpu_ctx* ctx, int kind, Pmft3Args a);
void launch_add_bins(fgpu_ctx* ctx, const uint32_t* bins, uint32_t n, uint32_t* hist);
void launch_add_hist(fgpu_ctx* ctx, const uint32_t* frame, uint32_t n, uint32_t* hist);
struct BondOrderArgs
{
    AxisDev at, ap;
    int mode;
    const uint32_t* neighbors;
    const float* vectors;
    uint64_t n_bonds;
    // ... or the bonds still in the search's bag, grouped by query row: bag != nullptr (as Pmft3Args)
    const float4* bag;
    const uint32_t* row_bag_start;
    const uint32_t* row_counts;
    uint32_t n_rows;
    uint32_t group;
    const float4* orientations;       // per point, (s, x, y, z)
    const float4* query_orientations; // per query point
    uint32_t* hist;
    uint4* deferred;
    float* deferred_z;
    uint32_t deferred_cap;
    uint32_t* deferred_count;
    int use_shared;
};
void launch_bond_order(fgpu_ctx* ctx, BondOrderArgs a);
void launch_correlation(fgpu_ctx* ctx, const uint32_t* neighbors, const float* distances, uint64_t n_bonds,
                        const double* values, const double* query_values, AxisDev axis, uint32_t* counts, double* sums);
void launch_local_density_rows(fgpu_ctx* ctx, const float4* bag, const uint32_t* row_bag_start, const uint32_t* row_counts,
                               uint32_t n_query, float r_max, float diameter, float measure, float* num_neighbors,
                               float* density);
void launch_correlation_rows(fgpu_ctx* ctx, const float4* bag, const uint32_t* row_bag_start, const uint32_t* row_counts,
                             uint32_t n_query, const double* values, const double* query_values, AxisDev axis,
                             uint32_t* counts, double* sums);
void launch_local_density(fgpu_ctx* ctx, const uint32_t* row_start, const float* distances, uint32_t n_query, float r_max,
                          float diameter, float measure, float* num_neighbors, float* density);

struct SteinhardtArgs
{
    BoxDev box;
    float rcp_lx, rcp_ly, rcp_lz; // RN(1 / L), rounded on the host (wrap_quick)
    const float* xyz;    // original order, n_points x 3
    const float4* xyz4;  // the same, padded to 16 bytes (fgpu_points::xyz4)
    uint32_t n;
    uint32_t row_offset; // row k of the list is particle k + row_offset (a rank's shard of the rows)
    const uint32_t* neighbors;
    const float* distances;
    const float* weights;
    const uint32_t* row_start;
    int weighted;
    uint32_t n_total;
    float* ql;       // n x n_ls
    float* qlm;      // concatenated per l
    double* sys_qlm; // concatenated per l, fp64 accumulators (re, im)
    double* sys_partials; // set by the launcher: one row of block sums per block (single-l kernel)
};
void launch_steinhardt(fgpu_ctx* ctx, const SteinhardtArgs& args, const std::vector<uint32_t>& ls);
void launch_pad_positions(fgpu_ctx* ctx, const float* xyz, uint32_t n, float4* out);
// k nearest of every bag row -> Y_lm, one kernel (steinhardt.cu k_knn_ylm); single l in {2, 4, .., 12}, k <= 16
bool knn_ylm_supported(const std::vector<uint32_t>& ls, uint32_t k);
void launch_knn_ylm(fgpu_ctx* ctx, const SteinhardtArgs& a, uint32_t l, const KnnSelectArgs& src);

// follow-up kernels over the per-particle q_lm array (the l tables of launch_steinhardt must be resident)
struct SteinhardtAveArgs
{
    uint32_t n;
    const uint32_t* neighbors;
    const uint32_t* row_start;
    const float* qlm;     // per l blocks, complex64[n][2l+1]
    float* qlm_ave;       // same layout
    float* ql_ave;        // n x n_ls
    double* sys_qlm;      // fp64 accumulators of the averaged q_lm (m >= 0), may be nullptr
    double* sys_partials; // set by the launcher: per-block partial sums
};
void launch_steinhardt_average(fgpu_ctx* ctx, const SteinhardtAveArgs& a, int n_ls, uint32_t tot_m);

struct SteinhardtWlArgs
{
    uint32_t n;
    const float* qlm;        // source: q_lm or the averaged q_lm
    const float* ql;         // normalisation source: q_l or the averaged q_l (n x n_ls)
    const float* w3j;        // concatenated Wigner 3j tables
    const uint32_t* w3j_off; // n_ls offsets into w3j (device)
    int normalize;
    float* wl;               // n x n_ls
};
void launch_steinhardt_wl(fgpu_ctx* ctx, const SteinhardtWlArgs& a, int n_ls);
std::vector<float> wigner3j_table(uint32_t l);

void launch_rdf_push(fgpu_ctx* ctx, const PeerBox& pb, const uint32_t* hist, uint32_t bins);
void launch_rdf_wait(fgpu_ctx* ctx, const PeerBox& pb, uint32_t bins, uint32_t* reduced, int* timeout);

// NCCL (loaded with dlopen)
int nccl_available(std::string* why);

} // namespace fgpu
