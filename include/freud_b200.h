/* freud_b200 -- C ABI of the B200-native neighbour-query + pair-accumulation path.
 *
 * The reference (glotzerlab/freud @ e4272dbe) has no C ABI or plugin interface for this path: the
 * boundary is the set of C++ methods its nanobind modules bind.  This header is
 * the thin C layer those methods are re-implemented on (freud_b200/host/ holds the C++ classes with the
 * reference's signatures; INTEGRATION.md shows the binding a maintainer would add).  Each entry point
 * cites the reference method it replaces.
 *
 * Conventions
 *   - every function returns FGPU_OK (0) or a negative FGPU_E* code; fgpu_last_error() returns the
 *     message of the last failure on the calling thread.  The C++ layer maps codes to the reference's
 *     exception types (INVALID -> std::invalid_argument, DOMAIN -> std::domain_error,
 *     RUNTIME/CUDA/NCCL -> std::runtime_error).
 *   - plain pointers and sizes only; "host" pointers are ordinary (pageable or pinned) host memory,
 *     "dev" pointers are device memory on the context's GPU.
 *   - points are (n, 3) C-contiguous float32, exactly the layout freud's bindings accept
 *     (freud/locality/export-NeighborQuery.cc:19-31).
 *   - box6 = {Lx, Ly, Lz, xy, xz, yz}; is2d != 0 selects freud's 2-D box (Lz ignored, z forced to 0).
 *   - all arithmetic on pair distances is un-fused IEEE float32 in the reference's operation order, so
 *     neighbour lists and RDF bin counts are bit-identical to the reference's x86-64 build.
 *   - there is no CPU fallback: without a CUDA device every call fails with FGPU_ECUDA.
 */
#ifndef FREUD_B200_H
#define FREUD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGPU_OK 0
#define FGPU_EINVALID (-1) /* std::invalid_argument upstream */
#define FGPU_EDOMAIN (-2)  /* std::domain_error upstream    */
#define FGPU_ERUNTIME (-3) /* std::runtime_error upstream   */
#define FGPU_ECUDA (-4)    /* CUDA runtime failure / no device */
#define FGPU_ENCCL (-5)    /* NCCL failure / library missing */
#define FGPU_ENOMEM (-6)
#define FGPU_ERANGE (-7)   /* std::out_of_range upstream    */

/* Pair-distance arithmetic ("flavour"), SURVEY.md fact 2 */
#define FGPU_FLAVOUR_WRAP 0  /* LinkCell:  r = Box::wrap(p_j - q)        freud/locality/LinkCell.cc:522 */
#define FGPU_FLAVOUR_IMAGE 1 /* AABBQuery: r = p_j - (q + image_k)       freud/locality/AABBQuery.cc:93,125 */
#define FGPU_FLAVOUR_GHOST 2 /* CellQuery: r = (p_j + shift_k) - q       freud/locality/CellQuery.cc:107,
                                freud/locality/CellIterator.h:167; ball queries only, every point inside the box */

typedef struct fgpu_ctx fgpu_ctx;       /* one GPU + one stream + scratch memory            */
typedef struct fgpu_points fgpu_points; /* device-resident reference points + box + cell list */
typedef struct fgpu_nlist fgpu_nlist;   /* device-resident NeighborList (SoA, CSR)            */
typedef struct fgpu_rdf fgpu_rdf;       /* device-resident RDF histogram accumulator          */
typedef struct fgpu_pmftxy fgpu_pmftxy; /* device-resident PMFTXY histogram                      */
typedef struct fgpu_pmft fgpu_pmft;     /* device-resident PMFTXYZ / PMFTXYT / PMFTR12 histogram  */
typedef struct fgpu_bondorder fgpu_bondorder; /* device-resident BondOrder histogram              */
typedef struct fgpu_corr fgpu_corr;     /* device-resident CorrelationFunction accumulators  */
typedef struct fgpu_comm fgpu_comm;     /* NCCL communicator (one rank per process / GPU)     */
typedef struct fgpu_buffer fgpu_buffer; /* a result array left on the device for a later read  */

const char* fgpu_last_error(void);
/* library + device description, e.g. "freud_b200 0.1 sm_100a NVIDIA B200 148 SMs"; safe without a GPU */
const char* fgpu_version(void);
int fgpu_device_count(void);

/* ---- context -------------------------------------------------------------------------------------- */
int fgpu_ctx_create(int device, fgpu_ctx** out);
void fgpu_ctx_destroy(fgpu_ctx* ctx);
/* Hands the context's grow-only scratch (hit bags, the arrays kept from the last destroyed NeighborList, kNN scratch)
 * back to the device's memory pool and trims the pool: after a frame far larger than the ones to come.  Everything is
 * re-grown on demand; live objects (points, lists, histograms) are untouched. */
int fgpu_ctx_trim(fgpu_ctx* ctx);
int fgpu_ctx_synchronize(fgpu_ctx* ctx);
/* the context's cudaStream_t, for callers that time with CUDA events on the launching stream */
void* fgpu_ctx_stream(fgpu_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t fgpu_ctx_launch_count(fgpu_ctx* ctx);
/* cumulative pair-distance evaluations ("pair evals") performed by search kernels launched since the last
 * reset; counted on the device.  Enabled with fgpu_ctx_count_pair_evals(ctx, 1) (off by default: the
 * counting variant costs one extra atomic per block). */
int fgpu_ctx_count_pair_evals(fgpu_ctx* ctx, int enable);
int fgpu_ctx_pair_evals(fgpu_ctx* ctx, uint64_t* out, int reset);

/* Two kernel families implement the ball search: the warp-cooperative one (regular grids with >= 3 cells per
 * periodic axis and all points inside the box -- the normal case) and a general thread-per-query one that
 * also covers tiny boxes, points outside the box and rows longer than the warp buffer.  The choice is
 * automatic; enable != 0 forces the general family (the parity tests run every family against the oracle). */
int fgpu_ctx_force_general_search(fgpu_ctx* ctx, int enable);

/* Experiment / test hooks (the library reads no environment variables).  Keys: "span" (> 0: cells per home tile of
 * the tile-walk search, 0: automatic), "no_symmetry" (1: self-query IMAGE RDF without the symmetric walk),
 * "lanes_over_queries" (NeighborList search mapping: -1 automatic, 0 tile walk, 1 one query per lane), "lq_blocks"
 * (> 0: resident blocks per SM of that mapping), "pmft_cluster" (1: PMFT histograms too large for one block's shared
 * memory count in the distributed shared memory of a thread-block cluster instead of with global atomics -- measured
 * 2.3x slower on B200, kept as a tested alternative).  Results never depend on them; the parity tests run every
 * mapping against the oracle. */
int fgpu_ctx_set_tuning(fgpu_ctx* ctx, const char* key, int value);

/* Per-kernel device timing with CUDA events on the context's stream (bench.py's roofline leg).  While enabled,
 * every kernel launch is bracketed by an event pair; fgpu_ctx_kernel_time synchronises and returns the summed
 * duration [ms] and launch count of the kernels whose name starts with `prefix` ("" = all) since the last
 * reset.  Names: cell_assign, cell_scatter, scan, search_nl, search_rdf, emit, segments, knn, knn_emit,
 * rdf_distances, local_density, local_density_rows, correlation, correlation_rows, pmft3, pmft3_rows, pmft_add_hist, bond_order, bond_order_rows, pmft_add_bins, steinhardt, steinhardt_average, steinhardt_wl, knn_rows, knn_select (general family: search_count, search_fill, search_rdf_general, emit_general). */
int fgpu_ctx_profile(fgpu_ctx* ctx, int enable);
int fgpu_ctx_kernel_time(fgpu_ctx* ctx, const char* prefix, double* ms_out, uint64_t* launches_out, int reset);
/* The same records as a timeline: one line "name begin_us end_us" per profiled launch since the last reset, times
 * relative to the first launch's begin (device clock), into `out` (NUL-terminated, truncated to `cap` bytes).  What
 * the idle time between dependent launches of a step is read from (tools/timeline.py). */
int fgpu_ctx_kernel_timeline(fgpu_ctx* ctx, char* out, uint64_t cap);

/* ---- points + periodic cell list -------------------------------------------------------------------
 * Replaces the constructors LinkCell(box, points, n, cell_width) freud/locality/LinkCell.cc:222-260,
 * AABBQuery(box, points, n) freud/locality/AABBQuery.cc:17-26 and RawPoints (RawPoints.h:35-37):
 * copies the points to the GPU.  The cell list itself (cell index, counting sort, prefix scan, cell-ordered
 * float4 positions; replaces LinkCell::computeCellList freud/locality/LinkCell.cc:316-336 and
 * AABBQuery::buildTree :53-69) is built on the first query, because its cell width follows r_max; it is
 * cached per (r_max) and rebuilt only when a query needs a different width.
 * Errors: n == 0 -> FGPU_EINVALID (NeighborQuery.h:97-100); 2-D box with |z| > 1e-6 -> FGPU_EINVALID
 * (NeighborQuery.h:103-112). */
int fgpu_points_create(fgpu_ctx* ctx, const float* box6, int is2d, const float* points_host, uint32_t n,
                       fgpu_points** out);
/* same, points already resident on the device (n x 3 float32) */
int fgpu_points_create_dev(fgpu_ctx* ctx, const float* box6, int is2d, const float* points_dev, uint32_t n,
                           fgpu_points** out);
/* same for a frame every rank of `comm` holds in host memory (one process per GPU, SURVEY.md section 8e: the points
 * are replicated): rank r uploads rows [r n / W, (r + 1) n / W) only and the ranks exchange their blocks over NVLink
 * (grouped ncclBroadcast on the context's stream), so the frame crosses PCIe once per node, not once per GPU.
 * Collective: every rank calls it with the same box, n and point values. */
int fgpu_points_create_replicated(fgpu_ctx* ctx, fgpu_comm* comm, const float* box6, int is2d, const float* points_host,
                                  uint32_t n, fgpu_points** out);
void fgpu_points_destroy(fgpu_points* pts);
/* Force the cell-list build for search radius r (what the first query would do); exposed so the build can
 * be timed and tested on its own.  out_dims[3] receives the cell grid, may be NULL. */
int fgpu_points_build_cells(fgpu_points* pts, float r_search, uint32_t* out_dims);
/* Multi-GPU self-query RDF (BASELINE.json configs[3]): the points are replicated on every
 * rank, the HOME TILES of the search -- hence the query points -- are dealt to the ranks in contiguous runs, and
 * rank `shard` of `n_shards` builds only the slab of the cell list its tiles can see (their cell layers plus one
 * halo layer on each side), so neither the search nor the build stays serial.  fgpu_rdf_accumulate with
 * query_points == NULL then bins the pairs of this rank's tiles; fgpu_rdf_allreduce sums the ranks.  Counts are
 * integers, so the total is bit-identical to the single-GPU histogram.  Sharded points serve that call only
 * (everything else returns FGPU_ERUNTIME), need a grid of >= 3 cells per periodic axis and every point inside
 * the box; n_shards == 1 restores the normal behaviour. */
int fgpu_points_set_shard(fgpu_points* pts, int shard, int n_shards);
/* The arithmetic behind it, host only (no device needed; tests/test_multirank_gloo.py): for a grid of dims[3] cells
 * holding n_points, out[8] = {ticket_begin, ticket_end, n_tickets, cell_begin, cell_end, slab_axis, slab_lo,
 * slab_len (0xffffffff: every layer)} of shard `shard` of `n_shards`. */
int fgpu_shard_plan(const uint32_t* dims, uint32_t n_points, int shard, int n_shards, uint32_t* out);
/* Copies out the cell list for tests: cell_start[n_cells + 1] and the point index of every cell-ordered
 * slot (order[n]).  Either pointer may be NULL. */
int fgpu_points_read_cells(fgpu_points* pts, uint32_t* cell_start_host, uint32_t* order_host);

/* ---- ball query -> NeighborList ---------------------------------------------------------------------
 * Replaces NeighborQuery::query(...)->toNeighborList(sort_by_distance) for mode == ball:
 * freud/locality/NeighborQuery.h:130-142, 434-481 with LinkCellQueryBallIterator::next
 * (freud/locality/LinkCell.cc:496-573, flavour WRAP) or AABBQueryBallIterator::next
 * (freud/locality/AABBQuery.cc:77-150, flavour IMAGE).
 * query_points_host == NULL means "the query points are the reference points themselves".
 * q_index_offset is added to the local query index when comparing against point indices for exclude_ii
 * and is NOT added to the emitted query index (rows are local); it lets one rank query a contiguous
 * shard of the points.
 * Errors: r_max <= 0 or r_max <= r_min -> FGPU_EINVALID (NeighborQuery.h:321-328); IMAGE flavour with
 * a plane distance <= 2 r_max -> FGPU_ERUNTIME (NeighborQuery.h:503-510); n_query == 0 yields an empty
 * list (tests/test_locality_neighbor_list.py:253-256). */
int fgpu_ball_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, uint32_t q_index_offset,
                    int flavour, float r_max, float r_min, int exclude_ii, int sort_by_distance, fgpu_nlist** out);
/* same with query points resident on the device */
int fgpu_ball_query_dev(fgpu_points* pts, const float* query_points_dev, uint32_t n_query,
                        uint32_t q_index_offset, int flavour, float r_max, float r_min, int exclude_ii,
                        int sort_by_distance, fgpu_nlist** out);

/* ---- k-nearest-neighbour query -> NeighborList ------------------------------------------------------
 * Replaces query(mode == nearest)->toNeighborList().
 *   FGPU_FLAVOUR_IMAGE: AABBQueryIterator::next, freud/locality/AABBQuery.cc:152-281 (E3 in SURVEY.md: the k
 *     smallest closest-image distances in the IMAGE arithmetic, d >= r_min, r_sq < r_max^2; r_guess/scale do
 *     not influence the result, tests/test_locality_neighbor_query.py:635-656).
 *   FGPU_FLAVOUR_WRAP: LinkCellQueryIterator::next, freud/locality/LinkCell.cc:575-679: the k smallest wrapped
 *     distances r = Box::wrap(p_j - q) with r_min^2 <= r_sq < r_max^2 (the cell width only steers its early
 *     exit, never the answer).
 * Ties at the k-th place are unspecified upstream (unstable std::sort on the distance) and resolved by
 * (r_sq, point index) here.  r_max may be INFINITY. */
int fgpu_knn_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, uint32_t q_index_offset,
                   int flavour, uint32_t num_neighbors, float r_max, float r_min, int exclude_ii,
                   int sort_by_distance, fgpu_nlist** out);

/* ---- NeighborList -----------------------------------------------------------------------------------
 * Layout = freud::locality::NeighborList (freud/locality/NeighborList.h:139-157): neighbors u32[nb][2],
 * distances f32[nb], weights f32[nb], vectors f32[nb][3], segments/counts u32[n_query]; bonds sorted by
 * (i, j) or (i, d, j) (NeighborBond.h:80-112); segments of empty rows are 0 (NeighborList.cc:199-232). */
uint64_t fgpu_nlist_num_bonds(const fgpu_nlist* nl);
uint32_t fgpu_nlist_num_query_points(const fgpu_nlist* nl);
uint32_t fgpu_nlist_num_points(const fgpu_nlist* nl);
/* device -> host copy of any subset (NULL pointers are skipped) */
int fgpu_nlist_copy(const fgpu_nlist* nl, uint32_t* neighbors_host, float* distances_host, float* weights_host,
                    float* vectors_host, uint32_t* segments_host, uint32_t* counts_host);
/* The same copy of the four bond arrays, split in two: _begin enqueues the transfers on the context's stream and
 * returns (truly asynchronous into page-locked memory, fgpu_host_alloc), _wait blocks until the arrays named in `which`
 * (bit 0 neighbors, 1 distances, 2 weights, 3 vectors) have landed -- unit weights of a query-built list are written
 * by host threads at that point instead of crossing the link.  The host class behind freud's NeighborList starts all
 * four on the first getter and waits per array: `nlist.distances` costs one array's latency, touching everything costs
 * one pass over the link with the ones filled meanwhile.  The destinations must stay valid until the list is destroyed
 * or every array was waited for. */
int fgpu_nlist_copy_begin(const fgpu_nlist* nl, uint32_t* neighbors_host, float* distances_host, float* weights_host,
                          float* vectors_host);
int fgpu_nlist_copy_wait(const fgpu_nlist* nl, unsigned which);
/* upload a host NeighborList (already sorted by query index) so RDF / Steinhardt can consume it:
 * replaces passing a NeighborList* into accumulate/compute (NeighborComputeFunctional.h:180-193, 121-135) */
int fgpu_nlist_from_host(fgpu_ctx* ctx, uint64_t n_bonds, uint32_t n_query, uint32_t n_points,
                         const uint32_t* neighbors_host, const float* distances_host, const float* weights_host,
                         const float* vectors_host, fgpu_nlist** out);
void fgpu_nlist_destroy(fgpu_nlist* nl);

/* ---- RDF ---------------------------------------------------------------------------------------------
 * Device half of freud::density::RDF (freud/density/RDF.cc:25-110) + BondHistogramCompute
 * (freud/locality/BondHistogramCompute.h:29-140): a resident u32[bins] histogram with
 * RegularAxis::bin semantics (freud/util/Histogram.h:126-174).  Normalisation to g(r), n(r)
 * (RDF::reduce, RDF.cc:73-99) is host arithmetic in freud_b200/host/RDF.cc. */
int fgpu_rdf_create(fgpu_ctx* ctx, uint32_t bins, float r_max, float r_min, fgpu_rdf** out);
void fgpu_rdf_destroy(fgpu_rdf* rdf);
int fgpu_rdf_reset(fgpu_rdf* rdf); /* BondHistogramCompute::reset :39-49 */
/* accumulate one frame by querying on the fly (no NeighborList is materialised):
 * RDF::accumulate with nlist == nullptr, RDF.cc:101-110 -> NeighborComputeFunctional.h:195-217.
 * The query window (q_r_max, q_r_min) and the histogram range (r_max, r_min of create) are independent,
 * as upstream. */
int fgpu_rdf_accumulate(fgpu_rdf* rdf, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                        uint32_t q_index_offset, int flavour, float q_r_max, float q_r_min, int exclude_ii);
int fgpu_rdf_accumulate_dev(fgpu_rdf* rdf, fgpu_points* pts, const float* query_points_dev, uint32_t n_query,
                            uint32_t q_index_offset, int flavour, float q_r_max, float q_r_min, int exclude_ii);
/* accumulate from an existing NeighborList: one increment per stored distance
 * (NeighborComputeFunctional.h:180-193) */
int fgpu_rdf_accumulate_nlist(fgpu_rdf* rdf, const fgpu_nlist* nl);
/* device -> host copy of the raw bin counts (BondHistogramCompute::getBinCounts :74-77) */
int fgpu_rdf_read(fgpu_rdf* rdf, uint32_t* counts_host);
/* Sum the histograms of all ranks: one ncclAllReduce(u32[bins], sum), OUT OF PLACE -- the
 * rank's own counts stay as they are, fgpu_rdf_read returns the sum until the next accumulate / reset.  So
 * accumulate -> allreduce -> read -> accumulate -> allreduce -> read (compute(..., reset=False) over a trajectory
 * with intermediate reads) never counts a frame twice. */
int fgpu_rdf_allreduce(fgpu_rdf* rdf, fgpu_comm* comm);
/* Peer-memory transport for that sum (collective over `comm`, once per RDF; all ranks on one node, <= 8): every rank
 * allocates a small mailbox, the CUDA IPC handles travel over NCCL, and from then on fgpu_rdf_allreduce /
 * fgpu_rdf_accumulate_reduce add the rank's counts straight into every rank's mailbox with red.add over NVLink and
 * wait on the stream until all ranks' counts are in (freud_b200/csrc/peer.cuh) -- no NCCL launch on the data path.
 * Returns FGPU_OK when attached, 1 when peer access / CUDA IPC is unavailable (the RDF stays on the NCCL route; every
 * rank gets the same answer).  Destroy the RDF before its communicator. */
int fgpu_rdf_attach_comm(fgpu_rdf* rdf, fgpu_comm* comm);
/* 1: reductions of this RDF go through ncclAllReduce, 2: through the peer mailbox */
int fgpu_rdf_reduce_transport(const fgpu_rdf* rdf);
/* BASELINE.json configs[3] in one call: self-query accumulation of (sharded, fgpu_points_set_shard) points AND the sum
 * over the ranks.  With a peer mailbox the search kernel itself is the collective: the block that merges its
 * histogram last pushes the finished counts to every rank (compute + exchange in one launch), and a one-block wait
 * kernel closes the epoch.  Afterwards fgpu_rdf_read returns the sum; the rank's own counts are untouched. */
int fgpu_rdf_accumulate_reduce(fgpu_rdf* rdf, fgpu_points* pts, fgpu_comm* comm, int flavour, float q_r_max,
                               float q_r_min, int exclude_ii);

/* ---- PMFTXY ----------------------------------------------------------------------------------------------
 * Device half of freud::pmft::PMFTXY (freud/pmft/PMFTXY.cc:25-87): a u32[n_x][n_y] histogram of the bond vectors
 * of a NeighborList rotated into the frame of their query particle, resident across accumulate calls.
 * query_orientations_host[n_query] are angles in radians.  The rotation uses cosf / sinf of the host libm upstream
 * (rotmat2::fromAngle, VectorMath.h:912-921) and CUDA's differ from them in the last place: the kernel rotates with
 * double sincos rounded to float and accepts a bin only when the rotated coordinate clears every bin edge by more
 * than a last-place change of (cos, sin) can move it; the other bonds (a few in 10^4) are binned by the host with
 * its libm (freud_b200/csrc/pmft.cu) -- the counts are identical to the reference's.  Normalisation to the PCF
 * (PMFT::reduce, freud/pmft/PMFT.h:73-83) is host arithmetic in freud_b200/host/PMFT.h.
 * Errors: n_x, n_y < 1 or x_max, y_max < 0 -> FGPU_EINVALID (PMFTXY.cc:27-42). */
int fgpu_pmftxy_create(fgpu_ctx* ctx, float x_max, float y_max, uint32_t n_x, uint32_t n_y, fgpu_pmftxy** out);
void fgpu_pmftxy_destroy(fgpu_pmftxy* pmft);
int fgpu_pmftxy_reset(fgpu_pmftxy* pmft);
int fgpu_pmftxy_accumulate_nlist(fgpu_pmftxy* pmft, const fgpu_nlist* nl, const float* query_orientations_host);
/* query + histogram in one call, no NeighborList (see fgpu_pmft_accumulate) */
int fgpu_pmftxy_accumulate(fgpu_pmftxy* pmft, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                           int flavour, float r_max, float r_min, int exclude_ii, const float* query_orientations_host);
int fgpu_pmftxy_read(fgpu_pmftxy* pmft, uint32_t* counts_host);

/* ---- PMFTXYZ, PMFTXYT, PMFTR12 ---------------------------------------------------------------------------
 * Device half of the three-axis PMFTs: a u32[n0][n1][n2] histogram over the bonds of a NeighborList, resident across
 * accumulate calls.
 *   FGPU_PMFT_XYZ  freud/pmft/PMFTXYZ.cc:24-147: axes x, y, z in [-max, max]; the bond vector rotated by
 *                  conj(query_orientations[i]) and by each of the n_equiv equivalent orientations (one count per
 *                  orientation); quaternions are (s, x, y, z) float[4]; `orientations` is not read.
 *   FGPU_PMFT_XYT  freud/pmft/PMFTXYT.cc:28-101: axes x, y in [-max, max] and t in [0, 2 pi); `orientations[n_points]`
 *                  and `query_orientations[n_query]` are angles in radians; max2 is not read.
 *   FGPU_PMFT_R12  freud/pmft/PMFTR12.cc:28-113: axes r in [0, max0], t1 and t2 in [0, 2 pi); max1, max2 not read.
 * Bin counts are bit-identical to the reference's: the float arithmetic runs on the GPU in the reference's operation
 * order; cosf / sinf of the query angles (XYT) and atan2f of the bond angle (XYT, R12) are bracketed on the GPU and
 * the few bonds whose bin could depend on libm's last place are binned by the host (freud_b200/csrc/pmft.cu).  fgpu_pmft_deferred reports how many bonds took that route since the last reset.
 * Errors: a bin count < 1 or a negative maximum -> FGPU_EINVALID with the reference's messages; XYZ with n_equiv = 0
 * or null quaternions, XYT / R12 with null angles -> FGPU_EINVALID. */
#define FGPU_PMFT_XYZ 0
#define FGPU_PMFT_XYT 1
#define FGPU_PMFT_R12 2
int fgpu_pmft_create(fgpu_ctx* ctx, int kind, float max0, float max1, float max2, uint32_t n0, uint32_t n1, uint32_t n2,
                     fgpu_pmft** out);
void fgpu_pmft_destroy(fgpu_pmft* pmft);
int fgpu_pmft_reset(fgpu_pmft* pmft);
int fgpu_pmft_accumulate_nlist(fgpu_pmft* pmft, const fgpu_nlist* nl, const float* orientations_host, uint32_t n_points,
                               const float* query_orientations_host, const float* equiv_orientations_host,
                               uint32_t n_equiv);
/* The ball query (fgpu_ball_query's arguments; query_points_host = NULL: the points themselves) and the histogram in
 * one call: the bonds go from the search's hit bag straight into the histogram kernel, no NeighborList is built
 * (NeighborComputeFunctional.h:137-178, the nlist == nullptr branch).  Same counts as the two-call route. */
int fgpu_pmft_accumulate(fgpu_pmft* pmft, fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour,
                         float r_max, float r_min, int exclude_ii, const float* orientations_host,
                         const float* query_orientations_host, const float* equiv_orientations_host, uint32_t n_equiv);
int fgpu_pmft_read(fgpu_pmft* pmft, uint32_t* counts_host);
int fgpu_pmft_deferred(const fgpu_pmft* pmft, uint64_t* bonds);

/* ---- BondOrder -------------------------------------------------------------------------------------------
 * Device half of freud::environment::BondOrder (freud/environment/BondOrder.cc:30-153): a u32[n_theta][n_phi]
 * histogram of bond directions over the bonds of a NeighborList, theta in [0, 2 pi), phi in [0, pi), resident across
 * accumulate calls.  mode = BondOrderMode (BondOrder.h:23-29).  orientations_host[n_points] and
 * query_orientations_host[n_query] are quaternions (s, x, y, z) float[4]; mode bod reads neither (both may be NULL).
 * Counts are bit-identical to the reference's: rotations in its float order on the GPU, atan2f / acosf bracketed on
 * the GPU with the bonds next to a bin edge binned by the host's libm (as for fgpu_pmft).  The normalisation by the
 * solid angle of the bins (BondOrder::reduce) is host arithmetic in freud_b200/host/BondOrder.h.
 * Errors: n_theta or n_phi < 2, unknown mode -> FGPU_EINVALID (BondOrder.cc:34-41). */
#define FGPU_BOND_ORDER_BOD 0
#define FGPU_BOND_ORDER_LBOD 1
#define FGPU_BOND_ORDER_OBCD 2
#define FGPU_BOND_ORDER_OOCD 3
int fgpu_bondorder_create(fgpu_ctx* ctx, uint32_t n_theta, uint32_t n_phi, int mode, fgpu_bondorder** out);
void fgpu_bondorder_destroy(fgpu_bondorder* bo);
int fgpu_bondorder_reset(fgpu_bondorder* bo);
int fgpu_bondorder_accumulate_nlist(fgpu_bondorder* bo, const fgpu_nlist* nl, const float* orientations_host,
                                    uint32_t n_points, const float* query_orientations_host);
/* the ball query and the histogram in one call, the bonds read from the search's hit bag (see fgpu_pmft_accumulate) */
int fgpu_bondorder_accumulate(fgpu_bondorder* bo, fgpu_points* pts, const float* query_points_host, uint32_t n_query,
                              int flavour, float r_max, float r_min, int exclude_ii, const float* orientations_host,
                              const float* query_orientations_host);
int fgpu_bondorder_read(fgpu_bondorder* bo, uint32_t* counts_host);
int fgpu_bondorder_deferred(const fgpu_bondorder* bo, uint64_t* bonds);

/* ---- CorrelationFunction ---------------------------------------------------------------------------------
 * Device half of freud::density::CorrelationFunction (freud/density/CorrelationFunction.cc:26-95): per bin of
 * RegularAxis(bins, 0, r_max) a u32 bond count and the complex<double> sum of conj(values[j]) * query_values[i]
 * over the bonds of a NeighborList, resident across accumulate calls (compute(..., reset=False)).  The division by
 * the counts (reduce, :49-59) is host arithmetic in freud_b200/host/CorrelationFunction.h.  Double sums are added in
 * no fixed order, as upstream's thread-local histograms are: agreement is to double rounding.
 * values_host: complex128[n_points]; query_values_host: complex128[n_query] (re, im interleaved). */
int fgpu_corr_create(fgpu_ctx* ctx, uint32_t bins, float r_max, fgpu_corr** out);
void fgpu_corr_destroy(fgpu_corr* corr);
int fgpu_corr_reset(fgpu_corr* corr);
int fgpu_corr_accumulate_nlist(fgpu_corr* corr, const fgpu_nlist* nl, const double* values_host,
                               const double* query_values_host);
/* query + accumulation in one call, no NeighborList (see fgpu_pmft_accumulate); identical bin counts, sums to double
 * rounding like every accumulation order */
int fgpu_corr_accumulate(fgpu_corr* corr, fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour,
                         float r_max, float r_min, int exclude_ii, const double* values_host,
                         const double* query_values_host);
int fgpu_corr_read(fgpu_corr* corr, uint32_t* counts_host, double* sums_host);

/* ---- LocalDensity --------------------------------------------------------------------------------------
 * Replaces LocalDensity::compute (freud/density/LocalDensity.cc:38-84) once the neighbours are a NeighborList (the
 * host class runs the ball query of the reference's default arguments, r_max + diameter / 2, when none is given):
 * per query point the fractional neighbour count -- summed in float in list order, so a list sorted like the one
 * upstream was handed gives the same bits -- and count / (pi r_max^2) in 2-D boxes or / (4/3 pi r_max^3).
 * Errors: r_max <= 0 or diameter < 0 -> FGPU_EINVALID (LocalDensity.cc:25-36).  Outputs: f32[n_query] each. */
int fgpu_local_density(const fgpu_nlist* nl, float r_max, float diameter, int is2d, float* num_neighbors_host,
                       float* density_host);
/* The ball query (q_r_max, q_r_min, exclude_ii; query_points_host = NULL: the points themselves) and the count in one
 * call, the bonds read from the search's hit bag instead of a NeighborList (the nlist == nullptr branch of
 * LocalDensity::compute).  Sums in bag order: agrees with the list route to float rounding. */
int fgpu_local_density_query(fgpu_points* pts, const float* query_points_host, uint32_t n_query, int flavour, float q_r_max,
                             float q_r_min, int exclude_ii, float r_max, float diameter, float* num_neighbors_host,
                             float* density_host);

/* ---- Steinhardt --------------------------------------------------------------------------------------
 * Replaces Steinhardt::compute (freud/order/Steinhardt.cc:85-118): baseCompute :120-222 with
 * fsph::PointSPHEvaluator (extern/fsph/src/spherical_harmonics.hpp:155-290), computeAve :224-289, aggregatewl
 * :329-359 with reduceWigner3j (freud/order/Wigner3j.cc:22-57) and normalizeSystem :291-327.  `flags` are the
 * constructor's booleans (Steinhardt.h:66-76).  The neighbour list must have been built against pts
 * (n_query == n_points == n).  Outputs (host pointers, any may be NULL):
 *   ql      f32[n][n_ls]   what getQl() returns: q_l, or the second-shell averaged q_l with FGPU_ST_AVERAGE
 *   wl      f32[n][n_ls]   w_l (FGPU_ST_WL only): from q_lm, or from the averaged q_lm with FGPU_ST_AVERAGE;
 *                          scaled by (sqrt(4 pi / (2l+1)) / q_l)^3 with FGPU_ST_WL_NORMALIZE
 *   qlm     per l, concatenated: complex64[n][2l+1] (re, im interleaved), m order 0..l, -1..-l; always the
 *           un-averaged q_lm(i), as getQlm()
 *   sys_qlm per l, concatenated: complex64[2l+1] = sum_i qlm_i / n (of the averaged q_lm with
 *           FGPU_ST_AVERAGE), accumulated in fp64 on the device and rounded once: a documented deviation, the
 *           reference's float32 thread-order sum is not reproducible, SURVEY.md section 7 "hard parts"
 *   order   f32[n_ls] getOrder(): system-wide q_l, or system-wide w_l with FGPU_ST_WL
 * Errors: FGPU_ST_WL with an l > 20 -> FGPU_ERANGE (Wigner3j.cc, "implemented for l <= 20").
 * comm != NULL: sys_qlm/order are reduced over all ranks (each rank holds a shard of the rows; n_total is
 * the global particle count used in the 1/N normalisation); FGPU_ST_AVERAGE needs every row on the rank. */
#define FGPU_ST_WEIGHTED 1
#define FGPU_ST_AVERAGE 2
#define FGPU_ST_WL 4
#define FGPU_ST_WL_NORMALIZE 8
int fgpu_steinhardt_compute(fgpu_points* pts, const fgpu_nlist* nl, const uint32_t* ls, uint32_t n_ls, int flags,
                            uint32_t n_total, fgpu_comm* comm, float* ql_host, float* wl_host, float* qlm_host,
                            float* sys_qlm_host, float* order_host);

/* The same, with the per-particle q_lm (the bulk of the output: 104 MB at l = 6, N = 1e6) left on the device:
 * *qlm_dev_out owns them until fgpu_buffer_destroy; fgpu_buffer_read copies any byte range out (layout as qlm above).
 * freud::order::Steinhardt::getQlm() (freud/order/Steinhardt.h:113-117) reads them on first access, so
 * compute(...).particle_order does not pay for arrays nobody asked for. */
int fgpu_steinhardt_compute_keep(fgpu_points* pts, const fgpu_nlist* nl, const uint32_t* ls, uint32_t n_ls, int flags,
                                 uint32_t n_total, fgpu_comm* comm, float* ql_host, float* wl_host,
                                 fgpu_buffer** qlm_dev_out, float* sys_qlm_host, float* order_host);
/* Steinhardt::compute(nlist = nullptr, points, {mode nearest, num_neighbors}) in one call (BASELINE.json configs[2];
 * loopOverNeighborsIterator's query-on-the-fly branch, NeighborComputeFunctional.h:137-178): the k-nearest-neighbour
 * search of the points themselves and the Y_lm sums over its result, with no NeighborList in between -- the window
 * search leaves every row's hits in device memory and one kernel picks the k nearest bond vectors of a row and
 * accumulates q_lm over them (freud_b200/csrc/steinhardt.cu k_knn_ylm).  A single l in {2, 4, ..., 12}, k <= 16 and
 * flags without FGPU_ST_AVERAGE / FGPU_ST_WL take that route; everything else (and frames the warp-cooperative search
 * does not take) builds the list and calls the kernels of fgpu_steinhardt_compute_keep: same results either way
 * (q_l to float summation order).  Outputs as fgpu_steinhardt_compute_keep; qlm_dev_out may be NULL. */
int fgpu_steinhardt_knn(fgpu_points* pts, int flavour, uint32_t num_neighbors, float r_max, float r_min, int exclude_ii,
                        const uint32_t* ls, uint32_t n_ls, int flags, float* ql_host, float* wl_host,
                        fgpu_buffer** qlm_dev_out, float* sys_qlm_host, float* order_host);
uint64_t fgpu_buffer_bytes(const fgpu_buffer* buf);
int fgpu_buffer_read(fgpu_buffer* buf, void* host, uint64_t offset_bytes, uint64_t bytes);
void fgpu_buffer_destroy(fgpu_buffer* buf);

/* ---- page-locked host memory ----------------------------------------------------------------------------
 * Backing store of util::ManagedArray (freud/util/ManagedArray.h:37-333) in freud_b200/host: blocks are page-locked
 * (cudaHostAlloc) once and cached by size class, so the arrays of every frame after the first cost no allocation and
 * device -> host copies into them run at the link's rate.  Without a CUDA device the blocks are ordinary aligned
 * memory (the container classes stay usable in host-only tests).  fgpu_host_free(NULL) is a no-op; fgpu_host_trim
 * releases the cache. */
int fgpu_host_alloc(uint64_t bytes, void** out);
void fgpu_host_free(void* p);
int fgpu_host_trim(void);

/* ---- multi-GPU plumbing (one process per GPU) ---------------------------------------------------------
 * NCCL is loaded at run time (libnccl.so.2); unique_id is NCCL's 128-byte ncclUniqueId, produced on rank 0
 * and distributed by the caller (bench.py uses torch.distributed / a file for that). */
#define FGPU_UNIQUE_ID_BYTES 128
int fgpu_comm_unique_id(uint8_t* unique_id_out);
int fgpu_comm_create(fgpu_ctx* ctx, const uint8_t* unique_id, int rank, int n_ranks, fgpu_comm** out);
void fgpu_comm_destroy(fgpu_comm* comm);
int fgpu_comm_rank(const fgpu_comm* comm);
int fgpu_comm_size(const fgpu_comm* comm);
int fgpu_comm_barrier(fgpu_comm* comm);
/* generic in-place sum over host buffers of u32 / f64 (staged through device memory, one ncclAllReduce) */
int fgpu_comm_allreduce_u32(fgpu_comm* comm, uint32_t* host_inout, uint64_t count);
int fgpu_comm_allreduce_f64(fgpu_comm* comm, double* host_inout, uint64_t count);

#ifdef __cplusplus
}
#endif
#endif /* FREUD_B200_H */
