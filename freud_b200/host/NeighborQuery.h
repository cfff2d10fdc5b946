// freud::locality::NeighborQuery and its engines, re-implemented on the C ABI of libfreud_b200.so.
//
// The class signatures are the ones the reference's binding layer calls
// (freud/locality/export-NeighborQuery.cc:26-134): constructors LinkCell(box, points, n, cell_width)
// (LinkCell.h:188), AABBQuery(box, points, n) (AABBQuery.h:42), RawPoints(box, points, n)
// (RawPoints.h:35-37); query(query_points, n, QueryArgs) -> NeighborQueryIterator (NeighborQuery.h:130-142);
// NeighborQueryIterator::toNeighborList(sort_by_distance) / next() (NeighborQuery.h:392-481).
// Argument validation, mode inference and error types follow NeighborQuery.h:195-292, 315-329.
//
// What differs is where the work happens: the constructor copies the points to the GPU (the reference keeps
// a raw host pointer), query() stays lazy exactly as upstream, and toNeighborList() runs the cell-list build,
// the 27-cell search and the sorted CSR emit on the device.  The engine type only selects the float32
// arithmetic ("flavour") whose results it reproduces bit for bit: LinkCell -> Box::wrap(p_j - q),
// AABBQuery / RawPoints -> p_j - (q + image).
#pragma once
#include <atomic>
#include <cstring>
#include <thread>
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>
#include <algorithm>

#include "Box.h"
#include "Context.h"
#include "NeighborList.h"
#include "VectorMath.h"

namespace freud { namespace locality {

enum class QueryType
{
    none,
    ball,
    nearest,
};

constexpr auto DEFAULT_MODE = QueryType::none;
constexpr unsigned int DEFAULT_NUM_NEIGHBORS(0xffffffff);
constexpr float DEFAULT_R_MAX(-1.0);
constexpr float DEFAULT_R_MIN(0);
constexpr float DEFAULT_R_GUESS(-1.0);
constexpr float DEFAULT_SCALE(-1.0);
constexpr bool DEFAULT_EXCLUDE_II(false);

// NeighborQuery.h:60-73, field for field
struct QueryArgs
{
    QueryArgs() = default;
    QueryType mode {DEFAULT_MODE};
    unsigned int num_neighbors {DEFAULT_NUM_NEIGHBORS};
    float r_max {DEFAULT_R_MAX};
    float r_min {DEFAULT_R_MIN};
    float r_guess {DEFAULT_R_GUESS};
    float scale {DEFAULT_SCALE};
    bool exclude_ii {DEFAULT_EXCLUDE_II};
};

// NeighborBond.h:25-60: (query index, point index, distance, weight, vector)
struct NeighborBond
{
    unsigned int query_point_idx {0xffffffffU}, point_idx {0xffffffffU};
    float distance {0}, weight {0};
    vec3<float> vector;
    bool operator==(const NeighborBond& o) const
    {
        return query_point_idx == o.query_point_idx && point_idx == o.point_idx && distance == o.distance
            && weight == o.weight;
    }
    bool operator!=(const NeighborBond& o) const { return !(*this == o); }
    // accessors the binding layer exposes (freud/locality/NeighborBond.h:62-78, export-NeighborList.cc:90-96)
    unsigned int getQueryPointIdx() const { return query_point_idx; }
    unsigned int getPointIdx() const { return point_idx; }
    float getDistance() const { return distance; }
    float getWeight() const { return weight; }
    const vec3<float>& getVector() const { return vector; }
};
inline NeighborBond iterator_terminator() // ITERATOR_TERMINATOR, NeighborQuery.h:51-52
{
    return NeighborBond();
}

// freud/locality/NeighborPerPointIterator.h:38-58: the neighbours of one query point, one bond per next() until end()
class NeighborPerPointIterator
{
public:
    NeighborPerPointIterator() = default;
    explicit NeighborPerPointIterator(unsigned int query_point_idx) : m_query_point_idx(query_point_idx) {}
    virtual ~NeighborPerPointIterator() = default;
    virtual bool end() const = 0;
    virtual NeighborBond next() = 0;

protected:
    unsigned int m_query_point_idx {0};
};

class NeighborQueryIterator;
class NeighborQueryPerPointIterator;
class NeighborQuery;
inline const float* selfOrHost(const NeighborQuery& nq, const vec3<float>* query_points, unsigned int n_query_points);

class NeighborQuery
{
public:
    NeighborQuery(box::Box box, const vec3<float>* points, unsigned int n_points, int flavour)
        : m_box(box), m_points(points), m_n_points(n_points), m_flavour(flavour)
    {
        // NeighborQuery.h:97-112 (the C ABI repeats both checks; they are made here first so that no device
        // is needed to reject bad input)
        if (m_n_points == 0)
        {
            throw std::invalid_argument("Cannot create a NeighborQuery with 0 particles.");
        }
        if (m_box.is2D())
        {
            for (unsigned int i = 0; i < n_points; ++i)
            {
                if (std::abs(m_points[i].z) > 1e-6)
                {
                    throw std::invalid_argument("A point with z != 0 was provided in a 2D box.");
                }
            }
        }
    }
    virtual ~NeighborQuery() = default;

    std::shared_ptr<NeighborQueryIterator> query(const vec3<float>* query_points, unsigned int n_query_points,
                                                 QueryArgs query_args) const;

    // The neighbours of ONE query point (NeighborQuery.h:144-154; what loopOverNeighborsIterator hands a compute that
    // normalises per row, NeighborComputeFunctional.h:112-150): a one-row query on the device, its bonds served in
    // (j) or, for nearest-neighbour mode, distance order.  query_point_idx is the row index the bonds carry and the
    // point index exclude_ii compares against.
    std::shared_ptr<NeighborQueryPerPointIterator> querySingle(const vec3<float> query_point, unsigned int query_point_idx,
                                                               QueryArgs args) const;

    const box::Box& getBox() const { return m_box; }
    const vec3<float>* getPoints() const { return m_points; }
    unsigned int getNPoints() const { return m_n_points; }
    int getFlavour() const { return m_flavour; }

    vec3<float> operator[](unsigned int index) const
    {
        if (index >= m_n_points)
        {
            throw std::runtime_error("NeighborQuery attempted to access a point with index >= n_points.");
        }
        return m_points[index];
    }

    // Device-resident copy of the points (+ cached cell list), created on first use.
    fgpu_points* device() const
    {
        if (!m_dev)
        {
            float b6[6];
            m_box.toArray6(b6);
            fgpu_points* h = nullptr;
            gpu::check(fgpu_points_create(gpu::context(), b6, m_box.is2D() ? 1 : 0, reinterpret_cast<const float*>(m_points),
                                          m_n_points, &h));
            m_dev = std::shared_ptr<fgpu_points>(h, fgpu_points_destroy);
        }
        return m_dev.get();
    }

    // NeighborQuery.h:195-292 (public so that the compute classes can resolve default arguments the same way)
    virtual void validateQueryArgs(QueryArgs& args) const
    {
        inferMode(args);
        if (args.mode == QueryType::ball)
        {
            if (args.r_max == DEFAULT_R_MAX)
            {
                throw std::runtime_error("You must set r_max in the query arguments when performing ball queries.");
            }
            if (args.num_neighbors != DEFAULT_NUM_NEIGHBORS)
            {
                throw std::runtime_error(
                    "You cannot set num_neighbors in the query arguments when performing ball queries.");
            }
        }
        else if (args.mode == QueryType::nearest)
        {
            if (args.num_neighbors == DEFAULT_NUM_NEIGHBORS)
            {
                throw std::runtime_error("You must set num_neighbors in the query arguments when performing "
                                         "number of neighbor queries.");
            }
            if (args.r_max == DEFAULT_R_MAX)
            {
                args.r_max = std::numeric_limits<float>::infinity();
            }
            // validateNearestNeighborArgs, NeighborQuery.h:235-270: scale / r_guess steer the reference's search
            // only, never its result (tests/test_locality_neighbor_query.py:635-656); scale is still validated
            if (args.scale != DEFAULT_SCALE && args.scale <= 1.0F)
            {
                throw std::runtime_error("The scale query argument must be greater than 1.");
            }
        }
        else
        {
            throw std::runtime_error("Unknown mode");
        }
        // NeighborQueryPerPointIterator ctor, NeighborQuery.h:321-328
        if (args.r_max <= 0)
        {
            throw std::invalid_argument("NeighborQuery requires r_max to be positive.");
        }
        if (args.r_max <= args.r_min)
        {
            throw std::invalid_argument("NeighborQuery requires that r_max must be greater than r_min.");
        }
    }

protected:
    virtual void inferMode(QueryArgs& args) const
    {
        if (args.mode == QueryType::none)
        {
            if (args.num_neighbors != DEFAULT_NUM_NEIGHBORS)
            {
                args.mode = QueryType::nearest;
            }
            else if (args.r_max != DEFAULT_R_MAX)
            {
                args.mode = QueryType::ball;
            }
        }
    }

    const box::Box m_box;
    const vec3<float>* m_points;
    unsigned int m_n_points;
    int m_flavour;
    mutable std::shared_ptr<fgpu_points> m_dev;
};

// Lazy result of query(): nothing runs until toNeighborList() or the first next()
// (NeighborQuery.h:364-493).
class NeighborQueryIterator
{
public:
    NeighborQueryIterator(const NeighborQuery* nq, const vec3<float>* query_points, unsigned int n_query_points,
                          QueryArgs qargs)
        : m_nq(nq), m_query_points(query_points), m_n_query_points(n_query_points), m_qargs(qargs)
    {}

    // One bond per call in (i, j) order, then ITERATOR_TERMINATOR forever (NeighborQuery.h:392-419 yields them in
    // engine traversal order; upstream only ever compares them as sets).
    NeighborBond next()
    {
        if (!m_iter_list)
        {
            m_iter_list = toNeighborList(false);
            m_cursor = 0;
        }
        if (m_cursor >= m_iter_list->getNumBonds())
        {
            return iterator_terminator();
        }
        size_t const b = m_cursor++;
        NeighborBond nb;
        nb.query_point_idx = (*m_iter_list->getNeighbors())[2 * b];
        nb.point_idx = (*m_iter_list->getNeighbors())[2 * b + 1];
        nb.distance = (*m_iter_list->getDistances())[b];
        nb.weight = (*m_iter_list->getWeights())[b];
        nb.vector = vec3<float>((*m_iter_list->getVectors())[3 * b], (*m_iter_list->getVectors())[3 * b + 1],
                                (*m_iter_list->getVectors())[3 * b + 2]);
        return nb;
    }

    // NeighborQuery.h:434-481: the whole query on the device, result sorted by (i, j) or (i, d, j)
    std::shared_ptr<NeighborList> toNeighborList(bool sort_by_distance = false)
    {
        fgpu_points* pts = m_nq->device();
        fgpu_nlist* out = nullptr;
        // queries == the reference points themselves (by value): skip the upload and the second cell sort
        const float* q = selfOrHost(*m_nq, m_query_points, m_n_query_points);
        if (m_qargs.mode == QueryType::ball)
        {
            gpu::check(fgpu_ball_query(pts, q, m_n_query_points, 0, m_nq->getFlavour(), m_qargs.r_max, m_qargs.r_min,
                                       m_qargs.exclude_ii ? 1 : 0, sort_by_distance ? 1 : 0, &out));
        }
        else
        {
            gpu::check(fgpu_knn_query(pts, q, m_n_query_points, 0, m_nq->getFlavour(), m_qargs.num_neighbors,
                                      m_qargs.r_max, m_qargs.r_min, m_qargs.exclude_ii ? 1 : 0,
                                      sort_by_distance ? 1 : 0, &out));
        }
        return std::make_shared<NeighborList>(out, gpu::context());
    }

    const QueryArgs& getQueryArgs() const { return m_qargs; }

    // NeighborQuery.h:380-390, 392-419
    bool end() const { return m_iter_list && m_cursor >= m_iter_list->getNumBonds(); }
    std::shared_ptr<NeighborQueryPerPointIterator> query(unsigned int i)
    {
        if (i >= m_n_query_points)
        {
            throw std::out_of_range("query point index out of range");
        }
        return m_nq->querySingle(m_query_points[i], i, m_qargs);
    }

private:
    const NeighborQuery* m_nq;
    const vec3<float>* m_query_points;
    unsigned int m_n_query_points;
    QueryArgs m_qargs;
    std::shared_ptr<NeighborList> m_iter_list;
    size_t m_cursor {0};
};

// The query points as the C ABI takes them: nullptr when they are the reference points themselves (no upload, no
// second cell sort), else the packed floats.
// NeighborQuery.h:309-350.  One device query of a single row; next() walks its bonds.
class NeighborQueryPerPointIterator : public NeighborPerPointIterator
{
public:
    NeighborQueryPerPointIterator(const NeighborQuery* neighbor_query, const vec3<float>& query_point,
                                  unsigned int query_point_idx, const QueryArgs& qargs)
        : NeighborPerPointIterator(query_point_idx), m_neighbor_query(neighbor_query), m_query_point(query_point),
          m_qargs(qargs)
    {
        if (qargs.r_max <= 0)
        {
            throw std::invalid_argument("NeighborQuery requires r_max to be positive.");
        }
        if (qargs.r_max <= qargs.r_min)
        {
            throw std::invalid_argument("NeighborQuery requires that r_max must be greater than r_min.");
        }
    }

    bool end() const override { return m_list && m_cursor >= m_list->getNumBonds(); }

    NeighborBond next() override
    {
        if (!m_list)
        {
            fgpu_nlist* out = nullptr;
            const float* q = reinterpret_cast<const float*>(&m_query_point);
            if (m_qargs.mode == QueryType::nearest)
            {
                gpu::check(fgpu_knn_query(m_neighbor_query->device(), q, 1, m_query_point_idx, m_neighbor_query->getFlavour(),
                                          m_qargs.num_neighbors, m_qargs.r_max, m_qargs.r_min, m_qargs.exclude_ii ? 1 : 0,
                                          1, &out));
            }
            else
            {
                gpu::check(fgpu_ball_query(m_neighbor_query->device(), q, 1, m_query_point_idx,
                                           m_neighbor_query->getFlavour(), m_qargs.r_max, m_qargs.r_min,
                                           m_qargs.exclude_ii ? 1 : 0, 0, &out));
            }
            m_list = std::make_shared<NeighborList>(out, gpu::context());
        }
        if (m_cursor >= m_list->getNumBonds())
        {
            return iterator_terminator();
        }
        size_t const b = m_cursor++;
        NeighborBond nb;
        nb.query_point_idx = m_query_point_idx;
        nb.point_idx = (*m_list->getNeighbors())[2 * b + 1];
        nb.distance = (*m_list->getDistances())[b];
        nb.weight = (*m_list->getWeights())[b];
        nb.vector = vec3<float>((*m_list->getVectors())[3 * b], (*m_list->getVectors())[3 * b + 1],
                                (*m_list->getVectors())[3 * b + 2]);
        return nb;
    }

protected:
    const NeighborQuery* m_neighbor_query;
    vec3<float> m_query_point;
    QueryArgs m_qargs;
    std::shared_ptr<NeighborList> m_list;
    size_t m_cursor {0};
};

inline std::shared_ptr<NeighborQueryPerPointIterator>
NeighborQuery::querySingle(const vec3<float> query_point, unsigned int query_point_idx, QueryArgs args) const
{
    this->validateQueryArgs(args);
    return std::make_shared<NeighborQueryPerPointIterator>(this, query_point, query_point_idx, args);
}

// Bitwise equality of two point arrays: a few probes first (different query sets fail at once), then the whole
// array, on a few threads when it is large.
inline bool samePoints(const vec3<float>* a, const vec3<float>* b, unsigned int n)
{
    if (a == b)
    {
        return true;
    }
    if (a == nullptr || b == nullptr)
    {
        return false;
    }
    size_t const bytes = (size_t) n * sizeof(vec3<float>);
    for (unsigned int k = 0; k < 64 && n != 0; ++k)
    {
        size_t const i = (size_t) k * (n - 1) / 63;
        if (std::memcmp(&a[i], &b[i], sizeof(vec3<float>)) != 0)
        {
            return false;
        }
    }
    auto const* pa = reinterpret_cast<const unsigned char*>(a);
    auto const* pb = reinterpret_cast<const unsigned char*>(b);
    unsigned int const n_threads = bytes >= (2U << 20) ? 4 : 1;
    if (n_threads == 1)
    {
        return std::memcmp(pa, pb, bytes) == 0;
    }
    std::atomic<bool> same {true};
    std::vector<std::thread> pool;
    size_t const chunk = (bytes + n_threads - 1) / n_threads;
    for (unsigned int t = 0; t < n_threads; ++t)
    {
        size_t const lo = std::min(bytes, (size_t) t * chunk), hi = std::min(bytes, lo + chunk);
        pool.emplace_back([=, &same] {
            if (std::memcmp(pa + lo, pb + lo, hi - lo) != 0)
            {
                same = false;
            }
        });
    }
    for (auto& th : pool)
    {
        th.join();
    }
    return same;
}

// The query points as the C ABI takes them: nullptr when they are the reference points themselves (no upload, no
// second cell sort), else the packed floats.  "Themselves" is decided by value, not by address: freud's Python layer
// keeps a private copy of the points (freud/locality.py:867-868), so nq.query(points, ...) with the array the engine
// was built from arrives here as a different pointer to the same numbers.
inline const float* selfOrHost(const NeighborQuery& nq, const vec3<float>* query_points, unsigned int n_query_points)
{
    bool const self = n_query_points == nq.getNPoints() && samePoints(query_points, nq.getPoints(), n_query_points);
    return self ? nullptr : reinterpret_cast<const float*>(query_points);
}

inline std::shared_ptr<NeighborQueryIterator> NeighborQuery::query(const vec3<float>* query_points,
                                                                   unsigned int n_query_points,
                                                                   QueryArgs query_args) const
{
    vec3<bool> const periodic = m_box.getPeriodic();
    if (!(periodic.x && periodic.y && periodic.z))
    {
        throw std::domain_error("Pair queries in a non-periodic box are not implemented.");
    }
    this->validateQueryArgs(query_args);
    return std::make_shared<NeighborQueryIterator>(this, query_points, n_query_points, query_args);
}

// LinkCell (freud/locality/LinkCell.h:188, LinkCell.cc:222-260).  cell_width is validated like upstream and
// otherwise ignored: by SURVEY.md E1 it never influences the result, and the GPU grid follows r_max.
// Nearest-neighbour queries are LinkCellQueryIterator::next (LinkCell.cc:575-679): the k smallest wrapped
// distances, in the same WRAP arithmetic as the ball query.
class LinkCell : public NeighborQuery
{
public:
    LinkCell(const box::Box& box, const vec3<float>* points, unsigned int n_points, float cell_width = 0)
        : NeighborQuery(box, points, n_points, FGPU_FLAVOUR_WRAP), m_cell_width(cell_width)
    {
        if (cell_width == 0)
        {
            unsigned int const desired = std::max(n_points / 10U, 1U);
            m_cell_width = std::cbrt(box.getVolume() / static_cast<float>(desired));
        }
        vec3<float> const pd = box.getNearestPlaneDistance();
        if ((m_cell_width * 2.0 > pd.x) || (m_cell_width * 2.0 > pd.y) || (!box.is2D() && m_cell_width * 2.0 > pd.z))
        {
            throw std::runtime_error("Cannot generate a cell list where cell_width is larger than half the box.");
        }
    }
    float getCellWidth() const { return m_cell_width; }

private:
    float m_cell_width;
};

// AABBQuery (freud/locality/AABBQuery.h:42).  The BVH is not rebuilt: a conservative cell list with the
// image arithmetic yields the identical bond list.
class AABBQuery : public NeighborQuery
{
public:
    AABBQuery(const box::Box& box, const vec3<float>* points, unsigned int n_points)
        : NeighborQuery(box, points, n_points, FGPU_FLAVOUR_IMAGE)
    {}
};

// CellQuery (freud/locality/CellQuery.h:29): ghost particles on a Cartesian grid upstream; by equivalence E5 its
// ball query is the bond set r = (p_j + shift_w) - q over the lattice displacements, reproduced bit for bit by the
// GHOST flavour.  Ball queries only (CellQuery.h:180-184).
class CellQuery : public NeighborQuery
{
public:
    CellQuery(const box::Box& box, const vec3<float>* points, unsigned int n_points)
        : NeighborQuery(box, points, n_points, FGPU_FLAVOUR_GHOST)
    {}

    // CellQuery::validateQueryArgs, CellQuery.h:177-203: raised by query(), before anything runs
    void validateQueryArgs(QueryArgs& args) const override
    {
        NeighborQuery::validateQueryArgs(args);
        if (args.mode == QueryType::nearest)
        {
            throw std::runtime_error("CellQuery only supports ball queries (r_max), not nearest queries "
                                     "(num_neighbors). Use AABBQuery for nearest neighbor queries.");
        }
        vec3<float> const pd = m_box.getNearestPlaneDistance();
        if ((pd.x <= args.r_max * 2.0F) || (pd.y <= args.r_max * 2.0F) || (!m_box.is2D() && pd.z <= args.r_max * 2.0F))
        {
            throw std::runtime_error("The CellQuery r_max is too large for this box.");
        }
    }

    // ---- grid introspection (CellQuery.h:84-171, CellQuery.cc:55-137) ------------------------------------------
    // Upstream answers ball queries from a Cartesian grid of width r_cut over the box's bounding cuboid, with a
    // ghost copy of every point that lies within r_cut of a periodic face; its binding layer exposes that grid
    // (export-NeighborQuery.cc:96-111).  Queries here run on the GPU's own cell list (the grid only generates
    // candidates), so these members describe the grid upstream WOULD build for r_cut --
    // dimensions, origin, per-cell populations with and without ghosts, cell offsets -- computed on the host when
    // asked, for callers that inspect it.
    void setupGrid(const float r_cut) const
    {
        m_cell_inverse_length = 1.0F / r_cut;
        float const lx = m_box.getLx(), ly = m_box.getLy(), lz = m_box.getLz();
        float const xy = m_box.getTiltFactorXY(), xz = m_box.getTiltFactorXZ(), yz = m_box.getTiltFactorYZ();
        // extent of the box along the Cartesian axes, and its lowest corner
        volatile float wx = lx + ly * std::abs(xy) + lz * std::abs(xz);
        volatile float wy = ly + lz * std::abs(yz);
        m_nx = (unsigned int) (int) ((float) wx * m_cell_inverse_length) + 3U;
        m_ny = (unsigned int) (int) ((float) wy * m_cell_inverse_length) + 3U;
        m_nz = (unsigned int) (int) (lz * m_cell_inverse_length) + 3U;
        volatile float min_x = ly * std::min(0.0F, xy) + lz * std::min(0.0F, xz);
        volatile float min_y = lz * std::min(0.0F, yz);
        vec3<float> const origin = m_box.makeAbsolute(vec3<float>(0.0F, 0.0F, 0.0F));
        volatile float px = (float) min_x - r_cut, py = (float) min_y - r_cut, pz = -r_cut;
        m_min_pos = vec3<float>((float) px + origin.x, (float) py + origin.y, (float) pz + origin.z);
    }

    void buildGrid(const float r_cut) const
    {
        if (r_cut <= 0)
        {
            throw std::runtime_error("CellQuery::buildGrid called with invalid r_cut (must be positive).");
        }
        setupGrid(r_cut);
        size_t const n_cells = (size_t) m_nx * m_ny * m_nz;
        m_counts.assign(n_cells, 0U);
        m_counts_real.assign(n_cells, 0U);
        m_cell_starts.assign(n_cells, 0U);
        m_n_total = 0;
        vec3<float> const pd = m_box.getNearestPlaneDistance();
        vec3<float> const fr(r_cut / pd.x, r_cut / pd.y, r_cut / pd.z);
        vec3<float> const a = m_box.getLatticeVector(0), b = m_box.getLatticeVector(1);
        vec3<float> const c = m_box.is2D() ? vec3<float>(0.0F, 0.0F, 0.0F) : m_box.getLatticeVector(2);
        auto cell_of = [&](const vec3<float>& p, size_t& idx) {
            int const cx = (int) std::floor((p.x - m_min_pos.x) * m_cell_inverse_length);
            int const cy = (int) std::floor((p.y - m_min_pos.y) * m_cell_inverse_length);
            int const cz = (int) std::floor((p.z - m_min_pos.z) * m_cell_inverse_length);
            if (cx < 0 || cy < 0 || cz < 0 || cx >= (int) m_nx || cy >= (int) m_ny || cz >= (int) m_nz)
            {
                return false;
            }
            idx = ((size_t) cz * m_ny + cy) * m_nx + cx;
            return true;
        };
        for (unsigned int i = 0; i < m_n_points; ++i)
        {
            vec3<float> const p = m_points[i];
            vec3<float> const f = m_box.makeFractional(p);
            // +1: near the low face (its image appears beyond the high face), -1: near the high face
            int const s[3] = {(int) (f.x <= fr.x) - (int) (f.x >= 1.0 - fr.x), (int) (f.y <= fr.y) - (int) (f.y >= 1.0 - fr.y),
                              m_box.is2D() ? 0 : (int) (f.z <= fr.z) - (int) (f.z >= 1.0 - fr.z)};
            // one ghost per non-empty subset of the faces the point is close to
            for (int mask = 1; mask < 8; ++mask)
            {
                bool const ux = (mask & 1) != 0, uy = (mask & 2) != 0, uz = (mask & 4) != 0;
                if ((ux && s[0] == 0) || (uy && s[1] == 0) || (uz && s[2] == 0))
                {
                    continue;
                }
                vec3<float> g = p;
                vec3<float> shift(0.0F, 0.0F, 0.0F);
                auto add = [&](const vec3<float>& v, int sign) {
                    shift = vec3<float>(shift.x + (sign > 0 ? v.x : -v.x), shift.y + (sign > 0 ? v.y : -v.y),
                                        shift.z + (sign > 0 ? v.z : -v.z));
                };
                if (ux)
                {
                    add(a, s[0]);
                }
                if (uy)
                {
                    add(b, s[1]);
                }
                if (uz)
                {
                    add(c, s[2]);
                }
                g = vec3<float>(p.x + shift.x, p.y + shift.y, p.z + shift.z);
                size_t idx = 0;
                if (cell_of(g, idx))
                {
                    m_counts[idx] += 1;
                    m_n_total += 1;
                }
            }
            size_t idx = 0;
            if (cell_of(p, idx))
            {
                m_counts[idx] += 1;
                m_counts_real[idx] += 1;
                m_n_total += 1;
            }
        }
        unsigned int acc = 0;
        for (size_t k = 0; k < n_cells; ++k)
        {
            m_cell_starts[k] = acc;
            acc += m_counts[k];
        }
    }

    float getCellWidth() const { return 1.0F / m_cell_inverse_length; }
    float getCellInverseWidth() const { return m_cell_inverse_length; }
    const std::vector<unsigned int>& getCountsReal() const { return m_counts_real; }
    const std::vector<unsigned int>& getCounts() const { return m_counts; }
    const std::vector<unsigned int>& getCellStarts() const { return m_cell_starts; }
    std::vector<float> getMinPos() const { return {m_min_pos.x, m_min_pos.y, m_min_pos.z}; }
    unsigned int getNx() const { return m_nx; }
    unsigned int getNy() const { return m_ny; }
    unsigned int getNz() const { return m_nz; }
    unsigned int getNTotal() const { return m_n_total; }

private:
    mutable float m_cell_inverse_length {0};
    mutable vec3<float> m_min_pos;
    mutable unsigned int m_nx {0}, m_ny {0}, m_nz {0}, m_n_total {0};
    mutable std::vector<unsigned int> m_counts, m_counts_real, m_cell_starts;
};

// RawPoints (freud/locality/RawPoints.h:35-73): upstream builds an AABBQuery lazily on the first query.
class RawPoints : public NeighborQuery
{
public:
    RawPoints(const box::Box& box, const vec3<float>* points, unsigned int n_points)
        : NeighborQuery(box, points, n_points, FGPU_FLAVOUR_IMAGE)
    {}
};

}} // namespace freud::locality
