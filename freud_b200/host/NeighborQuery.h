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
};
inline NeighborBond iterator_terminator() // ITERATOR_TERMINATOR, NeighborQuery.h:51-52
{
    return NeighborBond();
}

class NeighborQueryIterator;

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
        // queries == the reference points themselves: skip the upload and the second cell sort
        const float* q = reinterpret_cast<const float*>(m_query_points);
        if (m_query_points == m_nq->getPoints() && m_n_query_points == m_nq->getNPoints())
        {
            q = nullptr;
        }
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
