// freud::locality::NeighborList backed by a device-resident list.
//
// Same container contract as upstream (freud/locality/NeighborList.h:30-157, NeighborList.cc:191-232):
// neighbors u32 (n, 2), distances f32, weights f32, vectors f32 (n, 3), lazily derived counts/segments with
// segments of empty rows equal to 0.  A list produced by a query stays on the GPU (fgpu_nlist) so that
// RDF::accumulate / Steinhardt::compute consume it without a PCIe round trip; the host arrays are
// materialised (one D2H copy per array group) the first time a getter asks for them.  A list built from
// host arrays (from_arrays upstream, NeighborList.cc:62-101) is uploaded the first time a compute needs it.
#pragma once
#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>
#include <vector>

#include "Box.h"
#include "Context.h"
#include "ManagedArray.h"
#include "VectorMath.h"

namespace freud { namespace locality {

class NeighborList
{
public:
    NeighborList() : NeighborList(0) {}

    explicit NeighborList(unsigned int num_bonds)
        : m_num_bonds(num_bonds), m_neighbors(make<unsigned int>({num_bonds, 2})),
          m_distances(make<float>({num_bonds})), m_weights(make<float>({num_bonds})),
          m_vectors(make<float>({num_bonds, 3})), m_host_valid(true)
    {}

    // from_arrays (NeighborList.cc:62-101): validates ordering and index ranges like upstream
    NeighborList(unsigned int num_bonds, const unsigned int* query_point_index, unsigned int num_query_points,
                 const unsigned int* point_index, unsigned int num_points, const vec3<float>* vectors,
                 const float* weights)
        : NeighborList(num_bonds)
    {
        m_num_query_points = num_query_points;
        m_num_points = num_points;
        unsigned int last = 0;
        for (unsigned int b = 0; b < num_bonds; ++b)
        {
            unsigned int const i = query_point_index[b], j = point_index[b];
            if (i < last)
            {
                throw std::invalid_argument("NeighborList query_point_index must be sorted.");
            }
            if (i >= num_query_points)
            {
                throw std::invalid_argument("NeighborList query_point_index values must be less than num_query_points.");
            }
            if (j >= num_points)
            {
                throw std::invalid_argument("NeighborList point_index values must be less than num_points.");
            }
            last = i;
            (*m_neighbors)[2 * (size_t) b] = i;
            (*m_neighbors)[2 * (size_t) b + 1] = j;
            (*m_weights)[b] = weights != nullptr ? weights[b] : 1.0F;
            vec3<float> const v = vectors[b];
            (*m_vectors)[3 * (size_t) b] = v.x;
            (*m_vectors)[3 * (size_t) b + 1] = v.y;
            (*m_vectors)[3 * (size_t) b + 2] = v.z;
            volatile float xx = v.x * v.x, yy = v.y * v.y, zz = v.z * v.z;
            volatile float s = xx + yy;
            volatile float d2 = s + zz;
            (*m_distances)[b] = std::sqrt((float) d2);
        }
    }

    // all pairs (NeighborList.cc:84-128, NeighborList.all_pairs upstream): every (query point, point) bond, the vector
    // being box.wrap(query_points[i] - points[j]) as upstream writes it; O(N Nq) on the host, for small systems
    NeighborList(const vec3<float>* points, const vec3<float>* query_points, const box::Box& box, const bool exclude_ii,
                 const unsigned int num_points, const unsigned int num_query_points)
        : NeighborList(num_points * num_query_points - (exclude_ii ? std::min(num_points, num_query_points) : 0U))
    {
        m_num_points = num_points;
        m_num_query_points = num_query_points;
        size_t b = 0;
        for (unsigned int i = 0; i < num_query_points; ++i)
        {
            for (unsigned int j = 0; j < num_points; ++j)
            {
                if (exclude_ii && i == j)
                {
                    continue;
                }
                volatile float dx = query_points[i].x - points[j].x, dy = query_points[i].y - points[j].y,
                               dz = query_points[i].z - points[j].z;
                vec3<float> const dr = box.wrap(vec3<float>(dx, dy, dz));
                (*m_neighbors)[2 * b] = i;
                (*m_neighbors)[2 * b + 1] = j;
                (*m_weights)[b] = 1.0F;
                volatile float xx = dr.x * dr.x, yy = dr.y * dr.y, zz = dr.z * dr.z;
                volatile float s = xx + yy;
                volatile float d2 = s + zz;
                (*m_distances)[b] = std::sqrt((float) d2);
                (*m_vectors)[3 * b] = dr.x;
                (*m_vectors)[3 * b + 1] = dr.y;
                (*m_vectors)[3 * b + 2] = dr.z;
                ++b;
            }
        }
    }

    // adopt a device list produced by a query (takes ownership of the handle)
    NeighborList(fgpu_nlist* dev, fgpu_ctx* ctx)
        : m_num_bonds((unsigned int) fgpu_nlist_num_bonds(dev)), m_num_query_points(fgpu_nlist_num_query_points(dev)),
          m_num_points(fgpu_nlist_num_points(dev)), m_dev(dev, fgpu_nlist_destroy), m_ctx(ctx)
    {}

    unsigned int getNumBonds() const { return m_num_bonds; }
    unsigned int getNumQueryPoints() const { return m_num_query_points; }
    unsigned int getNumPoints() const { return m_num_points; }

    // each getter brings over its own array only (a caller that reads the distances does not pay for the vectors)
    std::shared_ptr<const util::ManagedArray<unsigned int>> getNeighbors() const
    {
        materialise(kNeighbors);
        return m_neighbors;
    }
    std::shared_ptr<const util::ManagedArray<float>> getDistances() const
    {
        materialise(kDistances);
        return m_distances;
    }
    std::shared_ptr<const util::ManagedArray<float>> getWeights() const
    {
        materialise(kWeights);
        return m_weights;
    }
    std::shared_ptr<const util::ManagedArray<float>> getVectors() const
    {
        materialise(kVectors);
        return m_vectors;
    }
    std::shared_ptr<const util::ManagedArray<unsigned int>> getCounts() const
    {
        updateSegmentCounts();
        return m_counts;
    }
    std::shared_ptr<const util::ManagedArray<unsigned int>> getSegments() const
    {
        updateSegmentCounts();
        return m_segments;
    }

    // NeighborList.cc:199-232: rows without bonds keep segment 0
    void updateSegmentCounts() const
    {
        if (m_segments_valid)
        {
            return;
        }
        // a list that came from a query still has its device twin (reorder() drops it): the kernels already wrote
        // counts and segments, empty rows 0 as upstream
        bool const from_device = (bool) m_dev;
        m_counts = from_device ? make_raw<unsigned int>({m_num_query_points}) : make<unsigned int>({m_num_query_points});
        m_segments = from_device ? make_raw<unsigned int>({m_num_query_points}) : make<unsigned int>({m_num_query_points});
        if (from_device)
        {
            gpu::check(fgpu_nlist_copy(m_dev.get(), nullptr, nullptr, nullptr, nullptr, m_segments->data(),
                                       m_counts->data()));
        }
        else
        {
            for (unsigned int b = 0; b < m_num_bonds; ++b)
            {
                unsigned int const i = (*m_neighbors)[2 * (size_t) b];
                if ((*m_counts)[i] == 0)
                {
                    (*m_segments)[i] = b;
                }
                (*m_counts)[i] += 1;
            }
        }
        m_segments_valid = true;
    }

    // NeighborList.cc:247-262: index of the first bond whose query index is >= i
    unsigned int find_first_index(unsigned int i) const
    {
        materialise();
        unsigned int lo = 0, hi = m_num_bonds;
        while (lo < hi)
        {
            unsigned int const mid = lo + (hi - lo) / 2;
            if ((*m_neighbors)[2 * (size_t) mid] < i)
            {
                lo = mid + 1;
            }
            else
            {
                hi = mid;
            }
        }
        return lo;
    }

    // keep the bonds whose flag is set (NeighborList.h:113, NeighborList.cc:264-318); returns the number removed
    unsigned int filter(const bool* keep)
    {
        materialise();
        std::vector<unsigned int> idx;
        for (unsigned int b = 0; b < m_num_bonds; ++b)
        {
            if (keep[b])
            {
                idx.push_back(b);
            }
        }
        unsigned int const removed = m_num_bonds - (unsigned int) idx.size();
        reorder(idx);
        return removed;
    }

    unsigned int filter_r(float r_max, float r_min = 0)
    {
        materialise();
        std::vector<char> keep(m_num_bonds);
        for (unsigned int b = 0; b < m_num_bonds; ++b)
        {
            float const d = (*m_distances)[b];
            keep[b] = d >= r_min && d < r_max;
        }
        return filter(reinterpret_cast<const bool*>(keep.data()));
    }

    // NeighborList.cc:371-399: (i, j, weight, distance) or (i, distance, j, weight)
    void sort(bool by_distance)
    {
        materialise();
        std::vector<unsigned int> idx(m_num_bonds);
        std::iota(idx.begin(), idx.end(), 0U);
        const unsigned int* nb = m_neighbors->data();
        const float* d = m_distances->data();
        const float* w = m_weights->data();
        std::stable_sort(idx.begin(), idx.end(), [&](unsigned int a, unsigned int b) {
            if (nb[2 * (size_t) a] != nb[2 * (size_t) b])
            {
                return nb[2 * (size_t) a] < nb[2 * (size_t) b];
            }
            if (by_distance)
            {
                if (d[a] != d[b])
                {
                    return d[a] < d[b];
                }
                if (nb[2 * (size_t) a + 1] != nb[2 * (size_t) b + 1])
                {
                    return nb[2 * (size_t) a + 1] < nb[2 * (size_t) b + 1];
                }
                return w[a] < w[b];
            }
            if (nb[2 * (size_t) a + 1] != nb[2 * (size_t) b + 1])
            {
                return nb[2 * (size_t) a + 1] < nb[2 * (size_t) b + 1];
            }
            if (w[a] != w[b])
            {
                return w[a] < w[b];
            }
            return d[a] < d[b];
        });
        reorder(idx);
    }

    void copy(const NeighborList& other)
    {
        other.materialise();
        m_num_bonds = other.m_num_bonds;
        m_num_query_points = other.m_num_query_points;
        m_num_points = other.m_num_points;
        m_neighbors = std::make_shared<util::ManagedArray<unsigned int>>(*other.m_neighbors);
        m_distances = std::make_shared<util::ManagedArray<float>>(*other.m_distances);
        m_weights = std::make_shared<util::ManagedArray<float>>(*other.m_weights);
        m_vectors = std::make_shared<util::ManagedArray<float>>(*other.m_vectors);
        m_host_valid = true;
        m_segments_valid = false;
        m_dev.reset();
    }

    // NeighborList.cc:330-341
    void validate(unsigned int num_query_points, unsigned int num_points) const
    {
        if (num_query_points != m_num_query_points)
        {
            throw std::runtime_error("NeighborList found inconsistent array sizes.");
        }
        if (num_points != m_num_points)
        {
            throw std::runtime_error("NeighborList found inconsistent array sizes.");
        }
    }

    // Device view for the compute classes: uploads a host-built list on first use.
    const fgpu_nlist* device(fgpu_ctx* ctx) const
    {
        if (!m_dev || m_ctx != ctx)
        {
            materialise();
            fgpu_nlist* h = nullptr;
            gpu::check(fgpu_nlist_from_host(ctx, m_num_bonds, m_num_query_points, m_num_points, m_neighbors->data(),
                                            m_distances->data(), m_weights->data(), m_vectors->data(), &h));
            m_dev = std::shared_ptr<fgpu_nlist>(h, fgpu_nlist_destroy);
            m_ctx = ctx;
        }
        return m_dev.get();
    }

private:
    template<typename T> static std::shared_ptr<util::ManagedArray<T>> make(std::vector<size_t> shape)
    {
        return std::make_shared<util::ManagedArray<T>>(std::move(shape));
    }

    enum : unsigned
    {
        kNeighbors = 1,
        kDistances = 2,
        kWeights = 4,
        kVectors = 8,
        kAll = 15
    };
    template<typename T> static std::shared_ptr<util::ManagedArray<T>> make_raw(std::vector<size_t> shape)
    {
        return std::make_shared<util::ManagedArray<T>>(std::move(shape), util::Uninitialized {});
    }

    // Device -> host copy of the bond arrays: the first getter starts all four transfers (page-locked destinations the
    // copies overwrite entirely: no zero fill) and every getter waits for its own array only.
    void materialise(unsigned want = kAll) const
    {
        if (m_host_valid)
        {
            return;
        }
        if (!m_copy_started)
        {
            m_neighbors = make_raw<unsigned int>({m_num_bonds, 2});
            m_distances = make_raw<float>({m_num_bonds});
            m_weights = make_raw<float>({m_num_bonds});
            m_vectors = make_raw<float>({m_num_bonds, 3});
            gpu::check(fgpu_nlist_copy_begin(m_dev.get(), m_neighbors->data(), m_distances->data(), m_weights->data(),
                                             m_vectors->data()));
            m_copy_started = true;
        }
        want &= ~m_have;
        if (want == 0)
        {
            return;
        }
        gpu::check(fgpu_nlist_copy_wait(m_dev.get(), want));
        m_have |= want;
        m_host_valid = m_have == kAll;
    }

    void reorder(const std::vector<unsigned int>& idx)
    {
        auto nb = make<unsigned int>({idx.size(), 2});
        auto d = make<float>({idx.size()});
        auto w = make<float>({idx.size()});
        auto v = make<float>({idx.size(), 3});
        for (size_t k = 0; k < idx.size(); ++k)
        {
            size_t const b = idx[k];
            (*nb)[2 * k] = (*m_neighbors)[2 * b];
            (*nb)[2 * k + 1] = (*m_neighbors)[2 * b + 1];
            (*d)[k] = (*m_distances)[b];
            (*w)[k] = (*m_weights)[b];
            for (int c = 0; c < 3; ++c)
            {
                (*v)[3 * k + c] = (*m_vectors)[3 * b + c];
            }
        }
        m_neighbors = nb;
        m_distances = d;
        m_weights = w;
        m_vectors = v;
        m_num_bonds = (unsigned int) idx.size();
        m_segments_valid = false;
        m_dev.reset(); // the device copy is stale
    }

    unsigned int m_num_bonds {0}, m_num_query_points {0}, m_num_points {0};
    mutable std::shared_ptr<util::ManagedArray<unsigned int>> m_neighbors;
    mutable std::shared_ptr<util::ManagedArray<float>> m_distances, m_weights, m_vectors;
    mutable std::shared_ptr<util::ManagedArray<unsigned int>> m_counts, m_segments;
    mutable bool m_host_valid {false}, m_segments_valid {false};
    mutable unsigned m_have {0}; // bond arrays already on the host (device-built lists)
    mutable bool m_copy_started {false};
    mutable std::shared_ptr<fgpu_nlist> m_dev;
    mutable fgpu_ctx* m_ctx {nullptr};
};

}} // namespace freud::locality
