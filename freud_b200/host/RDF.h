// freud::density::RDF on the device-resident histogram of libfreud_b200.so.
//
// Signatures: RDF(bins, r_max, r_min = 0) (freud/density/RDF.h:33), accumulate(neighbor_query, query_points,
// n_query_points, nlist /*nullable*/, qargs) (RDF.h:43-46), getRDF / getNr (RDF.h:55-68), and the
// BondHistogramCompute getters the bindings expose (freud/locality/BondHistogramCompute.h:29-140,
// export-BondHistogramCompute.cc:69-76).  Accumulation state follows upstream: the bin counts keep adding
// across accumulate() calls until reset(); the frame counter divides at reduce time; reduce is lazy and runs
// on the first getter after an accumulate (BondHistogramCompute.h:61-69).  The u32[bins] counters live on the
// GPU between frames; only reduce() copies them (bins * 4 bytes) to the host.
#pragma once
#include <cmath>
#include <memory>
#include <utility>
#include <vector>

#include "Context.h"
#include "ManagedArray.h"
#include "NeighborList.h"
#include "NeighborQuery.h"

namespace freud { namespace locality {

class BondHistogramCompute
{
public:
    BondHistogramCompute() = default;
    virtual ~BondHistogramCompute() = default;

    // BondHistogramCompute.h:39-49: new arrays, so views handed out earlier are not invalidated
    virtual void reset()
    {
        if (m_dev)
        {
            gpu::check(fgpu_rdf_reset(m_dev.get()));
        }
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_bins});
        m_frame_counter = 0;
        m_reduce = true;
    }

    virtual void reduce() = 0;

    const box::Box& getBox() const { return m_box; }

    template<typename U> std::shared_ptr<U> reduceAndReturn(std::shared_ptr<U> thing_to_return)
    {
        if (m_reduce)
        {
            reduce();
        }
        m_reduce = false;
        return thing_to_return;
    }

    std::shared_ptr<const util::ManagedArray<unsigned int>> getBinCounts()
    {
        if (m_reduce)
        {
            reduce();
        }
        m_reduce = false;
        return m_bin_counts;
    }

    // RegularAxis, freud/util/Histogram.h:126-138 and Axis::getBinCenters :87-95
    std::vector<std::vector<float>> getBinEdges() const { return {m_edges}; }
    std::vector<std::vector<float>> getBinCenters() const
    {
        std::vector<float> c(m_bins);
        for (size_t i = 0; i < m_bins; ++i)
        {
            volatile float s = m_edges[i] + m_edges[i + 1];
            c[i] = s / 2.0F;
        }
        return {c};
    }
    std::vector<std::pair<float, float>> getBounds() const { return {{m_r_min, m_r_max}}; }
    std::vector<size_t> getAxisSizes() const { return {m_bins}; }

protected:
    void setAxis(unsigned int bins, float r_min, float r_max)
    {
        m_bins = bins;
        m_r_min = r_min;
        m_r_max = r_max;
        volatile float span = r_max - r_min;
        volatile float width = span / static_cast<float>(bins);
        m_edges.resize((size_t) bins + 1);
        for (size_t i = 0; i <= bins; ++i)
        {
            volatile float t = static_cast<float>(i) * width;
            m_edges[i] = r_min + t;
        }
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_bins});
    }

    // created with the first accumulate so that constructing the object needs no device
    fgpu_rdf* device()
    {
        if (!m_dev)
        {
            fgpu_rdf* h = nullptr;
            gpu::check(fgpu_rdf_create(gpu::context(), (uint32_t) m_bins, m_r_max, m_r_min, &h));
            m_dev = std::shared_ptr<fgpu_rdf>(h, fgpu_rdf_destroy);
        }
        return m_dev.get();
    }

    // accumulateGeneral (BondHistogramCompute.h:115-126) with loopOverNeighbors
    // (NeighborComputeFunctional.h:173-218) specialised to "bin the bond distance"
    void accumulateDistances(const std::shared_ptr<NeighborQuery>& neighbor_query, const vec3<float>* query_points,
                             unsigned int n_query_points, const std::shared_ptr<NeighborList>& nlist, QueryArgs qargs)
    {
        m_box = neighbor_query->getBox();
        fgpu_rdf* rdf = device();
        if (nlist)
        {
            nlist->validate(n_query_points, neighbor_query->getNPoints());
            gpu::check(fgpu_rdf_accumulate_nlist(rdf, nlist->device(gpu::context())));
        }
        else
        {
            neighbor_query->validateQueryArgs(qargs);
            if (qargs.mode == QueryType::ball)
            {
                const float* q = reinterpret_cast<const float*>(query_points);
                if (query_points == neighbor_query->getPoints() && n_query_points == neighbor_query->getNPoints())
                {
                    q = nullptr; // self query: no second upload
                }
                gpu::check(fgpu_rdf_accumulate(rdf, neighbor_query->device(), q, n_query_points, 0,
                                               neighbor_query->getFlavour(), qargs.r_max, qargs.r_min,
                                               qargs.exclude_ii ? 1 : 0));
            }
            else
            {
                auto list = neighbor_query->query(query_points, n_query_points, qargs)->toNeighborList();
                gpu::check(fgpu_rdf_accumulate_nlist(rdf, list->device(gpu::context())));
            }
        }
        m_frame_counter++;
        m_n_points = neighbor_query->getNPoints();
        m_n_query_points = n_query_points;
        m_reduce = true;
    }

    void readCounts()
    {
        if (m_dev)
        {
            gpu::check(fgpu_rdf_read(m_dev.get(), m_bin_counts->data()));
        }
    }

    box::Box m_box;
    unsigned int m_frame_counter {0};
    unsigned int m_n_points {0};
    unsigned int m_n_query_points {0};
    bool m_reduce {true};
    size_t m_bins {0};
    float m_r_min {0}, m_r_max {0};
    std::vector<float> m_edges;
    std::shared_ptr<util::ManagedArray<unsigned int>> m_bin_counts;
    std::shared_ptr<fgpu_rdf> m_dev;
};

}} // namespace freud::locality

namespace freud { namespace density {

enum class NormalizationMode
{
    exact,
    finite_size
};

class RDF : public locality::BondHistogramCompute
{
public:
    NormalizationMode mode {NormalizationMode::exact};

    RDF(unsigned int bins, float r_max, float r_min = 0)
    {
        // RDF.cc:27-42
        if (bins == 0)
        {
            throw std::invalid_argument("RDF requires a nonzero number of bins.");
        }
        if (r_max <= 0)
        {
            throw std::invalid_argument("RDF requires r_max to be positive.");
        }
        if (r_min < 0)
        {
            throw std::invalid_argument("RDF requires r_min to be non-negative.");
        }
        if (r_max <= r_min)
        {
            throw std::invalid_argument("RDF requires that r_max must be greater than r_min.");
        }
        setAxis(bins, r_min, r_max);
        m_pcf = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_bins});
        m_N_r = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_bins});
        // shell volumes from the float32 bin edges (RDF.cc:50-63); upstream allocates bins^2 / bins^3 floats for
        // them and uses the first `bins` -- only those are kept here
        m_vol2D.resize(m_bins);
        m_vol3D.resize(m_bins);
        float const volume_prefactor = (float) ((double) (4.0F / 3.0F) * M_PI);
        for (size_t i = 0; i < m_bins; ++i)
        {
            float const r = m_edges[i], nextr = m_edges[i + 1];
            volatile float a2 = nextr * nextr, c2 = r * r;
            volatile float d2 = a2 - c2;
            m_vol2D[i] = (float) (M_PI * (double) d2);
            volatile float a3 = a2 * nextr, c3 = c2 * r;
            volatile float d3 = a3 - c3;
            volatile float v3 = volume_prefactor * d3;
            m_vol3D[i] = v3;
        }
    }
    ~RDF() override = default;

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const vec3<float>* query_points,
                    unsigned int n_query_points, const std::shared_ptr<locality::NeighborList>& nlist,
                    const locality::QueryArgs& qargs)
    {
        accumulateDistances(neighbor_query, query_points, n_query_points, nlist, qargs);
    }

    // RDF.cc:73-99, same float32 operation order
    void reduce() override
    {
        readCounts();
        float const nqp = static_cast<float>(m_n_query_points);
        volatile float number_density = nqp / m_box.getVolume();
        if (mode == NormalizationMode::finite_size)
        {
            volatile float ratio = static_cast<float>(m_n_query_points - 1) / static_cast<float>(m_n_query_points);
            number_density = number_density * ratio;
        }
        float const np = static_cast<float>(m_n_points);
        float const nf = static_cast<float>(m_frame_counter);
        volatile float den = np * number_density;
        den = den * nf;
        volatile float prefactor = 1.0F / den;
        const std::vector<float>& vol = m_box.is2D() ? m_vol2D : m_vol3D;
        for (size_t i = 0; i < m_bins; ++i)
        {
            volatile float t = static_cast<float>((*m_bin_counts)[i]) * prefactor;
            (*m_pcf)[i] = t / vol[i];
        }
        volatile float nn = nqp * nf;
        volatile float pre2 = 1.0F / nn;
        volatile float first = static_cast<float>((*m_bin_counts)[0]) * pre2;
        (*m_N_r)[0] = first;
        for (size_t i = 1; i < m_bins; ++i)
        {
            volatile float t = static_cast<float>((*m_bin_counts)[i]) * pre2;
            (*m_N_r)[i] = (*m_N_r)[i - 1] + t;
        }
    }

    // RDF.cc:66-71
    void reset() override
    {
        BondHistogramCompute::reset();
        m_pcf = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_bins});
        m_N_r = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_bins});
    }

    std::shared_ptr<util::ManagedArray<float>> getRDF() { return reduceAndReturn(m_pcf); }
    std::shared_ptr<util::ManagedArray<float>> getNr() { return reduceAndReturn(m_N_r); }

private:
    std::shared_ptr<util::ManagedArray<float>> m_pcf, m_N_r;
    std::vector<float> m_vol2D, m_vol3D;
};

}} // namespace freud::density
