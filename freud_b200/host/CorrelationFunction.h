// freud::density::CorrelationFunction on the GPU path.
//
// Signatures: CorrelationFunction(bins, r_max) (freud/density/CorrelationFunction.h:52, .cc:26-46),
// accumulate(neighbor_query, values, query_points, query_values, n_query_points, nlist /*nullable*/, qargs)
// (.cc:81-95), reset (.cc:61-66), getCorrelation and the BondHistogramCompute getters the bindings expose
// (export-CorrelationFunction.cc, freud/locality/BondHistogramCompute.h:29-140).  The bonds are the list handed in
// or the query over the points, materialised as a NeighborList on the device; counts and complex<double> sums live
// on the GPU across accumulate calls and reduce() (.cc:49-59) divides on the host.  Like upstream's thread-local
// histograms the double sums have no fixed order: results agree to double rounding.
#pragma once
#include <complex>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "Context.h"
#include "ManagedArray.h"
#include "NeighborList.h"
#include "NeighborQuery.h"

namespace freud { namespace density {

class CorrelationFunction
{
public:
    CorrelationFunction(unsigned int bins, float r_max) : m_bins(bins), m_r_max(r_max)
    {
        if (bins == 0)
        {
            throw std::invalid_argument("CorrelationFunction  requires a nonzero number of bins.");
        }
        if (r_max <= 0)
        {
            throw std::invalid_argument("CorrelationFunction requires r_max to be positive.");
        }
        // RegularAxis(bins, 0, r_max), freud/util/Histogram.h:126-138
        volatile float width = r_max / static_cast<float>(bins);
        m_edges.resize((size_t) bins + 1);
        for (size_t i = 0; i <= bins; ++i)
        {
            volatile float t = static_cast<float>(i) * width;
            m_edges[i] = 0.0F + t;
        }
        allocate();
    }

    // CorrelationFunction.cc:61-66 / BondHistogramCompute.h:39-49: new arrays, earlier views stay valid
    void reset()
    {
        if (m_dev)
        {
            gpu::check(fgpu_corr_reset(m_dev.get()));
        }
        allocate();
        m_frame_counter = 0;
        m_reduce = true;
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const std::complex<double>* values,
                    const vec3<float>* query_points, const std::complex<double>* query_values,
                    unsigned int n_query_points, const std::shared_ptr<locality::NeighborList>& nlist,
                    locality::QueryArgs qargs)
    {
        m_box = neighbor_query->getBox();
        if (!m_dev)
        {
            fgpu_corr* h = nullptr;
            gpu::check(fgpu_corr_create(gpu::context(), (uint32_t) m_bins, m_r_max, &h));
            m_dev = std::shared_ptr<fgpu_corr>(h, fgpu_corr_destroy);
        }
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            auto query = neighbor_query->query(query_points, n_query_points, qargs); // validates, infers the mode
            locality::QueryArgs const& args = query->getQueryArgs();
            if (args.mode == locality::QueryType::ball)
            {
                // a ball query made for this compute alone: the bonds are binned where the search left them
                gpu::check(fgpu_corr_accumulate(m_dev.get(), neighbor_query->device(),
                                                locality::selfOrHost(*neighbor_query, query_points, n_query_points),
                                                n_query_points, neighbor_query->getFlavour(), args.r_max, args.r_min,
                                                args.exclude_ii ? 1 : 0, reinterpret_cast<const double*>(values),
                                                reinterpret_cast<const double*>(query_values)));
            }
            else
            {
                list = query->toNeighborList();
            }
        }
        else
        {
            list->validate(n_query_points, neighbor_query->getNPoints());
        }
        if (list)
        {
            gpu::check(fgpu_corr_accumulate_nlist(m_dev.get(), list->device(gpu::context()),
                                                  reinterpret_cast<const double*>(values),
                                                  reinterpret_cast<const double*>(query_values)));
        }
        m_frame_counter++;
        m_reduce = true;
    }

    // CorrelationFunction.cc:49-59
    void reduce()
    {
        if (m_dev)
        {
            std::vector<double> sums(2 * m_bins);
            gpu::check(fgpu_corr_read(m_dev.get(), m_bin_counts->data(), sums.data()));
            for (size_t i = 0; i < m_bins; ++i)
            {
                std::complex<double> v(sums[2 * i], sums[2 * i + 1]);
                if ((*m_bin_counts)[i] != 0)
                {
                    v /= (*m_bin_counts)[i];
                }
                (*m_correlation)[i] = v;
            }
        }
        m_reduce = false;
    }

    std::shared_ptr<util::ManagedArray<std::complex<double>>> getCorrelation()
    {
        if (m_reduce)
        {
            reduce();
        }
        return m_correlation;
    }
    std::shared_ptr<util::ManagedArray<unsigned int>> getBinCounts()
    {
        if (m_reduce)
        {
            reduce();
        }
        return m_bin_counts;
    }
    const box::Box& getBox() const { return m_box; }
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
    std::vector<std::pair<float, float>> getBounds() const { return {{0.0F, m_r_max}}; }
    std::vector<size_t> getAxisSizes() const { return {m_bins}; }

private:
    void allocate()
    {
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_bins});
        m_correlation = std::make_shared<util::ManagedArray<std::complex<double>>>(std::vector<size_t> {m_bins});
    }

    size_t m_bins;
    float m_r_max;
    box::Box m_box;
    unsigned int m_frame_counter {0};
    bool m_reduce {true};
    std::vector<float> m_edges;
    std::shared_ptr<util::ManagedArray<unsigned int>> m_bin_counts;
    std::shared_ptr<util::ManagedArray<std::complex<double>>> m_correlation;
    std::shared_ptr<fgpu_corr> m_dev;
};

}} // namespace freud::density
