// freud::environment::BondOrder on the GPU path.
//
// Signatures: BondOrder(n_bins_theta, n_bins_phi, mode) (freud/environment/BondOrder.h:32-66, BondOrder.cc:30-78),
// accumulate(neighbor_query, orientations, query_points, query_orientations, n_query_points, nlist /*nullable*/,
// qargs) (BondOrder.cc:100-153), reset / getBondOrder / getMode and the BondHistogramCompute getters the bindings
// expose (freud/locality/BondHistogramCompute.h:29-140).  The bonds are the list handed in or the query over the
// points, materialised as a NeighborList on the device; rotation and binning run on the GPU (fgpu_bondorder_*,
// csrc/pmft.cu), bit-identical to the reference's counts.  The solid angle of the bins and BondOrder::reduce
// (BondOrder.cc:62-69, 86-93) are host float arithmetic in the reference's order, with the host libm's cosf.
#pragma once
#include <cmath>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "Context.h"
#include "ManagedArray.h"
#include "NeighborList.h"
#include "NeighborQuery.h"

namespace freud { namespace environment {

typedef enum
{
    bod = 0,
    lbod = 1,
    obcd = 2,
    oocd = 3
} BondOrderMode; // BondOrder.h:23-29

class BondOrder
{
public:
    BondOrder(unsigned int n_bins_theta, unsigned int n_bins_phi, BondOrderMode mode)
        : m_nt(n_bins_theta), m_np(n_bins_phi), m_mode(mode)
    {
        if (n_bins_theta < 2)
        {
            throw std::invalid_argument("BondOrder requires at least 2 bins in theta.");
        }
        if (n_bins_phi < 2)
        {
            throw std::invalid_argument("BondOrder requires at least 2 bins in phi.");
        }
        float const two_pi = static_cast<float>(2.0 * M_PI); // constants::TWO_PI, Box.h:24
        volatile float dt = two_pi / static_cast<float>(n_bins_theta);
        volatile float dp = static_cast<float>(M_PI / static_cast<float>(n_bins_phi)); // double division, BondOrder.cc:49
        m_sa.resize((size_t) n_bins_theta * n_bins_phi);
        for (size_t i = 0; i < n_bins_theta; ++i)
        {
            for (size_t j = 0; j < n_bins_phi; ++j)
            {
                volatile float phi = static_cast<float>(j) * dp;
                volatile float phi2 = phi + dp;
                volatile float diff = std::cos(static_cast<float>(phi)) - std::cos(static_cast<float>(phi2));
                m_sa[i * n_bins_phi + j] = dt * diff;
            }
        }
        m_edges_t = edges(n_bins_theta, 0.0F, two_pi);
        m_edges_p = edges(n_bins_phi, 0.0F, static_cast<float>(M_PI));
        allocate();
    }

    void reset() // BondOrder.cc:80-84
    {
        if (m_dev)
        {
            gpu::check(fgpu_bondorder_reset(m_dev.get()));
        }
        allocate();
        m_frame_counter = 0;
        m_reduce = true;
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const quat<float>* orientations,
                    const vec3<float>* query_points, const quat<float>* query_orientations, unsigned int n_query_points,
                    const std::shared_ptr<locality::NeighborList>& nlist, locality::QueryArgs qargs)
    {
        m_box = neighbor_query->getBox();
        if (!m_dev)
        {
            fgpu_bondorder* h = nullptr;
            gpu::check(fgpu_bondorder_create(gpu::context(), (uint32_t) m_nt, (uint32_t) m_np, (int) m_mode, &h));
            m_dev = std::shared_ptr<fgpu_bondorder>(h, fgpu_bondorder_destroy);
        }
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            auto query = neighbor_query->query(query_points, n_query_points, qargs); // validates, infers the mode
            locality::QueryArgs const& args = query->getQueryArgs();
            if (args.mode == locality::QueryType::ball)
            {
                // a ball query made for this diagram alone: the bonds are binned where the search left them
                gpu::check(fgpu_bondorder_accumulate(m_dev.get(), neighbor_query->device(),
                                                     locality::selfOrHost(*neighbor_query, query_points, n_query_points),
                                                     n_query_points, neighbor_query->getFlavour(), args.r_max, args.r_min,
                                                     args.exclude_ii ? 1 : 0, reinterpret_cast<const float*>(orientations),
                                                     reinterpret_cast<const float*>(query_orientations)));
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
            gpu::check(fgpu_bondorder_accumulate_nlist(m_dev.get(), list->device(gpu::context()),
                                                       reinterpret_cast<const float*>(orientations),
                                                       neighbor_query->getNPoints(),
                                                       reinterpret_cast<const float*>(query_orientations)));
        }
        m_frame_counter++;
        m_reduce = true;
    }

    void reduce() // BondOrder.cc:86-93
    {
        if (m_dev)
        {
            gpu::check(fgpu_bondorder_read(m_dev.get(), m_bin_counts->data()));
            for (size_t i = 0; i < m_nt * m_np; ++i)
            {
                volatile float t = static_cast<float>((*m_bin_counts)[i]) / m_sa[i];
                (*m_bo)[i] = t / static_cast<float>(m_frame_counter);
            }
        }
        m_reduce = false;
    }

    std::shared_ptr<util::ManagedArray<float>> getBondOrder()
    {
        if (m_reduce)
        {
            reduce();
        }
        return m_bo;
    }
    std::shared_ptr<util::ManagedArray<unsigned int>> getBinCounts()
    {
        if (m_reduce)
        {
            reduce();
        }
        return m_bin_counts;
    }
    BondOrderMode getMode() const { return m_mode; }
    const box::Box& getBox() const { return m_box; }
    std::vector<std::vector<float>> getBinEdges() const { return {m_edges_t, m_edges_p}; }
    std::vector<std::vector<float>> getBinCenters() const { return {centers(m_edges_t), centers(m_edges_p)}; }
    std::vector<std::pair<float, float>> getBounds() const
    {
        return {{0.0F, static_cast<float>(2.0 * M_PI)}, {0.0F, static_cast<float>(M_PI)}};
    }
    std::vector<size_t> getAxisSizes() const { return {m_nt, m_np}; }
    //! bonds since the last reset whose bin the host's libm decided (csrc/pmft.cu)
    unsigned long long getHostBinnedBonds() const
    {
        uint64_t n = 0;
        if (m_dev)
        {
            gpu::check(fgpu_bondorder_deferred(m_dev.get(), &n));
        }
        return n;
    }

private:
    static std::vector<float> edges(unsigned int bins, float lo, float hi) // RegularAxis, Histogram.h:126-138
    {
        volatile float span = hi - lo;
        volatile float width = span / static_cast<float>(bins);
        std::vector<float> e((size_t) bins + 1);
        for (size_t i = 0; i <= bins; ++i)
        {
            volatile float t = static_cast<float>(i) * width;
            e[i] = lo + t;
        }
        return e;
    }
    static std::vector<float> centers(const std::vector<float>& e) // Axis::getBinCenters, Histogram.h:87-95
    {
        std::vector<float> c(e.size() - 1);
        for (size_t i = 0; i + 1 < e.size(); ++i)
        {
            volatile float s = e[i] + e[i + 1];
            c[i] = s / 2.0F;
        }
        return c;
    }
    void allocate()
    {
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_nt, m_np});
        m_bo = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_nt, m_np});
    }

    size_t m_nt, m_np;
    BondOrderMode m_mode;
    std::vector<float> m_sa; // solid angle of the bins
    box::Box m_box;
    unsigned int m_frame_counter {0};
    bool m_reduce {true};
    std::vector<float> m_edges_t, m_edges_p;
    std::shared_ptr<util::ManagedArray<unsigned int>> m_bin_counts;
    std::shared_ptr<util::ManagedArray<float>> m_bo;
    std::shared_ptr<fgpu_bondorder> m_dev;
};

}} // namespace freud::environment
