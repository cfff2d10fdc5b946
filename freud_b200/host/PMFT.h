// freud::pmft::PMFTXY on the GPU path.
//
// Signatures: PMFTXY(x_max, y_max, n_x, n_y) (freud/pmft/PMFTXY.h, PMFTXY.cc:25-57), accumulate(neighbor_query,
// query_orientations, query_points, n_query_points, nlist /*nullable*/, qargs) (PMFTXY.cc:65-87), reset / getPCF
// (freud/pmft/PMFT.h:36-49) and the BondHistogramCompute getters the bindings expose
// (freud/locality/BondHistogramCompute.h:29-140).  The bonds are the list handed in or the query over the points,
// materialised as a NeighborList on the device.  cos/sin of the orientations are evaluated on the host (inside
// fgpu_pmftxy_accumulate_nlist) with the same libm the reference calls (rotmat2::fromAngle, VectorMath.h:912-921)
// and the rotation + binning run on the GPU in the reference's float operation order: bit-identical bin counts; PMFT::reduce
// (PMFT.h:73-83) is host float arithmetic in the same order.
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

namespace freud { namespace pmft {

class PMFTXY
{
public:
    PMFTXY(float x_max, float y_max, unsigned int n_x, unsigned int n_y) : m_x_max(x_max), m_y_max(y_max), m_nx(n_x), m_ny(n_y)
    {
        if (n_x < 1)
        {
            throw std::invalid_argument("PMFTXY requires at least 1 bin in X.");
        }
        if (n_y < 1)
        {
            throw std::invalid_argument("PMFTXY requires at least 1 bin in Y.");
        }
        if (x_max < 0)
        {
            throw std::invalid_argument("PMFTXY requires that x_max must be positive.");
        }
        if (y_max < 0)
        {
            throw std::invalid_argument("PMFTXY requires that y_max must be positive.");
        }
        volatile float dx = 2.0F * x_max / static_cast<float>(n_x); // PMFTXY.cc:49-51
        volatile float dy = 2.0F * y_max / static_cast<float>(n_y);
        volatile float jac = dx * dy;
        m_jacobian = jac;
        m_edges_x = edges(n_x, -x_max, x_max);
        m_edges_y = edges(n_y, -y_max, y_max);
        allocate();
    }

    void reset() // BondHistogramCompute.h:39-49, PMFT.h:42-49: new arrays, earlier views stay valid
    {
        if (m_dev)
        {
            gpu::check(fgpu_pmftxy_reset(m_dev.get()));
        }
        allocate();
        m_frame_counter = 0;
        m_reduce = true;
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const float* query_orientations,
                    const vec3<float>* query_points, unsigned int n_query_points,
                    const std::shared_ptr<locality::NeighborList>& nlist, locality::QueryArgs qargs)
    {
        if (!neighbor_query->getBox().is2D()) // Box::enforce2D, Box.h:571-577
        {
            throw std::invalid_argument("A 3D box was provided to a class that only supports 2D systems.");
        }
        m_box = neighbor_query->getBox();
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            list = neighbor_query->query(query_points, n_query_points, qargs)->toNeighborList();
        }
        else
        {
            list->validate(n_query_points, neighbor_query->getNPoints());
        }
        if (!m_dev)
        {
            fgpu_pmftxy* h = nullptr;
            gpu::check(fgpu_pmftxy_create(gpu::context(), m_x_max, m_y_max, m_nx, m_ny, &h));
            m_dev = std::shared_ptr<fgpu_pmftxy>(h, fgpu_pmftxy_destroy);
        }
        gpu::check(fgpu_pmftxy_accumulate_nlist(m_dev.get(), list->device(gpu::context()), query_orientations));
        m_frame_counter++;
        m_n_points = neighbor_query->getNPoints();
        m_n_query_points = n_query_points;
        m_reduce = true;
    }

    // PMFT::reduce with PMFTXY's constant Jacobian factor (PMFT.h:73-83, PMFTXY.cc:59-63)
    void reduce()
    {
        if (m_dev)
        {
            gpu::check(fgpu_pmftxy_read(m_dev.get(), m_bin_counts->data()));
            volatile float inv_num_dens = m_box.getVolume() / static_cast<float>(m_n_query_points);
            volatile float den = static_cast<float>(m_frame_counter) * static_cast<float>(m_n_points);
            volatile float norm_factor = 1.0F / den;
            volatile float prefactor = inv_num_dens * norm_factor;
            volatile float jacobian_factor = 1.0F / m_jacobian;
            for (size_t i = 0; i < (size_t) m_nx * m_ny; ++i)
            {
                volatile float t = static_cast<float>((*m_bin_counts)[i]) * prefactor;
                (*m_pcf)[i] = t * jacobian_factor;
            }
        }
        m_reduce = false;
    }

    std::shared_ptr<util::ManagedArray<float>> getPCF()
    {
        if (m_reduce)
        {
            reduce();
        }
        return m_pcf;
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
    std::vector<std::vector<float>> getBinEdges() const { return {m_edges_x, m_edges_y}; }
    std::vector<std::vector<float>> getBinCenters() const { return {centers(m_edges_x), centers(m_edges_y)}; }
    std::vector<std::pair<float, float>> getBounds() const { return {{-m_x_max, m_x_max}, {-m_y_max, m_y_max}}; }
    std::vector<size_t> getAxisSizes() const { return {m_nx, m_ny}; }

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
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_nx, m_ny});
        m_pcf = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_nx, m_ny});
    }

    float m_x_max, m_y_max;
    size_t m_nx, m_ny;
    float m_jacobian {1.0F};
    box::Box m_box;
    unsigned int m_frame_counter {0}, m_n_points {0}, m_n_query_points {0};
    bool m_reduce {true};
    std::vector<float> m_edges_x, m_edges_y;
    std::shared_ptr<util::ManagedArray<unsigned int>> m_bin_counts;
    std::shared_ptr<util::ManagedArray<float>> m_pcf;
    std::shared_ptr<fgpu_pmftxy> m_dev;
};

}} // namespace freud::pmft
