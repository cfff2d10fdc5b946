// freud::pmft::PMFTXY, PMFTXYZ, PMFTXYT and PMFTR12 on the GPU path (the three-axis classes: second half of the file).
//
// Signatures: PMFTXY(x_max, y_max, n_x, n_y) (freud/pmft/PMFTXY.h, PMFTXY.cc:25-57), accumulate(neighbor_query,
// query_orientations, query_points, n_query_points, nlist /*nullable*/, qargs) (PMFTXY.cc:65-87), reset / getPCF
// (freud/pmft/PMFT.h:36-49) and the BondHistogramCompute getters the bindings expose
// (freud/locality/BondHistogramCompute.h:29-140).  The bonds are the list handed in or the query over the points,
// materialised as a NeighborList on the device.  Rotation and binning run on the GPU in the reference's float
// operation order; wherever libm (cosf / sinf / atan2f) could decide a bin by its last place the bond is handed to
// the host's libm instead (csrc/pmft.cu): bit-identical bin counts.  PMFT::reduce (PMFT.h:73-83) is host float
// arithmetic in the same order.
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
        if (!m_dev)
        {
            fgpu_pmftxy* h = nullptr;
            gpu::check(fgpu_pmftxy_create(gpu::context(), m_x_max, m_y_max, m_nx, m_ny, &h));
            m_dev = std::shared_ptr<fgpu_pmftxy>(h, fgpu_pmftxy_destroy);
        }
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            auto query = neighbor_query->query(query_points, n_query_points, qargs); // validates, infers the mode
            locality::QueryArgs const& args = query->getQueryArgs();
            if (args.mode == locality::QueryType::ball)
            {
                // a ball query made for this histogram alone: the bonds go from the search straight into the bins
                gpu::check(fgpu_pmftxy_accumulate(m_dev.get(), neighbor_query->device(),
                                                  locality::selfOrHost(*neighbor_query, query_points, n_query_points),
                                                  n_query_points, neighbor_query->getFlavour(), args.r_max, args.r_min,
                                                  args.exclude_ii ? 1 : 0, query_orientations));
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
            gpu::check(fgpu_pmftxy_accumulate_nlist(m_dev.get(), list->device(gpu::context()), query_orientations));
        }
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

// ---- three-axis PMFTs ------------------------------------------------------------------------------------------
// Shared state of PMFTXYZ / PMFTXYT / PMFTR12: the device histogram (fgpu_pmft), the RegularAxis triple and
// PMFT::reduce (PMFT.h:73-83) with a per-bin Jacobian factor, all host float arithmetic in the reference's order.
class PMFT3
{
public:
    void reset() // BondHistogramCompute.h:39-49, PMFT.h:42-49: new arrays, earlier views stay valid
    {
        if (m_dev)
        {
            gpu::check(fgpu_pmft_reset(m_dev.get()));
        }
        allocate();
        m_frame_counter = 0;
        m_reduce = true;
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
    std::vector<std::vector<float>> getBinEdges() const { return {m_edges[0], m_edges[1], m_edges[2]}; }
    std::vector<std::vector<float>> getBinCenters() const
    {
        return {centers(m_edges[0]), centers(m_edges[1]), centers(m_edges[2])};
    }
    std::vector<std::pair<float, float>> getBounds() const
    {
        return {{m_lo[0], m_hi[0]}, {m_lo[1], m_hi[1]}, {m_lo[2], m_hi[2]}};
    }
    std::vector<size_t> getAxisSizes() const { return {m_n[0], m_n[1], m_n[2]}; }
    //! bonds of the frames since the last reset whose angle bin the host's libm decided (pmft.cu)
    unsigned long long getHostBinnedBonds() const
    {
        uint64_t n = 0;
        if (m_dev)
        {
            gpu::check(fgpu_pmft_deferred(m_dev.get(), &n));
        }
        return n;
    }

protected:
    PMFT3(int kind, float max0, float max1, float max2, unsigned int n0, unsigned int n1, unsigned int n2)
        : m_kind(kind), m_max {max0, max1, max2}, m_n {n0, n1, n2}
    {}

    void setAxis(int ax, float lo, float hi)
    {
        m_lo[ax] = lo;
        m_hi[ax] = hi;
        m_edges[ax] = edges(m_n[ax], lo, hi);
    }

    // the bonds: the list handed in, or the query over the points -- a ball query goes from the search straight into
    // the bins (fgpu_pmft_accumulate), anything else through a NeighborList on the device
    void accumulateDevice(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const vec3<float>* query_points,
                          unsigned int n_query_points, const std::shared_ptr<locality::NeighborList>& nlist,
                          const locality::QueryArgs& qargs, const float* orientations, const float* query_orientations,
                          const float* equiv, unsigned int n_equiv)
    {
        m_box = neighbor_query->getBox();
        if (!m_dev)
        {
            fgpu_pmft* h = nullptr;
            gpu::check(fgpu_pmft_create(gpu::context(), m_kind, m_max[0], m_max[1], m_max[2], (uint32_t) m_n[0],
                                        (uint32_t) m_n[1], (uint32_t) m_n[2], &h));
            m_dev = std::shared_ptr<fgpu_pmft>(h, fgpu_pmft_destroy);
        }
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            auto query = neighbor_query->query(query_points, n_query_points, qargs); // validates, infers the mode
            locality::QueryArgs const& args = query->getQueryArgs();
            if (args.mode == locality::QueryType::ball)
            {
                gpu::check(fgpu_pmft_accumulate(m_dev.get(), neighbor_query->device(),
                                                locality::selfOrHost(*neighbor_query, query_points, n_query_points),
                                                n_query_points, neighbor_query->getFlavour(), args.r_max, args.r_min,
                                                args.exclude_ii ? 1 : 0, orientations, query_orientations, equiv,
                                                n_equiv));
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
            gpu::check(fgpu_pmft_accumulate_nlist(m_dev.get(), list->device(gpu::context()), orientations,
                                                  neighbor_query->getNPoints(), query_orientations, equiv, n_equiv));
        }
        m_frame_counter++;
        m_n_points = neighbor_query->getNPoints();
        m_n_query_points = n_query_points;
        m_reduce = true;
    }

    virtual float normExtra() const { return 1.0F; }   // PMFTXYZ: the number of equivalent orientations
    virtual float jacobianFactor(size_t bin) const = 0; // 1 / volume element of the bin

    void reduce()
    {
        if (m_dev)
        {
            gpu::check(fgpu_pmft_read(m_dev.get(), m_bin_counts->data()));
            volatile float inv_num_dens = m_box.getVolume() / static_cast<float>(m_n_query_points);
            volatile float den = static_cast<float>(m_frame_counter) * static_cast<float>(m_n_points);
            if (m_kind == FGPU_PMFT_XYZ)
            {
                den = den * normExtra(); // PMFTXYZ.cc:89-91
            }
            volatile float norm_factor = 1.0F / den;
            volatile float prefactor = inv_num_dens * norm_factor;
            size_t const n_bins = m_n[0] * m_n[1] * m_n[2];
            for (size_t i = 0; i < n_bins; ++i)
            {
                volatile float t = static_cast<float>((*m_bin_counts)[i]) * prefactor;
                (*m_pcf)[i] = t * jacobianFactor(i);
            }
        }
        m_reduce = false;
    }

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
        m_bin_counts = std::make_shared<util::ManagedArray<unsigned int>>(std::vector<size_t> {m_n[0], m_n[1], m_n[2]});
        m_pcf = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {m_n[0], m_n[1], m_n[2]});
    }

    int m_kind;
    float m_max[3];
    size_t m_n[3];
    float m_lo[3] {0, 0, 0}, m_hi[3] {0, 0, 0};
    std::vector<float> m_edges[3];
    box::Box m_box;
    unsigned int m_frame_counter {0}, m_n_points {0}, m_n_query_points {0};
    bool m_reduce {true};
    std::shared_ptr<util::ManagedArray<unsigned int>> m_bin_counts;
    std::shared_ptr<util::ManagedArray<float>> m_pcf;
    std::shared_ptr<fgpu_pmft> m_dev;
};

static constexpr float kPmftTwoPi = static_cast<float>(2.0 * M_PI); // constants::TWO_PI, Box.h:24

// PMFTXYZ(x_max, y_max, z_max, n_x, n_y, n_z); accumulate(neighbor_query, query_orientations, query_points,
// n_query_points, equiv_orientations, num_equiv_orientations, nlist, qargs) -- freud/pmft/PMFTXYZ.h, PMFTXYZ.cc:24-147
class PMFTXYZ : public PMFT3
{
public:
    PMFTXYZ(float x_max, float y_max, float z_max, unsigned int n_x, unsigned int n_y, unsigned int n_z)
        : PMFT3(FGPU_PMFT_XYZ, x_max, y_max, z_max, n_x, n_y, n_z)
    {
        if (n_x < 1)
        {
            throw std::invalid_argument("PMFTXYZ requires at least 1 bin in X.");
        }
        if (n_y < 1)
        {
            throw std::invalid_argument("PMFTXYZ requires at least 1 bin in Y.");
        }
        if (n_z < 1)
        {
            throw std::invalid_argument("PMFTXYZ requires at least 1 bin in Z.");
        }
        if (x_max < 0)
        {
            throw std::invalid_argument("PMFTXYZ requires that x_max must be positive.");
        }
        if (y_max < 0)
        {
            throw std::invalid_argument("PMFTXYZ requires that y_max must be positive.");
        }
        if (z_max < 0)
        {
            throw std::invalid_argument("PMFTXYZ requires that z_max must be positive.");
        }
        volatile float dx = 2.0F * x_max / static_cast<float>(n_x); // PMFTXYZ.cc:53-72
        volatile float dy = 2.0F * y_max / static_cast<float>(n_y);
        volatile float dz = 2.0F * z_max / static_cast<float>(n_z);
        volatile float jac = dx * dy;
        jac = jac * dz;
        volatile float inv = 1.0F / jac;
        m_jacobian_factor = inv;
        setAxis(0, -x_max, x_max);
        setAxis(1, -y_max, y_max);
        setAxis(2, -z_max, z_max);
        allocate();
    }

    void reset()
    {
        PMFT3::reset();
        m_num_equiv_orientations = 0xffffffffU; // PMFTXYZ.cc:101-105
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const quat<float>* query_orientations,
                    const vec3<float>* query_points, unsigned int n_query_points, const quat<float>* equiv_orientations,
                    unsigned int num_equiv_orientations, const std::shared_ptr<locality::NeighborList>& nlist,
                    locality::QueryArgs qargs)
    {
        if (m_num_equiv_orientations == 0xffffffffU) // PMFTXYZ.cc:116-126
        {
            m_num_equiv_orientations = num_equiv_orientations;
        }
        else if (m_num_equiv_orientations != num_equiv_orientations)
        {
            throw std::runtime_error(
                "The number of equivalent orientations must be constant while accumulating data into PMFTXYZ.");
        }
        if (neighbor_query->getBox().is2D()) // Box::enforce3D, Box.h:579-585
        {
            throw std::invalid_argument("A 2D box was provided to a class that only supports 3D systems.");
        }
        accumulateDevice(neighbor_query, query_points, n_query_points, nlist, qargs, nullptr,
                         reinterpret_cast<const float*>(query_orientations),
                         reinterpret_cast<const float*>(equiv_orientations), num_equiv_orientations);
    }

private:
    float normExtra() const override { return static_cast<float>(m_num_equiv_orientations); }
    float jacobianFactor(size_t) const override { return m_jacobian_factor; }
    float m_jacobian_factor {1.0F};
    unsigned int m_num_equiv_orientations {0xffffffffU};
};

// PMFTXYT(x_max, y_max, n_x, n_y, n_t); accumulate(neighbor_query, orientations, query_points, query_orientations,
// n_query_points, nlist, qargs) -- freud/pmft/PMFTXYT.h, PMFTXYT.cc:28-101
class PMFTXYT : public PMFT3
{
public:
    PMFTXYT(float x_max, float y_max, unsigned int n_x, unsigned int n_y, unsigned int n_t)
        : PMFT3(FGPU_PMFT_XYT, x_max, y_max, 0.0F, n_x, n_y, n_t)
    {
        if (n_x < 1)
        {
            throw std::invalid_argument("PMFTXYT requires at least 1 bin in X.");
        }
        if (n_y < 1)
        {
            throw std::invalid_argument("PMFTXYT requires at least 1 bin in Y.");
        }
        if (n_t < 1)
        {
            throw std::invalid_argument("PMFTXYT requires at least 1 bin in T.");
        }
        if (x_max < 0)
        {
            throw std::invalid_argument("PMFTXYT requires that x_max must be positive.");
        }
        if (y_max < 0)
        {
            throw std::invalid_argument("PMFTXYT requires that y_max must be positive.");
        }
        volatile float dx = 2.0F * x_max / static_cast<float>(n_x); // PMFTXYT.cc:51-59
        volatile float dy = 2.0F * y_max / static_cast<float>(n_y);
        volatile float dt = 1 / static_cast<float>(n_t);
        volatile float jac = dx * dy;
        jac = jac * dt;
        volatile float inv = 1.0F / jac;
        m_jacobian_factor = inv;
        setAxis(0, -x_max, x_max);
        setAxis(1, -y_max, y_max);
        setAxis(2, 0.0F, kPmftTwoPi);
        allocate();
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const float* orientations,
                    const vec3<float>* query_points, const float* query_orientations, unsigned int n_query_points,
                    const std::shared_ptr<locality::NeighborList>& nlist, locality::QueryArgs qargs)
    {
        if (!neighbor_query->getBox().is2D()) // Box::enforce2D, Box.h:571-577
        {
            throw std::invalid_argument("A 3D box was provided to a class that only supports 2D systems.");
        }
        accumulateDevice(neighbor_query, query_points, n_query_points, nlist, qargs, orientations, query_orientations,
                         nullptr, 0);
    }

private:
    float jacobianFactor(size_t) const override { return m_jacobian_factor; }
    float m_jacobian_factor {1.0F};
};

// PMFTR12(r_max, n_r, n_t1, n_t2); accumulate as PMFTXYT -- freud/pmft/PMFTR12.h, PMFTR12.cc:28-113
class PMFTR12 : public PMFT3
{
public:
    PMFTR12(float r_max, unsigned int n_r, unsigned int n_t1, unsigned int n_t2)
        : PMFT3(FGPU_PMFT_R12, r_max, 0.0F, 0.0F, n_r, n_t1, n_t2)
    {
        if (n_r < 1)
        {
            throw std::invalid_argument("PMFTR12 requires at least 1 bin in R.");
        }
        if (n_t1 < 1)
        {
            throw std::invalid_argument("PMFTR12 requires at least 1 bin in T1.");
        }
        if (n_t2 < 1)
        {
            throw std::invalid_argument("PMFTR12 requires at least 1 bin in T2.");
        }
        if (r_max < 0)
        {
            throw std::invalid_argument("PMFTR12 requires that r_max must be positive.");
        }
        setAxis(0, 0.0F, r_max);
        setAxis(1, 0.0F, kPmftTwoPi);
        setAxis(2, 0.0F, kPmftTwoPi);
        // inverse Jacobian per r bin: 1 / (r_centre dr dt1 dt2), PMFTR12.cc:62-79
        std::vector<float> const r_centres = centers(m_edges[0]);
        volatile float dr = r_max / static_cast<float>(n_r);
        volatile float dt1 = kPmftTwoPi / static_cast<float>(n_t1);
        volatile float dt2 = 1 / static_cast<float>(n_t2);
        volatile float product = dr * dt1;
        product = product * dt2;
        m_inv_jacobian.resize(n_r);
        for (size_t i = 0; i < n_r; ++i)
        {
            volatile float rp = r_centres[i] * product;
            m_inv_jacobian[i] = 1.0F / rp;
        }
        allocate();
    }

    void accumulate(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const float* orientations,
                    const vec3<float>* query_points, const float* query_orientations, unsigned int n_query_points,
                    const std::shared_ptr<locality::NeighborList>& nlist, locality::QueryArgs qargs)
    {
        if (!neighbor_query->getBox().is2D())
        {
            throw std::invalid_argument("A 3D box was provided to a class that only supports 2D systems.");
        }
        accumulateDevice(neighbor_query, query_points, n_query_points, nlist, qargs, orientations, query_orientations,
                         nullptr, 0);
    }

private:
    float jacobianFactor(size_t bin) const override { return m_inv_jacobian[bin / (m_n[1] * m_n[2])]; }
    std::vector<float> m_inv_jacobian;
};

}} // namespace freud::pmft
