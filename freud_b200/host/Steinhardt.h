// freud::order::Steinhardt on the GPU kernels of libfreud_b200.so.
//
// Signatures: Steinhardt(ls, average, wl, weighted, wl_normalize) (freud/order/Steinhardt.h:66-90),
// compute(nlist /*nullable*/, points, qargs) (Steinhardt.h:153-155), getQl / getQlm / getParticleOrder /
// getOrder / getL / is* (Steinhardt.h:96-150; bound in export-Steinhardt.cc:22-33).
// compute() = reallocateArrays + baseCompute [+ computeAve] [+ aggregatewl] + normalizeSystem
// (Steinhardt.cc:54-118, 120-222, 224-289, 329-359, 291-327): per particle
// q_lm(i) = sum_j w_ij Y_lm(wrap(p_j - p_i)) / sum_j w_ij and q_l(i) = sqrt(4 pi / (2l+1) sum_m |q_lm|^2);
// average: q_lm averaged over i and its neighbours; wl: the Wigner-3j contraction of q_lm (or of the averaged
// q_lm), optionally normalised.  Getters follow upstream: getQl() is the averaged q_l when average is set,
// getParticleOrder() is w_l when wl is set, getQlm() is always the un-averaged q_lm(i).
// Deviation, documented: the system-wide q_lm is accumulated in fp64 on the device (the reference's float32
// thread-order sum is not reproducible run to run).
#pragma once
#include <complex>
#include <memory>
#include <vector>

#include "Context.h"
#include "ManagedArray.h"
#include "NeighborList.h"
#include "NeighborQuery.h"

namespace freud { namespace order {

class Steinhardt
{
public:
    explicit Steinhardt(const std::vector<unsigned int>& ls, bool average = false, bool wl = false,
                        bool weighted = false, bool wl_normalize = false)
        : m_ls(ls), m_average(average), m_wl(wl), m_weighted(weighted), m_wl_normalize(wl_normalize),
          m_qlmi(ls.size())
    {
        if (ls.empty())
        {
            throw std::invalid_argument("Steinhardt requires at least one l.");
        }
    }
    explicit Steinhardt(unsigned int l, bool average = false, bool wl = false, bool weighted = false,
                        bool wl_normalize = false)
        : Steinhardt(std::vector<unsigned int> {l}, average, wl, weighted, wl_normalize)
    {}

    unsigned int getNP() const { return m_Np; }
    const std::shared_ptr<util::ManagedArray<float>>& getParticleOrder() const { return m_wl ? m_wli : m_qli; }
    const std::shared_ptr<util::ManagedArray<float>>& getQl() const { return m_qli; }
    // The per-particle q_lm stay on the device after compute() (104 MB at l = 6, N = 1e6) and cross the link the
    // first time somebody asks for them: one copy into one page-locked block, the per-l arrays are slices of it.
    const std::vector<std::shared_ptr<util::ManagedArray<std::complex<float>>>>& getQlm() const
    {
        if (m_qlm_dev)
        {
            size_t tot_m = 0;
            for (unsigned int l : m_ls)
            {
                tot_m += 2 * (size_t) l + 1;
            }
            size_t const bytes = (size_t) m_Np * tot_m * sizeof(std::complex<float>);
            auto block = std::make_shared<util::HostBlock>(bytes);
            gpu::check(fgpu_buffer_read(m_qlm_dev.get(), block->get(), 0, bytes));
            auto* base = static_cast<std::complex<float>*>(block->get());
            size_t off = 0;
            for (size_t r = 0; r < m_ls.size(); ++r)
            {
                size_t const nm = 2 * (size_t) m_ls[r] + 1;
                m_qlmi[r] = std::make_shared<util::ManagedArray<std::complex<float>>>(
                    block, base + off, std::vector<size_t> {m_Np, nm});
                off += (size_t) m_Np * nm;
            }
            m_qlm_dev.reset();
        }
        return m_qlmi;
    }
    std::vector<float> getOrder() const { return m_norm; }
    bool isAverage() const { return m_average; }
    bool isWl() const { return m_wl; }
    bool isWeighted() const { return m_weighted; }
    bool isWlNormalized() const { return m_wl_normalize; }
    std::vector<unsigned int> getL() const { return m_ls; }

    void compute(const std::shared_ptr<locality::NeighborList>& nlist,
                 const std::shared_ptr<locality::NeighborQuery>& points, const locality::QueryArgs& qargs)
    {
        unsigned int const Np = points->getNPoints();
        // Neighbours: the list handed in, or the default query over the points themselves
        // (loopOverNeighborsIterator, NeighborComputeFunctional.h:112-150).  A nearest-neighbour query made for this
        // compute alone never becomes a list: fgpu_steinhardt_knn searches and accumulates in one call.
        std::shared_ptr<locality::NeighborList> list = nlist;
        locality::QueryArgs knn_args = qargs;
        bool fused = false;
        if (!list)
        {
            points->validateQueryArgs(knn_args); // infers the mode, fills the defaults, raises like query() would
            vec3<bool> const periodic = points->getBox().getPeriodic();
            if (!(periodic.x && periodic.y && periodic.z))
            {
                throw std::domain_error("Pair queries in a non-periodic box are not implemented.");
            }
            fused = knn_args.mode == locality::QueryType::nearest && !m_average && !m_wl
                && points->getFlavour() != FGPU_FLAVOUR_GHOST;
            if (!fused)
            {
                list = points->query(points->getPoints(), Np, qargs)->toNeighborList();
            }
        }
        else
        {
            list->validate(Np, Np);
        }
        // fresh outputs every call (Steinhardt.cc:54-83)
        m_Np = Np;
        size_t tot_m = 0;
        for (unsigned int l : m_ls)
        {
            tot_m += 2 * (size_t) l + 1;
        }
        // every element of these is written by the copy that fills them
        auto qli = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {Np, m_ls.size()}, util::Uninitialized {});
        std::shared_ptr<util::ManagedArray<float>> wli;
        if (m_wl)
        {
            wli = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {Np, m_ls.size()}, util::Uninitialized {});
        }
        std::vector<float> sys(tot_m * 2);
        std::vector<float> order(m_ls.size());
        int const flags = (m_weighted ? FGPU_ST_WEIGHTED : 0) | (m_average ? FGPU_ST_AVERAGE : 0)
            | (m_wl ? FGPU_ST_WL : 0) | (m_wl_normalize ? FGPU_ST_WL_NORMALIZE : 0);
        fgpu_buffer* keep = nullptr;
        if (fused)
        {
            gpu::check(fgpu_steinhardt_knn(points->device(), points->getFlavour(), knn_args.num_neighbors, knn_args.r_max,
                                           knn_args.r_min, knn_args.exclude_ii ? 1 : 0, m_ls.data(), (uint32_t) m_ls.size(),
                                           flags, qli->data(), nullptr, &keep, sys.data(), order.data()));
        }
        else
        {
            gpu::check(fgpu_steinhardt_compute_keep(points->device(), list->device(gpu::context()), m_ls.data(),
                                                    (uint32_t) m_ls.size(), flags, Np, nullptr, qli->data(),
                                                    wli ? wli->data() : nullptr, &keep, sys.data(), order.data()));
        }
        m_qlm_dev = std::shared_ptr<fgpu_buffer>(keep, fgpu_buffer_destroy);
        for (auto& arr : m_qlmi)
        {
            arr.reset(); // views handed out earlier keep their own block alive
        }
        m_qli = qli;
        m_wli = wli;
        m_norm = order;
    }

private:
    unsigned int m_Np {0};
    std::vector<unsigned int> m_ls;
    bool m_average, m_wl, m_weighted, m_wl_normalize;
    std::shared_ptr<util::ManagedArray<float>> m_qli; // q_l, or the averaged q_l (what getQl() returns upstream)
    std::shared_ptr<util::ManagedArray<float>> m_wli; // w_l when wl is set
    mutable std::vector<std::shared_ptr<util::ManagedArray<std::complex<float>>>> m_qlmi;
    mutable std::shared_ptr<fgpu_buffer> m_qlm_dev; // q_lm of the last compute(), not yet read
    std::vector<float> m_norm;
};

}} // namespace freud::order
