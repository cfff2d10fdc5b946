// freud::density::LocalDensity on the GPU path.
//
// Signatures: LocalDensity(r_max, diameter) (freud/density/LocalDensity.h:29, LocalDensity.cc:25-36),
// compute(neighbor_query, query_points, n_query_points, nlist /*nullable*/, qargs) (LocalDensity.h:51-54,
// LocalDensity.cc:38-84), getDensity / getNumNeighbors / getBox / getRMax / getDiameter (LocalDensity.h:33-70; bound in
// export-LocalDensity.cc:37-47).  The neighbours are the list handed in or the query over the points
// (loopOverNeighborsIterator, NeighborComputeFunctional.h:112-150) materialised as a NeighborList on the device;
// the fractional count is then one kernel over its rows (fgpu_local_density).  Upstream sums the bonds in the
// engine's traversal order when it queries on the fly; here the order is the sorted list's, so counts agree to
// float summation order (~1e-7 relative) and bit for bit when a NeighborList is passed in.
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include "Context.h"
#include "ManagedArray.h"
#include "NeighborList.h"
#include "NeighborQuery.h"

namespace freud { namespace density {

class LocalDensity
{
public:
    LocalDensity(float r_max, float diameter) : m_r_max(r_max), m_diameter(diameter)
    {
        if (r_max <= 0)
        {
            throw std::invalid_argument("LocalDensity requires r_max to be positive.");
        }
        if (diameter < 0)
        {
            throw std::invalid_argument("LocalDensity requires diameter to be non-negative.");
        }
    }

    const box::Box& getBox() const { return m_box; }
    float getRMax() const { return m_r_max; }
    float getDiameter() const { return m_diameter; }

    void compute(const std::shared_ptr<locality::NeighborQuery>& neighbor_query, const vec3<float>* query_points,
                 unsigned int n_query_points, const std::shared_ptr<locality::NeighborList>& nlist,
                 const locality::QueryArgs& qargs)
    {
        m_box = neighbor_query->getBox();
        // fresh outputs every call (LocalDensity.cc:45-46)
        auto density = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {n_query_points});
        auto counts = std::make_shared<util::ManagedArray<float>>(std::vector<size_t> {n_query_points});
        std::shared_ptr<locality::NeighborList> list = nlist;
        if (!list)
        {
            auto query = neighbor_query->query(query_points, n_query_points, qargs); // validates, infers the mode
            locality::QueryArgs const& args = query->getQueryArgs();
            if (args.mode == locality::QueryType::ball)
            {
                // a ball query made for this compute alone: the bonds are summed where the search left them
                gpu::check(fgpu_local_density_query(neighbor_query->device(),
                                                    locality::selfOrHost(*neighbor_query, query_points, n_query_points),
                                                    n_query_points, neighbor_query->getFlavour(), args.r_max, args.r_min,
                                                    args.exclude_ii ? 1 : 0, m_r_max, m_diameter, counts->data(),
                                                    density->data()));
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
            gpu::check(fgpu_local_density(list->device(gpu::context()), m_r_max, m_diameter, m_box.is2D() ? 1 : 0,
                                          counts->data(), density->data()));
        }
        m_density_array = density;
        m_num_neighbors_array = counts;
    }

    std::shared_ptr<const util::ManagedArray<float>> getDensity() const { return m_density_array; }
    std::shared_ptr<const util::ManagedArray<float>> getNumNeighbors() const { return m_num_neighbors_array; }

private:
    box::Box m_box;
    float m_r_max;
    float m_diameter;
    std::shared_ptr<util::ManagedArray<float>> m_density_array;
    std::shared_ptr<util::ManagedArray<float>> m_num_neighbors_array;
};

}} // namespace freud::density
