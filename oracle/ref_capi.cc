// TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).  Never linked into the product.
//
// C entry points over the UNMODIFIED reference sources (compiled where they lie under /root/reference
// by oracle/Makefile into oracle/_ref/libfreud_ref.so).  The wrappers only construct the reference's
// own classes and copy their outputs; no arithmetic happens here.
//
// Reference classes driven:
//   freud::box::Box                                  freud/box/Box.h:44
//   freud::locality::{LinkCell,AABBQuery,RawPoints}  freud/locality/LinkCell.h:188, AABBQuery.h:42, RawPoints.h:35
//   freud::locality::CellQuery                       freud/locality/CellQuery.h:29
//   NeighborQuery::query / toNeighborList            freud/locality/NeighborQuery.h:130,434
//   freud::density::RDF                              freud/density/RDF.h:33
//   freud::density::LocalDensity                     freud/density/LocalDensity.h:29
//   freud::density::CorrelationFunction              freud/density/CorrelationFunction.h:52
//   freud::pmft::PMFTXY                              freud/pmft/PMFTXY.h
//   freud::pmft::PMFTXYZ, PMFTXYT, PMFTR12           freud/pmft/PMFTXYZ.h, PMFTXYT.h, PMFTR12.h
//   freud::order::Steinhardt                         freud/order/Steinhardt.h:66
//   freud::environment::BondOrder                    freud/environment/BondOrder.h:32
//   freud::locality::PeriodicBuffer                  freud/locality/PeriodicBuffer.h:22
//   freud::parallel::setNumThreads                   freud/parallel/tbb_config.cc:25

#include <complex>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "AABBQuery.h"
#include "BondOrder.h"
#include "Box.h"
#include "CellQuery.h"
#include "CorrelationFunction.h"
#include "LinkCell.h"
#include "LocalDensity.h"
#include "NeighborList.h"
#include "NeighborQuery.h"
#include "PMFTR12.h"
#include "PMFTXY.h"
#include "PMFTXYT.h"
#include "PMFTXYZ.h"
#include "PeriodicBuffer.h"
#include "RDF.h"
#include "RawPoints.h"
#include "Steinhardt.h"
#include "tbb_config.h"

using freud::box::Box;
using freud::locality::NeighborList;
using freud::locality::NeighborQuery;
using freud::locality::QueryArgs;
using freud::locality::QueryType;

namespace {

thread_local std::string g_error;
thread_local int g_error_kind = 0; // 1 invalid_argument, 2 domain_error, 3 runtime_error, 4 out_of_range, 5 other

template<typename F> int guarded(F&& f)
{
    try
    {
        f();
        g_error.clear();
        g_error_kind = 0;
        return 0;
    }
    catch (const std::invalid_argument& e)
    {
        g_error = e.what();
        g_error_kind = 1;
    }
    catch (const std::domain_error& e)
    {
        g_error = e.what();
        g_error_kind = 2;
    }
    catch (const std::out_of_range& e)
    {
        g_error = e.what();
        g_error_kind = 4;
    }
    catch (const std::runtime_error& e)
    {
        g_error = e.what();
        g_error_kind = 3;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        g_error_kind = 5;
    }
    return g_error_kind;
}

Box makeBox(const float* b, int is2d)
{
    return Box(b[0], b[1], b[2], b[3], b[4], b[5], is2d != 0);
}

struct QueryHandle
{
    std::vector<vec3<float>> points; // the reference keeps a non-owning pointer (NeighborQuery.h:295)
    std::shared_ptr<NeighborQuery> nq;
};

QueryArgs makeArgs(int mode, unsigned num_neighbors, float r_max, float r_min, float r_guess, float scale,
                   int exclude_ii)
{
    QueryArgs a;
    a.mode = mode == 1 ? QueryType::ball : (mode == 2 ? QueryType::nearest : QueryType::none);
    a.num_neighbors = num_neighbors;
    a.r_max = r_max;
    a.r_min = r_min;
    a.r_guess = r_guess;
    a.scale = scale;
    a.exclude_ii = exclude_ii != 0;
    return a;
}

} // namespace

extern "C" {

const char* fref_last_error()
{
    return g_error.c_str();
}

int fref_last_error_kind()
{
    return g_error_kind;
}

void fref_set_num_threads(unsigned n)
{
    freud::parallel::setNumThreads(n);
}

// ---- Box ------------------------------------------------------------------------------------------
// op: 0 wrap, 1 makeFractional, 2 makeAbsolute
int fref_box_apply(const float* box6, int is2d, int op, const float* in, unsigned n, float* out)
{
    return guarded([&] {
        Box const box = makeBox(box6, is2d);
        const auto* v = reinterpret_cast<const vec3<float>*>(in);
        auto* o = reinterpret_cast<vec3<float>*>(out);
        for (unsigned i = 0; i < n; ++i)
        {
            o[i] = op == 0 ? box.wrap(v[i]) : (op == 1 ? box.makeFractional(v[i]) : box.makeAbsolute(v[i]));
        }
    });
}

int fref_box_info(const float* box6, int is2d, float* volume, float* plane_dist3)
{
    return guarded([&] {
        Box const box = makeBox(box6, is2d);
        *volume = box.getVolume();
        vec3<float> const d = box.getNearestPlaneDistance();
        plane_dist3[0] = d.x;
        plane_dist3[1] = d.y;
        plane_dist3[2] = d.z;
    });
}

// ---- NeighborQuery --------------------------------------------------------------------------------
// engine: 0 LinkCell, 1 AABBQuery, 2 RawPoints, 3 CellQuery
void* fref_nq_create(int engine, const float* box6, int is2d, const float* pts, unsigned n, float cell_width)
{
    QueryHandle* h = nullptr;
    int const rc = guarded([&] {
        auto handle = std::make_unique<QueryHandle>();
        const auto* p = reinterpret_cast<const vec3<float>*>(pts);
        handle->points.assign(p, p + n);
        Box const box = makeBox(box6, is2d);
        if (engine == 0)
        {
            handle->nq = std::make_shared<freud::locality::LinkCell>(box, handle->points.data(), n, cell_width);
        }
        else if (engine == 1)
        {
            handle->nq = std::make_shared<freud::locality::AABBQuery>(box, handle->points.data(), n);
        }
        else if (engine == 3)
        {
            handle->nq = std::make_shared<freud::locality::CellQuery>(box, handle->points.data(), n);
        }
        else
        {
            handle->nq = std::make_shared<freud::locality::RawPoints>(box, handle->points.data(), n);
        }
        h = handle.release();
    });
    return rc == 0 ? h : nullptr;
}

void fref_nq_destroy(void* nq)
{
    delete static_cast<QueryHandle*>(nq);
}

void* fref_query_nlist(void* nq, const float* qpts, unsigned n_query, int mode, unsigned num_neighbors,
                       float r_max, float r_min, float r_guess, float scale, int exclude_ii,
                       int sort_by_distance)
{
    std::shared_ptr<NeighborList>* out = nullptr;
    int const rc = guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        QueryArgs const args = makeArgs(mode, num_neighbors, r_max, r_min, r_guess, scale, exclude_ii);
        auto iter = h->nq->query(reinterpret_cast<const vec3<float>*>(qpts), n_query, args);
        out = new std::shared_ptr<NeighborList>(iter->toNeighborList(sort_by_distance != 0));
    });
    return rc == 0 ? out : nullptr;
}

unsigned fref_nlist_num_bonds(void* nl)
{
    return (*static_cast<std::shared_ptr<NeighborList>*>(nl))->getNumBonds();
}

void fref_nlist_copy(void* nl, unsigned* neighbors, float* distances, float* weights, float* vectors)
{
    auto& l = *static_cast<std::shared_ptr<NeighborList>*>(nl);
    size_t const nb = l->getNumBonds();
    if (nb == 0)
    {
        return;
    }
    std::memcpy(neighbors, l->getNeighbors()->data(), nb * 2 * sizeof(unsigned));
    std::memcpy(distances, l->getDistances()->data(), nb * sizeof(float));
    std::memcpy(weights, l->getWeights()->data(), nb * sizeof(float));
    std::memcpy(vectors, l->getVectors()->data(), nb * 3 * sizeof(float));
}

void fref_nlist_segments(void* nl, unsigned* segments, unsigned* counts)
{
    auto& l = *static_cast<std::shared_ptr<NeighborList>*>(nl);
    size_t const nq = l->getNumQueryPoints();
    if (nq == 0)
    {
        return;
    }
    std::memcpy(segments, l->getSegments()->data(), nq * sizeof(unsigned));
    std::memcpy(counts, l->getCounts()->data(), nq * sizeof(unsigned));
}

void fref_nlist_destroy(void* nl)
{
    delete static_cast<std::shared_ptr<NeighborList>*>(nl);
}

// ---- RDF ------------------------------------------------------------------------------------------
void* fref_rdf_create(unsigned bins, float r_max, float r_min, int finite_size_mode)
{
    freud::density::RDF* r = nullptr;
    int const rc = guarded([&] {
        r = new freud::density::RDF(bins, r_max, r_min);
        r->mode = finite_size_mode != 0 ? freud::density::NormalizationMode::finite_size
                                        : freud::density::NormalizationMode::exact;
    });
    return rc == 0 ? r : nullptr;
}

void fref_rdf_destroy(void* rdf)
{
    delete static_cast<freud::density::RDF*>(rdf);
}

int fref_rdf_reset(void* rdf)
{
    return guarded([&] { static_cast<freud::density::RDF*>(rdf)->reset(); });
}

int fref_rdf_accumulate(void* rdf, void* nq, const float* qpts, unsigned n_query, void* nlist_or_null, int mode,
                        unsigned num_neighbors, float r_max, float r_min, float r_guess, float scale,
                        int exclude_ii)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(mode, num_neighbors, r_max, r_min, r_guess, scale, exclude_ii);
        static_cast<freud::density::RDF*>(rdf)->accumulate(h->nq, reinterpret_cast<const vec3<float>*>(qpts),
                                                           n_query, nl, args);
    });
}

int fref_rdf_get(void* rdf, unsigned* bin_counts, float* g_r, float* n_r, float* bin_edges, float* bin_centers)
{
    return guarded([&] {
        auto* r = static_cast<freud::density::RDF*>(rdf);
        size_t const bins = r->getAxisSizes()[0];
        // order matters: each getter triggers the lazy reduce (BondHistogramCompute.h:61-69)
        auto counts = r->getBinCounts();
        auto g = r->getRDF();
        auto n = r->getNr();
        std::memcpy(bin_counts, counts->data(), bins * sizeof(unsigned));
        std::memcpy(g_r, g->data(), bins * sizeof(float));
        std::memcpy(n_r, n->data(), bins * sizeof(float));
        if (bin_edges != nullptr)
        {
            auto e = r->getBinEdges()[0];
            std::memcpy(bin_edges, e.data(), (bins + 1) * sizeof(float));
        }
        if (bin_centers != nullptr)
        {
            auto c = r->getBinCenters()[0];
            std::memcpy(bin_centers, c.data(), bins * sizeof(float));
        }
    });
}

// ---- CorrelationFunction --------------------------------------------------------------------------
// CorrelationFunction(bins, r_max).accumulate(nq, values, query_points, query_values, n, nlist /*nullable*/, qargs)
// once; outputs: correlation complex128[bins] (re, im interleaved), bin counts u32[bins]
int fref_correlation(void* nq, const double* values, const float* qpts, const double* query_values, unsigned n_query,
                     void* nlist_or_null, unsigned bins, float r_max, int exclude_ii, double* correlation,
                     unsigned* bin_counts)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(/*ball*/ 1, 0xffffffffU, r_max, 0.0F, -1.0F, -1.0F, exclude_ii);
        freud::density::CorrelationFunction cf(bins, r_max);
        cf.accumulate(h->nq, reinterpret_cast<const std::complex<double>*>(values),
                      reinterpret_cast<const vec3<float>*>(qpts),
                      reinterpret_cast<const std::complex<double>*>(query_values), n_query, nl, args);
        std::memcpy(correlation, cf.getCorrelation()->data(), bins * sizeof(std::complex<double>));
        std::memcpy(bin_counts, cf.getBinCounts()->data(), bins * sizeof(unsigned));
    });
}

// ---- PMFTXY -----------------------------------------------------------------------------------------
// PMFTXY(x_max, y_max, n_x, n_y).accumulate(nq, query_orientations, query_points, n, nlist /*nullable*/, qargs) once;
// outputs: bin counts u32[n_x * n_y], pcf f32[n_x * n_y]
int fref_pmftxy(void* nq, const float* query_orientations, const float* qpts, unsigned n_query, void* nlist_or_null,
                float x_max, float y_max, unsigned n_x, unsigned n_y, float r_max, int exclude_ii, unsigned* bin_counts,
                float* pcf)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(/*ball*/ 1, 0xffffffffU, r_max, 0.0F, -1.0F, -1.0F, exclude_ii);
        freud::pmft::PMFTXY p(x_max, y_max, n_x, n_y);
        p.accumulate(h->nq, query_orientations, reinterpret_cast<const vec3<float>*>(qpts), n_query, nl, args);
        std::memcpy(bin_counts, p.getBinCounts()->data(), size_t(n_x) * n_y * sizeof(unsigned));
        std::memcpy(pcf, p.getPCF()->data(), size_t(n_x) * n_y * sizeof(float));
    });
}

// ---- PMFTXYZ (kind 0), PMFTXYT (1), PMFTR12 (2) -----------------------------------------------------
// One accumulate of a fresh object; outputs: bin counts u32[n0 * n1 * n2], pcf f32[n0 * n1 * n2].  Orientations are
// quaternions (s, x, y, z) for XYZ (`orientations` unused, `equiv` = n_equiv quaternions) and angles otherwise.
int fref_pmft3(int kind, void* nq, const float* orientations, const float* query_orientations, const float* qpts,
               unsigned n_query, const float* equiv, unsigned n_equiv, void* nlist_or_null, float max0, float max1,
               float max2, unsigned n0, unsigned n1, unsigned n2, float r_max, int exclude_ii, unsigned* bin_counts,
               float* pcf)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(/*ball*/ 1, 0xffffffffU, r_max, 0.0F, -1.0F, -1.0F, exclude_ii);
        auto const* qp = reinterpret_cast<const vec3<float>*>(qpts);
        size_t const n_bins = size_t(n0) * n1 * n2;
        auto copy_out = [&](auto& p) {
            std::memcpy(bin_counts, p.getBinCounts()->data(), n_bins * sizeof(unsigned));
            std::memcpy(pcf, p.getPCF()->data(), n_bins * sizeof(float));
        };
        if (kind == 0)
        {
            freud::pmft::PMFTXYZ p(max0, max1, max2, n0, n1, n2);
            p.accumulate(h->nq, reinterpret_cast<const quat<float>*>(query_orientations), qp, n_query,
                         reinterpret_cast<const quat<float>*>(equiv), n_equiv, nl, args);
            copy_out(p);
        }
        else if (kind == 1)
        {
            freud::pmft::PMFTXYT p(max0, max1, n0, n1, n2);
            p.accumulate(h->nq, orientations, qp, query_orientations, n_query, nl, args);
            copy_out(p);
        }
        else
        {
            freud::pmft::PMFTR12 p(max0, n0, n1, n2);
            p.accumulate(h->nq, orientations, qp, query_orientations, n_query, nl, args);
            copy_out(p);
        }
    });
}

// ---- BondOrder --------------------------------------------------------------------------------------
// BondOrder(n_theta, n_phi, mode).accumulate(nq, orientations, query_points, query_orientations, n, nlist, qargs) once;
// outputs: bin counts u32[n_theta * n_phi], bond order f32[n_theta * n_phi]
int fref_bond_order(int mode, void* nq, const float* orientations, const float* qpts, const float* query_orientations,
                    unsigned n_query, void* nlist_or_null, unsigned n_theta, unsigned n_phi, int query_mode,
                    unsigned num_neighbors, float r_max, int exclude_ii, unsigned* bin_counts, float* bond_order)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(query_mode, num_neighbors, r_max, 0.0F, -1.0F, -1.0F, exclude_ii);
        freud::environment::BondOrder bo(n_theta, n_phi, static_cast<freud::environment::BondOrderMode>(mode));
        bo.accumulate(h->nq, reinterpret_cast<const quat<float>*>(orientations),
                      reinterpret_cast<const vec3<float>*>(qpts), reinterpret_cast<const quat<float>*>(query_orientations),
                      n_query, nl, args);
        std::memcpy(bin_counts, bo.getBinCounts()->data(), size_t(n_theta) * n_phi * sizeof(unsigned));
        std::memcpy(bond_order, bo.getBondOrder()->data(), size_t(n_theta) * n_phi * sizeof(float));
    });
}

// ---- LocalDensity ---------------------------------------------------------------------------------
// LocalDensity(r_max, diameter).compute(nq, query_points, n, nlist /*nullable*/, qargs); outputs n_query floats each
int fref_local_density(void* nq, const float* qpts, unsigned n_query, void* nlist_or_null, float r_max, float diameter,
                       float q_r_max, int exclude_ii, float* num_neighbors, float* density)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(/*ball*/ 1, 0xffffffffU, q_r_max, 0.0F, -1.0F, -1.0F, exclude_ii);
        freud::density::LocalDensity ld(r_max, diameter);
        ld.compute(h->nq, reinterpret_cast<const vec3<float>*>(qpts), n_query, nl, args);
        std::memcpy(num_neighbors, ld.getNumNeighbors()->data(), n_query * sizeof(float));
        std::memcpy(density, ld.getDensity()->data(), n_query * sizeof(float));
    });
}

// ---- PeriodicBuffer -------------------------------------------------------------------------------
// PeriodicBuffer().compute(nq, buffer, images, include_input_points).  Returns an owning handle; the count, the
// grown box and the arrays are read with the calls below.
void* fref_pbuff_compute(void* nq, const float* buffer3, int images, int include_input_points)
{
    freud::locality::PeriodicBuffer* pb = nullptr;
    int const rc = guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        auto fresh = std::make_unique<freud::locality::PeriodicBuffer>();
        fresh->compute(h->nq, {buffer3[0], buffer3[1], buffer3[2]}, images != 0, include_input_points != 0);
        pb = fresh.release();
    });
    return rc == 0 ? pb : nullptr;
}

unsigned fref_pbuff_size(void* pb)
{
    return static_cast<unsigned>(static_cast<freud::locality::PeriodicBuffer*>(pb)->getBufferIds()->size());
}

void fref_pbuff_copy(void* pb, float* points, unsigned* ids, float* box6)
{
    auto* b = static_cast<freud::locality::PeriodicBuffer*>(pb);
    auto const& pts = *b->getBufferPoints();
    auto const& id = *b->getBufferIds();
    if (!id.empty())
    {
        std::memcpy(points, pts.data(), pts.size() * sizeof(vec3<float>));
        std::memcpy(ids, id.data(), id.size() * sizeof(unsigned));
    }
    Box const& box = b->getBufferBox();
    box6[0] = box.getLx();
    box6[1] = box.getLy();
    box6[2] = box.getLz();
    box6[3] = box.getTiltFactorXY();
    box6[4] = box.getTiltFactorXZ();
    box6[5] = box.getTiltFactorYZ();
}

void fref_pbuff_destroy(void* pb)
{
    delete static_cast<freud::locality::PeriodicBuffer*>(pb);
}

// ---- Steinhardt -----------------------------------------------------------------------------------
void* fref_steinhardt_create(const unsigned* ls, unsigned n_ls, int average, int wl, int weighted, int wl_normalize)
{
    freud::order::Steinhardt* s = nullptr;
    int const rc = guarded([&] {
        s = new freud::order::Steinhardt(std::vector<unsigned>(ls, ls + n_ls), average != 0, wl != 0,
                                         weighted != 0, wl_normalize != 0);
    });
    return rc == 0 ? s : nullptr;
}

void fref_steinhardt_destroy(void* st)
{
    delete static_cast<freud::order::Steinhardt*>(st);
}

int fref_steinhardt_compute(void* st, void* nq, void* nlist_or_null, int mode, unsigned num_neighbors, float r_max,
                            float r_min, float r_guess, float scale, int exclude_ii)
{
    return guarded([&] {
        auto* h = static_cast<QueryHandle*>(nq);
        std::shared_ptr<NeighborList> nl;
        if (nlist_or_null != nullptr)
        {
            nl = *static_cast<std::shared_ptr<NeighborList>*>(nlist_or_null);
        }
        QueryArgs const args = makeArgs(mode, num_neighbors, r_max, r_min, r_guess, scale, exclude_ii);
        static_cast<freud::order::Steinhardt*>(st)->compute(nl, h->nq, args);
    });
}

// particle_order: N x n_ls (ql, or wl when wl=true); ql: N x n_ls; order: n_ls
int fref_steinhardt_get(void* st, float* particle_order, float* ql, float* order)
{
    return guarded([&] {
        auto* s = static_cast<freud::order::Steinhardt*>(st);
        size_t const n = size_t(s->getNP()) * s->getL().size();
        if (particle_order != nullptr)
        {
            std::memcpy(particle_order, s->getParticleOrder()->data(), n * sizeof(float));
        }
        if (ql != nullptr)
        {
            std::memcpy(ql, s->getQl()->data(), n * sizeof(float));
        }
        if (order != nullptr)
        {
            auto o = s->getOrder();
            std::memcpy(order, o.data(), o.size() * sizeof(float));
        }
    });
}

// qlm for one l index: N x (2l+1) complex64, m order 0..l,-1..-l (Steinhardt.cc:31-52)
int fref_steinhardt_get_qlm(void* st, unsigned l_index, float* out_complex)
{
    return guarded([&] {
        auto* s = static_cast<freud::order::Steinhardt*>(st);
        auto const& qlm = s->getQlm()[l_index];
        std::memcpy(out_complex, qlm->data(), qlm->size() * sizeof(std::complex<float>));
    });
}

} // extern "C"
