// Python bindings of the C++ host classes (module freud_b200._freud_b200).
//
// The reference binds the same methods with nanobind (freud/locality/export-NeighborQuery.cc,
// export-NeighborList.cc, export-BondHistogramCompute.cc, freud/density/export-RDF.cc,
// freud/order/export-Steinhardt.cc, freud/util/export-ManagedArray.h); nanobind is not installed in this
// image, so the binding layer is written against pybind11 with the same conventions: inputs are C-contiguous
// float32 (N, 3) arrays reinterpreted as vec3<float>*, outputs are zero-copy read-only numpy views whose
// base object keeps the owning ManagedArray alive, C++ exceptions surface as ValueError / RuntimeError /
// IndexError.  Submodules are named after the reference's extension modules (_box, _locality, _density,
// _order) so that freud's Python layer maps onto them one to one.
#include <array>
#include <thread>

#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "Box.h"
#include "NeighborList.h"
#include "NeighborQuery.h"
#include "CorrelationFunction.h"
#include "LocalDensity.h"
#include "BondOrder.h"
#include "PMFT.h"
#include "RDF.h"
#include "Steinhardt.h"

namespace py = pybind11;
using namespace freud;

namespace {

using points_array = py::array_t<float, py::array::c_style | py::array::forcecast>;

const vec3<float>* as_vec3(const points_array& a, unsigned int& n)
{
    if (a.ndim() != 2 || a.shape(1) != 3)
    {
        throw std::invalid_argument("points must have shape (N, 3)");
    }
    n = (unsigned int) a.shape(0);
    return reinterpret_cast<const vec3<float>*>(a.data());
}

// zero-copy, read-only view; the capsule owns a shared_ptr copy (export-ManagedArray.h:22-52)
template<typename T> py::array to_numpy(const std::shared_ptr<const util::ManagedArray<T>>& arr)
{
    if (!arr)
    {
        return py::array_t<T>(0);
    }
    auto* holder = new std::shared_ptr<const util::ManagedArray<T>>(arr);
    py::capsule owner(holder, [](void* p) { delete static_cast<std::shared_ptr<const util::ManagedArray<T>>*>(p); });
    std::vector<py::ssize_t> shape(arr->shape().begin(), arr->shape().end());
    py::array_t<T> out(shape, arr->data(), owner);
    py::detail::array_proxy(out.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
    return std::move(out);
}
template<typename T> py::array to_numpy(const std::shared_ptr<util::ManagedArray<T>>& arr)
{
    return to_numpy<T>(std::shared_ptr<const util::ManagedArray<T>>(arr));
}

} // namespace

PYBIND11_MODULE(_freud_b200, m)
{
    m.doc() = "C++ host classes of freud_b200 (reference-compatible signatures on the sm_100a C ABI)";
    m.def("version", [] { return std::string(fgpu_version()); });
    m.def("device_count", [] { return fgpu_device_count(); });
    // The private copy freud's Python layer keeps of the points (freud/locality.py:867-868), taken into a block of the
    // page-locked host cache: the upload that follows runs at the link's rate instead of through the driver's bounce
    // buffers.  Writable (N, 3) float32; the block returns to the cache when the array dies.
    m.def("private_points", [](points_array a) {
        unsigned int n = 0;
        const vec3<float>* src = as_vec3(a, n);
        auto* block = new std::shared_ptr<util::HostBlock>(std::make_shared<util::HostBlock>((size_t) n * 3 * sizeof(float)));
        py::capsule owner(block, [](void* p) { delete static_cast<std::shared_ptr<util::HostBlock>*>(p); });
        size_t const bytes = (size_t) n * 3 * sizeof(float);
        if (bytes >= (4U << 20))
        {
            // a 12 MB memcpy on one core costs more than the upload it precedes
            py::gil_scoped_release release;
            std::vector<std::thread> pool;
            unsigned int const n_threads = 4;
            size_t const chunk = (bytes / n_threads + 63) & ~(size_t) 63;
            for (unsigned int t = 0; t < n_threads; ++t)
            {
                size_t const lo = std::min(bytes, (size_t) t * chunk), hi = t + 1 == n_threads ? bytes : std::min(bytes, lo + chunk);
                pool.emplace_back([=] {
                    std::memcpy(static_cast<unsigned char*>((*block)->get()) + lo,
                                reinterpret_cast<const unsigned char*>(src) + lo, hi - lo);
                });
            }
            for (auto& th : pool)
            {
                th.join();
            }
        }
        else if (n != 0)
        {
            std::memcpy((*block)->get(), src, bytes);
        }
        return py::array_t<float>({(py::ssize_t) n, (py::ssize_t) 3}, static_cast<const float*>((*block)->get()), owner);
    });
    m.def("host_trim", [] { fgpu_host_trim(); });

    // ---- _box ------------------------------------------------------------------------------------------
    auto mbox = m.def_submodule("_box");
    py::class_<box::Box>(mbox, "Box")
        .def(py::init<float, float, float, float, float, float, bool>(), py::arg("Lx"), py::arg("Ly"), py::arg("Lz"),
             py::arg("xy") = 0.0F, py::arg("xz") = 0.0F, py::arg("yz") = 0.0F, py::arg("is2D") = false)
        .def("getLx", &box::Box::getLx)
        .def("getLy", &box::Box::getLy)
        .def("getLz", &box::Box::getLz)
        .def("getTiltFactorXY", &box::Box::getTiltFactorXY)
        .def("getTiltFactorXZ", &box::Box::getTiltFactorXZ)
        .def("getTiltFactorYZ", &box::Box::getTiltFactorYZ)
        .def("is2D", &box::Box::is2D)
        .def("getVolume", &box::Box::getVolume)
        .def("setPeriodic", &box::Box::setPeriodic)
        .def("getPeriodic",
             [](const box::Box& b) {
                 auto p = b.getPeriodic();
                 return py::make_tuple(p.x, p.y, p.z);
             })
        .def("getNearestPlaneDistance", [](const box::Box& b) {
            auto d = b.getNearestPlaneDistance();
            return py::make_tuple(d.x, d.y, d.z);
        });

    // ---- _locality -------------------------------------------------------------------------------------
    auto mloc = m.def_submodule("_locality");
    py::enum_<locality::QueryType>(mloc, "QueryType")
        .value("none", locality::QueryType::none)
        .value("ball", locality::QueryType::ball)
        .value("nearest", locality::QueryType::nearest);
    py::class_<locality::QueryArgs>(mloc, "QueryArgs")
        .def(py::init<>())
        .def_readwrite("mode", &locality::QueryArgs::mode)
        .def_readwrite("num_neighbors", &locality::QueryArgs::num_neighbors)
        .def_readwrite("r_max", &locality::QueryArgs::r_max)
        .def_readwrite("r_min", &locality::QueryArgs::r_min)
        .def_readwrite("r_guess", &locality::QueryArgs::r_guess)
        .def_readwrite("scale", &locality::QueryArgs::scale)
        .def_readwrite("exclude_ii", &locality::QueryArgs::exclude_ii);
    mloc.def("get_iterator_terminator", [] {
        auto t = locality::iterator_terminator();
        return py::make_tuple(t.query_point_idx, t.point_idx, t.distance, t.weight);
    });

    py::class_<locality::NeighborList, std::shared_ptr<locality::NeighborList>>(mloc, "NeighborList")
        .def(py::init<>())
        .def(py::init([](py::array_t<unsigned int, py::array::c_style | py::array::forcecast> qidx, unsigned int nq,
                         py::array_t<unsigned int, py::array::c_style | py::array::forcecast> pidx, unsigned int np,
                         points_array vectors, py::object weights) {
                 unsigned int nv = 0;
                 const vec3<float>* v = as_vec3(vectors, nv);
                 if (qidx.size() != pidx.size() || (py::ssize_t) nv != qidx.size())
                 {
                     throw std::invalid_argument("NeighborList arrays must have the same length.");
                 }
                 py::array_t<float, py::array::c_style | py::array::forcecast> w;
                 const float* wp = nullptr;
                 if (!weights.is_none())
                 {
                     w = weights.cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
                     if (w.size() != qidx.size())
                     {
                         throw std::invalid_argument("NeighborList arrays must have the same length.");
                     }
                     wp = w.data();
                 }
                 return std::make_shared<locality::NeighborList>((unsigned int) qidx.size(), qidx.data(), nq, pidx.data(),
                                                                 np, v, wp);
             }),
             py::arg("query_point_indices"), py::arg("num_query_points"), py::arg("point_indices"),
             py::arg("num_points"), py::arg("vectors"), py::arg("weights") = py::none())
        // all pairs (export-NeighborList.cc:41-51)
        .def(py::init([](points_array points, points_array query_points, const box::Box& box, bool exclude_ii) {
                 unsigned int np = 0, nq = 0;
                 const vec3<float>* p = as_vec3(points, np);
                 const vec3<float>* q = as_vec3(query_points, nq);
                 return std::make_shared<locality::NeighborList>(p, q, box, exclude_ii, np, nq);
             }),
             py::arg("points"), py::arg("query_points"), py::arg("box"), py::arg("exclude_ii"))
        .def("getNumBonds", &locality::NeighborList::getNumBonds)
        .def("getNumQueryPoints", &locality::NeighborList::getNumQueryPoints)
        .def("getNumPoints", &locality::NeighborList::getNumPoints)
        .def("getNeighbors", [](const locality::NeighborList& nl) { return to_numpy<unsigned int>(nl.getNeighbors()); })
        .def("getDistances", [](const locality::NeighborList& nl) { return to_numpy<float>(nl.getDistances()); })
        .def("getWeights", [](const locality::NeighborList& nl) { return to_numpy<float>(nl.getWeights()); })
        .def("getVectors", [](const locality::NeighborList& nl) { return to_numpy<float>(nl.getVectors()); })
        .def("getCounts", [](const locality::NeighborList& nl) { return to_numpy<unsigned int>(nl.getCounts()); })
        .def("getSegments", [](const locality::NeighborList& nl) { return to_numpy<unsigned int>(nl.getSegments()); })
        .def("find_first_index", &locality::NeighborList::find_first_index)
        .def("filter",
             [](locality::NeighborList& nl, py::array_t<bool, py::array::c_style | py::array::forcecast> keep) {
                 if (keep.size() != (py::ssize_t) nl.getNumBonds())
                 {
                     throw std::invalid_argument("filter mask must have one entry per bond");
                 }
                 return nl.filter(keep.data());
             })
        .def("filter_r", &locality::NeighborList::filter_r, py::arg("r_max"), py::arg("r_min") = 0.0F)
        .def("sort", &locality::NeighborList::sort)
        .def("copy", &locality::NeighborList::copy)
        .def("validate", &locality::NeighborList::validate);

    py::class_<locality::NeighborQueryIterator, std::shared_ptr<locality::NeighborQueryIterator>>(mloc,
                                                                                                  "NeighborQueryIterator")
        .def("next",
             [](locality::NeighborQueryIterator& it) {
                 auto b = it.next();
                 return py::make_tuple(b.query_point_idx, b.point_idx, b.distance, b.weight);
             })
        .def("toNeighborList", &locality::NeighborQueryIterator::toNeighborList, py::arg("sort_by_distance") = false);

    py::class_<locality::NeighborQuery, std::shared_ptr<locality::NeighborQuery>>(mloc, "NeighborQuery")
        .def("query",
             [](std::shared_ptr<locality::NeighborQuery> nq, points_array qp, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 return nq->query(q, n, qargs);
             },
             py::keep_alive<0, 1>(), py::keep_alive<0, 2>())
        .def("getBox", &locality::NeighborQuery::getBox)
        .def("getNPoints", &locality::NeighborQuery::getNPoints)
        // the bonds of one query point as a list of (i, j, distance, weight) (NeighborQuery::querySingle)
        .def("querySingle",
             [](std::shared_ptr<locality::NeighborQuery> nq, std::array<float, 3> q, unsigned int idx,
                const locality::QueryArgs& qargs) {
                 auto it = nq->querySingle(vec3<float>(q[0], q[1], q[2]), idx, qargs);
                 py::list out;
                 for (auto b = it->next(); !(b == locality::iterator_terminator()); b = it->next())
                 {
                     out.append(py::make_tuple(b.query_point_idx, b.point_idx, b.distance, b.weight));
                 }
                 return out;
             });
    py::class_<locality::LinkCell, locality::NeighborQuery, std::shared_ptr<locality::LinkCell>>(mloc, "LinkCell")
        .def(py::init([](const box::Box& b, points_array pts, float cell_width) {
                 unsigned int n = 0;
                 const vec3<float>* p = as_vec3(pts, n);
                 return std::make_shared<locality::LinkCell>(b, p, n, cell_width);
             }),
             py::arg("box"), py::arg("points"), py::arg("cell_width") = 0.0F, py::keep_alive<1, 3>())
        .def("getCellWidth", &locality::LinkCell::getCellWidth);
    py::class_<locality::AABBQuery, locality::NeighborQuery, std::shared_ptr<locality::AABBQuery>>(mloc, "AABBQuery")
        .def(py::init([](const box::Box& b, points_array pts) {
                 unsigned int n = 0;
                 const vec3<float>* p = as_vec3(pts, n);
                 return std::make_shared<locality::AABBQuery>(b, p, n);
             }),
             py::arg("box"), py::arg("points"), py::keep_alive<1, 3>());
    py::class_<locality::CellQuery, locality::NeighborQuery, std::shared_ptr<locality::CellQuery>>(mloc, "CellQuery")
        .def(py::init([](const box::Box& b, points_array pts) {
                 unsigned int n = 0;
                 const vec3<float>* p = as_vec3(pts, n);
                 return std::make_shared<locality::CellQuery>(b, p, n);
             }),
             py::arg("box"), py::arg("points"), py::keep_alive<1, 3>())
        // grid introspection, as export-NeighborQuery.cc:96-111 binds it
        .def("getCellWidth", &locality::CellQuery::getCellWidth)
        .def("getCountsReal", &locality::CellQuery::getCountsReal)
        .def("getCounts", &locality::CellQuery::getCounts)
        .def("getMinPos", &locality::CellQuery::getMinPos)
        .def("getCellInverseWidth", &locality::CellQuery::getCellInverseWidth)
        .def("getNx", &locality::CellQuery::getNx)
        .def("getNy", &locality::CellQuery::getNy)
        .def("getNz", &locality::CellQuery::getNz)
        .def("getNTotal", &locality::CellQuery::getNTotal)
        .def("getCellStarts", &locality::CellQuery::getCellStarts)
        .def("setupGrid", &locality::CellQuery::setupGrid)
        .def("buildGrid", &locality::CellQuery::buildGrid);
    py::class_<locality::RawPoints, locality::NeighborQuery, std::shared_ptr<locality::RawPoints>>(mloc, "RawPoints")
        .def(py::init([](const box::Box& b, points_array pts) {
                 unsigned int n = 0;
                 const vec3<float>* p = as_vec3(pts, n);
                 return std::make_shared<locality::RawPoints>(b, p, n);
             }),
             py::arg("box"), py::arg("points"), py::keep_alive<1, 3>());

    // ---- _density --------------------------------------------------------------------------------------
    auto mden = m.def_submodule("_density");
    py::enum_<density::NormalizationMode>(mden, "NormalizationMode")
        .value("exact", density::NormalizationMode::exact)
        .value("finite_size", density::NormalizationMode::finite_size);
    py::class_<density::RDF, std::shared_ptr<density::RDF>>(mden, "RDF")
        .def(py::init<unsigned int, float, float>(), py::arg("bins"), py::arg("r_max"), py::arg("r_min") = 0.0F)
        .def("accumulateRDF",
             [](density::RDF& rdf, std::shared_ptr<locality::NeighborQuery> nq, points_array qp,
                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 rdf.accumulate(nq, q, n, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("query_points"), py::arg("nlist").none(true), py::arg("qargs"))
        .def("getRDF", [](density::RDF& r) { return to_numpy<float>(r.getRDF()); })
        .def("getNr", [](density::RDF& r) { return to_numpy<float>(r.getNr()); })
        .def("getBinCounts", [](density::RDF& r) { return to_numpy<unsigned int>(r.getBinCounts()); })
        .def("getBinEdges", &density::RDF::getBinEdges)
        .def("getBinCenters", &density::RDF::getBinCenters)
        .def("getBounds", &density::RDF::getBounds)
        .def("getAxisSizes", &density::RDF::getAxisSizes)
        .def("getBox", &density::RDF::getBox)
        .def("reset", &density::RDF::reset)
        .def_readwrite("mode", &density::RDF::mode);
    // freud/density/export-CorrelationFunction.cc
    py::class_<density::CorrelationFunction, std::shared_ptr<density::CorrelationFunction>>(mden, "CorrelationFunction")
        .def(py::init<unsigned int, float>(), py::arg("bins"), py::arg("r_max"))
        .def("accumulateCF",
             [](density::CorrelationFunction& cf, std::shared_ptr<locality::NeighborQuery> nq,
                py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> values, points_array qp,
                py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> query_values,
                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 if ((size_t) values.size() != nq->getNPoints() || (size_t) query_values.size() != n)
                 {
                     throw std::invalid_argument("values / query_values must hold one number per point / query point");
                 }
                 cf.accumulate(nq, values.data(), q, query_values.data(), n, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("values"), py::arg("query_points"), py::arg("query_values"),
             py::arg("nlist").none(true), py::arg("qargs"))
        .def("getCorrelation",
             [](density::CorrelationFunction& c) { return to_numpy<std::complex<double>>(c.getCorrelation()); })
        .def("getBinCounts", [](density::CorrelationFunction& c) { return to_numpy<unsigned int>(c.getBinCounts()); })
        .def("getBinEdges", &density::CorrelationFunction::getBinEdges)
        .def("getBinCenters", &density::CorrelationFunction::getBinCenters)
        .def("getBounds", &density::CorrelationFunction::getBounds)
        .def("getAxisSizes", &density::CorrelationFunction::getAxisSizes)
        .def("getBox", &density::CorrelationFunction::getBox)
        .def("reset", &density::CorrelationFunction::reset);
    // freud/density/export-LocalDensity.cc:37-47
    py::class_<density::LocalDensity, std::shared_ptr<density::LocalDensity>>(mden, "LocalDensity")
        .def(py::init<float, float>(), py::arg("r_max"), py::arg("diameter"))
        .def("compute",
             [](density::LocalDensity& ld, std::shared_ptr<locality::NeighborQuery> nq, points_array qp,
                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 ld.compute(nq, q, n, nlist, qargs);
             },
             py::arg("points"), py::arg("query_points"), py::arg("nlist").none(true), py::arg("qargs"))
        .def("getRMax", &density::LocalDensity::getRMax)
        .def("getDiameter", &density::LocalDensity::getDiameter)
        .def_property_readonly("box", &density::LocalDensity::getBox)
        .def_property_readonly("density", [](const density::LocalDensity& ld) { return to_numpy<float>(ld.getDensity()); })
        .def_property_readonly("num_neighbors",
                               [](const density::LocalDensity& ld) { return to_numpy<float>(ld.getNumNeighbors()); });

    // ---- _pmft -----------------------------------------------------------------------------------------
    auto mpm = m.def_submodule("_pmft");
    py::class_<pmft::PMFTXY, std::shared_ptr<pmft::PMFTXY>>(mpm, "PMFTXY")
        .def(py::init<float, float, unsigned int, unsigned int>(), py::arg("x_max"), py::arg("y_max"), py::arg("n_x"),
             py::arg("n_y"))
        .def("accumulate",
             [](pmft::PMFTXY& p, std::shared_ptr<locality::NeighborQuery> nq,
                py::array_t<float, py::array::c_style | py::array::forcecast> query_orientations, points_array qp,
                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 if ((size_t) query_orientations.size() != n)
                 {
                     throw std::invalid_argument("query_orientations must hold one angle per query point");
                 }
                 p.accumulate(nq, query_orientations.data(), q, n, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("query_orientations"), py::arg("query_points"),
             py::arg("nlist").none(true), py::arg("qargs"))
        .def("getPCF", [](pmft::PMFTXY& p) { return to_numpy<float>(p.getPCF()); })
        .def("getBinCounts", [](pmft::PMFTXY& p) { return to_numpy<unsigned int>(p.getBinCounts()); })
        .def("getBinEdges", &pmft::PMFTXY::getBinEdges)
        .def("getBinCenters", &pmft::PMFTXY::getBinCenters)
        .def("getBounds", &pmft::PMFTXY::getBounds)
        .def("getAxisSizes", &pmft::PMFTXY::getAxisSizes)
        .def("getBox", &pmft::PMFTXY::getBox)
        .def("reset", &pmft::PMFTXY::reset);

    auto bind_pmft3 = [](auto& cls) {
        using T = typename std::remove_reference_t<decltype(cls)>::type;
        cls.def("getPCF", [](T& p) { return to_numpy<float>(p.getPCF()); })
            .def("getBinCounts", [](T& p) { return to_numpy<unsigned int>(p.getBinCounts()); })
            .def("getBinEdges", &T::getBinEdges)
            .def("getBinCenters", &T::getBinCenters)
            .def("getBounds", &T::getBounds)
            .def("getAxisSizes", &T::getAxisSizes)
            .def("getBox", &T::getBox)
            .def("getHostBinnedBonds", &T::getHostBinnedBonds)
            .def("reset", &T::reset);
    };
    using float_array = py::array_t<float, py::array::c_style | py::array::forcecast>;
    py::class_<pmft::PMFTXYZ, std::shared_ptr<pmft::PMFTXYZ>> xyz(mpm, "PMFTXYZ");
    xyz.def(py::init<float, float, float, unsigned int, unsigned int, unsigned int>(), py::arg("x_max"), py::arg("y_max"),
            py::arg("z_max"), py::arg("n_x"), py::arg("n_y"), py::arg("n_z"))
        .def("accumulate",
             [](pmft::PMFTXYZ& p, std::shared_ptr<locality::NeighborQuery> nq, float_array query_orientations,
                points_array qp, float_array equiv_orientations, std::shared_ptr<locality::NeighborList> nlist,
                const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 if (query_orientations.ndim() != 2 || query_orientations.shape(1) != 4
                     || (size_t) query_orientations.shape(0) != n)
                 {
                     throw std::invalid_argument("query_orientations must hold one quaternion per query point");
                 }
                 if (equiv_orientations.ndim() != 2 || equiv_orientations.shape(1) != 4)
                 {
                     throw std::invalid_argument("equiv_orientations must be an (N, 4) array of quaternions");
                 }
                 p.accumulate(nq, reinterpret_cast<const quat<float>*>(query_orientations.data()), q, n,
                              reinterpret_cast<const quat<float>*>(equiv_orientations.data()),
                              (unsigned int) equiv_orientations.shape(0), nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("query_orientations"), py::arg("query_points"),
             py::arg("equiv_orientations"), py::arg("nlist").none(true), py::arg("qargs"));
    bind_pmft3(xyz);
    auto accumulate_angles = [](auto& p, std::shared_ptr<locality::NeighborQuery> nq, float_array orientations,
                                points_array qp, float_array query_orientations,
                                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
        unsigned int n = 0;
        const vec3<float>* q = as_vec3(qp, n);
        if ((size_t) orientations.size() != nq->getNPoints())
        {
            throw std::invalid_argument("orientations must hold one angle per point");
        }
        if ((size_t) query_orientations.size() != n)
        {
            throw std::invalid_argument("query_orientations must hold one angle per query point");
        }
        p.accumulate(nq, orientations.data(), q, query_orientations.data(), n, nlist, qargs);
    };
    py::class_<pmft::PMFTXYT, std::shared_ptr<pmft::PMFTXYT>> xyt(mpm, "PMFTXYT");
    xyt.def(py::init<float, float, unsigned int, unsigned int, unsigned int>(), py::arg("x_max"), py::arg("y_max"),
            py::arg("n_x"), py::arg("n_y"), py::arg("n_t"))
        .def("accumulate",
             [accumulate_angles](pmft::PMFTXYT& p, std::shared_ptr<locality::NeighborQuery> nq, float_array orientations,
                                 points_array qp, float_array query_orientations,
                                 std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 accumulate_angles(p, nq, orientations, qp, query_orientations, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("orientations"), py::arg("query_points"), py::arg("query_orientations"),
             py::arg("nlist").none(true), py::arg("qargs"));
    bind_pmft3(xyt);
    py::class_<pmft::PMFTR12, std::shared_ptr<pmft::PMFTR12>> r12(mpm, "PMFTR12");
    r12.def(py::init<float, unsigned int, unsigned int, unsigned int>(), py::arg("r_max"), py::arg("n_r"), py::arg("n_t1"),
            py::arg("n_t2"))
        .def("accumulate",
             [accumulate_angles](pmft::PMFTR12& p, std::shared_ptr<locality::NeighborQuery> nq, float_array orientations,
                                 points_array qp, float_array query_orientations,
                                 std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 accumulate_angles(p, nq, orientations, qp, query_orientations, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("orientations"), py::arg("query_points"), py::arg("query_orientations"),
             py::arg("nlist").none(true), py::arg("qargs"));
    bind_pmft3(r12);

    // ---- _environment ----------------------------------------------------------------------------------
    auto menv = m.def_submodule("_environment");
    py::enum_<environment::BondOrderMode>(menv, "BondOrderMode")
        .value("bod", environment::bod)
        .value("lbod", environment::lbod)
        .value("obcd", environment::obcd)
        .value("oocd", environment::oocd)
        .export_values();
    py::class_<environment::BondOrder, std::shared_ptr<environment::BondOrder>>(menv, "BondOrder")
        .def(py::init<unsigned int, unsigned int, environment::BondOrderMode>(), py::arg("n_bins_theta"),
             py::arg("n_bins_phi"), py::arg("mode"))
        .def("accumulate",
             [](environment::BondOrder& b, std::shared_ptr<locality::NeighborQuery> nq,
                py::array_t<float, py::array::c_style | py::array::forcecast> orientations, points_array qp,
                py::array_t<float, py::array::c_style | py::array::forcecast> query_orientations,
                std::shared_ptr<locality::NeighborList> nlist, const locality::QueryArgs& qargs) {
                 unsigned int n = 0;
                 const vec3<float>* q = as_vec3(qp, n);
                 if (orientations.ndim() != 2 || orientations.shape(1) != 4
                     || (size_t) orientations.shape(0) != nq->getNPoints())
                 {
                     throw std::invalid_argument("orientations must hold one quaternion per point");
                 }
                 if (query_orientations.ndim() != 2 || query_orientations.shape(1) != 4
                     || (size_t) query_orientations.shape(0) != n)
                 {
                     throw std::invalid_argument("query_orientations must hold one quaternion per query point");
                 }
                 b.accumulate(nq, reinterpret_cast<const quat<float>*>(orientations.data()), q,
                              reinterpret_cast<const quat<float>*>(query_orientations.data()), n, nlist, qargs);
             },
             py::arg("neighbor_query"), py::arg("orientations"), py::arg("query_points"), py::arg("query_orientations"),
             py::arg("nlist").none(true), py::arg("qargs"))
        .def("getBondOrder", [](environment::BondOrder& b) { return to_numpy<float>(b.getBondOrder()); })
        .def("getBinCounts", [](environment::BondOrder& b) { return to_numpy<unsigned int>(b.getBinCounts()); })
        .def("getBinEdges", &environment::BondOrder::getBinEdges)
        .def("getBinCenters", &environment::BondOrder::getBinCenters)
        .def("getBounds", &environment::BondOrder::getBounds)
        .def("getAxisSizes", &environment::BondOrder::getAxisSizes)
        .def("getBox", &environment::BondOrder::getBox)
        .def("getMode", &environment::BondOrder::getMode)
        .def("getHostBinnedBonds", &environment::BondOrder::getHostBinnedBonds)
        .def("reset", &environment::BondOrder::reset);

    // ---- _order ----------------------------------------------------------------------------------------
    auto mord = m.def_submodule("_order");
    py::class_<order::Steinhardt, std::shared_ptr<order::Steinhardt>>(mord, "Steinhardt")
        .def(py::init<const std::vector<unsigned int>&, bool, bool, bool, bool>(), py::arg("ls"),
             py::arg("average") = false, py::arg("wl") = false, py::arg("weighted") = false,
             py::arg("wl_normalize") = false)
        .def("compute",
             [](order::Steinhardt& s, std::shared_ptr<locality::NeighborList> nlist,
                std::shared_ptr<locality::NeighborQuery> nq,
                const locality::QueryArgs& qargs) { s.compute(nlist, nq, qargs); },
             py::arg("nlist").none(true), py::arg("points"), py::arg("qargs"))
        .def("getQl", [](const order::Steinhardt& s) { return to_numpy<float>(s.getQl()); })
        .def("getParticleOrder", [](const order::Steinhardt& s) { return to_numpy<float>(s.getParticleOrder()); })
        .def("getQlm",
             [](const order::Steinhardt& s) {
                 py::list out;
                 for (const auto& a : s.getQlm())
                 {
                     out.append(to_numpy<std::complex<float>>(a));
                 }
                 return out;
             })
        .def("getOrder", &order::Steinhardt::getOrder)
        .def("getL", &order::Steinhardt::getL)
        .def("getNP", &order::Steinhardt::getNP)
        .def("isAverage", &order::Steinhardt::isAverage)
        .def("isWl", &order::Steinhardt::isWl)
        .def("isWeighted", &order::Steinhardt::isWeighted)
        .def("isWlNormalized", &order::Steinhardt::isWlNormalized);
}
