// TEST INFRASTRUCTURE ONLY -- a compile-time stand-in for nanobind (not installed in this image).
//
// tests/test_binding_compile.py compiles the reference's UNMODIFIED binding translation units
// (freud/locality/export-NeighborQuery.cc, export-NeighborList.cc, export-BondHistogramCompute.cc,
// freud/density/export-RDF.cc, freud/order/export-Steinhardt.cc) against the C++ classes of freud_b200/host with this
// header in place of nanobind's: every `.def(...)` swallows its arguments, so what the compiler really checks is that
// each member pointer the bindings take (&NeighborQuery::query, &RDF::getRDF, &CellQuery::getCountsReal, ...) exists
// and that each wrapper function body (constructor calls, accumulate(...) argument lists) type-checks against the
// replacement headers -- the claim INTEGRATION.md makes ("the binding files compile unchanged").
#pragma once
#include <cstddef>
#include <utility>

namespace nanobind {

struct handle
{};
struct object : handle
{};
struct list : object
{
    template<typename T> void append(T&&) {}
};
struct tuple : object
{};
template<typename... A> tuple make_tuple(A&&...)
{
    return {};
}

enum class rv_policy
{
    automatic,
    automatic_reference,
    take_ownership,
    copy,
    move,
    reference,
    reference_internal,
    none
};

struct arg
{
    explicit arg(const char*) {}
    arg& none(bool = true) { return *this; }
    template<typename V> arg& operator=(V&&) { return *this; }
};

template<typename... A> struct init
{};

struct self_t
{};
struct self_expr
{};
inline self_expr operator==(const self_t&, const self_t&)
{
    return {};
}
inline self_expr operator!=(const self_t&, const self_t&)
{
    return {};
}
static const self_t self {};

struct module_
{
    template<typename... A> module_& def(A&&...) { return *this; }
    template<typename... A> module_ def_submodule(A&&...) { return *this; }
};

template<typename T, typename... Bases> struct class_
{
    template<typename... A> explicit class_(A&&...) {}
    template<typename... A> class_& def(A&&...) { return *this; }
    template<typename... A> class_& def_rw(A&&...) { return *this; }
    template<typename... A> class_& def_ro(A&&...) { return *this; }
    template<typename... A> class_& def_prop_ro(A&&...) { return *this; }
    template<typename... A> class_& def_prop_rw(A&&...) { return *this; }
    template<typename... A> class_& def_static(A&&...) { return *this; }
};

template<typename E> struct enum_
{
    template<typename... A> explicit enum_(A&&...) {}
    enum_& value(const char*, E) { return *this; }
    enum_& export_values() { return *this; }
};

} // namespace nanobind
