// TEST INFRASTRUCTURE ONLY -- see nanobind.h in this directory.
#pragma once
#include <cstddef>

namespace nanobind {

template<long... N> struct shape
{};
template<int N> struct ndim
{};
struct c_contig
{};
struct numpy
{};
namespace device {
struct cpu
{};
} // namespace device

template<typename T, typename... Config> struct ndarray
{
    T* data() const { return nullptr; }
    std::size_t shape(std::size_t) const { return 0; }
    std::size_t size() const { return 0; }
    std::size_t ndim() const { return 0; }
};

} // namespace nanobind
