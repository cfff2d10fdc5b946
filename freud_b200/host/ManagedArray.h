// util::ManagedArray<T>: owning, shaped, zero-initialised array shared between a compute object and any
// number of Python views (freud/util/ManagedArray.h:37-333, export-ManagedArray.h:22-52).  Every
// compute()/reset() allocates NEW arrays, so views handed out earlier stay valid and unchanged
// (tests/test_managedarray.py:25-53 upstream).
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

namespace freud { namespace util {

template<typename T> class ManagedArray
{
public:
    ManagedArray() = default;
    explicit ManagedArray(std::vector<size_t> shape) : m_shape(std::move(shape))
    {
        size_t n = 1;
        for (size_t s : m_shape)
        {
            n *= s;
        }
        m_data.assign(m_shape.empty() ? 0 : n, T());
    }
    explicit ManagedArray(size_t n) : ManagedArray(std::vector<size_t> {n}) {}

    T* data() { return m_data.data(); }
    const T* data() const { return m_data.data(); }
    size_t size() const { return m_data.size(); }
    const std::vector<size_t>& shape() const { return m_shape; }

    T& operator[](size_t i)
    {
        if (i >= m_data.size())
        {
            throw std::out_of_range("ManagedArray index out of range"); // -> IndexError upstream
        }
        return m_data[i];
    }
    const T& operator[](size_t i) const
    {
        if (i >= m_data.size())
        {
            throw std::out_of_range("ManagedArray index out of range");
        }
        return m_data[i];
    }

private:
    std::vector<T> m_data;
    std::vector<size_t> m_shape;
};

}} // namespace freud::util
