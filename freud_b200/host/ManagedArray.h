// util::ManagedArray<T>: owning, shaped array shared between a compute object and any number of Python views
// (freud/util/ManagedArray.h:37-333, export-ManagedArray.h:22-52).  Every compute()/reset() allocates NEW arrays, so
// views handed out earlier stay valid and unchanged (tests/test_managedarray.py:25-53 upstream).
//
// Storage comes from the C ABI's cache of page-locked blocks (fgpu_host_alloc): the arrays on this path are the
// landing zones of device -> host copies (225 MB of NeighborList arrays per 1 M-point frame), which run at the
// link's rate only into page-locked memory, and page-locking per frame would cost what it saves.  Arrays are
// zero-initialised as upstream unless the caller is about to overwrite every element (Uninitialized), and several
// arrays may share one block (the per-l q_lm arrays of one Steinhardt::compute are slices of a single copy).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../../include/freud_b200.h"

namespace freud { namespace util {

struct Uninitialized
{};

// one block of the host cache; returns to the cache when the last array over it dies
class HostBlock
{
public:
    explicit HostBlock(size_t bytes)
    {
        if (bytes != 0 && fgpu_host_alloc(bytes, &m_ptr) != FGPU_OK)
        {
            throw std::bad_alloc();
        }
    }
    HostBlock(const HostBlock&) = delete;
    HostBlock& operator=(const HostBlock&) = delete;
    ~HostBlock() { fgpu_host_free(m_ptr); }
    void* get() const { return m_ptr; }

private:
    void* m_ptr = nullptr;
};

template<typename T> class ManagedArray
{
public:
    ManagedArray() = default;
    explicit ManagedArray(std::vector<size_t> shape) : ManagedArray(std::move(shape), Uninitialized {})
    {
        if (m_size != 0)
        {
            std::memset(static_cast<void*>(m_data), 0, m_size * sizeof(T)); // T() of every element type used here
        }
    }
    explicit ManagedArray(size_t n) : ManagedArray(std::vector<size_t> {n}) {}
    // every element is about to be overwritten (a device -> host copy lands here)
    ManagedArray(std::vector<size_t> shape, Uninitialized) : m_shape(std::move(shape))
    {
        m_size = m_shape.empty() ? 0 : 1;
        for (size_t s : m_shape)
        {
            m_size *= s;
        }
        m_block = std::make_shared<HostBlock>(m_size * sizeof(T));
        m_data = static_cast<T*>(m_block->get());
    }
    // a slice of a block another array also lives in
    ManagedArray(std::shared_ptr<HostBlock> block, T* data, std::vector<size_t> shape)
        : m_block(std::move(block)), m_data(data), m_shape(std::move(shape))
    {
        m_size = m_shape.empty() ? 0 : 1;
        for (size_t s : m_shape)
        {
            m_size *= s;
        }
    }
    // copies are deep, as upstream's (NeighborList::copy relies on it)
    ManagedArray(const ManagedArray& other) : ManagedArray(other.m_shape, Uninitialized {})
    {
        if (m_size != 0)
        {
            std::memcpy(static_cast<void*>(m_data), static_cast<const void*>(other.m_data), m_size * sizeof(T));
        }
    }
    ManagedArray& operator=(const ManagedArray& other)
    {
        if (this != &other)
        {
            ManagedArray tmp(other);
            std::swap(m_block, tmp.m_block);
            std::swap(m_data, tmp.m_data);
            std::swap(m_size, tmp.m_size);
            std::swap(m_shape, tmp.m_shape);
        }
        return *this;
    }

    T* data() { return m_data; }
    const T* data() const { return m_data; }
    size_t size() const { return m_size; }
    const std::vector<size_t>& shape() const { return m_shape; }
    const std::shared_ptr<HostBlock>& block() const { return m_block; }

    T& operator[](size_t i)
    {
        if (i >= m_size)
        {
            throw std::out_of_range("ManagedArray index out of range"); // -> IndexError upstream
        }
        return m_data[i];
    }
    const T& operator[](size_t i) const
    {
        if (i >= m_size)
        {
            throw std::out_of_range("ManagedArray index out of range");
        }
        return m_data[i];
    }

private:
    std::shared_ptr<HostBlock> m_block;
    T* m_data = nullptr;
    size_t m_size = 0;
    std::vector<size_t> m_shape;
};

}} // namespace freud::util
