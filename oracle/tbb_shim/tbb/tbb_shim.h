// TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/README.md).
//
// A small std::thread stand-in for the handful of oneTBB constructs the freud hot path uses, so that
// the UNMODIFIED reference sources under /root/reference can be compiled here (oneTBB is not installed
// and there is no network).  It changes scheduling only; every value the reference computes is
// independent of the scheduler except the float32 summation order of Steinhardt's system-wide q_lm.
//
// Constructs provided (usage sites in the reference):
//   blocked_range, blocked_range2d, parallel_for   freud/util/utils.h:54-90
//   enumerable_thread_specific, flatten2d          freud/util/Histogram.h:207-283, freud/util/ThreadStorage.h:23-123,
//                                                  freud/locality/NeighborQuery.h:436-458, freud/locality/NeighborList.cc:137-139
//   parallel_sort                                  freud/locality/NeighborQuery.h:459-466, freud/locality/NeighborList.cc:371-399
//   concurrent_hash_map                            freud/locality/LinkCell.h:253, freud/locality/LinkCell.cc:369-380,472-476
//   global_control                                 freud/parallel/tbb_config.cc:25-35
#ifndef FREUD_ORACLE_TBB_SHIM_H
#define FREUD_ORACLE_TBB_SHIM_H

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <iterator>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace tbb {

namespace shim_detail {

inline std::atomic<size_t>& parallelism_cap()
{
    static std::atomic<size_t> cap {0}; // 0 == hardware concurrency
    return cap;
}

inline size_t effective_threads()
{
    size_t hw = std::thread::hardware_concurrency();
    if (hw == 0)
    {
        hw = 1;
    }
    size_t const cap = parallelism_cap().load();
    return cap == 0 ? hw : std::min(cap, hw);
}

inline bool& inside_parallel_region()
{
    static thread_local bool inside = false;
    return inside;
}

// Persistent worker pool; a job is "call fn(chunk) for chunk in [0, n_chunks)" with dynamic chunk pickup.
class Pool
{
public:
    static Pool& instance()
    {
        static Pool pool;
        return pool;
    }

    void run(size_t n_chunks, size_t n_threads, const std::function<void(size_t)>& fn)
    {
        std::unique_lock<std::mutex> job_lock(m_job_mutex); // one job at a time
        ensureWorkers(n_threads > 0 ? n_threads - 1 : 0);
        {
            std::lock_guard<std::mutex> lk(m_mutex);
            m_fn = &fn;
            m_n_chunks = n_chunks;
            m_next.store(0);
            m_active_limit = n_threads > 0 ? n_threads - 1 : 0;
            m_pending = m_active_limit;
            ++m_generation;
        }
        m_cv.notify_all();
        work(fn, n_chunks); // the calling thread participates
        std::unique_lock<std::mutex> lk(m_mutex);
        m_done_cv.wait(lk, [this] { return m_pending == 0; });
        m_fn = nullptr;
    }

private:
    Pool() = default;
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(m_mutex);
            m_stop = true;
            ++m_generation;
        }
        m_cv.notify_all();
        for (auto& t : m_workers)
        {
            t.join();
        }
    }

    void work(const std::function<void(size_t)>& fn, size_t n_chunks)
    {
        bool const was_inside = inside_parallel_region();
        inside_parallel_region() = true;
        for (size_t c = m_next.fetch_add(1); c < n_chunks; c = m_next.fetch_add(1))
        {
            fn(c);
        }
        inside_parallel_region() = was_inside;
    }

    void ensureWorkers(size_t n)
    {
        while (m_workers.size() < n)
        {
            size_t const id = m_workers.size();
            size_t const start_generation = m_generation;
            m_workers.emplace_back([this, id, start_generation] { workerLoop(id, start_generation); });
        }
    }

    void workerLoop(size_t id, size_t seen)
    {
        while (true)
        {
            const std::function<void(size_t)>* fn = nullptr;
            size_t n_chunks = 0;
            bool participate = false;
            {
                std::unique_lock<std::mutex> lk(m_mutex);
                m_cv.wait(lk, [&] { return m_generation != seen; });
                seen = m_generation;
                if (m_stop)
                {
                    return;
                }
                participate = id < m_active_limit;
                fn = m_fn;
                n_chunks = m_n_chunks;
            }
            if (participate && fn != nullptr)
            {
                work(*fn, n_chunks);
                std::lock_guard<std::mutex> lk(m_mutex);
                if (--m_pending == 0)
                {
                    m_done_cv.notify_all();
                }
            }
        }
    }

    std::mutex m_job_mutex;
    std::mutex m_mutex;
    std::condition_variable m_cv;
    std::condition_variable m_done_cv;
    std::vector<std::thread> m_workers;
    const std::function<void(size_t)>* m_fn {nullptr};
    size_t m_n_chunks {0};
    std::atomic<size_t> m_next {0};
    size_t m_active_limit {0};
    size_t m_pending {0};
    size_t m_generation {0};
    bool m_stop {false};
};

} // namespace shim_detail

template<typename T> class blocked_range
{
public:
    using const_iterator = T;
    blocked_range(T b, T e, size_t grain = 1) : m_begin(b), m_end(e), m_grain(grain) {}
    T begin() const
    {
        return m_begin;
    }
    T end() const
    {
        return m_end;
    }
    size_t size() const
    {
        return size_t(m_end - m_begin);
    }
    size_t grainsize() const
    {
        return m_grain;
    }
    bool empty() const
    {
        return !(m_begin < m_end);
    }

private:
    T m_begin, m_end;
    size_t m_grain;
};

template<typename R, typename C = R> class blocked_range2d
{
public:
    blocked_range2d(R rb, R re, C cb, C ce) : m_rows(rb, re), m_cols(cb, ce) {}
    blocked_range2d(const blocked_range<R>& rows, const blocked_range<C>& cols) : m_rows(rows), m_cols(cols) {}
    const blocked_range<R>& rows() const
    {
        return m_rows;
    }
    const blocked_range<C>& cols() const
    {
        return m_cols;
    }

private:
    blocked_range<R> m_rows;
    blocked_range<C> m_cols;
};

template<typename T, typename Body> void parallel_for(const blocked_range<T>& range, const Body& body)
{
    if (range.empty())
    {
        return;
    }
    size_t const n = range.size();
    size_t const n_threads = shim_detail::effective_threads();
    if (n_threads <= 1 || n == 1 || shim_detail::inside_parallel_region())
    {
        body(range);
        return;
    }
    // ~16 chunks per thread gives dynamic load balance without much pickup overhead.
    size_t const chunk = std::max<size_t>(1, n / (n_threads * 16));
    size_t const n_chunks = (n + chunk - 1) / chunk;
    T const first = range.begin();
    std::function<void(size_t)> const fn = [&](size_t c) {
        T const b = first + T(c * chunk);
        T const e = first + T(std::min(n, (c + 1) * chunk));
        body(blocked_range<T>(b, e));
    };
    shim_detail::Pool::instance().run(n_chunks, n_threads, fn);
}

template<typename R, typename C, typename Body>
void parallel_for(const blocked_range2d<R, C>& range, const Body& body)
{
    // Split rows only; the column range is handed through whole.
    parallel_for(range.rows(), [&](const blocked_range<R>& rows) { body(blocked_range2d<R, C>(rows, range.cols())); });
}

template<typename It, typename Cmp> void parallel_sort(It first, It last, const Cmp& cmp)
{
    size_t const n = size_t(last - first);
    size_t const n_threads = shim_detail::effective_threads();
    if (n < (1U << 14) || n_threads <= 1 || shim_detail::inside_parallel_region())
    {
        std::sort(first, last, cmp);
        return;
    }
    size_t parts = 1;
    while (parts < n_threads)
    {
        parts <<= 1;
    }
    size_t const chunk = (n + parts - 1) / parts;
    auto bound = [&](size_t p) { return first + std::ptrdiff_t(std::min(n, p * chunk)); };
    parallel_for(blocked_range<size_t>(0, parts), [&](const blocked_range<size_t>& r) {
        for (size_t p = r.begin(); p < r.end(); ++p)
        {
            std::sort(bound(p), bound(p + 1), cmp);
        }
    });
    for (size_t width = 1; width < parts; width <<= 1)
    {
        size_t const n_merges = parts / (2 * width);
        parallel_for(blocked_range<size_t>(0, n_merges), [&](const blocked_range<size_t>& r) {
            for (size_t m = r.begin(); m < r.end(); ++m)
            {
                size_t const p = m * 2 * width;
                std::inplace_merge(bound(p), bound(p + width), bound(p + 2 * width), cmp);
            }
        });
    }
}

template<typename It> void parallel_sort(It first, It last)
{
    parallel_sort(first, last, std::less<typename std::iterator_traits<It>::value_type>());
}

//! One lazily constructed T per thread that touches it; element addresses are stable (std::list).
template<typename T> class enumerable_thread_specific
{
public:
    using value_type = T;
    using reference = T&;
    using const_reference = const T&;
    using iterator = typename std::list<T>::iterator;
    using const_iterator = typename std::list<T>::const_iterator;

    enumerable_thread_specific() : m_factory([]() { return T(); }) {}

    template<typename F, typename = decltype(T(std::declval<F&>()())),
             typename = typename std::enable_if<!std::is_convertible<F, T>::value>::type>
    explicit enumerable_thread_specific(F factory) : m_factory(std::move(factory))
    {}

    // NOLINTNEXTLINE(google-explicit-constructor): the reference relies on "ets<unsigned> x = 0;"
    enumerable_thread_specific(const T& exemplar) : m_factory([exemplar]() { return exemplar; }) {}

    enumerable_thread_specific(const enumerable_thread_specific& other)
        : m_factory(other.m_factory), m_items(other.m_items)
    {
        // thread ownership of the copied items is dropped: they stay enumerable but local() makes new ones
    }

    enumerable_thread_specific& operator=(const enumerable_thread_specific& other)
    {
        if (this != &other)
        {
            std::lock_guard<std::mutex> lk(m_mutex);
            m_factory = other.m_factory;
            m_items = other.m_items;
            m_index.clear();
        }
        return *this;
    }

    reference local()
    {
        std::thread::id const me = std::this_thread::get_id();
        std::lock_guard<std::mutex> lk(m_mutex);
        auto it = m_index.find(me);
        if (it != m_index.end())
        {
            return *it->second;
        }
        m_items.push_back(m_factory());
        T* p = &m_items.back();
        m_index.emplace(me, p);
        return *p;
    }

    iterator begin()
    {
        return m_items.begin();
    }
    iterator end()
    {
        return m_items.end();
    }
    const_iterator begin() const
    {
        return m_items.begin();
    }
    const_iterator end() const
    {
        return m_items.end();
    }
    size_t size() const
    {
        return m_items.size();
    }
    bool empty() const
    {
        return m_items.empty();
    }
    void clear()
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        m_items.clear();
        m_index.clear();
    }

private:
    std::function<T()> m_factory;
    std::list<T> m_items;
    std::unordered_map<std::thread::id, T*> m_index;
    mutable std::mutex m_mutex;
};

//! Flat view over an enumerable_thread_specific<Container>.
template<typename Ets> class flattened2d
{
public:
    using inner_container = typename Ets::value_type;
    using value_type = typename inner_container::value_type;

    class const_iterator
    {
    public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = typename flattened2d::value_type;
        using difference_type = std::ptrdiff_t;
        using pointer = const value_type*;
        using reference = const value_type&;

        const_iterator(typename Ets::const_iterator outer, typename Ets::const_iterator outer_end)
            : m_outer(outer), m_outer_end(outer_end)
        {
            if (m_outer != m_outer_end)
            {
                m_inner = m_outer->begin();
                skipEmpty();
            }
        }
        reference operator*() const
        {
            return *m_inner;
        }
        pointer operator->() const
        {
            return &*m_inner;
        }
        const_iterator& operator++()
        {
            ++m_inner;
            skipEmpty();
            return *this;
        }
        const_iterator operator++(int)
        {
            const_iterator tmp(*this);
            ++(*this);
            return tmp;
        }
        bool operator==(const const_iterator& o) const
        {
            if (m_outer != o.m_outer)
            {
                return false;
            }
            return m_outer == m_outer_end || m_inner == o.m_inner;
        }
        bool operator!=(const const_iterator& o) const
        {
            return !(*this == o);
        }

    private:
        void skipEmpty()
        {
            while (m_outer != m_outer_end && m_inner == m_outer->end())
            {
                ++m_outer;
                if (m_outer != m_outer_end)
                {
                    m_inner = m_outer->begin();
                }
            }
        }
        typename Ets::const_iterator m_outer, m_outer_end;
        typename inner_container::const_iterator m_inner;
    };

    explicit flattened2d(const Ets& ets) : m_ets(&ets) {}
    const_iterator begin() const
    {
        return const_iterator(m_ets->begin(), m_ets->end());
    }
    const_iterator end() const
    {
        return const_iterator(m_ets->end(), m_ets->end());
    }
    size_t size() const
    {
        size_t n = 0;
        for (auto it = m_ets->begin(); it != m_ets->end(); ++it)
        {
            n += it->size();
        }
        return n;
    }

private:
    const Ets* m_ets;
};

template<typename Ets> flattened2d<Ets> flatten2d(const Ets& ets)
{
    return flattened2d<Ets>(ets);
}

//! Mutex-guarded unordered_map; references stay valid after the accessor is gone (node-based map).
template<typename K, typename V> class concurrent_hash_map
{
public:
    using value_type = std::pair<const K, V>;

    class const_accessor
    {
    public:
        const value_type& operator*() const
        {
            return *m_ptr;
        }
        const value_type* operator->() const
        {
            return m_ptr;
        }
        bool empty() const
        {
            return m_ptr == nullptr;
        }
        void release()
        {
            m_ptr = nullptr;
        }

    protected:
        friend class concurrent_hash_map;
        value_type* m_ptr {nullptr};
    };

    class accessor : public const_accessor
    {
    public:
        value_type& operator*() const
        {
            return *this->m_ptr;
        }
        value_type* operator->() const
        {
            return this->m_ptr;
        }
    };

    bool find(const_accessor& a, const K& key) const
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        auto it = m_map.find(key);
        if (it == m_map.end())
        {
            a.m_ptr = nullptr;
            return false;
        }
        a.m_ptr = const_cast<value_type*>(&*it);
        return true;
    }

    bool insert(accessor& a, const K& key)
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        auto res = m_map.emplace(key, V());
        a.m_ptr = &*res.first;
        return res.second;
    }

    size_t size() const
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        return m_map.size();
    }

    void clear()
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        m_map.clear();
    }

private:
    mutable std::mutex m_mutex;
    mutable std::unordered_map<K, V> m_map;
};

template<typename T> using concurrent_vector = std::vector<T>;

class global_control
{
public:
    enum parameter
    {
        max_allowed_parallelism,
        thread_stack_size
    };

    global_control(parameter p, size_t value) : m_param(p), m_previous(shim_detail::parallelism_cap().load())
    {
        if (p == max_allowed_parallelism)
        {
            shim_detail::parallelism_cap().store(value);
        }
    }
    ~global_control()
    {
        if (m_param == max_allowed_parallelism)
        {
            shim_detail::parallelism_cap().store(m_previous);
        }
    }
    global_control(const global_control&) = delete;
    global_control& operator=(const global_control&) = delete;

    static size_t active_value(parameter p)
    {
        return p == max_allowed_parallelism ? shim_detail::effective_threads() : 0;
    }

private:
    parameter m_param;
    size_t m_previous;
};

} // namespace tbb

namespace oneapi {
namespace tbb = ::tbb;
}

#endif // FREUD_ORACLE_TBB_SHIM_H
