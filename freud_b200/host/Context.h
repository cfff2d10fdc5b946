// Process-wide GPU context for the C++ classes: one fgpu_ctx per device, created on first use.  The device
// is LOCAL_RANK (one process per GPU under torchrun) unless FREUD_B200_DEVICE overrides it.  There is no CPU
// fallback: without a CUDA device the first query throws std::runtime_error.
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/freud_b200.h"

namespace freud { namespace gpu {

// Maps a C-ABI status to the exception type the reference throws (and nanobind/pybind11 translate):
// EINVALID/EDOMAIN -> ValueError, ERUNTIME/ECUDA/ENCCL -> RuntimeError, ENOMEM -> MemoryError,
// ERANGE -> IndexError.
inline void check(int rc)
{
    if (rc == FGPU_OK)
    {
        return;
    }
    std::string const msg = fgpu_last_error();
    switch (rc)
    {
    case FGPU_EINVALID:
        throw std::invalid_argument(msg);
    case FGPU_EDOMAIN:
        throw std::domain_error(msg);
    case FGPU_ENOMEM:
        throw std::bad_alloc();
    case FGPU_ERANGE:
        throw std::out_of_range(msg);
    default:
        throw std::runtime_error(msg);
    }
}

inline int default_device()
{
    const char* e = std::getenv("FREUD_B200_DEVICE");
    if (e == nullptr)
    {
        e = std::getenv("LOCAL_RANK");
    }
    int const n = fgpu_device_count();
    int const want = e != nullptr ? std::atoi(e) : 0;
    return n > 0 ? want % n : 0;
}

inline fgpu_ctx* context(int device = -1)
{
    static std::mutex mtx;
    static std::map<int, fgpu_ctx*> ctxs;
    std::lock_guard<std::mutex> lock(mtx);
    if (device < 0)
    {
        device = default_device();
    }
    auto it = ctxs.find(device);
    if (it == ctxs.end())
    {
        fgpu_ctx* c = nullptr;
        check(fgpu_ctx_create(device, &c));
        it = ctxs.emplace(device, c).first;
    }
    return it->second;
}

}} // namespace freud::gpu
