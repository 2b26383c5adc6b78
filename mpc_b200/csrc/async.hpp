// async.hpp -- asynchronous staging of HOST buffers through one or more devices for the
// host-pointer entry points of the C ABI (gcb_garble / gcb_eval and their _begin forms, the IKNP
// and MiTCCRH host calls).
//
// A call becomes a JOB: one part per selected device (gcb_set_devices), each part cut into slices of
// whole kernel waves.  Everything is queued on the part's three streams -- inputs host->device on
// `h2d`, the kernel of a slice on `k` behind its inputs, results device->host on `d2h` behind the
// kernel -- and the call returns; gcb_job_wait() blocks until the results are in the caller's
// buffers.  One host thread therefore keeps the copy engines of every device busy in both
// directions.  Page-locked user memory (gcb_host_alloc) is DMA'd in place; pageable memory goes
// through a pinned arena with one extra host memcpy each way (inputs inside begin, results inside
// wait).  Streams, events and arenas are pooled per device and reused by later calls.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gcb200.h"

namespace gcb {

int fail(int code, const char* fmt, ...);

struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
    bool pinned_host = false;
    ~Arena() { release(); }
    void release() {
        if (base) { if (pinned_host) cudaFreeHost(base); else cudaFree(base); }
        base = nullptr; cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        bytes = (bytes + (1u << 20)) & ~((size_t)(1u << 20) - 1);
        cudaError_t e = pinned_host ? cudaHostAlloc((void**)&base, bytes, cudaHostAllocPortable)
                                    : cudaMalloc((void**)&base, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    size_t take(size_t bytes) {                      // returns the offset
        const size_t off = used;
        used += (bytes + 255) & ~(size_t)255;
        return off;
    }
};

inline bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// What one part of a job holds while it is in flight.
struct JobRes {
    int device = -1;
    cudaStream_t h2d = nullptr, k = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    Arena dev, pin;

    cudaError_t init(int dev_index) {
        device = dev_index;
        pin.pinned_host = true;
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&k, cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking);
    }
    cudaError_t event(cudaEvent_t* out) {
        if (ev_used == events.size()) {
            cudaEvent_t e;
            cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (rc != cudaSuccess) return rc;
            events.push_back(e);
        }
        *out = events[ev_used++];
        return cudaSuccess;
    }
    // Everything queued has run (or failed): nothing refers to caller memory any more.
    cudaError_t quiesce() {
        cudaError_t e = cudaSuccess, r;
        for (cudaStream_t s : {h2d, k, d2h})
            if (s && (r = cudaStreamSynchronize(s)) != cudaSuccess && e == cudaSuccess) e = r;
        ev_used = 0; dev.used = 0; pin.used = 0;
        return e;
    }
    ~JobRes() {
        if (device >= 0) cudaSetDevice(device);
        for (cudaStream_t s : {h2d, k, d2h}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};

class ResPool {
public:
    std::unique_ptr<JobRes> lease(int device, cudaError_t* err) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (auto& r : retired_) {                    // finished fan-out parts (retire() may not touch CUDA)
                r->ev_used = 0; r->dev.used = 0; r->pin.used = 0;
                pool_[r->device].push_back(std::move(r));
            }
            retired_.clear();
            auto& v = pool_[device];
            if (!v.empty()) { auto r = std::move(v.back()); v.pop_back(); return r; }
        }
        auto r = std::make_unique<JobRes>();
        *err = r->init(device);
        if (*err != cudaSuccess) return nullptr;
        return r;
    }
    void give_back(std::unique_ptr<JobRes> r) {          // callers quiesce() first
        if (!r) return;
        std::lock_guard<std::mutex> lk(mu_);
        auto& v = pool_[r->device];
        if (v.size() < kKeep) v.push_back(std::move(r));  // else: destroyed here (frees its arenas)
    }
    // From a stream callback (no CUDA calls allowed there): the part's work has completed in stream order.
    void retire(std::unique_ptr<JobRes> r) {
        std::lock_guard<std::mutex> lk(mu_);
        retired_.push_back(std::move(r));
    }
private:
    static constexpr size_t kKeep = 64;
    std::vector<std::unique_ptr<JobRes>> retired_;
    std::mutex mu_;
    std::map<int, std::vector<std::unique_ptr<JobRes>>> pool_;
};

// A result that lands in the pinned arena and still has to reach pageable caller memory.
struct LateCopy { void* dst; const uint8_t* src; size_t bytes; cudaEvent_t ready; };

struct JobPart {
    std::unique_ptr<JobRes> res;
    std::vector<LateCopy> late;
};

// Contiguous block [lo, hi) of `total` units for share `i` of `n`: sizes differ by at most one
// (the rule of mpc_b200/shard.py).
inline void share_range(uint64_t total, uint32_t i, uint32_t n, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = total / n, extra = total % n;
    *lo = i * base + (i < extra ? i : extra);
    *hi = *lo + base + (i < extra ? 1 : 0);
}

}  // namespace gcb

struct gcb_job {
    std::vector<gcb::JobPart> parts;
};
