// async.hpp -- asynchronous staging of HOST buffers through one or more devices for the
// host-pointer entry points of the C ABI (gcb_garble / gcb_eval and their _begin forms, the IKNP
// and MiTCCRH host calls).
//
// A call becomes a JOB: one part per selected device (gcb_set_devices), each part cut into slices of
// whole kernel waves.  Everything is queued on three of the device's streams -- inputs host->device on
// `h2d`, the kernel of a slice on `k` behind its inputs, results device->host on `d2h` behind the
// kernel, ordered by events -- and the call returns; gcb_job_wait() blocks until the results are in the
// caller's buffers.  One host thread therefore keeps the copy engines of every device busy in both
// directions.  Page-locked user memory (gcb_host_alloc) is DMA'd in place; pageable memory goes
// through a pinned arena with one extra host memcpy each way (inputs inside begin, results inside
// wait).  Streams, events and arenas are pooled per device and reused by later calls.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gcb200.h"

namespace gcb {

int fail(int code, const char* fmt, ...);

// Staging copies between pageable caller memory and the pinned arenas.  One thread moves about 10 GB/s, a fifth of what
// PCIe takes, so large copies are cut into 2 MB pieces that a few persistent workers (and the caller) pull from a shared
// counter.  GCB_COPY_THREADS overrides the worker count (default: half the hardware threads, at most 8; 1 = plain memcpy).
class HostCopier {
public:
    static HostCopier& get() { static HostCopier c; return c; }
    void copy(void* dst, const void* src, size_t n) {
        if (n < (8u << 20) || workers_.empty()) { memcpy(dst, src, n); return; }
        Task t;
        t.dst = static_cast<uint8_t*>(dst); t.src = static_cast<const uint8_t*>(src); t.n = n;
        t.pieces = (n + kPiece - 1) / kPiece;
        {
            std::lock_guard<std::mutex> lk(mu_);
            queue_.push_back(&t);
        }
        cv_.notify_all();
        work(t);
        {
            std::unique_lock<std::mutex> lk(mu_);
            for (size_t i = 0; i < queue_.size(); i++) if (queue_[i] == &t) { queue_.erase(queue_.begin() + (long)i); break; }
            done_.wait(lk, [&] { return t.finished.load() == t.pieces && t.inside == 0; });
        }
    }
    ~HostCopier() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (std::thread& w : workers_) w.join();
    }
private:
    static constexpr size_t kPiece = 2u << 20;
    struct Task {
        uint8_t* dst; const uint8_t* src; size_t n, pieces;
        std::atomic<size_t> next{0}, finished{0};
        int inside = 0;                                  // workers currently holding a pointer to the task (under mu_)
    };
    HostCopier() {
        unsigned n = std::thread::hardware_concurrency() / 2;
        if (n > 8) n = 8;
        if (const char* e = getenv("GCB_COPY_THREADS")) n = (unsigned)atoi(e);
        for (unsigned i = 1; i < n; i++) workers_.emplace_back([this] { loop(); });
    }
    static void work(Task& t) {
        for (;;) {
            const size_t i = t.next.fetch_add(1);
            if (i >= t.pieces) return;
            const size_t off = i * kPiece, len = t.n - off < kPiece ? t.n - off : kPiece;
            memcpy(t.dst + off, t.src + off, len);
            t.finished.fetch_add(1);
        }
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (stop_) return;
            Task* t = queue_.front();
            if (t->next.load() >= t->pieces) {           // nothing left to pull: leave it to its owner to dequeue
                queue_.erase(queue_.begin());
                continue;
            }
            t->inside++;
            lk.unlock();
            work(*t);
            lk.lock();
            t->inside--;
            done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<Task*> queue_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};
inline void host_copy(void* dst, const void* src, size_t n) { HostCopier::get().copy(dst, src, n); }

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

// The streams of one device, shared by every job on it.  Few on purpose: the hardware has a limited number
// of work queues per context (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default), and streams that share a queue block
// one another -- with three private streams per job, a result copy waiting for its kernel held up the input
// copies of other jobs queued behind it.  Bulk copies of the two directions have their own streams per job
// class (garble: small inputs up, tables down; eval: tables up, small results down), so neither class waits
// behind the other's bulk transfers; kernels rotate over two streams per class.
struct StreamSet {
    cudaStream_t h2d[2] = {nullptr, nullptr}, d2h[2] = {nullptr, nullptr}, k[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t next_k = 0, next_kc[2] = {0, 0};
    cudaError_t init() {
        cudaError_t e;
        for (cudaStream_t* s : {&h2d[0], &h2d[1], &d2h[0], &d2h[1], &k[0], &k[1], &k[2], &k[3]})
            if ((e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaSuccess;
    }
};

// What one part of a job holds while it is in flight: its slice of the device's streams, events, staging arenas.
struct JobRes {
    int device = -1;
    cudaStream_t h2d = nullptr, k = nullptr, d2h = nullptr;      // assigned from the device's StreamSet at lease time
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    cudaEvent_t tail[3] = {nullptr, nullptr, nullptr};             // behind everything this part queued on h2d / k / d2h
    bool sealed = false;
    Arena dev, pin;

    cudaError_t init(int dev_index) {
        device = dev_index;
        pin.pinned_host = true;
        cudaError_t e;
        for (cudaEvent_t& t : tail)
            if ((e = cudaEventCreateWithFlags(&t, cudaEventDisableTiming)) != cudaSuccess) return e;
        return cudaSuccess;
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
    // Marks the end of what this part queued (the streams are shared: later work of other jobs must not be waited for).
    cudaError_t seal() {
        cudaError_t e = cudaSuccess, r;
        const cudaStream_t st[3] = {h2d, k, d2h};
        for (int i = 0; i < 3; i++)
            if (st[i] && (r = cudaEventRecord(tail[i], st[i])) != cudaSuccess && e == cudaSuccess) e = r;
        sealed = true;
        return e;
    }
    // Everything this part queued has run (or failed): nothing refers to caller memory any more.
    cudaError_t quiesce() {
        cudaError_t e = cudaSuccess, r;
        if (!sealed) e = seal();
        for (cudaEvent_t t : tail)
            if (t && (r = cudaEventSynchronize(t)) != cudaSuccess && e == cudaSuccess) e = r;
        ev_used = 0; dev.used = 0; pin.used = 0; sealed = false;
        return e;
    }
    ~JobRes() {
        if (device >= 0) cudaSetDevice(device);
        if (sealed) for (cudaEvent_t t : tail) if (t) cudaEventSynchronize(t);
        for (cudaEvent_t t : tail) if (t) cudaEventDestroy(t);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};

class ResPool {
public:
    // cls: 0 = garble-like (bulk results), 1 = eval-like (bulk inputs).  want: device staging bytes the part needs;
    // the pooled entry that fits best is taken, and a new arena is sized for the largest request seen so far, so that
    // the pool converges to arenas any part can use (growing one costs a cudaFree, i.e. a device-wide synchronisation).
    std::unique_ptr<JobRes> lease(int device, int cls, size_t want, cudaError_t* err, size_t want_pin = 0) {
        std::unique_ptr<JobRes> r;
        StreamSet* ss = nullptr;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (auto& q : retired_) {                    // finished fan-out parts (retire() may not touch CUDA)
                q->ev_used = 0; q->dev.used = 0; q->pin.used = 0; q->sealed = false;
                pool_[q->device].push_back(std::move(q));
            }
            retired_.clear();
            auto& v = pool_[device];
            size_t best = v.size();
            for (size_t i = 0; i < v.size(); i++) {
                const bool fits = v[i]->dev.cap >= want && v[i]->pin.cap >= want_pin;
                if (best == v.size()) { best = i; continue; }
                const bool best_fits = v[best]->dev.cap >= want && v[best]->pin.cap >= want_pin;
                if ((fits && !best_fits) || (fits == best_fits && (fits ? v[i]->dev.cap < v[best]->dev.cap : v[i]->dev.cap > v[best]->dev.cap))) best = i;
            }
            if (best < v.size()) { r = std::move(v[best]); v.erase(v.begin() + (long)best); }
            size_t& hw = high_water_[device];
            if (want > hw) hw = want;
            want_alloc_ = hw;
            auto& slot = streams_[device];
            if (!slot) {
                slot = std::make_unique<StreamSet>();
                if ((*err = slot->init()) != cudaSuccess) { slot.reset(); return nullptr; }
            }
            ss = slot.get();
            if (!r) {
                r = std::make_unique<JobRes>();
                if ((*err = r->init(device)) != cudaSuccess) return nullptr;
            }
            // GCB_COPY_STREAMS=1: one stream per direction for both job classes (experiments)
            static const bool one = [] { const char* e = getenv("GCB_COPY_STREAMS"); return e && atoi(e) == 1; }();
            r->h2d = ss->h2d[one ? 0 : (cls & 1)];
            r->d2h = ss->d2h[one ? 0 : (cls & 1)];
            // kernels: two streams per job class.  With one rotation for both classes a garbling kernel could sit behind an
            // evaluation kernel that is still waiting for its tables to cross PCIe, and the device->host engine ran dry at
            // every step boundary of a pipelined caller (bench.py e2e).  GCB_KERNEL_STREAMS=shared restores the rotation.
            static const bool shared_k = [] { const char* e = getenv("GCB_KERNEL_STREAMS"); return e && !strcmp(e, "shared"); }();
            r->k = shared_k ? ss->k[ss->next_k++ & 3] : ss->k[(cls & 1) * 2 + (ss->next_kc[cls & 1]++ & 1)];
        }
        return r;
    }
    size_t alloc_hint(int device) {
        std::lock_guard<std::mutex> lk(mu_);
        return high_water_[device];
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
    std::map<int, std::unique_ptr<StreamSet>> streams_;
    std::map<int, size_t> high_water_;
    size_t want_alloc_ = 0;
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
