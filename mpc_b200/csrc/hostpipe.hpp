// hostpipe.hpp -- staging of HOST buffers through the device for the host-pointer
// entry points of the C ABI.
//
// A call is cut into slices; slice k uses slot k % 3.  Each slot owns a CUDA
// stream and a device arena, so the host->device copy of one slice, the kernel
// of the previous one and the device->host copy of the one before overlap.
// User memory that is already page-locked (allocated with gcb_host_alloc, or
// registered by the caller) is DMA'd in place; pageable memory goes through the
// slot's own pinned arena with one extra host memcpy each way.
#pragma once
#include <cstdint>
#include <cstring>
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
        cudaError_t e = pinned_host ? cudaHostAlloc((void**)&base, bytes, cudaHostAllocDefault)
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

struct HostPipe {
    static constexpr int kSlots = 3;
    static constexpr size_t kSliceBytes = 160u << 20;

    struct Region { const void* src; void* dst; size_t bytes, dev_off, pin_off; bool pinned; };

    struct Slot {
        cudaStream_t compute = nullptr;
        Arena dev, pin;
        std::vector<Region> ins, outs;
        bool busy = false;

        int check(cudaError_t e, const char* what) {
            if (e == cudaSuccess) return GCB_OK;
            return fail(GCB_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
        }
        // wait for the slot's previous slice and hand its results to the caller
        int drain() {
            if (!busy) return GCB_OK;
            int rc = check(cudaStreamSynchronize(compute), "slice synchronize");
            if (rc) return rc;
            for (const Region& r : outs)
                if (!r.pinned) memcpy(r.dst, pin.base + r.pin_off, r.bytes);
            busy = false;
            return GCB_OK;
        }
        int begin() {
            int rc = drain();
            if (rc) return rc;
            ins.clear(); outs.clear();
            dev.used = 0; pin.used = 0;
            return GCB_OK;
        }
        // Declare an input / output region; *dptr receives its device address
        // (valid once upload() has run, which may re-allocate the arenas).
        std::vector<void**> fixups;
        int in(const void* src, size_t bytes, const void** dptr) {
            Region r{src, nullptr, bytes, dev.take(bytes), 0, is_pinned(src)};
            if (!r.pinned) r.pin_off = pin.take(bytes);
            ins.push_back(r);
            fixups.push_back(const_cast<void**>(reinterpret_cast<const void**>(dptr)));
            *dptr = reinterpret_cast<const void*>(r.dev_off);
            return GCB_OK;
        }
        int out(void* dst, size_t bytes, void** dptr) {
            Region r{nullptr, dst, bytes, dev.take(bytes), 0, is_pinned(dst)};
            if (!r.pinned) r.pin_off = pin.take(bytes);
            outs.push_back(r);
            fixups.push_back(dptr);
            *dptr = reinterpret_cast<void*>(r.dev_off);
            return GCB_OK;
        }
        int upload() {
            int rc = check(dev.reserve(dev.used), "device staging allocation");
            if (rc) return rc;
            if (pin.used && (rc = check(pin.reserve(pin.used), "pinned staging allocation"))) return rc;
            for (void** f : fixups) *f = dev.base + reinterpret_cast<size_t>(*f);
            fixups.clear();
            for (const Region& r : ins) {
                const void* h = r.src;
                if (!r.pinned) { memcpy(pin.base + r.pin_off, r.src, r.bytes); h = pin.base + r.pin_off; }
                rc = check(cudaMemcpyAsync(dev.base + r.dev_off, h, r.bytes, cudaMemcpyHostToDevice, compute), "H2D copy");
                if (rc) return rc;
            }
            busy = true;
            return GCB_OK;
        }
        int download() {
            for (const Region& r : outs) {
                void* h = r.pinned ? r.dst : (void*)(pin.base + r.pin_off);
                int rc = check(cudaMemcpyAsync(h, dev.base + r.dev_off, r.bytes, cudaMemcpyDeviceToHost, compute), "D2H copy");
                if (rc) return rc;
            }
            return GCB_OK;
        }
    };

    Slot slots[kSlots];
    bool ready = false;

    int init() {
        if (ready) return GCB_OK;
        for (Slot& s : slots) {
            s.pin.pinned_host = true;
            cudaError_t e = cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking);
            if (e != cudaSuccess) return fail(GCB_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ready = true;
        return GCB_OK;
    }
    Slot& slot(uint32_t k) { return slots[k % kSlots]; }
    // Instances per slice: whole waves of `resident` instances, about kSliceBytes.
    uint32_t slice_for(size_t per_inst, uint32_t batch, uint32_t resident) const {
        if (resident == 0) resident = 1;
        const size_t wave_bytes = per_inst * resident;
        size_t waves = kSliceBytes / (wave_bytes ? wave_bytes : 1);
        if (waves == 0) waves = 1;
        const size_t s = waves * resident;
        return s >= batch ? batch : (uint32_t)s;
    }
    int finish() {
        int rc = GCB_OK;
        for (Slot& s : slots) {
            const int r2 = s.drain();
            if (r2 && !rc) rc = r2;
        }
        return rc;
    }
    ~HostPipe() {
        for (Slot& s : slots)
            if (s.compute) { cudaStreamSynchronize(s.compute); cudaStreamDestroy(s.compute); }
    }
};

}  // namespace gcb
