// hostpipe.hpp -- staging of HOST buffers through the device for the host-pointer
// entry points of the C ABI.
//
// A call is cut into slices; slice k uses slot k % kSlots.  Each slot owns a CUDA
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
        cudaError_t e = pinned_host ? cudaHostAlloc((void**)&base, bytes, cudaHostAllocMapped | cudaHostAllocPortable)
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
    static constexpr int kSlots = 4;
    static constexpr size_t kSliceBytes = 48u << 20;   // ~1 ms of PCIe per slice: copies of one slice hide the kernel of the next

    // Regions up to this size are not copied at all: the kernel reads / writes the page-locked
    // host memory directly (UVA).  Besides saving a copy this keeps small transfers out of the
    // copy-engine queues, where they would wait behind other callers' bulk table copies.
    static constexpr size_t kZeroCopyBytes = 4u << 20;

    struct Region { const void* src; void* dst; size_t bytes, dev_off, pin_off; bool pinned; bool zero_copy; };

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
        struct Fixup { void** where; bool from_pin; bool direct; void* direct_ptr; };
        std::vector<Fixup> fixups;
        static void* mapped_ptr(const void* host) {       // device view of page-locked host memory, or null
            void* d = nullptr;
            if (cudaHostGetDevicePointer(&d, const_cast<void*>(host), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            return d;
        }
        int in(const void* src, size_t bytes, const void** dptr) {
            Region r{src, nullptr, bytes, 0, 0, is_pinned(src), bytes <= kZeroCopyBytes};
            void* direct = (r.pinned && r.zero_copy) ? mapped_ptr(src) : nullptr;
            if (r.pinned && r.zero_copy && !direct) r.zero_copy = false;
            if (!r.pinned) r.pin_off = pin.take(bytes);
            if (!r.zero_copy) r.dev_off = dev.take(bytes);
            ins.push_back(r);
            fixups.push_back(Fixup{const_cast<void**>(reinterpret_cast<const void**>(dptr)), r.zero_copy && !r.pinned, direct != nullptr, direct});
            *dptr = reinterpret_cast<const void*>(r.zero_copy ? r.pin_off : r.dev_off);
            return GCB_OK;
        }
        int out(void* dst, size_t bytes, void** dptr) {
            Region r{nullptr, dst, bytes, 0, 0, is_pinned(dst), bytes <= kZeroCopyBytes};
            void* direct = (r.pinned && r.zero_copy) ? mapped_ptr(dst) : nullptr;
            if (r.pinned && r.zero_copy && !direct) r.zero_copy = false;
            if (!r.pinned) r.pin_off = pin.take(bytes);
            if (!r.zero_copy) r.dev_off = dev.take(bytes);
            outs.push_back(r);
            fixups.push_back(Fixup{dptr, r.zero_copy && !r.pinned, direct != nullptr, direct});
            *dptr = reinterpret_cast<void*>(r.zero_copy ? r.pin_off : r.dev_off);
            return GCB_OK;
        }
        int upload() {
            int rc = check(dev.reserve(dev.used), "device staging allocation");
            if (rc) return rc;
            if (pin.used && (rc = check(pin.reserve(pin.used), "pinned staging allocation"))) return rc;
            uint8_t* pin_dev = pin.base ? static_cast<uint8_t*>(mapped_ptr(pin.base)) : nullptr;
            for (const Fixup& f : fixups) {
                if (f.direct) *f.where = f.direct_ptr;
                else if (f.from_pin) *f.where = pin_dev + reinterpret_cast<size_t>(*f.where);
                else *f.where = dev.base + reinterpret_cast<size_t>(*f.where);
            }
            fixups.clear();
            for (const Region& r : ins) {
                const void* h = r.src;
                if (!r.pinned) { memcpy(pin.base + r.pin_off, r.src, r.bytes); h = pin.base + r.pin_off; }
                if (r.zero_copy) continue;                 // the kernel reads it in place
                rc = check(cudaMemcpyAsync(dev.base + r.dev_off, h, r.bytes, cudaMemcpyHostToDevice, compute), "H2D copy");
                if (rc) return rc;
            }
            busy = true;
            return GCB_OK;
        }
        int download() {
            for (const Region& r : outs) {
                if (r.zero_copy) continue;                 // the kernel wrote it in place
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
    // Instances per slice: about kSliceBytes, a multiple of `granule` (instances one SM holds).
    uint32_t slice_for(size_t per_inst, uint32_t batch, uint32_t granule) const {
        if (granule == 0) granule = 1;
        size_t n = kSliceBytes / (per_inst ? per_inst : 1);
        n = n / granule * granule;
        if (n < granule) n = granule;
        return n >= batch ? batch : (uint32_t)n;
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
