// stream_kernels.cuh -- device side of the streaming garbler around the gate
// kernel: the permanent wire file (circuit/stream_garble.go:78-157) and the
// serialiser that turns the dense garbled-row slab into the exact record
// stream garbleGate writes into conn.WriteBuf (stream_garble.go:391-446).
#pragma once
#include "gc_kernels.cuh"

namespace gcb {

// Streaming.Set / setWire for a list of ids: labels [batch][n].
__global__ void wf_set_kernel(uint4* const* pages, const uint32_t* ids, uint32_t n, const uint4* l0, uint32_t batch) {
    const size_t total = (size_t)batch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t inst = (uint32_t)(i / n), k = (uint32_t)(i % n);
        *wf_slot(pages, __ldg(ids + k), inst) = __ldg(l0 + i);
    }
}

// GetInput / GetInputs (stream_garble.go:117-128): wires [batch][n] = {L0, L0 ^ R}.
__global__ void wf_get_kernel(uint4* const* pages, const uint32_t* ids, uint32_t n, const uint4* r, uint4* wires,
                              uint32_t batch) {
    const size_t total = (size_t)batch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t inst = (uint32_t)(i / n), k = (uint32_t)(i % n);
        const uint4 l0 = *wf_slot(pages, __ldg(ids + k), inst);
        const uint4 rr = __ldg(r + inst);
        wires[2 * i] = l0;
        wires[2 * i + 1] = make_uint4(l0.x ^ rr.x, l0.y ^ rr.y, l0.z ^ rr.z, l0.w ^ rr.w);
    }
}

// StreamEval.Get for a list of ids: labels [batch][n].
__global__ void wf_get_labels_kernel(uint4* const* pages, const uint32_t* ids, uint32_t n, uint4* labels, uint32_t batch) {
    const size_t total = (size_t)batch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t inst = (uint32_t)(i / n), k = (uint32_t)(i % n);
        labels[i] = *wf_slot(pages, __ldg(ids + k), inst);
    }
}

// R with the S bit forced (stream_garble.go:44-50), once at NewStreaming.
__global__ void force_s_kernel(uint4* r, uint32_t batch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) r[i].y |= 0x80000000u;               // hi word of D0
}

constexpr int SER_TILE = 16384;
constexpr int SER_THREADS = 256;

struct SerParams {
    const uint8_t* tmpl;        // record stream with zeroed rows, padded to a 16-byte multiple
    uint32_t total;             // stream bytes per instance
    const uint32_t* row_pos;    // [n_rows] stream offset of each slab row (ascending)
    uint32_t n_rows;
    const uint4* slab;          // [batch][n_rows]
    uint8_t* dst;               // [batch][dst_stride], dst_stride % 16 == 0
    size_t dst_stride;
};

// grid = (tiles, batch).  One CTA assembles one 16 KiB tile of one instance's
// stream in shared memory -- header bytes from the template, rows scattered in
// as big-endian labels -- and writes it out with coalesced 16-byte stores.
__global__ void __launch_bounds__(SER_THREADS) serialize_kernel(const SerParams p) {
    __shared__ __align__(16) uint8_t tile[SER_TILE];
    const uint32_t t0 = blockIdx.x * SER_TILE;
    const uint32_t inst = blockIdx.y;
    const uint32_t len = p.total - t0 < (uint32_t)SER_TILE ? p.total - t0 : (uint32_t)SER_TILE;
    const uint32_t len16 = (len + 15) >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(p.tmpl + t0);
    for (uint32_t i = threadIdx.x; i < len16; i += SER_THREADS) reinterpret_cast<uint4*>(tile)[i] = __ldg(src + i);
    // rows that intersect [t0, t0 + len): row_pos in (t0 - 16, t0 + len)
    uint32_t lo = 0, hi = p.n_rows;
    {
        const uint32_t want = t0 >= 15 ? t0 - 15 : 0;     // first row_pos >= want
        uint32_t a = 0, b = p.n_rows;
        while (a < b) { const uint32_t m = (a + b) >> 1; if (__ldg(p.row_pos + m) < want) a = m + 1; else b = m; }
        lo = a;
        a = lo; b = p.n_rows;
        while (a < b) { const uint32_t m = (a + b) >> 1; if (__ldg(p.row_pos + m) < t0 + len) a = m + 1; else b = m; }
        hi = a;
    }
    __syncthreads();
    const uint4* rows = p.slab + (size_t)inst * p.n_rows;
    for (uint32_t r = lo + threadIdx.x; r < hi; r += SER_THREADS) {
        const int32_t q0 = (int32_t)(__ldg(p.row_pos + r) - t0);
        const uint4 m = __ldg(rows + r);
        // Label.GetData: BE64(D0) || BE64(D1); memory words are lo(D0) hi(D0) lo(D1) hi(D1)
        const uint32_t w[4] = {m.y, m.x, m.w, m.z};
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int32_t q = q0 + b;
            if (q >= 0 && q < (int32_t)len) tile[q] = (uint8_t)(w[b >> 2] >> (24 - 8 * (b & 3)));
        }
    }
    __syncthreads();
    uint4* out = reinterpret_cast<uint4*>(p.dst + (size_t)inst * p.dst_stride + t0);
    for (uint32_t i = threadIdx.x; i < len16; i += SER_THREADS) out[i] = reinterpret_cast<const uint4*>(tile)[i];
}

// The inverse of serialize_kernel for the streaming evaluator: gathers the garbled rows
// (16 big-endian bytes each, at arbitrary byte offsets of each instance's record stream)
// into the dense slab the eval kernel reads.  src rows are padded by 16 bytes.
struct DeserParams {
    const uint8_t* src;         // [batch][src_stride], src_stride % 16 == 0
    size_t src_stride;
    const uint32_t* row_pos;    // [n_rows]
    uint32_t n_rows;
    uint4* slab;                // [batch][n_rows]
    uint32_t batch;
};
__global__ void deserialize_kernel(const DeserParams p) {
    const size_t total = (size_t)p.batch * p.n_rows;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t inst = (uint32_t)(i / p.n_rows), r = (uint32_t)(i % p.n_rows);
        const uint32_t pos = __ldg(p.row_pos + r), off = pos & 15u;
        const uint4* a = reinterpret_cast<const uint4*>(p.src + (size_t)inst * p.src_stride + (pos - off));
        const uint4 lo = __ldg(a), hi = __ldg(a + 1);
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const uint32_t wo = off >> 2, sh = (off & 3u) * 8u;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t x0 = w[q], x1 = w[q + 1];
            if (wo == 1) { x0 = w[q + 1]; x1 = w[q + 2]; }
            else if (wo == 2) { x0 = w[q + 2]; x1 = w[q + 3]; }
            else if (wo == 3) { x0 = w[q + 3]; x1 = w[q + 4]; }
            v[q] = __byte_perm(__funnelshift_r(x0, x1, sh), 0, 0x0123);   // bytes of the stream, then big-endian -> word
        }
        // Label.SetData: D0 = BE64(bytes 0..7), D1 = BE64(bytes 8..15); memory words lo(D0) hi(D0) lo(D1) hi(D1)
        p.slab[i] = make_uint4(v[1], v[0], v[3], v[2]);
    }
}

}  // namespace gcb
