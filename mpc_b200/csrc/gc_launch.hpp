// gc_launch.hpp -- entry points of the two translation units that instantiate the gate
// kernels (gc_garble.cu, gc_eval.cu), so that they compile in parallel with the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "gc_kernels.cuh"

namespace gcb {

// Kernel variant: (AES blocks a thread interleaves, CTA size bound, resident T-tables).
//   ilp 2, nt 4: the default (wide levels: bound by the shared-memory pipe)
//   ilp 1, nt 4, up to 1024 threads: many one-warp teams (small circuits)
//   ilp 1, nt 2: deep, narrow circuits (twice the label space, four node rows in flight)
//   ilp 2, nt 2: wide circuits whose live labels do not fit beside four tables
//   ilp 2, nt 2, spill: circuits whose live labels do not fit on chip at all (the excess in global memory)
//   ilp 1, nt 2, spill 1: hot / cold plans with in-place scratch access (the first form: GCB_HOT_MODE=1)
//   ilp 1, nt 2, spill 2: hot / cold plans by live-range splitting (plan.cpp): per-phase evict / asynchronous reload lists,
//                         every gate and node access stays in shared memory
struct GcVariant { uint32_t ilp, nt; int spill; };     // spill: 0 none, 1 in-place scratch access, 2 live-range copies

cudaError_t gc_opt_in_garble(int smem_bytes);
cudaError_t gc_opt_in_eval(int smem_bytes);
// mode: GC_PLAIN / GC_FULL / GC_STREAM; keylen 16 / 24 / 32
void gc_launch_garble(int mode, uint32_t keylen, GcVariant v, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const GcParams& p);
void gc_launch_eval(int mode, uint32_t keylen, GcVariant v, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const GcParams& p);

// every (rounds, mode, variant) instantiation
#define GC_FOR_VARIANT(M, NR, MODE) \
    M(NR, MODE, 2, 512, 4, 0) M(NR, MODE, 1, 1024, 4, 0) M(NR, MODE, 1, 512, 2, 0) M(NR, MODE, 2, 512, 2, 0) \
    M(NR, MODE, 2, 512, 2, 1) M(NR, MODE, 1, 512, 2, 1) M(NR, MODE, 1, 512, 2, 2)
#define GC_FOR_NR(M, MODE) GC_FOR_VARIANT(M, 10, MODE) GC_FOR_VARIANT(M, 12, MODE) GC_FOR_VARIANT(M, 14, MODE)
#define GC_FOR_ALL(M) GC_FOR_NR(M, GC_PLAIN) GC_FOR_NR(M, GC_FULL) GC_FOR_NR(M, GC_STREAM)

}  // namespace gcb
