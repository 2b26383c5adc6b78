// gc_eval.cu -- instantiations and launch dispatch of eval_kernel (gc_kernels.cuh).
#include "gc_launch.hpp"

namespace gcb {

cudaError_t gc_opt_in_eval(int smem_bytes) {
    cudaError_t e = cudaSuccess;
#define GC_OPT(NR, MODE, ILP, MAXT, NT, SPILL) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(eval_kernel<NR, MODE, ILP, MAXT, NT, SPILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    GC_FOR_ALL(GC_OPT)
#undef GC_OPT
    return e;
}

template <int NR, int MODE>
static void launch(GcVariant v, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const GcParams& p) {
    if (v.spill == 2) eval_kernel<NR, MODE, 1, 512, 2, 2><<<grid, block, smem, s>>>(p);
    else if (v.spill && v.ilp == 1) eval_kernel<NR, MODE, 1, 512, 2, 1><<<grid, block, smem, s>>>(p);
    else if (v.spill) eval_kernel<NR, MODE, 2, 512, 2, 1><<<grid, block, smem, s>>>(p);
    else if (v.nt == 2 && v.ilp == 1) eval_kernel<NR, MODE, 1, 512, 2, 0><<<grid, block, smem, s>>>(p);
    else if (v.nt == 2) eval_kernel<NR, MODE, 2, 512, 2, 0><<<grid, block, smem, s>>>(p);
    else if (v.ilp == 1) eval_kernel<NR, MODE, 1, 1024, 4, 0><<<grid, block, smem, s>>>(p);
    else eval_kernel<NR, MODE, 2, 512, 4, 0><<<grid, block, smem, s>>>(p);
}
template <int MODE>
static void launch_nr(uint32_t keylen, GcVariant v, dim3 g, dim3 b, size_t sm, cudaStream_t s, const GcParams& p) {
    if (keylen == 16) launch<10, MODE>(v, g, b, sm, s, p);
    else if (keylen == 24) launch<12, MODE>(v, g, b, sm, s, p);
    else launch<14, MODE>(v, g, b, sm, s, p);
}
void gc_launch_eval(int mode, uint32_t keylen, GcVariant v, dim3 g, dim3 b, size_t sm, cudaStream_t s, const GcParams& p) {
    if (mode == GC_STREAM) launch_nr<GC_STREAM>(keylen, v, g, b, sm, s, p);
    else if (mode == GC_FULL) launch_nr<GC_FULL>(keylen, v, g, b, sm, s, p);
    else launch_nr<GC_PLAIN>(keylen, v, g, b, sm, s, p);
}

}  // namespace gcb
