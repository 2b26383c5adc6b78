// microbench_lds.cu -- what one B200 SM sustains for the T-table AES pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_lds tools/microbench_lds.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../mpc_b200/csrc/aes_core.cuh"
using namespace gcb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// A: dependent-free LDS.32 at random conflict-free addresses (the T-table pattern), U independent chains
template <int W>   // W = bytes per load: 4, 8, 16
__global__ void k_lds(uint32_t* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) uint8_t smem[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) ((uint32_t*)smem)[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t x = threadIdx.x * 77 + blockIdx.x, acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            x = x * 1664525u + 1013904223u;
            if (W == 4) {
                uint32_t a = tb + ((x >> 16) & 0xff) * 256 + ((j & 3) * 128 & 0x80) + lane * 4, v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
                acc ^= v;
            } else if (W == 8) {
                uint32_t a = tb + ((x >> 16) & 0x7f) * 256 + lane * 8; uint32_t v0, v1;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a));
                acc ^= v0 ^ v1;
            } else {
                uint32_t a = tb + ((x >> 16) & 0x3f) * 512 + lane * 16; uint32_t v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a));
                acc ^= v0 ^ v1 ^ v2 ^ v3;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// C: real AES-128 rounds, U interleaved blocks per thread, round keys in smem
template <int U>
__global__ void k_aes(uint4* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    uint32_t* rk = (uint32_t*)(smem + AES_TABLE_BYTES);
    if (threadIdx.x < 44) rk[threadIdx.x] = threadIdx.x * 0x9e3779b9u;
    __syncthreads();
    const AesLane a = aes_lane(smem);
    uint32_t s[U][4];
#pragma unroll
    for (int j = 0; j < U; j++) { s[j][0] = threadIdx.x + j; s[j][1] = blockIdx.x; s[j][2] = j * 17; s[j][3] = 99; }
    const uint4* k4 = (const uint4*)rk;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int r = 1; r < 10; r++) {
            const uint4 k = k4[r];
#pragma unroll
            for (int j = 0; j < U; j++) aes_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
        }
#pragma unroll
        for (int j = 0; j < U; j++) aes_last_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k4[10]);
    }
    long long t1 = clock64();
    uint4 o = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < U; j++) { o.x ^= s[j][0]; o.y ^= s[j][1]; o.z ^= s[j][2]; o.w ^= s[j][3]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// C2: AES-128 rounds with only T0/T1 resident (64 KiB); T2 = rot16(T0), T3 = rot16(T1) by PRMT
template <int T, int K> __device__ __forceinline__ uint32_t te2(const AesLane& a, uint32_t s) {
    const uint32_t e = __byte_perm(s, a.lb, 0x7604 | (K << 4));
    const uint32_t v = lds_u32_off<(T & 1) * 128>(e);
    return (T & 2) ? __byte_perm(v, 0, 0x1032) : v;
}
__device__ __forceinline__ void aes_round2(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3, const uint4 k) {
    const uint32_t t0 = te2<0, 3>(a, s0) ^ te2<1, 2>(a, s1) ^ te2<2, 1>(a, s2) ^ te2<3, 0>(a, s3) ^ k.x;
    const uint32_t t1 = te2<0, 3>(a, s1) ^ te2<1, 2>(a, s2) ^ te2<2, 1>(a, s3) ^ te2<3, 0>(a, s0) ^ k.y;
    const uint32_t t2 = te2<0, 3>(a, s2) ^ te2<1, 2>(a, s3) ^ te2<2, 1>(a, s0) ^ te2<3, 0>(a, s1) ^ k.z;
    const uint32_t t3 = te2<0, 3>(a, s3) ^ te2<1, 2>(a, s0) ^ te2<2, 1>(a, s1) ^ te2<3, 0>(a, s2) ^ k.w;
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}
template <int U>
__global__ void k_aes2(uint4* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    uint32_t* rk = (uint32_t*)(smem + AES_TABLE_BYTES);
    if (threadIdx.x < 44) rk[threadIdx.x] = threadIdx.x * 0x9e3779b9u;
    __syncthreads();
    const AesLane a = aes_lane(smem);
    uint32_t s[U][4];
#pragma unroll
    for (int j = 0; j < U; j++) { s[j][0] = threadIdx.x + j; s[j][1] = blockIdx.x; s[j][2] = j * 17; s[j][3] = 99; }
    const uint4* k4 = (const uint4*)rk;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int r = 1; r <= 10; r++) {
            const uint4 k = k4[r];
#pragma unroll
            for (int j = 0; j < U; j++) aes_round2(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
        }
    }
    long long t1 = clock64();
    uint4 o = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < U; j++) { o.x ^= s[j][0]; o.y ^= s[j][1]; o.z ^= s[j][2]; o.w ^= s[j][3]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// D: ALU only (LOP3 / PRMT chains)
__global__ void k_alu(uint32_t* out, int iters, long long* cycles) {
    uint32_t a = threadIdx.x, b = blockIdx.x + 1, c = 0x12345, d = 77;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            a = (a & b) ^ c; b = __byte_perm(b, d, 0x5140) ^ a; c = (c | a) ^ d; d = __byte_perm(d, a, 0x3210 + j) ^ b;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint4* out; long long* cyc; CK(cudaMalloc(&out, 148 * 1024 * 16)); CK(cudaMalloc(&cyc, 148 * 8));
    long long h[148];
    const int smem = 65536 + AES_TABLE_BYTES + 1024;
    CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_aes<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_aes<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_aes<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    auto avg = [&]() { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost)); double s = 0; for (int i = 0; i < sms; i++) s += h[i]; return s / sms; };
    printf("SMs %d\n", sms);
    for (int warps : {4, 8, 16, 32}) {
        const int it = 2000;
        k_lds<4><<<sms, warps * 32, smem>>>((uint32_t*)out, it, cyc); double c4 = avg();
        k_lds<8><<<sms, warps * 32, smem>>>((uint32_t*)out, it, cyc); double c8 = avg();
        k_lds<16><<<sms, warps * 32, smem>>>((uint32_t*)out, it, cyc); double c16 = avg();
        const double n = (double)warps * it * 16;
        printf("LDS warps=%2d: LDS.32 %.3f instr/clk/SM (%.0f B/clk)  LDS.64 %.3f (%.0f B/clk)  LDS.128 %.3f (%.0f B/clk)\n", warps,
               n / c4, n / c4 * 128, n / c8, n / c8 * 256, n / c16, n / c16 * 512);
    }
    for (int warps : {4, 6, 8, 12, 16, 24, 32}) {
        const int it = 200;
        double r[3]; int u = 0;
        k_aes<1><<<sms, warps * 32, smem>>>(out, it, cyc); r[u++] = (double)warps * 32 * it * 1 / avg();
        k_aes<2><<<sms, warps * 32, smem>>>(out, it, cyc); r[u++] = (double)warps * 32 * it * 2 / avg();
        if (warps * 32 <= 512) { k_aes<4><<<sms, warps * 32, smem>>>(out, it, cyc); r[u++] = (double)warps * 32 * it * 4 / avg(); } else r[u++] = 0;
        printf("AES-128 warps=%2d: blocks/clk/SM  U=1 %.4f  U=2 %.4f  U=4 %.4f   (x148 SMs x1.9GHz: %.1f / %.1f / %.1f G blocks/s)\n", warps,
               r[0], r[1], r[2], r[0] * 148 * 1.9, r[1] * 148 * 1.9, r[2] * 148 * 1.9);
    }
    CK(cudaFuncSetAttribute(k_aes2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_aes2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int warps : {8, 12, 16, 24, 32}) {
        const int it = 200;
        k_aes2<1><<<sms, warps * 32, smem>>>(out, it, cyc); const double r1 = (double)warps * 32 * it / avg();
        double r2 = 0;
        if (warps <= 16) { k_aes2<2><<<sms, warps * 32, smem>>>(out, it, cyc); r2 = (double)warps * 32 * it * 2 / avg(); }
        printf("AES-128 with 2 resident tables (64 KiB) warps=%2d: blocks/clk/SM U=1 %.4f U=2 %.4f\n", warps, r1, r2);
    }
    for (int warps : {4, 8, 16, 32}) {
        const int it = 2000;
        k_alu<<<sms, warps * 32>>>((uint32_t*)out, it, cyc);
        const double c = avg();
        printf("ALU warps=%2d: %.3f LOP3/PRMT instr/clk/SM\n", warps, (double)warps * it * 16 * 8 / c);
    }
    return 0;
}
