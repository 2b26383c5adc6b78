// microbench_bs.cu -- the bitsliced AES core BASELINE.json's north_star names, measured next to the
// T-table core the kernels use.  Device analogue of the reference's stand-alone
// AES benchmark circuit/aesni/c/aesni.c:119-147 / BenchmarkEncHalf (circuit/enc_test.go:73).
//
//   bitsliced:  one thread encrypts 32 blocks at once; the state is 128 bit planes (plane[i][j] = bit i of
//               state byte j of the thread's 32 blocks), SubBytes is the 113-gate Boyar-Peralta circuit
//               (eprint 2009/191) evaluated on whole planes, ShiftRows is a renaming, MixColumns and
//               AddRoundKey are XORs (nvcc merges the 2-input gates into LOP3).  ALU pipe only.
//   T-table:    aes_core.cuh (160 shared-memory lookups per block).  Shared-memory pipe.
//   hybrid:     both forms on the same SM at the same time (two kernels, one CTA of each per SM) -- they sit on
//               different pipes, so the question is how much of the sum survives sharing the issue slots and the
//               ALU pipe (the T-table form needs 24 ALU instructions per 16 lookups).
//   + the 32x32 bit transposes that move labels into and out of bit planes (4 per direction per 32 blocks).
//
// Every form is checked against a host AES-128 (FIPS-197) before it is timed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_bs tools/microbench_bs.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../mpc_b200/csrc/aes_core.cuh"
using namespace gcb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ---- Boyar-Peralta S-box on bit planes: x[0] = MSB ... x[7] = LSB, same for s ------------------------------
template <class W>
__host__ __device__ __forceinline__ void bp_sbox(const W x0, const W x1, const W x2, const W x3, const W x4, const W x5,
                                                 const W x6, const W x7, W& s0, W& s1, W& s2, W& s3, W& s4, W& s5, W& s6, W& s7) {
    const W y14 = x3 ^ x5, y13 = x0 ^ x6, y9 = x0 ^ x3, y8 = x0 ^ x5, t0 = x1 ^ x2, y1 = t0 ^ x7, y4 = y1 ^ x3;
    const W y12 = y13 ^ y14, y2 = y1 ^ x0, y5 = y1 ^ x6, y3 = y5 ^ y8, t1 = x4 ^ y12, y15 = t1 ^ x5, y20 = t1 ^ x1;
    const W y6 = y15 ^ x7, y10 = y15 ^ t0, y11 = y20 ^ y9, y7 = x7 ^ y11, y17 = y10 ^ y11, y19 = y10 ^ y8;
    const W y16 = t0 ^ y11, y21 = y13 ^ y16, y18 = x0 ^ y16;
    const W t2 = y12 & y15, t3 = y3 & y6, t4 = t3 ^ t2, t5 = y4 & x7, t6 = t5 ^ t2, t7 = y13 & y16, t8 = y5 & y1;
    const W t9 = t8 ^ t7, t10 = y2 & y7, t11 = t10 ^ t7, t12 = y9 & y11, t13 = y14 & y17, t14 = t13 ^ t12;
    const W t15 = y8 & y10, t16 = t15 ^ t12, t17 = t4 ^ t14, t18 = t6 ^ t16, t19 = t9 ^ t14, t20 = t11 ^ t16;
    const W t21 = t17 ^ y20, t22 = t18 ^ y19, t23 = t19 ^ y21, t24 = t20 ^ y18;
    const W t25 = t21 ^ t22, t26 = t21 & t23, t27 = t24 ^ t26, t28 = t25 & t27, t29 = t28 ^ t22, t30 = t23 ^ t24;
    const W t31 = t22 ^ t26, t32 = t31 & t30, t33 = t32 ^ t24, t34 = t23 ^ t33, t35 = t27 ^ t33, t36 = t24 & t35;
    const W t37 = t36 ^ t34, t38 = t27 ^ t36, t39 = t29 & t38, t40 = t25 ^ t39;
    const W t41 = t40 ^ t37, t42 = t29 ^ t33, t43 = t29 ^ t40, t44 = t33 ^ t37, t45 = t42 ^ t41;
    const W z0 = t44 & y15, z1 = t37 & y6, z2 = t33 & x7, z3 = t43 & y16, z4 = t40 & y1, z5 = t29 & y7, z6 = t42 & y11;
    const W z7 = t45 & y17, z8 = t41 & y10, z9 = t44 & y12, z10 = t37 & y3, z11 = t33 & y4, z12 = t43 & y13;
    const W z13 = t40 & y5, z14 = t29 & y2, z15 = t42 & y9, z16 = t45 & y14, z17 = t41 & y8;
    const W t46 = z15 ^ z16, t47 = z10 ^ z11, t48 = z5 ^ z13, t49 = z9 ^ z10, t50 = z2 ^ z12, t51 = z2 ^ z5;
    const W t52 = z7 ^ z8, t53 = z0 ^ z3, t54 = z6 ^ z7, t55 = z16 ^ z17, t56 = z12 ^ t48, t57 = t50 ^ t53;
    const W t58 = z4 ^ t46, t59 = z3 ^ t54, t60 = t46 ^ t57, t61 = z14 ^ t57, t62 = t52 ^ t58, t63 = t49 ^ t58;
    const W t64 = z4 ^ t59, t65 = t61 ^ t62, t66 = z1 ^ t63;
    s0 = t59 ^ t63; s6 = t56 ^ ~t62; s7 = t48 ^ ~t60;
    const W t67 = t64 ^ t65;
    s3 = t53 ^ t66; s4 = t51 ^ t66; s5 = t47 ^ t65; s1 = t64 ^ ~s3; s2 = t55 ^ ~t67;
}

// One round on the 128 planes.  b[i][j]: bit i (0 = LSB) of state byte j (j = row + 4 * column).
// rk: 128 words of 0 / ~0 for this round (plane order), shared memory.
template <bool LAST>
__device__ __forceinline__ void bs_round(uint32_t (&b)[8][16], const uint32_t* __restrict__ rk) {
    uint32_t s[8][16];
#pragma unroll
    for (int j = 0; j < 16; j++)
        bp_sbox<uint32_t>(b[7][j], b[6][j], b[5][j], b[4][j], b[3][j], b[2][j], b[1][j], b[0][j],
                          s[7][j], s[6][j], s[5][j], s[4][j], s[3][j], s[2][j], s[1][j], s[0][j]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        // ShiftRows: the byte of row r in column c comes from column (c + r) % 4
        uint32_t a[4][8];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[r][i] = s[i][r + 4 * ((c + r) & 3)];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int r1 = (r + 1) & 3, r2 = (r + 2) & 3, r3 = (r + 3) & 3;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t v;
                if (LAST) v = a[r][i];
                else {
                    // out_r = xtime(a_r ^ a_r1) ^ a_r1 ^ a_r2 ^ a_r3; xtime: bit i <- bit i-1, bits 0,1,3,4 ^= bit 7
                    uint32_t xt = i ? (a[r][i - 1] ^ a[r1][i - 1]) : 0u;
                    if (i == 0 || i == 1 || i == 3 || i == 4) xt ^= a[r][7] ^ a[r1][7];
                    v = xt ^ a[r1][i] ^ a[r2][i] ^ a[r3][i];
                }
                b[i][r + 4 * c] = v ^ rk[i * 16 + r + 4 * c];
            }
        }
    }
}

// 10 rounds; planes of round key 0 are XORed in first.  rk: [11][128] masks.
__device__ __forceinline__ void bs_encrypt(uint32_t (&b)[8][16], const uint32_t* __restrict__ rk) {
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) b[i][j] ^= rk[i * 16 + j];
#pragma unroll 1
    for (int r = 1; r < 10; r++) bs_round<false>(b, rk + 128 * r);
    bs_round<true>(b, rk + 1280);
}

// 32x32 bit transpose inside one thread (5 masked-swap stages, 80 swaps of 3 ops... the label <-> plane move)
__device__ __forceinline__ void transpose32(uint32_t (&m)[32]) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t mask = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k++)
            if ((k & j) == 0) {
                const uint32_t t = ((m[k] >> j) ^ m[k + j]) & mask;
                m[k] ^= t << j;
                m[k + j] ^= t;
            }
    }
}

__global__ void __launch_bounds__(256, 1) k_bs(uint32_t* io, const uint32_t* rk_g, int iters, int with_transpose, long long* cycles) {
    __shared__ uint32_t rk[11 * 128];
    for (int i = threadIdx.x; i < 11 * 128; i += blockDim.x) rk[i] = rk_g[i];
    __syncthreads();
    uint32_t b[8][16];
    uint32_t* mine = io + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) b[i][j] = mine[i * 16 + j];
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (with_transpose) {                     // planes -> labels -> planes: what every hash of a label would pay
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t m[32];
#pragma unroll
                for (int k = 0; k < 32; k++) m[k] = b[k & 7][4 * q + (k >> 3)];
                transpose32(m);
                transpose32(m);
#pragma unroll
                for (int k = 0; k < 32; k++) b[k & 7][4 * q + (k >> 3)] = m[k];
            }
        }
        bs_encrypt(b, rk);
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) mine[i * 16 + j] = b[i][j];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// T-table form, U interleaved blocks per thread (as tools/microbench_lds.cu), and the hybrid CTA
template <int U>
__device__ __forceinline__ void tt_loop(const AesLane& a, const uint32_t* rk, uint32_t (&s)[U][4], int iters) {
    const uint4* k4 = (const uint4*)rk;
    for (int it = 0; it < iters; it++) {
        {
            const uint4 k = k4[0];
#pragma unroll
            for (int j = 0; j < U; j++) { s[j][0] ^= k.x; s[j][1] ^= k.y; s[j][2] ^= k.z; s[j][3] ^= k.w; }
        }
#pragma unroll 1
        for (int r = 1; r < 10; r++) {
            const uint4 k = k4[r];
#pragma unroll
            for (int j = 0; j < U; j++) aes_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
        }
#pragma unroll
        for (int j = 0; j < U; j++) aes_last_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k4[10]);
    }
}
// The hybrid: the two forms as two kernels that share each SM (one CTA of each per SM: the register file holds
// 12 T-table warps at <= 80 registers beside 4 bitsliced warps at 255).  Each CTA records its SM and its time span
// (%globaltimer), so the host can see that the CTAs really ran side by side.
struct Span { unsigned long long t0, t1; unsigned smid, pad; };
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

__global__ void __launch_bounds__(384, 1) k_tt(uint4* tt_io, const uint32_t* rkw_g, int iters, Span* spans, long long* cycles) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    uint32_t* rkw = (uint32_t*)(smem + AES_TABLE_BYTES);
    if (threadIdx.x < 44) rkw[threadIdx.x] = rkw_g[threadIdx.x];
    __syncthreads();
    const AesLane a = aes_lane(smem);
    uint32_t s[2][4];
    uint4* mine = tt_io + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
#pragma unroll
    for (int j = 0; j < 2; j++) { const uint4 v = mine[j]; s[j][0] = v.x; s[j][1] = v.y; s[j][2] = v.z; s[j][3] = v.w; }
    const unsigned long long g0 = gtime();
    const long long t0 = clock64();
    tt_loop<2>(a, rkw, s, iters);
    const long long t1 = clock64();
    const unsigned long long g1 = gtime();
#pragma unroll
    for (int j = 0; j < 2; j++) mine[j] = make_uint4(s[j][0], s[j][1], s[j][2], s[j][3]);
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; spans[blockIdx.x] = Span{g0, g1, smid(), 0}; }
}
__global__ void __launch_bounds__(128, 1) k_bs_span(uint32_t* io, const uint32_t* rk_g, int iters, Span* spans, long long* cycles) {
    __shared__ uint32_t rk[11 * 128];
    for (int i = threadIdx.x; i < 11 * 128; i += blockDim.x) rk[i] = rk_g[i];
    __syncthreads();
    uint32_t b[8][16];
    uint32_t* mine = io + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) b[i][j] = mine[i * 16 + j];
    const unsigned long long g0 = gtime();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) bs_encrypt(b, rk);
    const long long t1 = clock64();
    const unsigned long long g1 = gtime();
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) mine[i * 16 + j] = b[i][j];
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; spans[blockIdx.x] = Span{g0, g1, smid(), 0}; }
}

// ---- host reference AES-128 (FIPS-197) --------------------------------------------------------------------
static void host_expand(const uint8_t key[16], uint8_t rk[176]) {
    memcpy(rk, key, 16);
    uint8_t rcon = 1;
    for (int i = 16; i < 176; i += 4) {
        uint8_t t[4] = {rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1]};
        if (i % 16 == 0) {
            const uint8_t u = t[0];
            t[0] = kAesTables.sbox[t[1]] ^ rcon; t[1] = kAesTables.sbox[t[2]]; t[2] = kAesTables.sbox[t[3]]; t[3] = kAesTables.sbox[u];
            rcon = gf_mul(rcon, 2);
        }
        for (int k = 0; k < 4; k++) rk[i + k] = rk[i - 16 + k] ^ t[k];
    }
}
static void host_encrypt(const uint8_t rk[176], uint8_t s[16]) {
    for (int k = 0; k < 16; k++) s[k] ^= rk[k];
    for (int r = 1; r <= 10; r++) {
        uint8_t t[16];
        for (int c = 0; c < 4; c++) for (int row = 0; row < 4; row++) t[row + 4 * c] = kAesTables.sbox[s[row + 4 * ((c + row) & 3)]];
        for (int c = 0; c < 4; c++) {
            uint8_t* a = t + 4 * c;
            for (int row = 0; row < 4; row++) {
                uint8_t v = a[row];
                if (r < 10) v = gf_mul(a[row], 2) ^ gf_mul(a[(row + 1) & 3], 3) ^ a[(row + 2) & 3] ^ a[(row + 3) & 3];
                s[row + 4 * c] = v ^ rk[16 * r + row + 4 * c];
            }
        }
    }
}

int main() {
    // the S-box circuit on all 256 inputs (bit planes of one "block" each)
    for (int x = 0; x < 256; x++) {
        uint32_t in[8], out[8];
        for (int i = 0; i < 8; i++) in[i] = (x >> (7 - i)) & 1 ? ~0u : 0u;
        bp_sbox<uint32_t>(in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
        int y = 0;
        for (int i = 0; i < 8; i++) y |= (out[i] & 1) << (7 - i);
        if (y != kAesTables.sbox[x]) { printf("S-box circuit wrong at %02x: %02x != %02x\n", x, y, kAesTables.sbox[x]); return 1; }
    }
    printf("Boyar-Peralta S-box circuit: all 256 inputs correct\n");
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint8_t key[16] = {0x2b, 0x7e, 0x15, 0x16, 0x28, 0xae, 0xd2, 0xa6, 0xab, 0xf7, 0x15, 0x88, 0x09, 0xcf, 0x4f, 0x3c};
    uint8_t rkb[176];
    host_expand(key, rkb);
    uint32_t rk_masks[11 * 128], rk_words[44];
    for (int r = 0; r < 11; r++)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 16; j++) rk_masks[r * 128 + i * 16 + j] = (rkb[16 * r + j] >> i) & 1 ? ~0u : 0u;
    for (int w = 0; w < 44; w++) rk_words[w] = ((uint32_t)rkb[4 * w] << 24) | ((uint32_t)rkb[4 * w + 1] << 16) | ((uint32_t)rkb[4 * w + 2] << 8) | rkb[4 * w + 3];
    const int max_threads = 512;
    const size_t n_words = (size_t)sms * max_threads * 128;
    uint32_t *io, *rk_d, *rkw_d; uint4* tt_io; long long* cyc;
    CK(cudaMalloc(&io, n_words * 4)); CK(cudaMalloc(&rk_d, sizeof rk_masks)); CK(cudaMalloc(&rkw_d, sizeof rk_words));
    CK(cudaMalloc(&tt_io, (size_t)sms * max_threads * 2 * 16)); CK(cudaMalloc(&cyc, sms * 16));
    CK(cudaMemcpy(rk_d, rk_masks, sizeof rk_masks, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rkw_d, rk_words, sizeof rk_words, cudaMemcpyHostToDevice));
    // correctness of the bitsliced core: thread 0's 32 blocks against the host AES
    {
        uint8_t blocks[32][16];
        uint32_t planes[128] = {0};
        for (int n = 0; n < 32; n++)
            for (int j = 0; j < 16; j++) {
                blocks[n][j] = (uint8_t)(n * 37 + j * 11 + 5);
                for (int i = 0; i < 8; i++) planes[i * 16 + j] |= (uint32_t)((blocks[n][j] >> i) & 1) << n;
            }
        CK(cudaMemset(io, 0, n_words * 4));
        CK(cudaMemcpy(io, planes, sizeof planes, cudaMemcpyHostToDevice));
        k_bs<<<1, 32>>>(io, rk_d, 1, 0, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(planes, io, sizeof planes, cudaMemcpyDeviceToHost));
        for (int n = 0; n < 32; n++) {
            host_encrypt(rkb, blocks[n]);
            for (int j = 0; j < 16; j++) {
                int v = 0;
                for (int i = 0; i < 8; i++) v |= ((planes[i * 16 + j] >> n) & 1) << i;
                if (v != blocks[n][j]) { printf("bitsliced AES wrong: block %d byte %d\n", n, j); return 1; }
            }
        }
        printf("bitsliced AES-128: 32 blocks equal the host AES (FIPS-197)\n");
    }
    long long h[2 * 148];
    auto avg = [&](int stride, int off) { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, sms * 8 * stride, cudaMemcpyDeviceToHost)); double s = 0; for (int i = 0; i < sms; i++) s += h[i * stride + off]; return s / sms; };
    printf("SMs %d\n", sms);
    for (int warps : {1, 2, 4, 8}) {
        const int it = 20;
        k_bs<<<sms, warps * 32>>>(io, rk_d, it, 0, cyc);
        const double c0 = avg(1, 0);
        k_bs<<<sms, warps * 32>>>(io, rk_d, it, 1, cyc);
        const double c1 = avg(1, 0);
        const double blocks = (double)warps * 32 * 32 * it;
        printf("bitsliced AES-128 warps=%d: %.4f blocks/clk/SM core only, %.4f with the label<->plane transposes  (%.1f / %.1f G blocks/s at 148 SMs x 1.9 GHz)\n",
               warps, blocks / c0, blocks / c1, blocks / c0 * 148 * 1.9, blocks / c1 * 148 * 1.9);
    }
    // hybrid: 12 T-table warps and 1 / 2 / 4 bitsliced warps per SM at the same time, as two kernels
    const int smem = 65536 + AES_TABLE_BYTES + 1024;
    CK(cudaFuncSetAttribute(k_tt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // kernels that ask for different shared-memory carve-outs cannot share an SM
    CK(cudaFuncSetAttribute(k_tt, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_bs_span, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    Span *sp_tt, *sp_bs; long long* cyc2;
    CK(cudaMalloc(&sp_tt, sms * sizeof(Span))); CK(cudaMalloc(&sp_bs, sms * sizeof(Span))); CK(cudaMalloc(&cyc2, sms * 8));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    std::vector<Span> ht(sms), hb(sms);
    const int tt_iters = 4000;
    k_tt<<<sms, 384, smem, s1>>>(tt_io, rkw_d, tt_iters, sp_tt, cyc);
    const double tt_alone = avg(1, 0);
    printf("T-table alone, 12 warps x 2 blocks: %.4f blocks/clk/SM\n", 12.0 * 32 * 2 * tt_iters / tt_alone);
    for (int bs_warps : {1, 2, 4}) {
        // the bitsliced warps run about as long as the T-table warps (sized from the stand-alone rates above)
        k_bs_span<<<sms, bs_warps * 32, 0, s2>>>(io, rk_d, 8, sp_bs, cyc2);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc2, sms * 8, cudaMemcpyDeviceToHost));
        double c8 = 0; for (int i = 0; i < sms; i++) c8 += h[i]; c8 /= sms;
        const int bs_iters = (int)(8 * tt_alone * 1.15 / c8) + 1;
        cudaEvent_t e0, e1, e2;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        CK(cudaStreamWaitEvent(s2, e0, 0));
        k_tt<<<sms, 384, smem, s1>>>(tt_io, rkw_d, tt_iters, sp_tt, cyc);
        k_bs_span<<<sms, bs_warps * 32, 0, s2>>>(io, rk_d, bs_iters, sp_bs, cyc2);
        CK(cudaEventRecord(e1, s1)); CK(cudaEventRecord(e2, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ht.data(), sp_tt, sms * sizeof(Span), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hb.data(), sp_bs, sms * sizeof(Span), cudaMemcpyDeviceToHost));
        const double ct = avg(1, 0);
        CK(cudaMemcpy(h, cyc2, sms * 8, cudaMemcpyDeviceToHost));
        double cb = 0; for (int i = 0; i < sms; i++) cb += h[i]; cb /= sms;
        // co-residency: for every SM, how much of the T-table CTA's span a bitsliced CTA on the same SM covered
        double covered = 0; int paired = 0;
        for (int i = 0; i < sms; i++)
            for (int j = 0; j < sms; j++)
                if (ht[i].smid == hb[j].smid) {
                    const double lo = (double)std::max(ht[i].t0, hb[j].t0), hi = (double)std::min(ht[i].t1, hb[j].t1);
                    if (hi > lo) { covered += (hi - lo) / (double)(ht[i].t1 - ht[i].t0); paired++; }
                }
        const double tt_blocks = 12.0 * 32 * 2 * tt_iters, bs_blocks = (double)bs_warps * 32 * 32 * bs_iters;
        printf("hybrid 12 T-table warps + %d bitsliced warps per SM: T-table %.4f blocks/clk/SM (alone %.4f), bitsliced %.4f; "
               "sum while both run %.4f; %d of %d SMs shared, %.0f %% of the T-table span covered\n", bs_warps, tt_blocks / ct,
               12.0 * 32 * 2 * tt_iters / tt_alone, bs_blocks / cb, tt_blocks / ct + bs_blocks / cb, paired, sms, 100.0 * covered / sms);
    }
    return 0;
}
