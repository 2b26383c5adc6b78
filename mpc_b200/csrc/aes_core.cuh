// aes_core.cuh -- the AES arithmetic of the hot path for sm_100a.
//
// B200 has no AES instruction, so the block cipher the reference reaches through
// Go's crypto/aes (FIPS-197; call sites circuit/garble.go:46,63,92,122,
// circuit/eval.go:20, ot/iknp.go:624, ot/mitccrh.go:82,117) is computed with
// four T-tables held in shared memory.  Each table is replicated 32x so that
// lane l of a warp only ever touches bank l: a warp-wide lookup is exactly one
// conflict-free shared-memory wavefront whatever the 32 indices are.  The
// kernels that use this core are bound by that wavefront rate (16 per round per
// 32 blocks), not by HBM; see DESIGN.md.
//
// State convention: four big-endian column words s0..s3 of the 16-byte block.
// A label {D0,D1} (ot/label.go:28-31, D0 = high half) serialises big-endian
// (GetData, ot/label.go:105-108), so s0 = D0>>32, s1 = (u32)D0, s2 = D1>>32,
// s3 = (u32)D1 -- no byte swaps anywhere on the device.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gcb {

// ---- shared-memory table layout ----------------------------------------------
// Entry x of table t for lane l lives at byte offset
//     (t >> 1) * 65536 + x * 256 + (t & 1) * 128 + l * 4
// i.e. two tables interleaved per 64 KiB region with a 256-byte entry stride, so
// that "x * 256 + l * 4" is a single PRMT of the state word with the lane base.
constexpr int AES_TABLE_BYTES = 4 * 256 * 32 * 4;   // 131072
constexpr int AES_MAX_RK_WORDS = 60;                 // AES-256: 15 round keys

// Te0 for the 256 byte values, big-endian column convention:
// Te0[x] = (02*S[x], S[x], S[x], 03*S[x]) from MSB to LSB.  Filled by the host
// at library initialisation (aes_tables_init) from the FIPS-197 definition.
extern __device__ uint32_t g_te0[256];

__device__ __forceinline__ uint32_t ror8(uint32_t x) { return __funnelshift_r(x, x, 8); }

// Cooperative fill of the replicated tables; call with all threads of the CTA,
// then __syncthreads().
__device__ __forceinline__ void aes_tables_to_smem(uint8_t* smem_tables) {
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        const int x = i >> 5, l = i & 31;
        const uint32_t t0 = g_te0[x];
        const uint32_t t1 = ror8(t0), t2 = ror8(t1), t3 = ror8(t2);
        uint8_t* e = smem_tables + x * 256 + l * 4;
        *reinterpret_cast<uint32_t*>(e) = t0;
        *reinterpret_cast<uint32_t*>(e + 128) = t1;
        *reinterpret_cast<uint32_t*>(e + 65536) = t2;
        *reinterpret_cast<uint32_t*>(e + 65536 + 128) = t3;
    }
}

// Per-thread lookup context: 32-bit shared-window address of this lane's column
// in region A, entry 0.
struct AesLane {
    uint32_t base;      // smem address of tables + lane*4
};

__device__ __forceinline__ AesLane aes_lane(const uint8_t* smem_tables) {
    AesLane a;
    a.base = static_cast<uint32_t>(__cvta_generic_to_shared(smem_tables)) + (threadIdx.x & 31) * 4;
    return a;
}

template <int OFF>
__device__ __forceinline__ uint32_t lds_off(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}

// byte k (0 = LSB) of s, scaled by the 256-byte entry stride
template <int K>
__device__ __forceinline__ uint32_t entry_off(uint32_t s) {
    // PRMT: result byte1 = byte K of s, other bytes zero
    return __byte_perm(s, 0, 0x4404 | (K << 4));
}

template <int T, int K>
__device__ __forceinline__ uint32_t te(const AesLane& a, uint32_t s) {
    constexpr int OFF = (T >> 1) * 65536 + (T & 1) * 128;
    return lds_off<OFF>(a.base + entry_off<K>(s));
}

// One full round on (s0..s3) with round-key words k0..k3.
__device__ __forceinline__ void aes_round(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2,
                                          uint32_t& s3, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    const uint32_t t0 = te<0, 3>(a, s0) ^ te<1, 2>(a, s1) ^ te<2, 1>(a, s2) ^ te<3, 0>(a, s3) ^ k0;
    const uint32_t t1 = te<0, 3>(a, s1) ^ te<1, 2>(a, s2) ^ te<2, 1>(a, s3) ^ te<3, 0>(a, s0) ^ k1;
    const uint32_t t2 = te<0, 3>(a, s2) ^ te<1, 2>(a, s3) ^ te<2, 1>(a, s0) ^ te<3, 0>(a, s1) ^ k2;
    const uint32_t t3 = te<0, 3>(a, s3) ^ te<1, 2>(a, s0) ^ te<2, 1>(a, s1) ^ te<3, 0>(a, s2) ^ k3;
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// Final round (SubBytes + ShiftRows + AddRoundKey): S[x] is picked out of the
// T-table entry whose byte at the wanted position is S[x].
__device__ __forceinline__ void aes_last_round(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2,
                                               uint32_t& s3, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    const uint32_t t0 = (te<2, 3>(a, s0) & 0xff000000u) ^ (te<3, 2>(a, s1) & 0x00ff0000u) ^
                        (te<0, 1>(a, s2) & 0x0000ff00u) ^ (te<1, 0>(a, s3) & 0x000000ffu) ^ k0;
    const uint32_t t1 = (te<2, 3>(a, s1) & 0xff000000u) ^ (te<3, 2>(a, s2) & 0x00ff0000u) ^
                        (te<0, 1>(a, s3) & 0x0000ff00u) ^ (te<1, 0>(a, s0) & 0x000000ffu) ^ k1;
    const uint32_t t2 = (te<2, 3>(a, s2) & 0xff000000u) ^ (te<3, 2>(a, s3) & 0x00ff0000u) ^
                        (te<0, 1>(a, s0) & 0x0000ff00u) ^ (te<1, 0>(a, s1) & 0x000000ffu) ^ k2;
    const uint32_t t3 = (te<2, 3>(a, s3) & 0xff000000u) ^ (te<3, 2>(a, s0) & 0x00ff0000u) ^
                        (te<0, 1>(a, s1) & 0x0000ff00u) ^ (te<1, 0>(a, s2) & 0x000000ffu) ^ k3;
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// Encrypt one block in place.  rk: round-key words in shared memory (uniform
// address across the warp -> broadcast loads), nr = 10/12/14.
__device__ __forceinline__ void aes_encrypt(const AesLane& a, const uint32_t* __restrict__ rk, int nr,
                                            uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    uint4 k = *reinterpret_cast<const uint4*>(rk);
    s0 ^= k.x; s1 ^= k.y; s2 ^= k.z; s3 ^= k.w;
#pragma unroll 1
    for (int r = 1; r < nr; r++) {
        k = *reinterpret_cast<const uint4*>(rk + 4 * r);
        aes_round(a, s0, s1, s2, s3, k.x, k.y, k.z, k.w);
    }
    k = *reinterpret_cast<const uint4*>(rk + 4 * nr);
    aes_last_round(a, s0, s1, s2, s3, k.x, k.y, k.z, k.w);
}

// S-box byte through the tables (Te0 = (2s, s, s, 3s): bits 8..15 hold S[x]).
__device__ __forceinline__ uint32_t sbox_byte(const uint8_t* smem_tables, uint32_t x) {
    return (*reinterpret_cast<const uint32_t*>(smem_tables + (x & 0xff) * 256) >> 8) & 0xff;
}
__device__ __forceinline__ uint32_t sub_word(const uint8_t* t, uint32_t w) {
    return (sbox_byte(t, w >> 24) << 24) | (sbox_byte(t, w >> 16) << 16) | (sbox_byte(t, w >> 8) << 8) |
           sbox_byte(t, w);
}

// FIPS-197 key expansion into big-endian words.  key: raw key bytes (global or
// shared), keylen 16/24/32.  One thread; 4*(nr+1) words written.  Returns nr.
__device__ __forceinline__ int aes_expand_key(const uint8_t* smem_tables, const uint8_t* key, int keylen,
                                              uint32_t* rk) {
    const int nk = keylen >> 2, nr = nk + 6;
    for (int i = 0; i < nk; i++)
        rk[i] = (uint32_t(key[4 * i]) << 24) | (uint32_t(key[4 * i + 1]) << 16) |
                (uint32_t(key[4 * i + 2]) << 8) | uint32_t(key[4 * i + 3]);
    uint32_t rcon = 0x01000000u;
    for (int i = nk; i < 4 * (nr + 1); i++) {
        uint32_t t = rk[i - 1];
        if (i % nk == 0) {
            t = sub_word(smem_tables, (t << 8) | (t >> 24)) ^ rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80000000u) ? 0x1b000000u : 0u);
        } else if (nk > 6 && i % nk == 4) {
            t = sub_word(smem_tables, t);
        }
        rk[i] = rk[i - nk] ^ t;
    }
    return nr;
}

// ---- label <-> state -----------------------------------------------------------
// A label in Go memory order is {u64 D0, u64 D1}; as a uint4 of little-endian
// words it is (lo(D0), hi(D0), lo(D1), hi(D1)).
struct Label {
    uint32_t w0, w1, w2, w3;     // big-endian column words: w0 = D0>>32 ... w3 = (u32)D1
};
__device__ __forceinline__ Label label_from_mem(uint4 m) { return Label{m.y, m.x, m.w, m.z}; }
__device__ __forceinline__ uint4 label_to_mem(Label l) { return make_uint4(l.w1, l.w0, l.w3, l.w2); }
__device__ __forceinline__ Label operator^(Label a, Label b) {
    return Label{a.w0 ^ b.w0, a.w1 ^ b.w1, a.w2 ^ b.w2, a.w3 ^ b.w3};
}
__device__ __forceinline__ Label label_and_mask(Label a, uint32_t m) {   // m = 0 or 0xffffffff
    return Label{a.w0 & m, a.w1 & m, a.w2 & m, a.w3 & m};
}
__device__ __forceinline__ uint32_t label_s(Label a) { return a.w0 >> 31; }      // S(), ot/label.go:65-67
// Mul2 / Mul4, ot/label.go:79-90: 128-bit left shift, top bits dropped.
__device__ __forceinline__ Label label_shl(Label a, int n) {
    return Label{__funnelshift_l(a.w1, a.w0, n), __funnelshift_l(a.w2, a.w1, n),
                 __funnelshift_l(a.w3, a.w2, n), a.w3 << n};
}

// H(K) = AES(K) ^ K  -- the tail shared by encryptHalf (circuit/garble.go:104-136)
// and encrypt/decrypt (circuit/garble.go:40-73).
__device__ __forceinline__ Label aes_hash_k(const AesLane& a, const uint32_t* rk, int nr, Label k) {
    uint32_t s0 = k.w0, s1 = k.w1, s2 = k.w2, s3 = k.w3;
    aes_encrypt(a, rk, nr, s0, s1, s2, s3);
    return Label{s0 ^ k.w0, s1 ^ k.w1, s2 ^ k.w2, s3 ^ k.w3};
}

}  // namespace gcb
