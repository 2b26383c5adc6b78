// aes_core.cuh -- the AES arithmetic of the hot path for sm_100a.
//
// B200 has no AES instruction, so the block cipher the reference reaches through
// Go's crypto/aes (FIPS-197; call sites circuit/garble.go:46,63,92,122,
// circuit/eval.go:20, ot/iknp.go:624, ot/mitccrh.go:82,117) is computed with
// four T-tables held in shared memory.  Each table is replicated 32x so that
// lane l of a warp only ever touches bank l: a warp-wide lookup is exactly one
// conflict-free shared-memory wavefront whatever the 32 indices are.  One
// lookup costs one PRMT (the complete address: 64 KiB-aligned table base in bits
// 16.., byte k of the state word in bits 8..15, the lane's bank offset in bits
// 0..7) and one LDS [R + imm]; one round is 16 PRMT + 16 LDS + 8 LOP3.  The kernels built on this core are bound by the
// shared-memory wavefront rate (16 per round per 32 blocks), not by HBM; see
// DESIGN.md.
//
// State convention: four big-endian column words s0..s3 of the 16-byte block.
// A label {D0,D1} (ot/label.go:28-31, D0 = high half) serialises big-endian
// (GetData, ot/label.go:105-108), so s0 = D0>>32, s1 = (u32)D0, s2 = D1>>32,
// s3 = (u32)D1 -- no byte swaps anywhere on the garble / eval path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gcb {

// ---- FIPS-197 tables, generated at compile time ------------------------------
struct AesConstTables {
    uint8_t sbox[256];
    uint32_t te0[256];     // (02*S, S, S, 03*S) from MSB to LSB
};

constexpr uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t r = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) r ^= a;
        const bool hi = a & 0x80;
        a = (uint8_t)(a << 1);
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return r;
}

constexpr AesConstTables make_aes_tables() {
    AesConstTables t{};
    // multiplicative inverse by walking the generator 3 and its inverse 0xf6
    uint8_t p = 1, q = 1;
    do {
        p = (uint8_t)(p ^ (uint8_t)(p << 1) ^ ((p & 0x80) ? 0x1b : 0));   // p *= 3
        q ^= (uint8_t)(q << 1);                                         // q /= 3
        q ^= (uint8_t)(q << 2);
        q ^= (uint8_t)(q << 4);
        if (q & 0x80) q ^= 0x09;
        const uint8_t x = (uint8_t)(q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6)) ^
                                    (uint8_t)((q << 3) | (q >> 5)) ^ (uint8_t)((q << 4) | (q >> 4)));
        t.sbox[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    t.sbox[0] = 0x63;
    for (int i = 0; i < 256; i++) {
        const uint8_t s = t.sbox[i];
        t.te0[i] = ((uint32_t)gf_mul(s, 2) << 24) | ((uint32_t)s << 16) | ((uint32_t)s << 8) | gf_mul(s, 3);
    }
    return t;
}

constexpr AesConstTables kAesTables = make_aes_tables();
static_assert(kAesTables.sbox[0x00] == 0x63 && kAesTables.sbox[0x01] == 0x7c &&
              kAesTables.sbox[0x53] == 0xed && kAesTables.sbox[0xff] == 0x16, "S-box generation");
static_assert(kAesTables.te0[0] == 0xc66363a5u && kAesTables.te0[1] == 0xf87c7c84u, "Te0 generation");

struct Te0Array { uint32_t v[256]; };
constexpr Te0Array make_te0() {
    Te0Array a{};
    for (int i = 0; i < 256; i++) a.v[i] = kAesTables.te0[i];
    return a;
}
__device__ const Te0Array g_te0 = make_te0();

// ---- shared-memory table layout ----------------------------------------------
// Entry x of table t for lane l lives at byte offset
//     (t >> 1) * 65536 + x * 256 + (t & 1) * 128 + l * 4
// (two tables interleaved per 64 KiB region with a 256-byte entry stride).
// Template parameter NT of everything below: NT = 4 keeps T0..T3 resident (128 KiB).  NT = 2
// keeps only T0/T1 (64 KiB); T2 = rot16(T0) and T3 = rot16(T1) cost one more PRMT per lookup
// of those tables (8 per round) and free 64 KiB of shared memory for wire labels -- the gate
// kernels use it for deep, narrow circuits, which are bound by latency (resident instances),
// not by lookup throughput.
__host__ __device__ constexpr int aes_table_bytes(int nt) { return nt * 256 * 32 * 4; }   // 131072 or 65536
constexpr int AES_TABLE_BYTES = aes_table_bytes(4);
constexpr int AES_MAX_RK_WORDS = 60;                 // AES-256: 15 round keys

__device__ __forceinline__ uint32_t ror8(uint32_t x) { return __funnelshift_r(x, x, 8); }

// Cooperative fill of the replicated tables; all threads of the CTA call it,
// then __syncthreads().
template <int NT = 4>
__device__ __forceinline__ void aes_tables_to_smem(uint8_t* smem_tables) {
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        const int x = i >> 5, l = i & 31;
        const uint32_t t0 = g_te0.v[x];
        const uint32_t t1 = ror8(t0), t2 = ror8(t1), t3 = ror8(t2);
        uint8_t* e = smem_tables + x * 256 + l * 4;
        *reinterpret_cast<uint32_t*>(e) = t0;
        *reinterpret_cast<uint32_t*>(e + 128) = t1;
        if (NT == 4) {
            *reinterpret_cast<uint32_t*>(e + 65536) = t2;
            *reinterpret_cast<uint32_t*>(e + 65536 + 128) = t3;
        }
    }
}

// The tables start at a shared-window address that is a multiple of 64 KiB, so that the
// whole lookup address -- table base, entry x * 256 and the lane's bank offset -- comes
// out of ONE PRMT of the state word with a per-lane constant (no add, no uniform-register
// operand), and the table selection is the immediate offset of the LDS.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint8_t* aes_align_tables(uint8_t* smem) {
    const uint32_t a = smem_addr(smem);
    return smem + (((a + 0xffffu) & ~0xffffu) - a);
}

// Per-thread lookup context: (table base | lane * 4); bits 8..15 are zero.
struct AesLane {
    uint32_t lb;
};
__device__ __forceinline__ AesLane aes_lane(const uint8_t* aligned_tables) {
    AesLane a;
    a.lb = smem_addr(aligned_tables) | ((threadIdx.x & 31u) * 4u);
    return a;
}

template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_off(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}

// Table T lookup of byte K (0 = LSB) of s.
template <int T, int K, int NT = 4>
__device__ __forceinline__ uint32_t te(const AesLane& a, uint32_t s) {
    // PRMT: byte1 = byte K of s, bytes 0, 2, 3 = those of the lane constant
    const uint32_t e = __byte_perm(s, a.lb, 0x7604 | (K << 4));
    if (NT == 4) return lds_u32_off<(T >> 1) * 65536 + (T & 1) * 128>(e);
    const uint32_t v = lds_u32_off<(T & 1) * 128>(e);
    return (T & 2) ? __byte_perm(v, 0, 0x1032) : v;
}

// One full round on (s0..s3) with round-key words k.
template <int NT = 4>
__device__ __forceinline__ void aes_round(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2,
                                          uint32_t& s3, const uint4 k) {
    const uint32_t t0 = te<0, 3, NT>(a, s0) ^ te<1, 2, NT>(a, s1) ^ te<2, 1, NT>(a, s2) ^ te<3, 0, NT>(a, s3) ^ k.x;
    const uint32_t t1 = te<0, 3, NT>(a, s1) ^ te<1, 2, NT>(a, s2) ^ te<2, 1, NT>(a, s3) ^ te<3, 0, NT>(a, s0) ^ k.y;
    const uint32_t t2 = te<0, 3, NT>(a, s2) ^ te<1, 2, NT>(a, s3) ^ te<2, 1, NT>(a, s0) ^ te<3, 0, NT>(a, s1) ^ k.z;
    const uint32_t t3 = te<0, 3, NT>(a, s3) ^ te<1, 2, NT>(a, s0) ^ te<2, 1, NT>(a, s1) ^ te<3, 0, NT>(a, s2) ^ k.w;
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// The same round for NT == 2 with a PRE-ROTATED round key kr = rot16(k) (aes_rotate_mid_keys).  T2 = rot16(T0),
// T3 = rot16(T1), and XOR commutes with the byte rotation, so a column is
//     T0[a] ^ T1[b] ^ rot16(T0[c] ^ T1[d] ^ rot16(k)):
// one rotation per column instead of one per T2 / T3 lookup, with the key folded into the first LOP3 -- 7 ALU
// instructions per column (4 address PRMTs, LOP3, PRMT, LOP3) instead of 8.
__device__ __forceinline__ uint32_t col_rot2(const AesLane& a, uint32_t x3, uint32_t x2, uint32_t x1, uint32_t x0, uint32_t kr) {
    const uint32_t u = te<0, 1, 2>(a, x1) ^ te<1, 0, 2>(a, x0) ^ kr;
    return te<0, 3, 2>(a, x3) ^ te<1, 2, 2>(a, x2) ^ __byte_perm(u, 0, 0x1032);
}
__device__ __forceinline__ void aes_round_rot2(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3,
                                               const uint4 kr) {
    const uint32_t t0 = col_rot2(a, s0, s1, s2, s3, kr.x);
    const uint32_t t1 = col_rot2(a, s1, s2, s3, s0, kr.y);
    const uint32_t t2 = col_rot2(a, s2, s3, s0, s1, kr.z);
    const uint32_t t3 = col_rot2(a, s3, s0, s1, s2, kr.w);
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}
// Round keys 1 .. nr-1 of an expanded schedule -> rot16 of themselves (round 0 and the final round keep theirs).
__device__ __forceinline__ void aes_rotate_mid_keys(uint32_t* rk, int nr) {
    for (int i = 4; i < 4 * nr; i++) rk[i] = __byte_perm(rk[i], 0, 0x1032);
}

// Final round (SubBytes + ShiftRows + AddRoundKey).  S[x] sits in two byte lanes
// of every T-table entry; T2 = (S,3S,2S,S) and T3 = (S,S,3S,2S) have it in the
// top and bottom bytes, so two PRMTs gather the four S-box bytes of a column.
template <int NT = 4>
__device__ __forceinline__ uint32_t last_col(const AesLane& a, uint32_t x3, uint32_t x2, uint32_t x1,
                                             uint32_t x0, uint32_t k) {
    const uint32_t b3 = te<2, 3, NT>(a, x3);        // S in byte 3 (and byte 0)
    const uint32_t b2 = te<3, 2, NT>(a, x2);        // S in bytes 3, 2
    const uint32_t b1 = te<0, 1, NT>(a, x1);        // (2S,S,S,3S): S in bytes 2, 1
    const uint32_t b0 = te<1, 0, NT>(a, x0);        // (3S,2S,S,S): S in bytes 1, 0
    const uint32_t hi = __byte_perm(b3, b2, 0x3600);   // byte3 = b3.3, byte2 = b2.2
    const uint32_t lo = __byte_perm(b1, b0, 0x0014);   // byte1 = b1.1, byte0 = b0.0
    return __byte_perm(hi, lo, 0x3254) ^ k;            // (hi.3, hi.2, lo.1, lo.0)
}
template <int NT = 4>
__device__ __forceinline__ void aes_last_round(const AesLane& a, uint32_t& s0, uint32_t& s1, uint32_t& s2,
                                               uint32_t& s3, const uint4 k) {
    const uint32_t t0 = last_col<NT>(a, s0, s1, s2, s3, k.x);
    const uint32_t t1 = last_col<NT>(a, s1, s2, s3, s0, k.y);
    const uint32_t t2 = last_col<NT>(a, s2, s3, s0, s1, k.z);
    const uint32_t t3 = last_col<NT>(a, s3, s0, s1, s2, k.w);
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// Encrypt one block in place with round keys in shared memory (warp-uniform
// address -> one broadcast LDS.128 per round).  NR = 10/12/14.
template <int NR, int NT = 4>
__device__ __forceinline__ void aes_encrypt_smem(const AesLane& a, const uint32_t* __restrict__ rk,
                                                 uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    const uint4* k4 = reinterpret_cast<const uint4*>(rk);
    uint4 k = k4[0];
    s0 ^= k.x; s1 ^= k.y; s2 ^= k.z; s3 ^= k.w;
#pragma unroll
    for (int r = 1; r < NR; r++) aes_round<NT>(a, s0, s1, s2, s3, k4[r]);
    aes_last_round<NT>(a, s0, s1, s2, s3, k4[NR]);
}

// Same with AES-128 round keys held in registers (IKNP: one key per thread).
template <int NT = 4>
__device__ __forceinline__ void aes128_encrypt_regs(const AesLane& a, const uint32_t (&rk)[44],
                                                    uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
#pragma unroll
    for (int r = 1; r < 10; r++)
        aes_round<NT>(a, s0, s1, s2, s3, make_uint4(rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3]));
    aes_last_round<NT>(a, s0, s1, s2, s3, make_uint4(rk[40], rk[41], rk[42], rk[43]));
}

// SubWord through this lane's table column (Te0 = (2s, s, s, 3s): bits 8..15).
template <int NT = 4>
__device__ __forceinline__ uint32_t sub_word(const AesLane& a, uint32_t w) {
    const uint32_t b3 = te<2, 3, NT>(a, w), b2 = te<3, 2, NT>(a, w), b1 = te<0, 1, NT>(a, w), b0 = te<1, 0, NT>(a, w);
    const uint32_t hi = __byte_perm(b3, b2, 0x3600);
    const uint32_t lo = __byte_perm(b1, b0, 0x0014);
    return __byte_perm(hi, lo, 0x3254);
}

// FIPS-197 key expansion into big-endian words.  key: raw key bytes, keylen
// 16/24/32.  One thread; 4*(nr+1) words written to rk (shared or local).
template <int NT = 4>
__device__ __forceinline__ void aes_expand_key(const AesLane& a, const uint8_t* key, int keylen, uint32_t* rk) {
    const int nk = keylen >> 2, nr = nk + 6;
    for (int i = 0; i < nk; i++)
        rk[i] = (uint32_t(key[4 * i]) << 24) | (uint32_t(key[4 * i + 1]) << 16) |
                (uint32_t(key[4 * i + 2]) << 8) | uint32_t(key[4 * i + 3]);
    uint32_t rcon = 0x01000000u;
    for (int i = nk; i < 4 * (nr + 1); i++) {
        uint32_t t = rk[i - 1];
        if (i % nk == 0) {
            t = sub_word<NT>(a, (t << 8) | (t >> 24)) ^ rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80000000u) ? 0x1b000000u : 0u);
        } else if (nk > 6 && i % nk == 4) {
            t = sub_word<NT>(a, t);
        }
        rk[i] = rk[i - nk] ^ t;
    }
}

// AES-128 key expansion from four big-endian key words into registers.
template <int NT = 4>
__device__ __forceinline__ void aes128_expand_regs(const AesLane& a, uint32_t k0, uint32_t k1, uint32_t k2,
                                                   uint32_t k3, uint32_t (&rk)[44]) {
    rk[0] = k0; rk[1] = k1; rk[2] = k2; rk[3] = k3;
    uint32_t rcon = 0x01000000u;
#pragma unroll
    for (int r = 1; r <= 10; r++) {
        const uint32_t t = rk[4 * r - 1];
        rk[4 * r] = rk[4 * r - 4] ^ sub_word<NT>(a, (t << 8) | (t >> 24)) ^ rcon;
        rk[4 * r + 1] = rk[4 * r - 3] ^ rk[4 * r];
        rk[4 * r + 2] = rk[4 * r - 2] ^ rk[4 * r + 1];
        rk[4 * r + 3] = rk[4 * r - 1] ^ rk[4 * r + 2];
        rcon = (r == 8) ? 0x1b000000u : (rcon << 1);
    }
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
template <int NR, int NT = 4>
__device__ __forceinline__ Label aes_hash_k(const AesLane& a, const uint32_t* rk, Label k) {
    uint32_t s0 = k.w0, s1 = k.w1, s2 = k.w2, s3 = k.w3;
    aes_encrypt_smem<NR, NT>(a, rk, s0, s1, s2, s3);
    return Label{s0 ^ k.w0, s1 ^ k.w1, s2 ^ k.w2, s3 ^ k.w3};
}

}  // namespace gcb
