// ot_kernels.cuh -- IKNP OT-extension, MiTCCRH and gate-hash kernels for sm_100a.
//
// K4  iknp_kernel      replaces the chunk loops of IKNPSender.send
//                      (ot/iknp.go:197-226) and IKNPReceiver.receive (:468-511):
//                      AES-128-CTR column PRG (newPrg/prg :622-637), the U matrix
//                      and the 128-wide bit transpose createLabels (:647-683).
// K5  mitccrh_kernel   replaces MITCCRH.Hash (ot/mitccrh.go:93-128) with one
//                      key schedule per key (renewKeys :70-89).
// K6  hash_half_kernel replaces the encryptHalf micro-benchmarks
//                      (circuit/garble.go:104-136, circuit/enc_test.go:73,
//                      circuit/aesni/c/aesni.c).
//
// IKNP execution model.  One thread owns one matrix column (one PRG key) for
// the whole launch and keeps that key's 44 round-key words in registers.  A
// GROUP of 128 (sender) or 256 (receiver: T0 and T1 halves) threads processes
// one 512-row chunk of the reference's wire format at a time: four (five when
// the byte-granular stream position is not block aligned) CTR blocks per thread,
// XOR with the chunk's U bytes, then a 32x32 bit transpose per warp with
// shuffles, staged through shared memory so that the 512 labels of the chunk
// leave as 8 KiB of coalesced 16-byte stores.
#pragma once
#include "aes_core.cuh"

namespace gcb {

// ------------------------------------------------------------- hash_half (K6) --
struct HashParams {
    const uint8_t* key;       // device pointer to the raw key
    uint32_t keylen;
    const uint4* x;
    uint4* out;
    uint32_t tweak0;
    uint64_t n;
};

template <int NR>
__global__ void __launch_bounds__(1024, 1) hash_half_kernel(const HashParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    uint32_t* rk = reinterpret_cast<uint32_t*>(smem + AES_TABLE_BYTES);
    aes_tables_to_smem(smem);
    __syncthreads();
    const AesLane lane = aes_lane(smem);
    if (threadIdx.x == 0) aes_expand_key(lane, p.key, (int)p.keylen, rk);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const Label x = label_from_mem(__ldg(p.x + i));
        Label K = label_shl(x, 1);                       // Mul2, garble.go:110-113
        K.w3 ^= p.tweak0 + (uint32_t)i;                   // NewTweak: low 32 bits, garble.go:114
        p.out[i] = label_to_mem(aes_hash_k<NR>(lane, rk, K));
    }
}

// --------------------------------------------------------------- MiTCCRH (K5) --
struct MitccrhParams {
    uint64_t seed_d0, seed_d1;
    uint64_t gid_start;
    uint4* blks;
    uint64_t nkeys;
    uint32_t h;
};

__global__ void __launch_bounds__(512, 1) mitccrh_kernel(const MitccrhParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    __syncthreads();
    const AesLane lane = aes_lane(smem);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nkeys; i += stride) {
        // key g = BE64(seed.D0 ^ g) || BE64(seed.D1), mitccrh.go:72-86
        const uint64_t d0 = p.seed_d0 ^ (p.gid_start + i);
        uint32_t rk[44];
        aes128_expand_regs(lane, (uint32_t)(d0 >> 32), (uint32_t)d0, (uint32_t)(p.seed_d1 >> 32),
                           (uint32_t)p.seed_d1, rk);
        uint4* b = p.blks + i * p.h;
        for (uint32_t j = 0; j < p.h; j++) {             // mitccrh.go:112-127: AES(x) ^ x in place
            const Label x = label_from_mem(b[j]);
            uint32_t s0 = x.w0, s1 = x.w1, s2 = x.w2, s3 = x.w3;
            aes128_encrypt_regs(lane, rk, s0, s1, s2, s3);
            b[j] = label_to_mem(Label{s0 ^ x.w0, s1 ^ x.w1, s2 ^ x.w2, s3 ^ x.w3});
        }
    }
}

// ------------------------------------------------------------------ IKNP (K4) --
constexpr int IKNP_CTA_THREADS = 512;
constexpr int IKNP_CHUNK_ROWS = 512;       // ot/iknp.go:64-77: 8 KiB chunks = 64 byte-rows
constexpr int IKNP_STAGE_BYTES = 8192;     // 512 labels

struct IknpParams {
    const uint4* k0;          // [128] PRG keys (sender: its keys; receiver: the L0 seeds)
    const uint4* k1;          // [128] receiver only: the L1 seeds
    const uint4* delta;       // sender only
    uint64_t stream_pos;      // keystream bytes already consumed by every column PRG
    const uint8_t* choice;    // receiver: n bytes of 0/1 (label variant)
    const uint32_t* choice_bits;   // receiver, bit variant: packed LSB-first words (the []uint64 of ReceiveBits)
    uint32_t* result_bits;    // bit variants: packed LSB-first output bits (column 0 of the matrix)
    const uint8_t* u_in;      // sender: received U, chunked layout
    uint8_t* u_out;           // receiver: U to send
    uint4* labels;            // [n] out
    uint64_t n;
    uint32_t* counter;        // next chunk to claim
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// 32x32 bit transpose across a warp: in, lane l holds row l; out, lane k holds
// column k (bit l = old lane l's bit k).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = (j == 16) ? 0x0000ffffu : (j == 8) ? 0x00ff00ffu : (j == 4) ? 0x0f0f0f0fu
                          : (j == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        const bool up = lane & j;
        const uint32_t t = __funnelshift_l(y, y, up ? 32 - j : j);   // rotate towards our half
        const uint32_t keep = up ? ~m : m;
        x = (x & keep) | (t & ~keep);
    }
    return x;
}

// Swizzled staging address of word w of row r (conflict-free for the writes of
// one warp: lane = r & 31, fixed w).
__device__ __forceinline__ uint32_t stage_word(uint32_t r, uint32_t w) { return r * 4 + (w ^ ((r >> 3) & 3)); }

template <int WOFS>
__device__ __forceinline__ void ks_extract(const uint32_t (&W)[20], uint32_t sh, uint32_t (&T)[16]) {
#pragma unroll
    for (int q = 0; q < 16; q++) T[q] = __funnelshift_r(W[q + WOFS], W[q + WOFS + 1], sh);
}

// RECEIVER = false: group of 128 threads, thread c = column c.
// RECEIVER = true : group of 256 threads, thread c < 128 -> T0 of column c,
//                   thread c >= 128 -> T1 of column c - 128.
// BITS: the bit-COT variants SendBits / ReceiveBits (ot/iknp.go:259-310, 554-620): same U
// exchange, but the output is only Bit(0) of every label, i.e. column 0 of the matrix, packed
// LSB-first; no transpose is needed.
template <bool RECEIVER, bool BITS>
__global__ void __launch_bounds__(IKNP_CTA_THREADS, 1) iknp_kernel(const IknpParams p) {
    constexpr int GT = RECEIVER ? 256 : 128;              // threads per group
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    __syncthreads();
    const AesLane al = aes_lane(smem);
    const uint32_t group = threadIdx.x / GT, gt = threadIdx.x % GT;
    const uint32_t col = gt & 127u, lane = threadIdx.x & 31u, wcol = (gt >> 5) & 3u;
    const bool second = RECEIVER && gt >= 128;
    constexpr int XCHG = RECEIVER ? 8192 : 0;                             // [16][128] T1 words (receiver only)
    uint8_t* gs = smem + AES_TABLE_BYTES + group * (IKNP_STAGE_BYTES + XCHG + 128);
    uint32_t* stage = reinterpret_cast<uint32_t*>(gs);                    // [512][4] swizzled
    uint32_t* xchg = reinterpret_cast<uint32_t*>(gs + IKNP_STAGE_BYTES);
    uint32_t* bbuf = reinterpret_cast<uint32_t*>(gs + IKNP_STAGE_BYTES + XCHG);   // [16] + claim
    volatile uint32_t* claim = bbuf + 16;
    const uint32_t bar_id = group + 1;

    // this thread's PRG: AES-128 keyed by BE(label), newPrg ot/iknp.go:622-630
    uint32_t rk[44];
    {
        const Label k = label_from_mem(__ldg((second ? p.k1 : p.k0) + col));
        aes128_expand_regs(al, k.w0, k.w1, k.w2, k.w3, rk);
    }
    bool flip = false;                                    // sender: Delta.Bit(col), ot/label.go:129-141
    if (!RECEIVER) {
        const uint4 d = __ldg(p.delta);                   // memory words (lo D0, hi D0, lo D1, hi D1)
        const uint32_t dw = col < 32 ? d.x : col < 64 ? d.y : col < 96 ? d.z : d.w;
        flip = (dw >> (col & 31u)) & 1u;
    }
    const uint64_t nchunks = (p.n + IKNP_CHUNK_ROWS - 1) / IKNP_CHUNK_ROWS;
    const uint32_t phase = (uint32_t)(p.stream_pos & 15u);

    for (;;) {
        if (gt == 0) *claim = atomicAdd(p.counter, 1u);
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GT) : "memory");
        const uint64_t c = *claim;
        if (c >= nchunks) break;
        const uint64_t row0 = c * IKNP_CHUNK_ROWS;
        const uint32_t rows = (uint32_t)((p.n - row0 < IKNP_CHUNK_ROWS) ? (p.n - row0) : IKNP_CHUNK_ROWS);
        const uint32_t w = (rows + 7) >> 3;               // byteRows of this chunk
        const size_t chunk_off = (size_t)c * (IKNP_CHUNK_ROWS / 8) * 128;   // all earlier chunks are full

        // ---- keystream: bytes [pos + 64c, +64) of this column's CTR stream
        const uint64_t off = p.stream_pos + 64 * c;
        const uint64_t j0 = off >> 4;
        uint32_t W[20];
#pragma unroll
        for (int b = 0; b < 5; b++) {
            if (b == 4 && phase == 0) { W[16] = W[17] = W[18] = W[19] = 0; break; }
            const uint64_t j = j0 + b;                    // 128-bit big-endian counter, iv = 0
            uint32_t s0 = 0, s1 = 0, s2 = (uint32_t)(j >> 32), s3 = (uint32_t)j;
            aes128_encrypt_regs(al, rk, s0, s1, s2, s3);
            W[4 * b] = bswap32(s0); W[4 * b + 1] = bswap32(s1);          // little-endian words of the
            W[4 * b + 2] = bswap32(s2); W[4 * b + 3] = bswap32(s3);      // keystream byte sequence
        }
        uint32_t T[16];
        const uint32_t sh = (phase & 3u) * 8u;
        switch (phase >> 2) {
            case 0: ks_extract<0>(W, sh, T); break;
            case 1: ks_extract<1>(W, sh, T); break;
            case 2: ks_extract<2>(W, sh, T); break;
            default: ks_extract<3>(W, sh, T); break;
        }
        if (w < 64) {                                     // prg() produced only w bytes
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const uint32_t lo = 4u * q;
                T[q] = (lo >= w) ? 0u : (lo + 4 <= w) ? T[q] : (T[q] & (0xffffffffu >> (8 * (lo + 4 - w))));
            }
        }

        if (RECEIVER) {
            if (BITS) {
                // ReceiveBits XORs the packed choice words in whole 64-bit units: words = byteRows / 8
                // (iknp.go:583-593), so the tail bytes of a short last chunk get no choice bits
                if (gt < 16) bbuf[gt] = (gt >> 1) < (w >> 3) ? __ldg(p.choice_bits + (row0 >> 5) + gt) : 0u;
            } else {
                // choice bits of the chunk packed LSB-first, iknp.go:472-477
                for (uint32_t r = gt; r < IKNP_CHUNK_ROWS; r += GT) {
                    const bool bit = (r < rows) && (__ldg(p.choice + row0 + r) != 0);
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                    if (lane == 0) bbuf[r >> 5] = bal;
                }
            }
            if (second) {
#pragma unroll
                for (int q = 0; q < 16; q++) xchg[q * 128 + col] = T[q];
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GT) : "memory");
            if (!second) {
                // U_i = T0_i ^ T1_i ^ b, iknp.go:491-497
                uint32_t U[16];
#pragma unroll
                for (int q = 0; q < 16; q++) U[q] = T[q] ^ xchg[q * 128 + col] ^ bbuf[q];
                uint8_t* dst = p.u_out + chunk_off + (size_t)col * w;
                if (w == 64) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                    for (int q = 0; q < 4; q++) d4[q] = make_uint4(U[4 * q], U[4 * q + 1], U[4 * q + 2], U[4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        for (uint32_t b = 0; b < 4; b++)
                            if (4u * q + b < w) dst[4 * q + b] = (uint8_t)(U[q] >> (8 * b));
                }
            }
        } else if (flip) {
            // t_i ^= chunk_i when Delta.Bit(i), iknp.go:215-218
            const uint8_t* src = p.u_in + chunk_off + (size_t)col * w;
            if (w == 64) {
                const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 v = __ldg(s4 + q);
                    T[4 * q] ^= v.x; T[4 * q + 1] ^= v.y; T[4 * q + 2] ^= v.z; T[4 * q + 3] ^= v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    uint32_t v = 0;
                    for (uint32_t b = 0; b < 4; b++)
                        if (4u * q + b < w) v |= (uint32_t)__ldg(src + 4 * q + b) << (8 * b);
                    T[q] ^= v;
                }
            }
        }

        if (BITS) {
            // Bit(0) of label r = bit r of column 0 (T0 for the receiver, q for the sender)
            if (!second && col == 0) {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const uint32_t lo = 32u * q;
                    if (lo < rows) p.result_bits[(row0 >> 5) + q] = (lo + 32 <= rows) ? T[q] : (T[q] & (0xffffffffu >> (lo + 32 - rows)));
                }
            }
            continue;                                      // the claim barrier orders the shared buffers
        }
        // ---- createLabels, iknp.go:647-683: label r = row r of the bit matrix
        if (!second) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const uint32_t x = warp_transpose32(T[q], lane);
                stage[stage_word(32u * q + lane, wcol)] = x;
            }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GT) : "memory");
        for (uint32_t r = gt; r < rows; r += GT) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 4);
            const uint32_t z = (r >> 3) & 3u;             // undo the swizzle: word k sits at k ^ z
            uint4 o;
            o.x = z == 0 ? v.x : z == 1 ? v.y : z == 2 ? v.z : v.w;
            o.y = z == 0 ? v.y : z == 1 ? v.x : z == 2 ? v.w : v.z;
            o.z = z == 0 ? v.z : z == 1 ? v.w : z == 2 ? v.x : v.y;
            o.w = z == 0 ? v.w : z == 1 ? v.z : z == 2 ? v.y : v.x;
            p.labels[row0 + r] = o;
        }
        // the next iteration's claim barrier orders these reads before the next writes
    }
}

// ------------------------------------------------- COT / ROT post-processing (K7) --
// Replaces the per-batch loops of COT.Send / COT.Receive (ot/cot.go:157-181, :201-233) and
// ROT.Send / ROT.Receive (ot/rot.go:155-172, :192-197): OT number j is hashed under MiTCCRH key
// j (NewMITCCRH(seed, 8) hands out keys gid 0, 1, 2, ... eight per renewal; Hash(pad, 8, h) lets
// key i encrypt blocks i*h .. i*h+h-1, ot/mitccrh.go:70-128), so the batches of eight are only
// buffer management and every OT is independent: one thread per OT, one key schedule in
// registers, one (receiver) or two (sender) blocks, the pad XORed with the payload on the way out.
// The labels arrive straight from the IKNP kernel's output in HBM -- no PCIe round trip between
// extension and hashing.
enum : int { COT_SEND = 0, COT_RECEIVE = 1, ROT_SEND = 2, ROT_RECEIVE = 3 };
struct CotParams {
    uint64_t seed_d0, seed_d1;
    uint64_t delta_d0, delta_d1;
    const uint4* data;        // sender: q_j; receiver: t_j
    const uint4* wires;       // COT_SEND: {L0, L1} per OT
    const uint8_t* flags;     // COT_RECEIVE: choice per OT (Go []bool)
    const uint4* msgs_in;     // COT_RECEIVE: the 2n labels taken off the wire
    uint4* out;               // COT_SEND: 2n messages; COT_RECEIVE / ROT_RECEIVE: n labels; ROT_SEND: n wires {L0, L1}
    uint64_t n;
    uint64_t first;           // number of the first OT of this launch (a row-range shard): OT i uses MiTCCRH key first + i
    uint32_t wire_bytes;      // 1: messages are in the SendLabel byte encoding (BE64(D0) || BE64(D1), ot/label.go:105-108)
};
// Label <-> the 16 bytes SendLabel / ReceiveLabel move (ot/io.go, ot/label.go:105-114)
__device__ __forceinline__ uint4 label_to_wire(Label l) { return make_uint4(bswap32(l.w0), bswap32(l.w1), bswap32(l.w2), bswap32(l.w3)); }
__device__ __forceinline__ Label label_from_wire(uint4 m) { return Label{bswap32(m.x), bswap32(m.y), bswap32(m.z), bswap32(m.w)}; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) cot_kernel(const CotParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    aes_tables_to_smem(smem);
    __syncthreads();
    const AesLane lane = aes_lane(smem);
    const Label delta = Label{(uint32_t)(p.delta_d0 >> 32), (uint32_t)p.delta_d0, (uint32_t)(p.delta_d1 >> 32), (uint32_t)p.delta_d1};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const uint64_t d0 = p.seed_d0 ^ (p.first + i);     // key j = BE(seed ^ {D0: j}), mitccrh.go:72-86
        uint32_t rk[44];
        aes128_expand_regs(lane, (uint32_t)(d0 >> 32), (uint32_t)d0, (uint32_t)(p.seed_d1 >> 32), (uint32_t)p.seed_d1, rk);
        const Label x = label_from_mem(__ldg(p.data + i));
        uint32_t s0 = x.w0, s1 = x.w1, s2 = x.w2, s3 = x.w3;
        aes128_encrypt_regs(lane, rk, s0, s1, s2, s3);
        const Label h0 = Label{s0 ^ x.w0, s1 ^ x.w1, s2 ^ x.w2, s3 ^ x.w3};
        if (MODE == ROT_RECEIVE) {                         // rot.go:192-197
            p.out[i] = label_to_mem(h0);
        } else if (MODE == COT_RECEIVE) {                  // cot.go:213-231
            const uint4 m = __ldg(p.msgs_in + 2 * i + (p.flags[i] ? 1 : 0));
            const Label r = p.wire_bytes ? label_from_wire(m) : label_from_mem(m);
            p.out[i] = label_to_mem(r ^ h0);
        } else {
            const Label y = x ^ delta;                     // cot.go:163-165, rot.go:161-163
            s0 = y.w0; s1 = y.w1; s2 = y.w2; s3 = y.w3;
            aes128_encrypt_regs(lane, rk, s0, s1, s2, s3);
            const Label h1 = Label{s0 ^ y.w0, s1 ^ y.w1, s2 ^ y.w2, s3 ^ y.w3};
            if (MODE == COT_SEND) {                        // cot.go:168-177
                const Label m0 = h0 ^ label_from_mem(__ldg(p.wires + 2 * i));
                const Label m1 = h1 ^ label_from_mem(__ldg(p.wires + 2 * i + 1));
                p.out[2 * i] = p.wire_bytes ? label_to_wire(m0) : label_to_mem(m0);
                p.out[2 * i + 1] = p.wire_bytes ? label_to_wire(m1) : label_to_mem(m1);
            } else {                                       // rot.go:166-169
                p.out[2 * i] = label_to_mem(h0);
                p.out[2 * i + 1] = label_to_mem(h1);
            }
        }
    }
}

// ------------------------------------- malicious-mode consistency sums of IKNP (K8) --
// Replaces the chi loops of IKNPSender.Send (ot/iknp.go:150-173) and IKNPReceiver.Receive
// (:408-451) with vectorInnPrdtSumNoRed / mul128 (ot/gf128.go:14-27, ot/mul128_generic.go,
// ot/mul128_amd64.s): chi_i is block chi_start + i of the AES-128-CTR stream keyed by seed2
// (prgLabels :639-645), the sums are (lo, hi) ^= chi_i (x) l_i as 256-bit carry-less products and,
// for the receiver, x ^= chi_i where the choice bit is set.  There is no carry-less multiply on
// the GPU: a 32x32 product is 16 integer multiplies of operands with every fourth bit kept (the
// partial sums of a bit position never exceed 8, so they cannot carry into the next kept bit),
// which run on the FMA pipe beside the AES lookups.
struct CheckParams {
    uint64_t seed_d0, seed_d1; // seed2; the PRG key is BE(seed2) (newPrg, iknp.go:622-631)
    uint64_t chi_start;
    const uint4* labels;
    const uint8_t* choice;    // nullable
    uint64_t n;
    uint32_t* acc;            // 12 words: lo (4), hi (4), x (4) as {D0 lo, D0 hi, D1 lo, D1 hi}; XOR-accumulated
};
__device__ __forceinline__ uint64_t clmul32(uint32_t x, uint32_t y) {
    const uint32_t x0 = x & 0x11111111u, x1 = x & 0x22222222u, x2 = x & 0x44444444u, x3 = x & 0x88888888u;
    const uint32_t y0 = y & 0x11111111u, y1 = y & 0x22222222u, y2 = y & 0x44444444u, y3 = y & 0x88888888u;
    auto m = [](uint32_t a, uint32_t b) { return (uint64_t)a * b; };
    const uint64_t z0 = m(x0, y0) ^ m(x1, y3) ^ m(x2, y2) ^ m(x3, y1);
    const uint64_t z1 = m(x0, y1) ^ m(x1, y0) ^ m(x2, y3) ^ m(x3, y2);
    const uint64_t z2 = m(x0, y2) ^ m(x1, y1) ^ m(x2, y0) ^ m(x3, y3);
    const uint64_t z3 = m(x0, y3) ^ m(x1, y2) ^ m(x2, y1) ^ m(x3, y0);
    return (z0 & 0x1111111111111111ull) | (z1 & 0x2222222222222222ull) | (z2 & 0x4444444444444444ull) |
           (z3 & 0x8888888888888888ull);
}
// a, b: four 32-bit limbs, limb 0 = coefficients x^0..x^31 (Label.Bit numbering: D0 low word first);
// r: eight limbs of the 256-bit product, XORed in.
__device__ __forceinline__ void clmul128_acc(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t z = clmul32(a[i], b[j]);
            r[i + j] ^= (uint32_t)z;
            r[i + j + 1] ^= (uint32_t)(z >> 32);
        }
}
__global__ void __launch_bounds__(512, 1) iknp_check_kernel(const CheckParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* const smem = aes_align_tables(smem_raw);
    uint32_t* rk = reinterpret_cast<uint32_t*>(smem + AES_TABLE_BYTES);
    uint32_t* red = rk + 64;                               // 12 words of block-level accumulator
    aes_tables_to_smem(smem);
    if (threadIdx.x < 12) red[threadIdx.x] = 0;
    __syncthreads();
    const AesLane lane = aes_lane(smem);
    if (threadIdx.x == 0) {
        uint8_t* kb = reinterpret_cast<uint8_t*>(red + 16);
        for (int b = 0; b < 8; b++) { kb[b] = (uint8_t)(p.seed_d0 >> (56 - 8 * b)); kb[8 + b] = (uint8_t)(p.seed_d1 >> (56 - 8 * b)); }
        aes_expand_key(lane, kb, 16, rk);
    }
    __syncthreads();
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xs[4] = {0, 0, 0, 0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const uint64_t ctr = p.chi_start + i;              // 128-bit big-endian counter, iv = 0 (newPrg :622-631)
        uint32_t s0 = 0, s1 = 0, s2 = (uint32_t)(ctr >> 32), s3 = (uint32_t)ctr;
        aes_encrypt_smem<10>(lane, rk, s0, s1, s2, s3);
        // SetBytes (label.go:111-114): D0 = (s0, s1), D1 = (s2, s3); limbs in Label.Bit order
        const uint32_t chi[4] = {s1, s0, s3, s2};
        const uint4 lm = __ldg(p.labels + i);              // Go memory order: (lo D0, hi D0, lo D1, hi D1)
        const uint32_t l[4] = {lm.x, lm.y, lm.z, lm.w};
        clmul128_acc(chi, l, acc);
        if (p.choice && p.choice[i]) { xs[0] ^= chi[0]; xs[1] ^= chi[1]; xs[2] ^= chi[2]; xs[3] ^= chi[3]; }
    }
    // warp XOR-reduction, then one shared and one global atomic per word
#pragma unroll
    for (int w = 0; w < 12; w++) {
        uint32_t v = w < 8 ? acc[w] : xs[w - 8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicXor(red + w, v);
    }
    __syncthreads();
    if (threadIdx.x < 12 && red[threadIdx.x]) atomicXor(p.acc + threadIdx.x, red[threadIdx.x]);
}

}  // namespace gcb
