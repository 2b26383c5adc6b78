// gc_kernels.cuh -- batch garble / eval kernels (K1, K2, and the gate loop of K3)
// for sm_100a.
//
// Replaces the gate loops of Circuit.Garble (circuit/garble.go:285-299 with
// Gate.garbleInto :311-482), Circuit.Eval (circuit/eval.go:28-112) and
// Streaming.Garble (circuit/stream_garble.go:179-190 with garbleGate :195-449).
//
// Execution model.  One persistent CTA per SM.  The CTA is split into TEAMS; a
// team owns one circuit instance at a time and keeps that instance's live wire
// labels in its slice of shared memory (slots assigned by the plan compiler), so
// intermediate labels never touch HBM: per instance HBM sees the input labels
// once, the garbled rows once (written by the garbler, read by the evaluator)
// and the output labels once.  A team walks the plan's dependency steps with a
// team-local barrier between steps (a __syncwarp for one-warp teams).  Inside a
// step every AES block is one task: an AND gate is a quad of tasks hashing
// (a0,j0) (a1,j0) (b0,j1) (b1,j1) on four adjacent lanes and combining with warp
// shuffles, INV a pair, OR a quad; XOR/XNOR gates are one lane and no AES
// (Free-XOR).  A thread runs up to ILP tasks at once (their AES rounds
// interleaved) so that a few warps per SM keep the shared-memory pipe busy, and
// it loads the gate records of the NEXT step before it starts on the current
// one, which takes the global-memory latency off the step-to-step chain.  All
// teams share the 128 KiB of replicated AES T-tables (aes_core.cuh).
#pragma once
#include "aes_core.cuh"
#include "plan.hpp"

namespace gcb {

constexpr int GC_MAX_TEAMS = 32;          // 32 one-warp teams, or <= 16 named-barrier teams
constexpr int GC_RK_BYTES = 256;          // 60 round-key words, padded

struct GcParams {
    const uint4* recs;                    // GateRec[]
    const uint4* steps;                   // StepRec[] followed by two zero records
    const uint32_t* out_wire;
    const uint2* live_in;                 // SlotRef[]
    const uint2* live_out;
    uint32_t n_steps, n_in, n_out, n_slots, n_rows, n_wires;
    const uint8_t* keys;
    uint32_t keylen, key_stride;
    uint32_t batch;
    const uint4* r;                       // garble: raw R per instance
    const uint4* in_labels;               // garble: L0 per input; eval: active input labels
    uint4* tables;                        // garble: out; eval: in
    uint4* io;                            // garble: io_wires (2 labels per wire); eval: out labels
    uint4* wires_full;                    // optional
    uint32_t* counter;                    // next instance to claim
    uint32_t team_threads, n_teams;
    // streaming mode (GC_STREAM): live-in / live-out labels come from and go to
    // the permanent wire file instead of in_labels / io
    const uint32_t* in_ids;               // [n_in] permanent wire id of live-in k
    const uint32_t* out_ids;              // [n_out] permanent wire id of live-out k
    uint4* const* pages;                  // wire-file page table
};

enum : int { GC_PLAIN = 0, GC_FULL = 1, GC_STREAM = 2 };

// Permanent wire file of the streaming garbler (circuit/stream_garble.go:78-114:
// wires[w>>16][w&0xffff] pages of ot.Wire).  Here a page covers WF_PAGE_IDS ids
// and holds the L0 label of each id for every instance, [instance][id]; L1 is
// always L0 ^ R.
constexpr uint32_t WF_PAGE_SHIFT = 12;
constexpr uint32_t WF_PAGE_IDS = 1u << WF_PAGE_SHIFT;
__device__ __forceinline__ uint4* wf_slot(uint4* const* pages, uint32_t id, uint32_t inst) {
    return pages[id >> WF_PAGE_SHIFT] + ((size_t)inst << WF_PAGE_SHIFT) + (id & (WF_PAGE_IDS - 1));
}

__device__ __forceinline__ void team_barrier(uint32_t team, uint32_t team_threads) {
    if (team_threads == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team), "r"(team_threads) : "memory");
}

__device__ __forceinline__ Label lds_label(const uint4* slots, uint32_t s) { return label_from_mem(slots[s]); }
__device__ __forceinline__ void sts_label(uint4* slots, uint32_t s, Label l) { slots[s] = label_to_mem(l); }

__device__ __forceinline__ Label shfl_xor_label(Label h, int m) {
    return Label{__shfl_xor_sync(0xffffffffu, h.w0, m), __shfl_xor_sync(0xffffffffu, h.w1, m),
                 __shfl_xor_sync(0xffffffffu, h.w2, m), __shfl_xor_sync(0xffffffffu, h.w3, m)};
}
__device__ __forceinline__ Label shfl_label(Label h, int src) {
    return Label{__shfl_sync(0xffffffffu, h.w0, src), __shfl_sync(0xffffffffu, h.w1, src),
                 __shfl_sync(0xffffffffu, h.w2, src), __shfl_sync(0xffffffffu, h.w3, src)};
}
__device__ __forceinline__ uint32_t mask_of(uint32_t bit) { return 0u - (bit & 1u); }

// H(K) = AES(K) ^ K for UU independent blocks, rounds interleaved.
template <int NR, int UU>
__device__ __forceinline__ void aes_hash_multi(const AesLane& a, const uint32_t* __restrict__ rk, const Label (&K)[UU],
                                               Label (&H)[UU]) {
    const uint4* k4 = reinterpret_cast<const uint4*>(rk);
    uint32_t s[UU][4];
    {
        const uint4 k = k4[0];
#pragma unroll
        for (int j = 0; j < UU; j++) {
            s[j][0] = K[j].w0 ^ k.x; s[j][1] = K[j].w1 ^ k.y; s[j][2] = K[j].w2 ^ k.z; s[j][3] = K[j].w3 ^ k.w;
        }
    }
    if (UU == 1) {
#pragma unroll
        for (int r = 1; r < NR; r++) aes_round(a, s[0][0], s[0][1], s[0][2], s[0][3], k4[r]);
    } else {
#pragma unroll 1
        for (int r = 1; r < NR; r++) {
            const uint4 k = k4[r];
#pragma unroll
            for (int j = 0; j < UU; j++) aes_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
        }
    }
    {
        const uint4 k = k4[NR];
#pragma unroll
        for (int j = 0; j < UU; j++) {
            aes_last_round(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
            H[j] = Label{s[j][0] ^ K[j].w0, s[j][1] ^ K[j].w1, s[j][2] ^ K[j].w2, s[j][3] ^ K[j].w3};
        }
    }
}

// Shared-memory carve-up common to both kernels.
struct TeamCtx {
    uint8_t* tables;
    uint32_t* rk;          // this team's round keys
    uint4* slots;          // this team's wire labels
    volatile uint32_t* claim;
    uint32_t team, ttid;
};

__device__ __forceinline__ TeamCtx team_ctx(uint8_t* smem, const GcParams& p) {
    TeamCtx c;
    c.tables = smem;
    c.team = threadIdx.x / p.team_threads;
    c.ttid = threadIdx.x - c.team * p.team_threads;
    uint8_t* q = smem + AES_TABLE_BYTES;
    c.rk = reinterpret_cast<uint32_t*>(q + c.team * GC_RK_BYTES);
    q += p.n_teams * GC_RK_BYTES;
    c.slots = reinterpret_cast<uint4*>(q + (size_t)c.team * p.n_slots * 16);
    q += (size_t)p.n_teams * p.n_slots * 16;
    c.claim = reinterpret_cast<volatile uint32_t*>(q) + c.team;
    return c;
}

// Gate index of cipher task t of a step.  Garble: 4 tasks per AND/OR, 2 per INV;
// eval: 2 per AND/OR (OR uses one), 1 per INV.
template <bool GARBLE>
__device__ __forceinline__ uint32_t task_gate(const uint4& st, uint32_t t, uint32_t& k) {
    constexpr uint32_t QS = GARBLE ? 2 : 1;           // log2 tasks per AND/OR
    constexpr uint32_t IS = GARBLE ? 1 : 0;           // log2 tasks per INV
    const uint32_t nq = st.z << QS, cbase = st.x + st.y;
    if (t < nq) { k = t & ((1u << QS) - 1); return cbase + (t >> QS); }
    const uint32_t u = t - nq;
    k = u & ((1u << IS) - 1);
    return cbase + st.z + (u >> IS);
}
template <bool GARBLE>
__device__ __forceinline__ uint32_t task_count(const uint4& st) {
    return GARBLE ? 4 * st.z + 2 * st.w : 2 * st.z + st.w;
}

// Records of the first pass of a step, loaded one step ahead.
template <bool GARBLE, int ILP>
__device__ __forceinline__ void prefetch_step(const GcParams& p, const uint4& st, uint32_t ttid, uint32_t TT,
                                              uint4& free_rec, uint4 (&cipher_rec)[ILP]) {
    free_rec = make_uint4(0, 0, 0, 0);
    if (ttid < st.y) free_rec = __ldg(p.recs + st.x + ttid);
    const uint32_t ntask = task_count<GARBLE>(st);
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        uint32_t k;
        const uint32_t t = j * TT + ttid;
        cipher_rec[j] = make_uint4(0, 0, 0, 0);
        if (t < ntask) cipher_rec[j] = __ldg(p.recs + task_gate<GARBLE>(st, t, k));
    }
}

// ------------------------------------------------------------------ garble ----
struct GarbleEnv {
    GcParams const* p;
    uint4* slots;
    const uint32_t* rk;
    uint4* tab;
    Label R;
    uint32_t inst;
};

// One pass of UU cipher tasks per thread: tasks base + j*TT + ttid.
template <int NR, int MODE, int UU, int ILP>
__device__ __forceinline__ void garble_pass(const AesLane& lane, const GarbleEnv& e, const uint4& st, uint32_t ntask,
                                            uint32_t base, uint32_t ttid, uint32_t TT, const uint4 (&pre)[ILP]) {
    static_assert(UU <= ILP, "pass wider than the prefetch");
    constexpr bool FULL = MODE == GC_FULL;
    const GcParams& p = *e.p;
    uint4* const slots = e.slots;
    const Label R = e.R;
    const bool first = base == 0;
    uint4 g[UU];
    Label a0[UU], K[UU], H[UU];
    uint32_t op[UU], kk[UU], pp[UU], gidx[UU];
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const uint32_t t = base + j * TT + ttid;
        const bool active = t < ntask;
        uint32_t k;
        const uint32_t gi = task_gate<true>(st, t, k);
        g[j] = pre[j];
        if (!first) { g[j] = make_uint4(0, 0, 0, 0); if (active) g[j] = __ldg(p.recs + gi); }
        const uint32_t sa = g[j].x & 0xffff, sb = g[j].x >> 16;
        op[j] = active ? ((g[j].y >> 16) & 0xff) : 0xffu;
        kk[j] = k; gidx[j] = gi;
        a0[j] = lds_label(slots, sa);
        const Label b0 = lds_label(slots, sb);
        const uint32_t pa = label_s(a0[j]), pb = label_s(b0);
        pp[j] = 2 * pa + pb;
        uint32_t tw = g[j].z;
        if (op[j] == OP_OR) {                              // K = 2a ^ 4b ^ t, task k = (a_i, b_j), k = 2i+j
            const Label xa = a0[j] ^ label_and_mask(R, mask_of(k >> 1));
            const Label xb = b0 ^ label_and_mask(R, mask_of(k));
            K[j] = label_shl(xa, 1) ^ label_shl(xb, 2);
        } else {                                           // AND: (a0,j0)(a1,j0)(b0,j1)(b1,j1); INV: a0, a1
            const bool use_b = (op[j] == OP_AND) && (k & 2);
            Label x = use_b ? b0 : a0[j];
            x = x ^ label_and_mask(R, mask_of(k));
            K[j] = label_shl(x, 1);
            if (op[j] == OP_AND) tw += k >> 1;
        }
        K[j].w3 ^= tw;
    }
    aes_hash_multi<NR, UU>(lane, e.rk, K, H);
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const Label h = H[j];
        const uint32_t k = kk[j], pa = pp[j] >> 1, pb = pp[j] & 1, sc = g[j].y & 0xffff;
        const Label u = h ^ shfl_xor_label(h, 1);
        Label v = Label{0, 0, 0, 0};
        if (op[j] == OP_AND) {
            if (k == 0) {                                  // generator half (garble.go:361-369)
                const Label tg = u ^ label_and_mask(R, mask_of(pb));
                v = h ^ label_and_mask(tg, mask_of(pa));
                e.tab[g[j].w] = label_to_mem(tg);
            } else if (k == 2) {                           // evaluator half (garble.go:372-380)
                const Label te = u ^ a0[j];
                v = h ^ label_and_mask(u, mask_of(pb));
                e.tab[g[j].w + 1] = label_to_mem(te);
            }
        } else if (op[j] == OP_INV) {                      // garble.go:446-474, row-reduced
            if (k == 0) {
                const Label c0 = pa ? (u ^ h) : (h ^ R);
                e.tab[g[j].w] = label_to_mem(u ^ R);
                sts_label(slots, sc, c0);
                if (FULL) {
                    uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.out_wire + gidx[j])) * 2;
                    w[0] = label_to_mem(c0);
                    w[1] = label_to_mem(c0 ^ R);
                }
            }
        }
        if (__any_sync(0xffffffffu, op[j] == OP_OR)) {     // garble.go:412-444, row-reduced
            const uint32_t l0i = pp[j];
            const Label t0 = shfl_label(h, (int)((threadIdx.x & 28u) + l0i));
            if (op[j] == OP_OR) {
                const Label c0 = (l0i == 0) ? t0 : (t0 ^ R);
                const Label c1 = c0 ^ R;
                const uint32_t pos = k ^ l0i;              // table position of this task's row
                if (pos != 0) e.tab[g[j].w + pos - 1] = label_to_mem(h ^ ((pos == l0i) ? c0 : c1));
                if (k == 0) {
                    sts_label(slots, sc, c0);
                    if (FULL) {
                        uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.out_wire + gidx[j])) * 2;
                        w[0] = label_to_mem(c0);
                        w[1] = label_to_mem(c1);
                    }
                }
            }
        }
        const Label w2 = v ^ shfl_xor_label(v, 2);
        if (op[j] == OP_AND && k == 0) {                   // combine halves (garble.go:383-392)
            sts_label(slots, sc, w2);
            if (FULL) {
                uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.out_wire + gidx[j])) * 2;
                w[0] = label_to_mem(w2);
                w[1] = label_to_mem(w2 ^ R);
            }
        }
    }
}

template <int NR, int MODE, int ILP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) garble_kernel(const GcParams p) {
    constexpr bool FULL = MODE == GC_FULL;
    constexpr bool STREAM = MODE == GC_STREAM;
    extern __shared__ __align__(16) uint8_t smem[];
    const TeamCtx tc = team_ctx(smem, p);
    aes_tables_to_smem(tc.tables);
    __syncthreads();
    const AesLane lane = aes_lane(tc.tables);
    const uint32_t TT = p.team_threads, ttid = tc.ttid;
    if (p.key_stride == 0) {
        if (ttid == 0) aes_expand_key(lane, p.keys, (int)p.keylen, tc.rk);
        team_barrier(tc.team, TT);
    }
    uint4* const slots = tc.slots;

    for (;;) {
        if (ttid == 0) *tc.claim = atomicAdd(p.counter, 1u);
        team_barrier(tc.team, TT);
        const uint32_t inst = *tc.claim;
        if (inst >= p.batch) break;
        if (p.key_stride != 0 && ttid == 0)
            aes_expand_key(lane, p.keys + (size_t)inst * p.key_stride, (int)p.keylen, tc.rk);
        Label R = label_from_mem(__ldg(p.r + inst));
        R.w0 |= 0x80000000u;                                   // r.SetS(true), garble.go:258
        // the records of steps 0 and 1 load while the inputs do
        uint4 st = __ldg(p.steps), st_n = __ldg(p.steps + 1);
        uint4 cf, cc[ILP];
        prefetch_step<true, ILP>(p, st, ttid, TT, cf, cc);
        // input wires: L0 from the caller's reader bytes, L1 = L0 ^ R (garble.go:271-278)
        for (uint32_t k = ttid; k < p.n_in; k += TT) {
            const uint2 ref = __ldg(p.live_in + k);
            uint4 m;
            if (STREAM) m = *wf_slot(p.pages, __ldg(p.in_ids + ref.y), inst);
            else m = __ldg(p.in_labels + (size_t)inst * p.n_in + ref.y);
            slots[ref.x] = m;
            if (!STREAM && p.io) {
                uint4* w = p.io + ((size_t)inst * (p.n_in + p.n_out) + ref.y) * 2;
                w[0] = m;
                w[1] = label_to_mem(label_from_mem(m) ^ R);
            }
            if (FULL) {
                uint4* w = p.wires_full + ((size_t)inst * p.n_wires + ref.y) * 2;
                w[0] = m;
                w[1] = label_to_mem(label_from_mem(m) ^ R);
            }
        }
        team_barrier(tc.team, TT);
        const GarbleEnv env{&p, slots, tc.rk, p.tables + (size_t)inst * p.n_rows, R, inst};

        for (uint32_t s = 0; s < p.n_steps; s++) {
            // ---- records of the next step, in flight while this one computes
            const uint4 st_nn = __ldg(p.steps + s + 2);
            uint4 nf, nc[ILP];
            prefetch_step<true, ILP>(p, st_n, ttid, TT, nf, nc);
            // ---- Free-XOR gates (garble.go:331-351): one lane each, no AES
            for (uint32_t j = ttid; j < st.y; j += TT) {
                uint4 g = cf;
                if (j != ttid) g = __ldg(p.recs + st.x + j);
                const uint32_t sa = g.x & 0xffff, sb = g.x >> 16, sc = g.y & 0xffff, op = (g.y >> 16) & 0xff;
                Label c0 = lds_label(slots, sa) ^ lds_label(slots, sb);
                if (op == OP_XNOR) c0 = c0 ^ R;                // XNOR swaps (L0, L1)
                sts_label(slots, sc, c0);
                if (FULL) {
                    uint4* w = p.wires_full + ((size_t)inst * p.n_wires + __ldg(p.out_wire + st.x + j)) * 2;
                    w[0] = label_to_mem(c0);
                    w[1] = label_to_mem(c0 ^ R);
                }
            }
            // ---- ciphered gates: one AES block per task, up to ILP tasks per thread at once
            const uint32_t ntask = task_count<true>(st);
            for (uint32_t base = 0; base < ntask; base += TT * ILP) {
                const uint32_t left = ntask - base;
                if (ILP >= 4 && left > 2 * TT)
                    garble_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP>(lane, env, st, ntask, base, ttid, TT, cc);
                else if (ILP >= 2 && left > TT)
                    garble_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP>(lane, env, st, ntask, base, ttid, TT, cc);
                else if (base + (ttid & ~31u) < ntask)
                    garble_pass<NR, MODE, 1, ILP>(lane, env, st, ntask, base, ttid, TT, cc);
            }
            team_barrier(tc.team, TT);
            st = st_n; st_n = st_nn; cf = nf;
#pragma unroll
            for (int j = 0; j < ILP; j++) cc[j] = nc[j];
        }
        // output wires (what circuit/garbler.go:153 and sha2pc/garbler.go:125 read)
        if (STREAM) {                                          // Streaming.Set, stream_garble.go:144-157
            for (uint32_t k = ttid; k < p.n_out; k += TT) {
                const uint2 ref = __ldg(p.live_out + k);
                *wf_slot(p.pages, __ldg(p.out_ids + ref.y), inst) = slots[ref.x];
            }
        } else if (p.io) {
            for (uint32_t k = ttid; k < p.n_out; k += TT) {
                const uint2 ref = __ldg(p.live_out + k);
                const Label l0 = lds_label(slots, ref.x);
                uint4* w = p.io + ((size_t)inst * (p.n_in + p.n_out) + p.n_in + ref.y) * 2;
                w[0] = label_to_mem(l0);
                w[1] = label_to_mem(l0 ^ R);
            }
        }
        team_barrier(tc.team, TT);
    }
}

// -------------------------------------------------------------------- eval ----
struct EvalEnv {
    GcParams const* p;
    uint4* slots;
    const uint32_t* rk;
    const uint4* tab;
    uint32_t inst;
};

template <int NR, int MODE, int UU, int ILP>
__device__ __forceinline__ void eval_pass(const AesLane& lane, const EvalEnv& e, const uint4& st, uint32_t ntask,
                                          uint32_t base, uint32_t ttid, uint32_t TT, const uint4 (&pre)[ILP]) {
    static_assert(UU <= ILP, "pass wider than the prefetch");
    constexpr bool FULL = MODE == GC_FULL;
    const GcParams& p = *e.p;
    uint4* const slots = e.slots;
    const bool first = base == 0;
    uint4 g[UU];
    Label K[UU], H[UU], row[UU];
    uint32_t op[UU], kk[UU], gidx[UU];
    bool act[UU];
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const uint32_t t = base + j * TT + ttid;
        act[j] = t < ntask;
        uint32_t k;
        const uint32_t gi = task_gate<false>(st, t, k);
        g[j] = pre[j];
        if (!first) { g[j] = make_uint4(0, 0, 0, 0); if (act[j]) g[j] = __ldg(p.recs + gi); }
        const uint32_t sa = g[j].x & 0xffff, sb = g[j].x >> 16;
        op[j] = act[j] ? ((g[j].y >> 16) & 0xff) : 0xffu;
        kk[j] = k; gidx[j] = gi;
        const Label a = lds_label(slots, sa);
        const Label b = lds_label(slots, sb);
        const uint32_t sA = label_s(a), sB = label_s(b);
        // the garbled row this task may need, fetched before the AES so the HBM
        // latency hides behind the rounds
        uint32_t ridx = g[j].w;
        bool need = false;
        if (op[j] == OP_AND) { ridx += k; need = k ? sB : sA; }
        else if (op[j] == OP_INV) { need = sA; }
        else if (op[j] == OP_OR) { const uint32_t ix = 2 * sA + sB; need = (ix > 0) && (k == 0); ridx += ix - 1; }
        row[j] = Label{0, 0, 0, 0};
        if (need) row[j] = label_from_mem(__ldg(e.tab + ridx));
        uint32_t tw = g[j].z;
        if (op[j] == OP_OR) {
            K[j] = label_shl(a, 1) ^ label_shl(b, 2);         // makeK, garble.go:75-83
        } else {
            K[j] = label_shl((op[j] == OP_AND && k) ? b : a, 1);
            if (op[j] == OP_AND) tw += k;
        }
        K[j].w3 ^= tw;
        // we ^= a when S(b) (eval.go:72-75): folded into the row term
        row[j] = row[j] ^ label_and_mask(a, mask_of((op[j] == OP_AND && k == 1) ? sB : 0u));
    }
    aes_hash_multi<NR, UU>(lane, e.rk, K, H);
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const Label v = H[j] ^ row[j];                         // decrypt (garble.go:58-73) / half-gate
        const Label o = v ^ shfl_xor_label(v, 1);
        if (act[j] && (kk[j] == 0)) {
            const Label res = (op[j] == OP_AND) ? o : v;
            sts_label(slots, g[j].y & 0xffff, res);
            if (FULL) p.wires_full[(size_t)e.inst * p.n_wires + __ldg(p.out_wire + gidx[j])] = label_to_mem(res);
        }
    }
}

template <int NR, int MODE, int ILP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) eval_kernel(const GcParams p) {
    constexpr bool FULL = MODE == GC_FULL;
    extern __shared__ __align__(16) uint8_t smem[];
    const TeamCtx tc = team_ctx(smem, p);
    aes_tables_to_smem(tc.tables);
    __syncthreads();
    const AesLane lane = aes_lane(tc.tables);
    const uint32_t TT = p.team_threads, ttid = tc.ttid;
    if (p.key_stride == 0) {
        if (ttid == 0) aes_expand_key(lane, p.keys, (int)p.keylen, tc.rk);
        team_barrier(tc.team, TT);
    }
    uint4* const slots = tc.slots;

    for (;;) {
        if (ttid == 0) *tc.claim = atomicAdd(p.counter, 1u);
        team_barrier(tc.team, TT);
        const uint32_t inst = *tc.claim;
        if (inst >= p.batch) break;
        if (p.key_stride != 0 && ttid == 0)
            aes_expand_key(lane, p.keys + (size_t)inst * p.key_stride, (int)p.keylen, tc.rk);
        uint4 st = __ldg(p.steps), st_n = __ldg(p.steps + 1);
        uint4 cf, cc[ILP];
        prefetch_step<false, ILP>(p, st, ttid, TT, cf, cc);
        for (uint32_t k = ttid; k < p.n_in; k += TT) {
            const uint2 ref = __ldg(p.live_in + k);
            const uint4 m = __ldg(p.in_labels + (size_t)inst * p.n_in + ref.y);
            slots[ref.x] = m;
            if (FULL) p.wires_full[(size_t)inst * p.n_wires + ref.y] = m;
        }
        team_barrier(tc.team, TT);
        const EvalEnv env{&p, slots, tc.rk, p.tables + (size_t)inst * p.n_rows, inst};

        for (uint32_t s = 0; s < p.n_steps; s++) {
            const uint4 st_nn = __ldg(p.steps + s + 2);
            uint4 nf, nc[ILP];
            prefetch_step<false, ILP>(p, st_n, ttid, TT, nf, nc);
            // XOR and XNOR are both a plain XOR on the evaluator side (eval.go:48-50)
            for (uint32_t j = ttid; j < st.y; j += TT) {
                uint4 g = cf;
                if (j != ttid) g = __ldg(p.recs + st.x + j);
                const uint32_t sa = g.x & 0xffff, sb = g.x >> 16, sc = g.y & 0xffff;
                const Label c = lds_label(slots, sa) ^ lds_label(slots, sb);
                sts_label(slots, sc, c);
                if (FULL) p.wires_full[(size_t)inst * p.n_wires + __ldg(p.out_wire + st.x + j)] = label_to_mem(c);
            }
            // ciphered gates: AND = 2 tasks (a with j0, b with j1); OR/INV = 1 hash.
            // OR shares the 2-task slot of its class (second task idle).
            const uint32_t ntask = task_count<false>(st);
            for (uint32_t base = 0; base < ntask; base += TT * ILP) {
                const uint32_t left = ntask - base;
                if (ILP >= 4 && left > 2 * TT)
                    eval_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP>(lane, env, st, ntask, base, ttid, TT, cc);
                else if (ILP >= 2 && left > TT)
                    eval_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP>(lane, env, st, ntask, base, ttid, TT, cc);
                else if (base + (ttid & ~31u) < ntask)
                    eval_pass<NR, MODE, 1, ILP>(lane, env, st, ntask, base, ttid, TT, cc);
            }
            team_barrier(tc.team, TT);
            st = st_n; st_n = st_nn; cf = nf;
#pragma unroll
            for (int j = 0; j < ILP; j++) cc[j] = nc[j];
        }
        for (uint32_t k = ttid; k < p.n_out; k += TT) {
            const uint2 ref = __ldg(p.live_out + k);
            p.io[(size_t)inst * p.n_out + ref.y] = slots[ref.x];
        }
        team_barrier(tc.team, TT);
    }
}

}  // namespace gcb
