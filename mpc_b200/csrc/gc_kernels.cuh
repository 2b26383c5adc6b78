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
// and the output labels once.
//
// A team walks the plan's PHASES.  A phase is (1) the free wires (XOR / XNOR
// results) that became computable after the previous cipher level and that
// something reads: the plan compiler has flattened the Free-XOR gates into
// NODES, each the XOR of up to 14 already existing labels, grouped into waves of
// mutually independent nodes (almost always one wave) -- one thread per node,
// one team barrier per wave.  The node records reach the kernel as ROWS of one
// record per team thread (plan.hpp); a thread streams its records a few
// rows ahead of the one it runs, across wave and phase boundaries, so the L2
// latency of the plan never sits on the dependency chain; and (2) one level of ciphered gates executed by the
// whole team: every AES block is a task; an AND gate is a quad of tasks hashing
// (a0,j0) (a1,j0) (b0,j1) (b1,j1) on four adjacent lanes and combining with warp
// shuffles, INV a pair, OR a quad.  A thread runs up to ILP tasks at once (AES
// rounds interleaved).  Node and gate records are loaded one phase ahead, so no
// global-memory latency sits on the phase-to-phase chain.  Teams start staggered
// so that their shared-memory-bound cipher levels and latency-bound node waves
// interleave.  All teams share the replicated AES T-tables (aes_core.cuh).
#pragma once
#include "aes_core.cuh"
#include "plan.hpp"

namespace gcb {

// 1: the single-block pass keeps its AES rounds in a loop (48 instructions) instead of unrolling them (430): with
// many one-warp teams per SM every warp is somewhere else in the kernel and the instruction caches, not the
// pipes, become the limit (ncu: no_inst 44 % of the stall samples with 16 teams of the unrolled form).
#ifndef GC_ROUNDS_UNROLL
#define GC_ROUNDS_UNROLL 1
#endif
#ifndef GC_ROLL_SINGLE
#define GC_ROLL_SINGLE 1
#endif
constexpr int GC_MAX_TEAMS = 32;          // 32 one-warp teams, or <= 16 named-barrier teams
constexpr int GC_RK_BYTES = 256;          // 60 round-key words, padded
// Node rows a thread keeps in flight ahead of the one it runs: the single-block 512-thread variant
// (deep, narrow circuits: short rows, nothing else to hide the L2 latency behind) has the registers
// for four.  The two-block variants sit at the 128-register cap: eval keeps two rows in flight, garble
// (four more words of state per block: R, the kept label) only one -- measured on aes_128 x 4096,
// a deeper pipeline costs more in register shuffling than it hides (garble 4.08 -> 4.00 ms).
__host__ __device__ constexpr uint32_t node_pipe(int ilp, int maxt, bool garble) {
    return (ilp == 1 && maxt == 512) ? 4u : (ilp == 2 && garble) ? 1u : 2u;
}

struct GcParams {
    const uint4* phases;                  // DevPhaseRec[] (two uint4 each) followed by two zero records
    const uint4* nodes;                   // NodeRec rows (two uint4 per record, team_threads records per row)
    const uint4* crecs;                   // GateRec[]
    const uint32_t* nout_wire;            // original output wire of nodes[i] / crecs[i] (GC_FULL)
    const uint32_t* cout_wire;
    const uint2* live_in;                 // SlotRef[]
    const uint2* live_out;
    uint32_t n_phases, n_in, n_out, n_slots, n_rows, n_wires;
    uint32_t n_smem;                      // wire slots held in shared memory (= n_slots unless the plan spills)
    uint4* spill;                         // spilling variant: [grid * n_teams][n_slots - n_smem] labels in global memory
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
    uint32_t twin;                        // one-warp teams run as lock-step pairs (team_ctx)
    const uint32_t* copies;               // SPILL == 2: evict / reload records, src | dst << 16 (DevPhaseRec::copy_first indexes it)
    uint32_t n_live_in;                   // entries of live_in (an input of a split plan may be loaded into a hot slot AND the scratch)
    uint32_t hdr_split;                   // shared-memory layout: team headers kept together (1) or in front of each label block (0)
    uint32_t stagger;                     // SM cycles by which consecutive teams start apart
    long long* trace;                     // optional: phase timestamps of block 0 / team 0 (tools/trace_phases.py)
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

// team: the team's barrier number, or -- for TWIN teams, see team_ctx -- 0x80000000 | the pair's barrier number.
__device__ __forceinline__ void team_barrier(uint32_t team, uint32_t team_threads) {
    if (team & 0x80000000u) asm volatile("bar.sync %0, 64;" ::"r"(team & 0xffu) : "memory");
    else if (team_threads == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team), "r"(team_threads) : "memory");
}

// Where a team's wire labels live.  Normally all of them are in the team's slice of shared memory.  A
// circuit that keeps more labels live than fit there (about 6,400 beside two tables) runs the SPILL
// variant: slots below n_sm stay in shared memory -- the plan hands out low slot numbers first, so they
// are the hot ones -- and the rest go to a per-team scratch in global memory (L2-resident; ordered
// by the same team barriers, and written and read by this SM only).
struct SlotsSmem {
    uint4* sm;
    __device__ __forceinline__ uint4 ldm(uint32_t s) const { return sm[s]; }
    __device__ __forceinline__ void stm(uint32_t s, uint4 v) const { sm[s] = v; }
};
struct SlotsSpill {
    uint4* sm;
    uint4* gl;
    uint32_t n_sm;
    __device__ __forceinline__ uint4 ldm(uint32_t s) const { return s < n_sm ? sm[s] : gl[s - n_sm]; }
    __device__ __forceinline__ void stm(uint32_t s, uint4 v) const { if (s < n_sm) sm[s] = v; else gl[s - n_sm] = v; }
};
template <class SL> __device__ __forceinline__ Label lds_label(const SL& slots, uint32_t s) { return label_from_mem(slots.ldm(s)); }
template <class SL> __device__ __forceinline__ void sts_label(const SL& slots, uint32_t s, Label l) { slots.stm(s, label_to_mem(l)); }

__device__ __forceinline__ Label shfl_xor_label(Label h, int m) {
    return Label{__shfl_xor_sync(0xffffffffu, h.w0, m), __shfl_xor_sync(0xffffffffu, h.w1, m),
                 __shfl_xor_sync(0xffffffffu, h.w2, m), __shfl_xor_sync(0xffffffffu, h.w3, m)};
}
__device__ __forceinline__ Label shfl_label(Label h, int src) {
    return Label{__shfl_sync(0xffffffffu, h.w0, src), __shfl_sync(0xffffffffu, h.w1, src),
                 __shfl_sync(0xffffffffu, h.w2, src), __shfl_sync(0xffffffffu, h.w3, src)};
}
__device__ __forceinline__ uint32_t mask_of(uint32_t bit) { return 0u - (bit & 1u); }

// A team's round keys (one thread): FIPS-197 schedule; with two resident tables the middle round keys are kept rotated.
template <int NT>
__device__ __forceinline__ void gate_expand_key(const AesLane& a, const uint8_t* key, int keylen, uint32_t* rk) {
    aes_expand_key<NT>(a, key, keylen, rk);
    if (NT == 2) aes_rotate_mid_keys(rk, (keylen >> 2) + 6);
}

// H(K) = AES(K) ^ K for UU independent blocks, rounds interleaved.
template <int NR, int UU, int NT>
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
    // NT == 2: the team's middle round keys are stored pre-rotated (gate_expand_key), see aes_round_rot2
    if (UU == 1 && !GC_ROLL_SINGLE) {
#pragma unroll
        for (int r = 1; r < NR; r++) {
            if (NT == 2) aes_round_rot2(a, s[0][0], s[0][1], s[0][2], s[0][3], k4[r]);
            else aes_round<NT>(a, s[0][0], s[0][1], s[0][2], s[0][3], k4[r]);
        }
    } else {
        // rolled: with 8-16 teams per SM every warp is somewhere else in the kernel and unrolled rounds thrash the
        // instruction caches (GC_ROUNDS_UNROLL: rounds per loop iteration of the two-block form, for experiments)
        constexpr int U = UU == 2 ? GC_ROUNDS_UNROLL : 1;
#pragma unroll U
        for (int r = 1; r < NR; r++) {
            const uint4 k = k4[r];
#pragma unroll
            for (int j = 0; j < UU; j++) {
                if (NT == 2) aes_round_rot2(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
                else aes_round<NT>(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
            }
        }
    }
    {
        const uint4 k = k4[NR];
#pragma unroll
        for (int j = 0; j < UU; j++) {
            aes_last_round<NT>(a, s[j][0], s[j][1], s[j][2], s[j][3], k);
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
    uint32_t bar;          // what team_barrier takes: the team, or the twin pair
    uint32_t sub;          // twin teams: 0 / 1 within the pair
};

// Shared-memory map of the gate kernels (plan.hpp): [blocks | 64 KiB-aligned tables | blocks], packed below the tables
// first.  A team has a header (round keys, GC_RK_BYTES, and a claim word, 16 B) and n_smem * 16 bytes of labels; the
// headers sit in front of their label blocks or all together at the start, whichever layout holds more teams.
template <int NT>
__device__ __forceinline__ TeamCtx team_ctx(uint8_t* smem, const GcParams& p) {
    TeamCtx c;
    c.tables = aes_align_tables(smem);
    c.team = threadIdx.x / p.team_threads;
    c.ttid = threadIdx.x - c.team * p.team_threads;
    constexpr uint32_t H = GC_RK_BYTES + 16;
    const uint32_t split = p.hdr_split;                                   // 1: headers together at the start, 0: in front of each block
    const uint32_t tb = p.n_smem * 16 + (split ? 0u : H), hdr = split ? p.n_teams * H : 0u;
    const uint32_t pad = (uint32_t)(c.tables - smem);
    const uint32_t in_a = (tb && pad > hdr) ? (pad - hdr) / tb : 0u;      // teams whose block fits below the tables
    auto block_of = [&](uint32_t t) { return t < in_a ? smem + hdr + t * tb : c.tables + aes_table_bytes(NT) + (t - in_a) * tb; };
    auto header_of = [&](uint32_t t) { return split ? smem + t * H : block_of(t); };
    uint8_t* blk = block_of(c.team);
    uint8_t* h = header_of(c.team);
    c.rk = reinterpret_cast<uint32_t*>(h);
    c.claim = reinterpret_cast<volatile uint32_t*>(h + GC_RK_BYTES);
    c.slots = reinterpret_cast<uint4*>(split ? blk : blk + H);
    c.bar = c.team;
    c.sub = 0;
    // TWIN teams: with 16 one-warp teams every warp scheduler holds four warps that are each somewhere else in a 60 KB
    // kernel, and its instruction cache thrashes (ncu: no_inst 26-44 % of the stall samples).  The plan drives the control
    // flow, not the data, so two instances started together stay together if their barriers are shared: warps w and w ^ 4
    // (same scheduler: warp w runs on scheduler w % 4) form a pair, claim two consecutive instances at once and replace
    // __syncwarp by a 64-thread named barrier.  The pair shares the leader's claim word.
    if (p.twin) {
        c.sub = (c.team >> 2) & 1u;
        c.bar = 0x80000000u | (1u + (c.team & 3u) + ((c.team >> 3) << 2));
        c.claim = reinterpret_cast<volatile uint32_t*>(header_of(c.team & ~4u) + GC_RK_BYTES);
    }
    return c;
}

// SPILL: 0 = every label in shared memory; 1 = slots from n_smem up in the instance's global scratch, read and written in
// place (SlotsSpill); 2 = live-range splitting (plan.cpp): every access is shared memory, and the plan's per-phase copy
// lists move values to the scratch after a burst of uses and back, asynchronously, before the next
template <int SPILL>
__device__ __forceinline__ auto team_slots(const TeamCtx& tc, const GcParams& p) {
    if constexpr (SPILL == 1) {
        const size_t team_index = (size_t)blockIdx.x * p.n_teams + tc.team;
        return SlotsSpill{tc.slots, p.spill + team_index * (p.n_slots - p.n_smem), p.n_smem};
    } else {
        return SlotsSmem{tc.slots};
    }
}

// Gate index of cipher task t of a phase.  Garble: 4 tasks per AND/OR, 2 per INV;
// eval: 2 per AND/OR (OR uses one), 1 per INV.
struct Phase {
    uint32_t n_rows, cipher_first, n_quad, n_inv;
    bool has_or;
};
__device__ __forceinline__ Phase load_phase(const uint4* phases, uint32_t i) {
    const uint4 a = __ldg(phases + 2 * i);
    return Phase{a.x & 0x7fffffffu, a.y, a.z, a.w, (a.x >> 31) != 0};
}
template <bool GARBLE>
__device__ __forceinline__ uint32_t task_gate(const Phase& ph, uint32_t t, uint32_t& k) {
    constexpr uint32_t QS = GARBLE ? 2 : 1;           // log2 tasks per AND/OR
    constexpr uint32_t IS = GARBLE ? 1 : 0;           // log2 tasks per INV
    const uint32_t nq = ph.n_quad << QS;
    if (t < nq) { k = t & ((1u << QS) - 1); return ph.cipher_first + (t >> QS); }
    const uint32_t u = t - nq;
    k = u & ((1u << IS) - 1);
    return ph.cipher_first + ph.n_quad + (u >> IS);
}
template <bool GARBLE>
__device__ __forceinline__ uint32_t task_count(const Phase& ph) {
    return GARBLE ? 4 * ph.n_quad + 2 * ph.n_inv : 2 * ph.n_quad + ph.n_inv;
}

// Cipher records of this thread's tasks (k0 + j)*TT + ttid, j < N, of a phase.
template <bool GARBLE, int N>
__device__ __forceinline__ void prefetch_cipher(const GcParams& p, const Phase& ph, uint32_t k0, uint32_t ttid, uint32_t TT,
                                                uint4 (&rec)[N]) {
    const uint32_t ntask = task_count<GARBLE>(ph);
#pragma unroll
    for (int j = 0; j < N; j++) {
        uint32_t k;
        const uint32_t t = (k0 + j) * TT + ttid;
        rec[j] = make_uint4(0, 0, 0, 0);
        if (t < ntask) rec[j] = __ldg(p.crecs + task_gate<GARBLE>(ph, t, k));
    }
}

// One node: dst = XOR of its leaves (^ R for an odd number of XNORs on the way,
// garbler only: garble.go:342-351 swaps the labels of an XNOR, eval.go:48-50 does not).
struct NodeRegs { uint4 lo, hi; };
__device__ __forceinline__ NodeRegs load_node(const uint4* nodes, uint32_t i) {
    return NodeRegs{__ldg(nodes + 2 * (size_t)i), __ldg(nodes + 2 * (size_t)i + 1)};
}

// One node per thread: dst = XOR of the leaves (^ R for an odd number of XNORs on the way).
// words: lo.x = dst | k<<16 | flags<<24 (NODE_PARITY / NODE_WAVE_END / NODE_ACTIVE); then 14 leaf
// slots, two per word.
// PAIR: two leaves (one record word) per step, both loads in flight before the first XOR needs its operand.  One-warp teams
// are bound by the latency of their rows and gain (sha256 x 2368: 2,644 -> 2,724 M AND/s, mul64: 3,756 -> 4,008); the
// multi-warp teams of wide circuits are bound by the shared-memory pipe, where the bursts of 128-bit loads delay the
// other teams' table lookups (aes_128 garble 3.61 -> 3.85 ms): they keep one leaf per step.
template <bool GARBLE, bool FULL, bool PAIR, class SL>
__device__ __forceinline__ void run_node(const GcParams& p, const SL& slots, const Label R, uint32_t inst, uint32_t index,
                                         const NodeRegs& n) {
    const bool active = (n.lo.x >> 24) & NODE_ACTIVE;
    const uint32_t k = active ? (n.lo.x >> 16) & 0xff : 0u;
    Label acc = Label{0, 0, 0, 0};
    if (GARBLE) acc = label_and_mask(R, mask_of(n.lo.x >> 24));
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, k);
    if (PAIR) {
#pragma unroll
        for (int j = 0; j < NODE_MAX_FANIN; j += 2) {
            if ((uint32_t)j >= kmax) break;                    // warp-uniform
            const uint32_t w = j < 2 ? n.lo.y : j < 4 ? n.lo.z : j < 6 ? n.lo.w : j < 8 ? n.hi.x
                             : j < 10 ? n.hi.y : j < 12 ? n.hi.z : n.hi.w;
            Label l0 = Label{0, 0, 0, 0}, l1 = Label{0, 0, 0, 0};
            if ((uint32_t)j < k) l0 = lds_label(slots, w & 0xffff);
            if ((uint32_t)j + 1 < k) l1 = lds_label(slots, w >> 16);
            acc = acc ^ l0 ^ l1;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NODE_MAX_FANIN; j++) {
            if ((uint32_t)j >= kmax) break;                    // warp-uniform
            const uint32_t w = j < 2 ? n.lo.y : j < 4 ? n.lo.z : j < 6 ? n.lo.w : j < 8 ? n.hi.x
                             : j < 10 ? n.hi.y : j < 12 ? n.hi.z : n.hi.w;
            const uint32_t s = (j & 1) ? (w >> 16) : (w & 0xffff);
            if ((uint32_t)j < k) acc = acc ^ lds_label(slots, s);
        }
    }
    if (!active) return;
    sts_label(slots, n.lo.x & 0xffff, acc);
    if (FULL) {
        const size_t ow = (size_t)inst * p.n_wires + __ldg(p.nout_wire + index);
        if (GARBLE) {
            p.wires_full[2 * ow] = label_to_mem(acc);
            p.wires_full[2 * ow + 1] = label_to_mem(acc ^ R);
        } else {
            p.wires_full[ow] = label_to_mem(acc);
        }
    }
}

// This thread's record of node row `row`.
__device__ __forceinline__ NodeRegs load_row(const GcParams& p, uint32_t row, uint32_t ttid) {
    return load_node(p.nodes, row * p.team_threads + ttid);
}

// The node rows of a phase.  pipe[k] holds this thread's record of the next row whose index is
// k mod D: running a row retires its register set and requests the record D rows ahead into the
// same set (the row array ends with GC_NODE_PIPE_MAX empty rows).  The sets are addressed
// statically -- rows run in aligned groups of D with warp-uniform guards -- because a pipeline that
// shifts registers would copy the newest load right after issuing it and wait for it there.  A
// team barrier follows the last row of a wave.
template <bool GARBLE, bool FULL, uint32_t D, bool PAIR, class SL>
__device__ __forceinline__ void run_rows(const GcParams& p, const SL& slots, const Label R, uint32_t inst, uint32_t n_rows,
                                         uint32_t team, uint32_t ttid, uint32_t TT, NodeRegs (&pipe)[D], uint32_t& row) {
    static_assert((D & (D - 1)) == 0, "pipeline depth must be a power of two");
    const uint32_t end = row + n_rows;
    while (row < end) {
        const uint32_t k0 = row & (D - 1);
#pragma unroll
        for (uint32_t k = 0; k < D; k++) {
            if (k >= k0 && row < end) {
                const NodeRegs cur = pipe[k];
                pipe[k] = load_row(p, row + D, ttid);
                run_node<GARBLE, FULL, PAIR>(p, slots, R, inst, row * TT + ttid, cur);
                row++;
                if ((cur.lo.x >> 24) & NODE_WAVE_END) team_barrier(team, TT);
            }
        }
    }
}

// Start teams `stagger` cycles apart.
__device__ __forceinline__ void stagger_start(uint32_t team, uint32_t stagger) {
    if (stagger == 0 || team == 0) return;
    const long long until = clock64() + (long long)team * stagger;
    while (clock64() < until) __nanosleep(2000);
}

// ---- live-range splitting: the copies queued at the top of a phase (SPILL == 2) ------------------------------------
__device__ __forceinline__ void cp_async16(uint4* smem_dst, const uint4* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
// What a thread keeps of a phase's copy lists: the header and its own first record of each list, requested a phase ahead.
struct CopyRegs { uint32_t first, counts, ev, rl; };          // counts = evicts | reloads << 16
__device__ __forceinline__ CopyRegs load_copies(const GcParams& p, uint32_t pi, uint32_t ttid) {
    const uint4 h = __ldg(p.phases + 2 * pi + 1);             // (row_first, copy_first, counts, -)
    CopyRegs c{h.y, h.z, 0u, 0u};
    const uint32_t n_ev = c.counts & 0xffffu, n_rl = c.counts >> 16;
    if (ttid < n_ev) c.ev = __ldg(p.copies + c.first + ttid);
    if (ttid < n_rl) c.rl = __ldg(p.copies + c.first + n_ev + ttid);
    return c;
}
// Evicts: hot slot -> scratch (the value was produced in the previous phase; a barrier lies between).  Reloads: scratch ->
// hot slot, asynchronously; they are complete when the phase ends (copies_done before its last barrier) and the cluster
// of uses starts in the next phase.  Returns the number of reloads queued.
__device__ __forceinline__ uint32_t run_copies(const GcParams& p, uint4* sm, uint4* scratch, const CopyRegs& c, uint32_t ttid, uint32_t TT) {
    const uint32_t n_ev = c.counts & 0xffffu, n_rl = c.counts >> 16;
    if (ttid < n_ev) scratch[c.ev >> 16] = sm[c.ev & 0xffffu];
    for (uint32_t k = TT + ttid; k < n_ev; k += TT) { const uint32_t r = __ldg(p.copies + c.first + k); scratch[r >> 16] = sm[r & 0xffffu]; }
    if (ttid < n_rl) cp_async16(sm + (c.rl >> 16), scratch + (c.rl & 0xffffu));
    for (uint32_t k = TT + ttid; k < n_rl; k += TT) { const uint32_t r = __ldg(p.copies + c.first + n_ev + k); cp_async16(sm + (r >> 16), scratch + (r & 0xffffu)); }
    if (n_rl) asm volatile("cp.async.commit_group;" ::: "memory");
    return n_rl;
}
__device__ __forceinline__ void copies_done() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------ garble ----
template <class SL>
struct GarbleEnv {
    GcParams const* p;
    SL slots;
    const uint32_t* rk;
    uint4* tab;
    Label R;
    uint32_t inst;
};

// One pass of UU cipher tasks per thread: tasks (k0 + j)*TT + ttid.
template <int NR, int MODE, int UU, int ILP, bool AND_ONLY, int NT, class SL>
__device__ __forceinline__ void garble_pass(const AesLane& lane, const GarbleEnv<SL>& e, const Phase& ph, uint32_t ntask,
                                            uint32_t k0, uint32_t ttid, uint32_t TT, const uint4 (&rec)[ILP]) {
    static_assert(UU <= ILP, "pass wider than the record buffer");
    constexpr bool FULL = MODE == GC_FULL;
    const GcParams& p = *e.p;
    const SL& slots = e.slots;
    const Label R = e.R;
    if constexpr (AND_ONLY) {
        // every task of the pass belongs to an AND gate: (a0,j0) (a1,j0) (b0,j1) (b1,j1) on four
        // adjacent lanes (garble.go:353-395; the reference hashes a0 and b0 twice)
        const uint32_t k = ttid & 3u;
        Label keep[UU], K[UU], H[UU];
        uint32_t pab[UU];
#pragma unroll
        for (int j = 0; j < UU; j++) {
            const uint4 g = rec[j];
            const Label a0 = lds_label(slots, g.x & 0xffff), b0 = lds_label(slots, g.x >> 16);
            pab[j] = 2 * label_s(a0) + label_s(b0);
            keep[j] = a0;
            Label x = (k & 2) ? b0 : a0;
            x = x ^ label_and_mask(R, mask_of(k));
            K[j] = label_shl(x, 1);
            K[j].w3 ^= g.z + (k >> 1);
        }
        aes_hash_multi<NR, UU, NT>(lane, e.rk, K, H);
#pragma unroll
        for (int j = 0; j < UU; j++) {
            const uint4 g = rec[j];
            const Label h = H[j];
            const uint32_t pa = pab[j] >> 1, pb = pab[j] & 1;
            const Label u = h ^ shfl_xor_label(h, 1);
            Label v = Label{0, 0, 0, 0};
            if (k == 0) {                                      // generator half (garble.go:361-369)
                const Label tg = u ^ label_and_mask(R, mask_of(pb));
                v = h ^ label_and_mask(tg, mask_of(pa));
                e.tab[g.w] = label_to_mem(tg);
            } else if (k == 2) {                               // evaluator half (garble.go:372-380)
                v = h ^ label_and_mask(u, mask_of(pb));
                e.tab[g.w + 1] = label_to_mem(u ^ keep[j]);
            }
            const Label w2 = v ^ shfl_xor_label(v, 2);
            if (k == 0) {                                      // combine halves (garble.go:383-392)
                sts_label(slots, g.y & 0xffff, w2);
                if (FULL) {
                    uint32_t kk;
                    const uint32_t gi = task_gate<true>(ph, (k0 + j) * TT + ttid, kk);
                    uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gi)) * 2;
                    w[0] = label_to_mem(w2);
                    w[1] = label_to_mem(w2 ^ R);
                }
            }
        }
        return;
    }
    uint4 g[UU];
    Label a0[UU], K[UU], H[UU];
    uint32_t op[UU], kk[UU], pp[UU], gidx[UU];
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const uint32_t t = (k0 + j) * TT + ttid;
        const bool active = t < ntask;
        uint32_t k;
        const uint32_t gi = task_gate<true>(ph, t, k);
        g[j] = rec[j];
        const uint32_t sa = g[j].x & 0xffff, sb = g[j].x >> 16;
        op[j] = active ? ((g[j].y >> 16) & 0xff) : 0xffu;
        kk[j] = k; gidx[j] = gi;
        // idle lanes of a partly filled pass read nothing (their zero record would name slot 0, which a
        // gate of this level may be writing)
        a0[j] = Label{0, 0, 0, 0};
        Label b0 = Label{0, 0, 0, 0};
        if (active) { a0[j] = lds_label(slots, sa); b0 = lds_label(slots, sb); }
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
    aes_hash_multi<NR, UU, NT>(lane, e.rk, K, H);
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
                    uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gidx[j])) * 2;
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
                        uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gidx[j])) * 2;
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
                uint4* w = p.wires_full + ((size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gidx[j])) * 2;
                w[0] = label_to_mem(w2);
                w[1] = label_to_mem(w2 ^ R);
            }
        }
    }
}

template <int NR, int MODE, int ILP, int MAXT, int NT, int SPILL>
__global__ void __launch_bounds__(MAXT, 1) garble_kernel(const GcParams p) {
    constexpr uint32_t D = node_pipe(ILP, MAXT, true);
    constexpr bool FULL = MODE == GC_FULL;
    constexpr bool STREAM = MODE == GC_STREAM;
    extern __shared__ __align__(16) uint8_t smem[];
    const TeamCtx tc = team_ctx<NT>(smem, p);
    aes_tables_to_smem<NT>(tc.tables);
    __syncthreads();
    const AesLane lane = aes_lane(tc.tables);
    const uint32_t TT = p.team_threads, ttid = tc.ttid;
    if (p.key_stride == 0) {
        if (ttid == 0) gate_expand_key<NT>(lane, p.keys, (int)p.keylen, tc.rk);
        team_barrier(tc.bar, TT);
    }
    const auto slots = team_slots<SPILL>(tc, p);
    constexpr bool COPIES = SPILL == 2;
    uint4* const scratch = COPIES ? p.spill + ((size_t)blockIdx.x * p.n_teams + tc.team) * (p.n_slots - p.n_smem) : nullptr;
    stagger_start(p.twin ? (tc.bar & 0xffu) : tc.team, p.stagger);

    for (;;) {
        // the next instance (twin teams: the pair's next two; an odd tail is run twice, by both warps, with the same result)
        if (ttid == 0 && tc.sub == 0) *tc.claim = atomicAdd(p.counter, p.twin ? 2u : 1u);
        team_barrier(tc.bar, TT);
        const uint32_t claimed = *tc.claim;
        if (claimed >= p.batch) break;
        const uint32_t inst = claimed + tc.sub < p.batch ? claimed + tc.sub : p.batch - 1;
        if (p.key_stride != 0 && ttid == 0)
            gate_expand_key<NT>(lane, p.keys + (size_t)inst * p.key_stride, (int)p.keylen, tc.rk);
        Label R = label_from_mem(__ldg(p.r + inst));
        R.w0 |= 0x80000000u;                                   // r.SetS(true), garble.go:258
        // plan records of the first phases load while the inputs do
        Phase ph = load_phase(p.phases, 0), ph_n = load_phase(p.phases, 1);
        NodeRegs pipe[D];
        uint32_t row = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) pipe[d] = load_row(p, d, ttid);
        CopyRegs cp_cur{0, 0, 0, 0};
        if (COPIES) cp_cur = load_copies(p, 0, ttid);
        // input wires: L0 from the caller's reader bytes, L1 = L0 ^ R (garble.go:271-278)
        for (uint32_t k = ttid; k < p.n_live_in; k += TT) {
            const uint2 ref = __ldg(p.live_in + k);
            const uint32_t src = ref.y & 0x7fffffffu;          // bit 31: the second entry of an input that goes to a hot slot and the scratch
            uint4 m;
            if (STREAM) m = *wf_slot(p.pages, __ldg(p.in_ids + src), inst);
            else m = __ldg(p.in_labels + (size_t)inst * p.n_in + src);
            if (COPIES && ref.x >= p.n_smem) scratch[ref.x - p.n_smem] = m;
            else slots.stm(ref.x, m);
            if (ref.y >> 31) continue;
            if (!STREAM && p.io) {
                uint4* w = p.io + ((size_t)inst * (p.n_in + p.n_out) + src) * 2;
                w[0] = m;
                w[1] = label_to_mem(label_from_mem(m) ^ R);
            }
            if (FULL) {
                uint4* w = p.wires_full + ((size_t)inst * p.n_wires + src) * 2;
                w[0] = m;
                w[1] = label_to_mem(label_from_mem(m) ^ R);
            }
        }
        team_barrier(tc.bar, TT);
        const GarbleEnv<decltype(slots)> env{&p, slots, tc.rk, p.tables + (size_t)inst * p.n_rows, R, inst};
        // first-pass gate records: requested one phase ahead where the registers allow it (single-block
        // variants), at the top of their own phase -- behind the node rows -- in the two-block variants
        constexpr bool AHEAD = ILP == 1;
        uint4 cur[ILP];
        if (AHEAD) prefetch_cipher<true, ILP>(p, ph, 0, ttid, TT, cur);

        for (uint32_t pi = 0; pi < p.n_phases; pi++) {
            const Phase ph_nn = load_phase(p.phases, pi + 2);  // two zero records of padding
            uint4 cur_n[AHEAD ? ILP : 1];
            if (AHEAD) prefetch_cipher<true, ILP>(p, ph_n, 0, ttid, TT, reinterpret_cast<uint4 (&)[ILP]>(cur_n));
            else prefetch_cipher<true, ILP>(p, ph, 0, ttid, TT, cur);
            uint32_t n_reload = 0;
            if (COPIES) {                                      // this phase's evicts and reloads; the next phase's records are requested
                n_reload = run_copies(p, tc.slots, scratch, cp_cur, ttid, TT);
                cp_cur = load_copies(p, pi + 1, ttid);
            }
            const bool tracing = p.trace && blockIdx.x == 0 && threadIdx.x == 0 && inst < gridDim.x * p.n_teams;
            if (tracing) p.trace[4 * pi] = clock64();
            // ---- free wires: waves of independent XOR nodes, one thread per node
            // (the single-block variants always pair: their circuits are narrow; the two-block variants decide by team width)
            if (ILP == 1 || TT == 32) run_rows<true, FULL, D, true>(p, slots, R, inst, ph.n_rows, tc.bar, ttid, TT, pipe, row);
            else run_rows<true, FULL, D, ILP == 1>(p, slots, R, inst, ph.n_rows, tc.bar, ttid, TT, pipe, row);
            if (tracing) p.trace[4 * pi + 1] = p.trace[4 * pi + 2] = clock64();
            // ---- ciphered gates: one AES block per task, up to ILP tasks per thread at once
            const uint32_t ntask = task_count<true>(ph);
            if (ntask) {
                const uint32_t per_thread = (ntask + TT - 1) / TT;        // tasks k = 0 .. per_thread-1
                const uint32_t and_tasks = ph.has_or ? 0u : 4 * ph.n_quad;  // leading tasks that are all AND
                for (uint32_t k0 = 0; k0 < per_thread;) {
                    const uint32_t left = per_thread - k0;
                    const uint32_t uu = (ILP >= 4 && left >= 3) ? 4u : (ILP >= 2 && left >= 2) ? 2u : 1u;
                    uint4 nxt[ILP];                                        // records of the next pass, in flight during this one
                    prefetch_cipher<true, ILP>(p, ph, k0 + uu, ttid, TT, nxt);
                    const bool and_only = (k0 + uu) * TT <= and_tasks;
                    if (ILP >= 4 && uu == 4) {
                        if (and_only) garble_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                        else garble_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                    } else if (ILP >= 2 && uu == 2 && (k0 + 1) * TT + (ttid & ~31u) < ntask) {
                        if (and_only) garble_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                        else garble_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                    } else if (k0 * TT + (ttid & ~31u) < ntask) {
                        if (and_only) garble_pass<NR, MODE, 1, ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                        else garble_pass<NR, MODE, 1, ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur);
                    }
                    k0 += uu;
#pragma unroll
                    for (int j = 0; j < ILP; j++) cur[j] = nxt[j];
                }
                if (tracing) p.trace[4 * pi + 3] = clock64();
                if (COPIES && n_reload) copies_done();
                team_barrier(tc.bar, TT);
            } else if (COPIES && n_reload) {
                copies_done();
                team_barrier(tc.bar, TT);
            }
            ph = ph_n; ph_n = ph_nn;
            if (AHEAD) {
#pragma unroll
                for (int j = 0; j < ILP; j++) cur[j] = cur_n[j];
            }
        }
        // output wires (what circuit/garbler.go:153 and sha2pc/garbler.go:125 read)
        if (STREAM) {                                          // Streaming.Set, stream_garble.go:144-157
            for (uint32_t k = ttid; k < p.n_out; k += TT) {
                const uint2 ref = __ldg(p.live_out + k);
                *wf_slot(p.pages, __ldg(p.out_ids + ref.y), inst) = (COPIES && ref.x >= p.n_smem) ? scratch[ref.x - p.n_smem] : slots.ldm(ref.x);
            }
        } else if (p.io) {
            for (uint32_t k = ttid; k < p.n_out; k += TT) {
                const uint2 ref = __ldg(p.live_out + k);
                const Label l0 = label_from_mem((COPIES && ref.x >= p.n_smem) ? scratch[ref.x - p.n_smem] : slots.ldm(ref.x));
                uint4* w = p.io + ((size_t)inst * (p.n_in + p.n_out) + p.n_in + ref.y) * 2;
                w[0] = label_to_mem(l0);
                w[1] = label_to_mem(l0 ^ R);
            }
        }
        team_barrier(tc.bar, TT);
    }
}

// -------------------------------------------------------------------- eval ----
template <class SL>
struct EvalEnv {
    GcParams const* p;
    SL slots;
    const uint32_t* rk;
    const uint4* tab;
    uint32_t inst;
};

template <int NR, int MODE, int UU, int ILP, bool AND_ONLY, int NT, class SL>
__device__ __forceinline__ void eval_pass(const AesLane& lane, const EvalEnv<SL>& e, const Phase& ph, uint32_t ntask,
                                          uint32_t k0, uint32_t ttid, uint32_t TT, const uint4 (&rec)[ILP],
                                          const uint4 (&rowpre)[ILP], bool have_rows) {
    static_assert(UU <= ILP, "pass wider than the record buffer");
    constexpr bool FULL = MODE == GC_FULL;
    const GcParams& p = *e.p;
    const SL& slots = e.slots;
    if constexpr (AND_ONLY) {
        // every task of the pass belongs to an AND gate: lane pair (a, j0) (b, j1), eval.go:52-78
        const uint32_t k = ttid & 1u;
        Label K[UU], H[UU], row[UU];
#pragma unroll
        for (int j = 0; j < UU; j++) {
            const uint4 g = rec[j];
            const Label a = lds_label(slots, g.x & 0xffff), b = lds_label(slots, g.x >> 16);
            const Label x = k ? b : a;
            // tg when S(a), te when S(b); the row was requested a phase ago when this is the first pass
            uint4 rm = rowpre[j];
            if (!have_rows) rm = __ldg(e.tab + g.w + k);
            row[j] = label_and_mask(label_from_mem(rm), mask_of(label_s(x)));
            if (k) row[j] = row[j] ^ label_and_mask(a, mask_of(label_s(b)));   // we ^= a when S(b)
            K[j] = label_shl(x, 1);
            K[j].w3 ^= g.z + k;
        }
        aes_hash_multi<NR, UU, NT>(lane, e.rk, K, H);
#pragma unroll
        for (int j = 0; j < UU; j++) {
            const uint4 g = rec[j];
            const Label v = H[j] ^ row[j];
            const Label o = v ^ shfl_xor_label(v, 1);
            if (k == 0) {
                sts_label(slots, g.y & 0xffff, o);
                if (FULL) {
                    uint32_t kk;
                    const uint32_t gi = task_gate<false>(ph, (k0 + j) * TT + ttid, kk);
                    p.wires_full[(size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gi)] = label_to_mem(o);
                }
            }
        }
        return;
    }
    uint4 g[UU];
    Label K[UU], H[UU], row[UU];
    uint32_t op[UU], kk[UU], gidx[UU];
    bool act[UU];
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const uint32_t t = (k0 + j) * TT + ttid;
        act[j] = t < ntask;
        uint32_t k;
        const uint32_t gi = task_gate<false>(ph, t, k);
        g[j] = rec[j];
        const uint32_t sa = g[j].x & 0xffff, sb = g[j].x >> 16;
        op[j] = act[j] ? ((g[j].y >> 16) & 0xff) : 0xffu;
        kk[j] = k; gidx[j] = gi;
        Label a = Label{0, 0, 0, 0}, b = Label{0, 0, 0, 0};     // idle lanes read nothing (see garble_pass)
        if (act[j]) { a = lds_label(slots, sa); b = lds_label(slots, sb); }
        const uint32_t sA = label_s(a), sB = label_s(b);
        // the garbled row this task may need, fetched before the AES so the HBM
        // latency hides behind the rounds
        uint32_t ridx = g[j].w;
        bool need = false;
        if (op[j] == OP_AND) { ridx += k; need = k ? sB : sA; }
        else if (op[j] == OP_INV) { need = sA; }
        else if (op[j] == OP_OR) { const uint32_t ix = 2 * sA + sB; need = (ix > 0) && (k == 0); ridx += ix - 1; }
        row[j] = Label{0, 0, 0, 0};
        if (need) row[j] = label_from_mem((have_rows && op[j] != OP_OR) ? rowpre[j] : __ldg(e.tab + ridx));
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
    aes_hash_multi<NR, UU, NT>(lane, e.rk, K, H);
#pragma unroll
    for (int j = 0; j < UU; j++) {
        const Label v = H[j] ^ row[j];                         // decrypt (garble.go:58-73) / half-gate
        const Label o = v ^ shfl_xor_label(v, 1);
        if (act[j] && (kk[j] == 0)) {
            const Label res = (op[j] == OP_AND) ? o : v;
            sts_label(slots, g[j].y & 0xffff, res);
            if (FULL) p.wires_full[(size_t)e.inst * p.n_wires + __ldg(p.cout_wire + gidx[j])] = label_to_mem(res);
        }
    }
}

template <int NR, int MODE, int ILP, int MAXT, int NT, int SPILL>
__global__ void __launch_bounds__(MAXT, 1) eval_kernel(const GcParams p) {
    constexpr uint32_t D = node_pipe(ILP, MAXT, false);
    constexpr bool FULL = MODE == GC_FULL;
    constexpr bool STREAM = MODE == GC_STREAM;
    extern __shared__ __align__(16) uint8_t smem[];
    const TeamCtx tc = team_ctx<NT>(smem, p);
    aes_tables_to_smem<NT>(tc.tables);
    __syncthreads();
    const AesLane lane = aes_lane(tc.tables);
    const uint32_t TT = p.team_threads, ttid = tc.ttid;
    if (p.key_stride == 0) {
        if (ttid == 0) gate_expand_key<NT>(lane, p.keys, (int)p.keylen, tc.rk);
        team_barrier(tc.bar, TT);
    }
    const auto slots = team_slots<SPILL>(tc, p);
    constexpr bool COPIES = SPILL == 2;
    uint4* const scratch = COPIES ? p.spill + ((size_t)blockIdx.x * p.n_teams + tc.team) * (p.n_slots - p.n_smem) : nullptr;
    stagger_start(p.twin ? (tc.bar & 0xffu) : tc.team, p.stagger);

    for (;;) {
        // the next instance (twin teams: the pair's next two; an odd tail is run twice, by both warps, with the same result)
        if (ttid == 0 && tc.sub == 0) *tc.claim = atomicAdd(p.counter, p.twin ? 2u : 1u);
        team_barrier(tc.bar, TT);
        const uint32_t claimed = *tc.claim;
        if (claimed >= p.batch) break;
        const uint32_t inst = claimed + tc.sub < p.batch ? claimed + tc.sub : p.batch - 1;
        if (p.key_stride != 0 && ttid == 0)
            gate_expand_key<NT>(lane, p.keys + (size_t)inst * p.key_stride, (int)p.keylen, tc.rk);
        Phase ph = load_phase(p.phases, 0), ph_n = load_phase(p.phases, 1);
        NodeRegs pipe[D];
        uint32_t row = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) pipe[d] = load_row(p, d, ttid);
        CopyRegs cp_cur{0, 0, 0, 0};
        if (COPIES) cp_cur = load_copies(p, 0, ttid);
        for (uint32_t k = ttid; k < p.n_live_in; k += TT) {
            const uint2 ref = __ldg(p.live_in + k);
            const uint32_t src = ref.y & 0x7fffffffu;
            uint4 m;
            if (STREAM) m = *wf_slot(p.pages, __ldg(p.in_ids + src), inst);      // StreamEval.Get, stream_evaluator.go:57-66
            else m = __ldg(p.in_labels + (size_t)inst * p.n_in + src);
            if (COPIES && ref.x >= p.n_smem) scratch[ref.x - p.n_smem] = m;
            else slots.stm(ref.x, m);
            if (FULL && !(ref.y >> 31)) p.wires_full[(size_t)inst * p.n_wires + src] = m;
        }
        team_barrier(tc.bar, TT);
        const EvalEnv<decltype(slots)> env{&p, slots, tc.rk, p.tables + (size_t)inst * p.n_rows, inst};
        uint4 cur[ILP];
        prefetch_cipher<false, ILP>(p, ph, 0, ttid, TT, cur);

        for (uint32_t pi = 0; pi < p.n_phases; pi++) {
            const Phase ph_nn = load_phase(p.phases, pi + 2);
            // the first-pass records were loaded a phase ago; request their garbled rows now, so the
            // HBM latency hides behind the node waves (AND: row k of the pair; INV: its one row)
            uint4 rows[ILP];
            {
                const uint32_t nt = task_count<false>(ph);
#pragma unroll
                for (int j = 0; j < ILP; j++) {
                    const uint32_t t = j * TT + ttid;
                    rows[j] = make_uint4(0, 0, 0, 0);
                    if (t < nt) rows[j] = __ldg(env.tab + cur[j].w + ((t < 2 * ph.n_quad) ? (t & 1u) : 0u));
                }
            }
            uint4 cur_n[ILP];
            prefetch_cipher<false, ILP>(p, ph_n, 0, ttid, TT, cur_n);
            uint32_t n_reload = 0;
            if (COPIES) {
                n_reload = run_copies(p, tc.slots, scratch, cp_cur, ttid, TT);
                cp_cur = load_copies(p, pi + 1, ttid);
            }
            if (ILP == 1 || TT == 32) run_rows<false, FULL, D, true>(p, slots, Label{0, 0, 0, 0}, inst, ph.n_rows, tc.bar, ttid, TT, pipe, row);
            else run_rows<false, FULL, D, ILP == 1>(p, slots, Label{0, 0, 0, 0}, inst, ph.n_rows, tc.bar, ttid, TT, pipe, row);
            // ciphered gates: AND = 2 tasks (a with j0, b with j1); OR/INV = 1 hash.
            // OR shares the 2-task slot of its class (second task idle).
            const uint32_t ntask = task_count<false>(ph);
            if (ntask) {
                const uint32_t per_thread = (ntask + TT - 1) / TT;
                const uint32_t and_tasks = ph.has_or ? 0u : 2 * ph.n_quad;  // leading tasks that are all AND
                for (uint32_t k0 = 0; k0 < per_thread;) {
                    const uint32_t left = per_thread - k0;
                    const uint32_t uu = (ILP >= 4 && left >= 3) ? 4u : (ILP >= 2 && left >= 2) ? 2u : 1u;
                    uint4 nxt[ILP];                                        // records of the next pass, in flight during this one
                    prefetch_cipher<false, ILP>(p, ph, k0 + uu, ttid, TT, nxt);
                    const bool and_only = (k0 + uu) * TT <= and_tasks;
                    if (ILP >= 4 && uu == 4) {
                        if (and_only) eval_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                        else eval_pass<NR, MODE, (ILP >= 4 ? 4 : 1), ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                    } else if (ILP >= 2 && uu == 2 && (k0 + 1) * TT + (ttid & ~31u) < ntask) {
                        if (and_only) eval_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                        else eval_pass<NR, MODE, (ILP >= 2 ? 2 : 1), ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                    } else if (k0 * TT + (ttid & ~31u) < ntask) {
                        if (and_only) eval_pass<NR, MODE, 1, ILP, true, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                        else eval_pass<NR, MODE, 1, ILP, false, NT>(lane, env, ph, ntask, k0, ttid, TT, cur, rows, k0 == 0);
                    }
                    k0 += uu;
#pragma unroll
                    for (int j = 0; j < ILP; j++) cur[j] = nxt[j];
                }
                if (COPIES && n_reload) copies_done();
                team_barrier(tc.bar, TT);
            } else if (COPIES && n_reload) {
                copies_done();
                team_barrier(tc.bar, TT);
            }
            ph = ph_n; ph_n = ph_nn;
#pragma unroll
            for (int j = 0; j < ILP; j++) cur[j] = cur_n[j];
        }
        for (uint32_t k = ttid; k < p.n_out; k += TT) {
            const uint2 ref = __ldg(p.live_out + k);
            const uint4 v = (COPIES && ref.x >= p.n_smem) ? scratch[ref.x - p.n_smem] : slots.ldm(ref.x);
            if (STREAM) *wf_slot(p.pages, __ldg(p.out_ids + ref.y), inst) = v;   // StreamEval.Set
            else p.io[(size_t)inst * p.n_out + ref.y] = v;
        }
        team_barrier(tc.bar, TT);
    }
}

}  // namespace gcb
