// plan.hpp -- host-side circuit plan compiler.
//
// Turns Circuit.Gates (circuit/circuit.go:120-131, 260-266) into the static
// schedule the kernels walk: dependency steps, per-gate tweak ids and slab row
// offsets in ORIGINAL gate order (circuit/garble.go:285-299,357-359,419-420,
// 451-452), and on-chip wire slots assigned from liveness.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gcb200.h"

namespace gcb { int fail(int code, const char* fmt, ...); }
// Exceptions must not cross the C boundary (a cgo caller would terminate).
#define GCB_TRY try {
#define GCB_CATCH                                                                                             \
    } catch (const std::bad_alloc&) { return gcb::fail(GCB_E_TOO_LARGE, "out of host memory"); }              \
    catch (const std::exception& ex_) { return gcb::fail(GCB_E_ARG, "internal error: %s", ex_.what()); }      \
    catch (...) { return gcb::fail(GCB_E_ARG, "internal error"); }

namespace gcb {

enum : uint8_t { OP_XOR = 0, OP_XNOR = 1, OP_AND = 2, OP_OR = 3, OP_INV = 4 };

// One ciphered gate (AND / OR / INV) as the kernels read it: 16 bytes, one
// 128-bit load.
struct GateRec {
    uint16_t a, b, c;     // wire slots of in0, in1, out
    uint8_t op;
    uint8_t pad;
    uint32_t tweak;       // first tweak id of the gate
    uint32_t row;         // first slab row
};
static_assert(sizeof(GateRec) == 16, "GateRec must be 16 bytes");

// One materialised free wire.  The plan compiler flattens the XOR / XNOR gates of a
// phase: a wire that a ciphered gate, a later phase or the caller reads becomes a
// NODE = the XOR of up to NODE_MAX_FANIN leaves (wires that exist when the phase
// starts, or nodes of an earlier wave of the same phase), plus R when an odd
// number of XNORs lie on the way (garbler only; the evaluator's XNOR is a plain
// XOR, eval.go:48-50).  Intermediate XOR wires that nothing else reads are never
// computed or stored.  32 bytes, two 128-bit loads.
constexpr int NODE_MAX_FANIN = 14;
struct NodeRec {
    uint16_t dst;         // wire slot of the result
    uint8_t k;            // number of leaves
    uint8_t parity;       // 1: XOR R in (garbler)
    uint16_t leaf[NODE_MAX_FANIN];
};
static_assert(sizeof(NodeRec) == 32, "NodeRec must be 32 bytes");

// Nodes of one wave are independent of each other; wave w+1 may read wave w.
struct WaveRec {
    uint32_t first, count;            // range in the NodeRec array
};

// One phase = the waves of free-wire nodes that become computable after the
// previous cipher level, followed by one level of ciphered gates: AND/OR first,
// then INV.
struct PhaseRec {
    uint32_t wave_first, n_waves;     // range in the WaveRec array; bit 31 of n_waves: the level has OR gates
    uint32_t cipher_first, n_quad;    // range in the GateRec array: n_quad AND/OR gates ...
    uint32_t n_inv;                   // ... then n_inv INV gates
    uint32_t w0_first;                // first node of the phase; its waves are contiguous in the NodeRec array
    uint16_t wave_count[4];           // node counts of waves 0..3 (saves dependent loads); 0xffff = see WaveRec
};
static_assert(sizeof(PhaseRec) == 32, "PhaseRec must be 32 bytes");

struct SlotRef {
    uint32_t slot;
    uint32_t index;       // index into the caller's live-in / live-out array
};

// ---- what the kernels read (built per device and team width by plan_on_device) ----
// The nodes are laid out as ROWS of team_threads records: lane t of the team runs record t of
// each row.  A wave occupies whole rows (padded with inactive records), rows follow each other
// in schedule order across waves and phases, so every thread streams its records with a fixed
// prefetch distance no matter where wave and phase boundaries fall.  Flags in NodeRec::parity:
enum : uint8_t { NODE_PARITY = 1, NODE_WAVE_END = 2, NODE_ACTIVE = 4 };   // WAVE_END: team barrier after the row
struct DevPhaseRec {
    uint32_t n_rows;                  // node rows of the phase; bit 31: the cipher level has OR gates
    uint32_t cipher_first, n_quad;    // range in the GateRec array: n_quad AND/OR gates ...
    uint32_t n_inv;                   // ... then n_inv INV gates
    uint32_t row_first;               // first node row of the phase
    uint32_t copy_first, n_evict, n_reload;   // copies queued at the top of the phase (hot / cold plans)
};
static_assert(sizeof(DevPhaseRec) == 32, "DevPhaseRec must be 32 bytes");
constexpr uint32_t GC_NODE_PIPE_MAX = 4;  // most rows any kernel variant keeps in flight ahead of the one it runs

struct PlanSpec {
    const gcb_gate* gates = nullptr;
    uint32_t num_gates = 0, num_wires = 0;
    // wire -> storage location (identity for a static circuit; in streaming mode
    // aliased in/out ids share a location).  Empty = identity.
    std::vector<uint32_t> loc;
    uint32_t num_locs = 0;
    std::vector<uint32_t> live_in;    // locations holding a value before gate 0; k-th entry loads source k
    std::vector<uint32_t> live_out;   // locations read back after the last gate; k-th entry stores dest k
};

struct DevicePlan {                   // per (device, team width) copy of the tables
    int device = -1;
    uint32_t team_threads = 0;
    DevPhaseRec* phases = nullptr;    // + two zero records of padding
    NodeRec* nodes = nullptr;         // rows of team_threads records + GC_NODE_PIPE_MAX empty rows
    GateRec* crecs = nullptr;
    uint32_t* nout_wire = nullptr;    // original output wire of each node record / ciphered gate
    uint32_t* cout_wire = nullptr;
    SlotRef* live_in = nullptr;
    SlotRef* live_out = nullptr;
    uint32_t* copies = nullptr;       // (src, dst) pairs of the evict / reload lists
    ~DevicePlan();
};

struct Plan {
    gcb_plan_info info{};
    uint32_t ilp = 1;                     // AES blocks a thread interleaves (kernel variant)
    uint32_t stagger = 0;                 // SM cycles between the starts of consecutive teams
    int policy = 0;                       // which schedule won (0 ASAP, 1 ALAP, 2 cipher-ASAP + free-ALAP, 3 / 4 balanced for the garbler's / evaluator's passes)
    std::vector<PhaseRec> phases;
    std::vector<WaveRec> waves;
    std::vector<NodeRec> nodes;           // free-wire nodes in schedule order
    std::vector<GateRec> crecs;           // ciphered gates in schedule order
    std::vector<uint32_t> nout_wire, cout_wire;   // original output wire of nodes[i] / crecs[i]
    uint32_t node_loads = 0;              // sum of node fan-ins (label loads of the free part)
    uint64_t cold_accesses = 0;           // label reads + writes that go to the global-memory scratch (hot / cold plans)
    // live-range splitting (plan.cpp): per phase (first copy, evicts, reloads); copies = (src, dst) pairs -- an evict copies
    // hot slot src to scratch index dst, a reload scratch index src to hot slot dst
    std::vector<uint32_t> copies;
    std::vector<std::array<uint32_t, 3>> phase_copy;
    std::vector<SlotRef> live_in, live_out;
    std::vector<uint32_t> row_off;        // num_gates+1, original order
    std::vector<uint8_t> ops;             // original order
    std::vector<gcb_gate> gates;          // kept for streaming (header templates)

    mutable std::mutex mu;
    mutable std::map<std::pair<int, uint32_t>, std::shared_ptr<DevicePlan>> dev;   // (device, team width)
};

// ---- shared-memory geometry of the gate kernels (gc_kernels.cuh), needed by the compiler's choices too ----
// [team headers | label blocks of the first teams | 64 KiB-aligned T-tables (nt * 32 KiB) | label blocks of the rest].
// A header is a team's round keys (256 B) and claim word (16 B); keeping the headers together, away from the label
// blocks, can let the blocks fill the two regions either side of the tables better (sha256's balanced schedule needs
// 1,272 labels: eight such blocks fit only this way).  smem_base: where dynamic shared memory starts in the shared window.
constexpr size_t kSmemOptin = 232448;                 // 227 KiB per CTA on sm_100
constexpr uint32_t kAssumedSmemBase = 1024;
constexpr size_t kTeamHeaderBytes = 256 + 16;
inline size_t table_pad(uint32_t smem_base) { return ((smem_base + 0xffffu) & ~0xffffu) - smem_base; }
// Two layouts, whichever holds more teams: split (below) or inline (each team's header directly in front of its labels).
// teams_below: team blocks that fit below the tables.
inline size_t team_block_bytes(uint32_t smem_slots, bool split) { return (size_t)smem_slots * 16 + (split ? 0 : kTeamHeaderBytes); }
inline size_t teams_below(uint32_t smem_slots, uint32_t n_teams, uint32_t smem_base, bool split) {
    const size_t per = team_block_bytes(smem_slots, split), pad = table_pad(smem_base), hdr = split ? (size_t)n_teams * kTeamHeaderBytes : 0;
    return (per && pad > hdr) ? (pad - hdr) / per : 0;
}
inline size_t teams_that_fit_layout(uint32_t smem_slots, uint32_t smem_base, uint32_t nt, bool split) {
    const size_t per = team_block_bytes(smem_slots, split), pad = table_pad(smem_base);
    if (!per) return 32;
    const size_t above = (kSmemOptin - pad - (size_t)nt * 32768) / per;
    for (uint32_t n = 32; n >= 1; n--)
        if ((!split || (size_t)n * kTeamHeaderBytes <= pad) && teams_below(smem_slots, n, smem_base, split) + above >= n) return n;
    return 0;
}
inline size_t teams_that_fit(uint32_t smem_slots, uint32_t smem_base, uint32_t nt, bool* split = nullptr) {
    const size_t a = teams_that_fit_layout(smem_slots, smem_base, nt, false), b = teams_that_fit_layout(smem_slots, smem_base, nt, true);
    if (split) *split = b > a;
    return b > a ? b : a;
}

// Returns GCB_OK or a negative status; message in err.  max_fanin = NODE_MAX_FANIN
// for the normal plan; 2 keeps every XOR / XNOR gate as its own node (the
// wires_full variant, where every wire label has to be produced).
// balance: 0 = the three slot-minimising schedules only; otherwise also the two schedules that fill the warp
// passes of the garbler / of the evaluator (plan.cpp: balanced_levels).
// hot_cap: 0 = every label lives in shared memory; otherwise at most hot_cap do (slots below info.num_hot_slots) and the
// values that idle longest per access live in the team's global-memory scratch (slots from num_hot_slots up).
int build_plan(const PlanSpec& spec, Plan& plan, std::string& err, int max_fanin = NODE_MAX_FANIN, int balance = 1,
               uint32_t hot_cap = 0, int only_policy = -1);

// build_plan; for deep, narrow circuits of which fewer than 16 instances fit per SM (sha512: 3, sha256: 8) also a second plan that
// keeps only a hot subset of the labels in shared memory so that 16 fit (hot_cap above), kept when at most a fifth of the
// label accesses are evicts / reloads.  GCB_HOT_TEAMS = 0 switches it off, N forces a target; see the measurement in plan.cpp.
// batch_hint: instances the plan will run on at once; the second plan is only considered when they do not fit the first
// plan's resident instances in one wave.
int build_best_plan(const PlanSpec& spec, Plan& plan, std::string& err, int max_fanin = NODE_MAX_FANIN, uint64_t batch_hint = ~0ull);

}  // namespace gcb

namespace gcb {
// Byte layout of the garbled tables as Garbler sends them (circuit/garbler.go:69-82) and
// Evaluator reads them (circuit/evaluator.go:40-66): u32 BE NumGates, then per gate in ORIGINAL
// order a u32 BE row count followed by the rows as 16-byte big-endian labels.  The template holds
// every count with the rows zeroed; row_pos[r] is the offset of slab row r.
struct DevWireLayout {
    uint8_t* tmpl = nullptr;
    uint32_t* row_pos = nullptr;
    ~DevWireLayout();
};
}  // namespace gcb

struct gcb_plan {
    uint64_t uid = 0;                         // unique per created plan (caches must not trust a recycled address)
    gcb::Plan p;                              // flattened free wires (the fast plan)
    mutable std::mutex wire_mu;
    mutable std::map<int, std::shared_ptr<gcb::DevWireLayout>> wire_dev;   // per device
    mutable std::mutex full_mu;
    mutable std::unique_ptr<gcb::Plan> full;  // every wire materialised (wires_full requests), built on demand
    mutable std::mutex many_mu;
    mutable std::unique_ptr<gcb::Plan> many;  // hot / cold plan for batches that overflow `p`'s resident instances, on demand
    mutable bool many_tried = false;
};

namespace gcb {
// Byte layout of the record stream Streaming.Garble emits for one sub-circuit
// (circuit/stream_garble.go:391-446): the template holds every header byte
// with the garbled rows zeroed; row_pos[r] is the stream offset of slab row r.
struct StreamLayout {
    std::vector<uint8_t> tmpl;
    std::vector<uint32_t> row_pos;
};
// One gate record parsed back from a record stream (the evaluator's view): wire indices with
// their tmp flags (stream_evaluator.go:272-343).
struct StreamGate { uint32_t a, b, c; uint8_t op, a_tmp, b_tmp, c_tmp; };
int parse_stream(const uint8_t* buf, size_t len, uint32_t ngates, std::vector<StreamGate>& gates,
                 std::vector<uint32_t>& row_pos, size_t* consumed, std::string& err);
int build_stream_layout(const std::vector<gcb_gate>& gates, uint32_t num_wires, const uint32_t* in, uint32_t nin,
                        const uint32_t* out, uint32_t nout, StreamLayout& lay, std::string& err);
}  // namespace gcb
