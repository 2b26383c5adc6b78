// plan.hpp -- host-side circuit plan compiler.
//
// Turns Circuit.Gates (circuit/circuit.go:120-131, 260-266) into the static
// schedule the kernels walk: dependency steps, per-gate tweak ids and slab row
// offsets in ORIGINAL gate order (circuit/garble.go:285-299,357-359,419-420,
// 451-452), and on-chip wire slots assigned from liveness.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gcb200.h"

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

// One free gate (XOR / XNOR): 8 bytes.  `wave` orders the gates of one 32-gate
// chunk: a gate of wave w may read what gates of waves < w of its chunk wrote.
struct FreeRec {
    uint16_t a, b, c;
    uint8_t op;
    uint8_t wave;
};
static_assert(sizeof(FreeRec) == 8, "FreeRec must be 8 bytes");

// One phase = every free gate that becomes computable after the previous cipher
// level (sorted by dependency sub-level, executed by ONE warp of the team with
// warp-level synchronisation only), followed by one level of ciphered gates
// (executed by the whole team): AND/OR first, then INV.
struct PhaseRec {
    uint32_t free_chunk, n_chunks;    // 32-record chunks of the FreeRec array (runs are padded with
                                      // op = FREE_PAD records to whole chunks)
    uint32_t cipher_first, n_quad;    // range in the GateRec array: n_quad AND/OR gates ...
    uint32_t n_inv;                   // ... then n_inv INV gates
    uint32_t n_free;                  // free gates of the phase (without padding)
    uint32_t pad[2];
};
constexpr uint8_t FREE_PAD = 0xff;
static_assert(sizeof(PhaseRec) == 32, "PhaseRec must be 32 bytes");

// One dependency step of the schedule (host-side bookkeeping: liveness, stats).
struct StepRec {
    uint32_t first, n_free, n_quad, n_inv;
};

struct SlotRef {
    uint32_t slot;
    uint32_t index;       // index into the caller's live-in / live-out array
};

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

struct DevicePlan {                   // per-device copy of the tables
    int device = -1;
    PhaseRec* phases = nullptr;       // + two zero records of padding
    FreeRec* frecs = nullptr;         // + 256 zero records of padding (chunk prefetch runs ahead)
    GateRec* crecs = nullptr;
    uint32_t* fout_wire = nullptr;    // original output wire of each free / ciphered gate
    uint32_t* cout_wire = nullptr;
    SlotRef* live_in = nullptr;
    SlotRef* live_out = nullptr;
    ~DevicePlan();
};

struct Plan {
    gcb_plan_info info{};
    uint32_t ilp = 1;                     // AES blocks a thread interleaves (kernel variant)
    uint32_t stagger = 0;                 // SM cycles between the starts of consecutive teams
    std::vector<PhaseRec> phases;
    std::vector<FreeRec> frecs;           // free gates in schedule order
    std::vector<GateRec> crecs;           // ciphered gates in schedule order
    std::vector<uint32_t> fout_wire, cout_wire;   // original output wire of frecs[i] / crecs[i]
    std::vector<StepRec> steps;
    std::vector<SlotRef> live_in, live_out;
    std::vector<uint32_t> row_off;        // num_gates+1, original order
    std::vector<uint8_t> ops;             // original order
    std::vector<gcb_gate> gates;          // kept for streaming (header templates)

    mutable std::mutex mu;
    mutable std::map<int, std::shared_ptr<DevicePlan>> dev;
};

// Returns GCB_OK or a negative status; message in err.
int build_plan(const PlanSpec& spec, Plan& plan, std::string& err);

}  // namespace gcb

struct gcb_plan {
    gcb::Plan p;
};

namespace gcb {
// Byte layout of the record stream Streaming.Garble emits for one sub-circuit
// (circuit/stream_garble.go:391-446): the template holds every header byte
// with the garbled rows zeroed; row_pos[r] is the stream offset of slab row r.
struct StreamLayout {
    std::vector<uint8_t> tmpl;
    std::vector<uint32_t> row_pos;
};
int build_stream_layout(const std::vector<gcb_gate>& gates, uint32_t num_wires, const uint32_t* in, uint32_t nin,
                        const uint32_t* out, uint32_t nout, StreamLayout& lay, std::string& err);
}  // namespace gcb
