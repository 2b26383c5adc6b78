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

// One gate as the kernels read it (16 bytes, one 128-bit load).
struct GateRec {
    uint16_t a, b, c;     // wire slots of in0, in1, out
    uint8_t op;
    uint8_t pad;
    uint32_t tweak;       // first tweak id of the gate
    uint32_t row;         // first slab row (or stream byte offset in streaming mode)
};
static_assert(sizeof(GateRec) == 16, "GateRec must be 16 bytes");

// One dependency step: gates [first, first+n_free+n_quad+n_inv) of the sorted
// gate array; free (XOR/XNOR) gates first, then AND/OR, then INV.
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

struct DevicePlan {                   // per-device copy of the tables (plan_dev.cu)
    int device = -1;
    GateRec* recs = nullptr;
    StepRec* steps = nullptr;
    uint32_t* out_wire = nullptr;
    SlotRef* live_in = nullptr;
    SlotRef* live_out = nullptr;
    ~DevicePlan();
};

struct Plan {
    gcb_plan_info info{};
    uint32_t ilp = 1;                     // AES blocks a thread interleaves (kernel variant)
    std::vector<GateRec> recs;            // sorted by (step, class, original index)
    std::vector<uint32_t> out_wire;       // original output wire of recs[i]
    std::vector<uint32_t> orig_index;     // original gate index of recs[i]
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
