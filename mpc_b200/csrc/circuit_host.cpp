// circuit_host.cpp -- the host-side circuit front end of the engine for callers that are not Go:
// the reference's two file formats, its level statistics and its plaintext evaluation, so that a
// circuit file can be turned into a plan (and checked) through the C ABI alone.
//
//   parse_bristol  follows circuit/parser.go:265-494 (text: "ngates nwires", inputs line, outputs
//                  line, one gate per line "nin nout in.. out OP"; blank lines are skipped)
//   parse_mpclc    follows circuit/parser.go:71-211 (big-endian binary: magic, ngates, nwires,
//                  ninputs, noutputs, IOArg*, then gates as op byte + 3 or 2 u32)
//   levels         follows Circuit.AssignLevels(TargetYao), circuit/circuit.go:206-254
//   compute        follows Circuit.Compute, circuit/computer.go:15-91, on wire bits
// No device code; error texts are the reference's where it has one.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gcb200.h"
#include "plan.hpp"

namespace gcb { int fail(int code, const char* fmt, ...); }
using gcb::fail;

struct gcb_circuit {
    std::vector<gcb_gate> gates;
    uint32_t num_wires = 0;
    std::vector<uint32_t> inputs, outputs;          // bits per argument
    uint32_t num_inputs = 0, num_outputs = 0;       // IO.Size(): total bits
    uint32_t num_levels = 0, max_width = 0;         // Stats[NumLevels], Stats[MaxWidth]
    uint32_t count[5] = {0, 0, 0, 0, 0};
};

namespace {

enum : uint8_t { OP_XOR = 0, OP_XNOR = 1, OP_AND = 2, OP_OR = 3, OP_INV = 4 };

// the parser's "wire seen" checks (parser.go:97-104,140-160,199-204) + levels
int finish(gcb_circuit& c) {
    const uint32_t nw = c.num_wires;
    c.num_inputs = c.num_outputs = 0;
    for (uint32_t b : c.inputs) c.num_inputs += b;
    for (uint32_t b : c.outputs) c.num_outputs += b;
    if (c.num_inputs > nw || c.num_outputs > nw) return fail(GCB_E_WIRE, "more I/O wires than wires");
    std::vector<uint8_t> seen(nw, 0);
    for (uint32_t i = 0; i < c.num_inputs; i++) seen[i] = 1;
    std::vector<uint32_t> level(nw, 0), by_level(nw ? nw : 1, 0);
    uint32_t max = 0;
    for (size_t i = 0; i < c.gates.size(); i++) {
        gcb_gate& g = c.gates[i];
        if (g.op > OP_INV) return fail(GCB_E_BADOP, "unsupported gate type %u", g.op);
        if (g.in0 >= nw || !seen[g.in0]) return fail(GCB_E_WIRE, "input %u of gate %zu not set", g.in0, i);
        if (g.op != OP_INV && (g.in1 >= nw || !seen[g.in1])) return fail(GCB_E_WIRE, "input %u of gate %zu not set", g.in1, i);
        if (g.out >= nw) return fail(GCB_E_WIRE, "wire %u out of range", g.out);
        seen[g.out] = 1;
        uint32_t l = level[g.in0];
        if (g.op != OP_INV && level[g.in1] > l) l = level[g.in1];
        g.level = l;
        by_level[l]++;
        level[g.out] = l + 1;                        // TargetYao
        if (l + 1 > max) max = l + 1;
        c.count[g.op]++;
    }
    for (uint32_t w = 0; w < nw; w++)
        if (!seen[w]) return fail(GCB_E_WIRE, "wire %u not assigned", w);
    c.num_levels = max;
    c.max_width = 0;
    for (uint32_t n : by_level) if (n > c.max_width) c.max_width = n;
    return GCB_OK;
}

struct Tokens {                                      // one non-blank line, split on white space
    std::vector<std::string> t;
};
bool to_u32(const std::string& s, uint32_t* v) {
    if (s.empty() || s.size() > 10) return false;
    uint64_t x = 0;
    for (char ch : s) { if (ch < '0' || ch > '9') return false; x = x * 10 + (uint64_t)(ch - '0'); }
    if (x > 0xffffffffull) return false;
    *v = (uint32_t)x;
    return true;
}

int parse_bristol(const char* p, size_t len, gcb_circuit& c) {
    std::vector<Tokens> lines;
    size_t i = 0;
    while (i < len) {
        Tokens ln;
        while (i < len && p[i] != '\n') {
            while (i < len && p[i] != '\n' && (p[i] == ' ' || p[i] == '\t' || p[i] == '\r')) i++;
            size_t s = i;
            while (i < len && p[i] != '\n' && p[i] != ' ' && p[i] != '\t' && p[i] != '\r') i++;
            if (i > s) ln.t.emplace_back(p + s, i - s);
        }
        if (i < len) i++;
        if (!ln.t.empty()) lines.push_back(std::move(ln));     // readLine skips blank lines
    }
    uint32_t ng, nw, niv, nov;
    if (lines.size() < 3 || lines[0].t.size() != 2 || !to_u32(lines[0].t[0], &ng) || !to_u32(lines[0].t[1], &nw))
        return fail(GCB_E_CORRUPT, "invalid 1st line");
    if (!to_u32(lines[1].t[0], &niv) || 1 + (size_t)niv != lines[1].t.size()) return fail(GCB_E_CORRUPT, "invalid inputs line");
    uint32_t total = 0;
    for (size_t k = 1; k < lines[1].t.size(); k++) {
        uint32_t b;
        if (!to_u32(lines[1].t[k], &b)) return fail(GCB_E_CORRUPT, "invalid inputs line");
        c.inputs.push_back(b);
        total += b;
    }
    if (total == 0) return fail(GCB_E_CORRUPT, "no inputs defined");
    if (!to_u32(lines[2].t[0], &nov) || 1 + (size_t)nov != lines[2].t.size()) return fail(GCB_E_CORRUPT, "invalid outputs line");
    for (size_t k = 1; k < lines[2].t.size(); k++) {
        uint32_t b;
        if (!to_u32(lines[2].t[k], &b)) return fail(GCB_E_CORRUPT, "invalid outputs line");
        c.outputs.push_back(b);
    }
    c.num_wires = nw;
    c.gates.reserve(ng);
    for (size_t li = 3; li < lines.size(); li++) {
        const std::vector<std::string>& t = lines[li].t;
        if (c.gates.size() >= ng) return fail(GCB_E_CORRUPT, "too many gates");
        uint32_t n1, n2;
        if (t.size() < 3 || !to_u32(t[0], &n1) || !to_u32(t[1], &n2) || 2 + (size_t)n1 + n2 + 1 != t.size())
            return fail(GCB_E_CORRUPT, "invalid gate: line %zu", li + 1);
        const std::string& name = t.back();
        uint8_t op;
        if (name == "XOR") op = OP_XOR; else if (name == "XNOR") op = OP_XNOR; else if (name == "AND") op = OP_AND;
        else if (name == "OR") op = OP_OR; else if (name == "INV") op = OP_INV;
        else return fail(GCB_E_BADOP, "invalid operation '%s'", name.c_str());
        if (n1 != (op == OP_INV ? 1u : 2u)) return fail(GCB_E_CORRUPT, "invalid number of inputs %u for %s", n1, name.c_str());
        if (n2 != 1) return fail(GCB_E_CORRUPT, "invalid number of outputs %u for %s", n2, name.c_str());
        gcb_gate g{};
        g.op = op;
        if (!to_u32(t[2], &g.in0) || (n1 > 1 && !to_u32(t[3], &g.in1)) || !to_u32(t[2 + n1], &g.out))
            return fail(GCB_E_CORRUPT, "invalid gate: line %zu", li + 1);
        c.gates.push_back(g);
    }
    if (c.gates.size() != ng) return fail(GCB_E_CORRUPT, "not enough gates: got %zu, expected %u", c.gates.size(), ng);
    return finish(c);
}

struct Reader {
    const uint8_t* p; size_t len, pos = 0; bool ok = true;
    uint32_t u32() {
        if (pos + 4 > len) { ok = false; return 0; }
        const uint32_t v = ((uint32_t)p[pos] << 24) | ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3];
        pos += 4;
        return v;
    }
    void skip_str() { const uint32_t n = u32(); if (!ok || pos + n > len) { ok = false; return; } pos += n; }
};
// IOArg (parser.go:213-262): name, type string, bits, compound count, compounds recursively
uint32_t read_ioarg(Reader& r, int depth = 0) {
    r.skip_str(); r.skip_str();
    const uint32_t bits = r.u32(), ncomp = r.u32();
    for (uint32_t k = 0; r.ok && k < ncomp && depth < 64; k++) read_ioarg(r, depth + 1);
    return bits;
}
int parse_mpclc(const uint8_t* p, size_t len, gcb_circuit& c) {
    Reader r{p, len};
    (void)r.u32();                                   // magic
    const uint32_t ng = r.u32(), nw = r.u32(), ni = r.u32(), no = r.u32();
    if (!r.ok) return fail(GCB_E_CORRUPT, "unexpected end of file");
    for (uint32_t k = 0; k < ni && r.ok; k++) c.inputs.push_back(read_ioarg(r));
    for (uint32_t k = 0; k < no && r.ok; k++) c.outputs.push_back(read_ioarg(r));
    if (!r.ok) return fail(GCB_E_CORRUPT, "unexpected end of file");
    c.num_wires = nw;
    c.gates.reserve(ng);
    while (r.pos < len) {
        gcb_gate g{};
        g.op = p[r.pos++];
        if (g.op <= OP_OR) { g.in0 = r.u32(); g.in1 = r.u32(); g.out = r.u32(); }
        else if (g.op == OP_INV) { g.in0 = r.u32(); g.out = r.u32(); }
        else return fail(GCB_E_BADOP, "unsupported gate type %u", g.op);
        if (!r.ok) return fail(GCB_E_CORRUPT, "unexpected end of file");
        c.gates.push_back(g);
    }
    if (c.gates.size() != ng) return fail(GCB_E_CORRUPT, "not enough gates: got %zu, expected %u", c.gates.size(), ng);
    return finish(c);
}

}  // namespace

extern "C" {

int gcb_circuit_parse(const void* data, size_t len, int format, gcb_circuit** out) {
    GCB_TRY
    if (!out || (!data && len)) return fail(GCB_E_ARG, "null argument");
    *out = nullptr;
    auto* c = new gcb_circuit();
    int rc;
    if (format == GCB_FORMAT_BRISTOL) rc = parse_bristol(static_cast<const char*>(data), len, *c);
    else if (format == GCB_FORMAT_MPCLC) rc = parse_mpclc(static_cast<const uint8_t*>(data), len, *c);
    else rc = fail(GCB_E_ARG, "unsupported circuit format");
    if (rc) { delete c; return rc; }
    *out = c;
    return GCB_OK;
    GCB_CATCH
}
int gcb_circuit_from_gates(const gcb_gate* gates, uint32_t num_gates, uint32_t num_wires, const uint32_t* inputs,
                           uint32_t n_input_args, const uint32_t* outputs, uint32_t n_output_args, gcb_circuit** out) {
    GCB_TRY
    if (!out || (!gates && num_gates) || (!inputs && n_input_args) || (!outputs && n_output_args))
        return fail(GCB_E_ARG, "null argument");
    *out = nullptr;
    auto* c = new gcb_circuit();
    c->gates.assign(gates, gates + num_gates);
    c->num_wires = num_wires;
    c->inputs.assign(inputs, inputs + n_input_args);
    c->outputs.assign(outputs, outputs + n_output_args);
    const int rc = finish(*c);
    if (rc) { delete c; return rc; }
    *out = c;
    return GCB_OK;
    GCB_CATCH
}
void gcb_circuit_destroy(gcb_circuit* c) { delete c; }

int gcb_circuit_get_info(const gcb_circuit* c, gcb_circuit_info* info) {
    GCB_TRY
    if (!c || !info) return fail(GCB_E_ARG, "null argument");
    *info = gcb_circuit_info{};
    info->num_gates = (uint32_t)c->gates.size();
    info->num_wires = c->num_wires;
    info->num_inputs = c->num_inputs; info->num_outputs = c->num_outputs;
    info->num_input_args = (uint32_t)c->inputs.size(); info->num_output_args = (uint32_t)c->outputs.size();
    info->num_xor = c->count[OP_XOR]; info->num_xnor = c->count[OP_XNOR]; info->num_and = c->count[OP_AND];
    info->num_or = c->count[OP_OR]; info->num_inv = c->count[OP_INV];
    info->num_levels = c->num_levels; info->max_width = c->max_width;
    return GCB_OK;
    GCB_CATCH
}
int gcb_circuit_get_gates(const gcb_circuit* c, gcb_gate* gates) {
    GCB_TRY
    if (!c || (!gates && !c->gates.empty())) return fail(GCB_E_ARG, "null argument");
    if (!c->gates.empty()) memcpy(gates, c->gates.data(), c->gates.size() * sizeof(gcb_gate));
    return GCB_OK;
    GCB_CATCH
}
int gcb_circuit_get_io(const gcb_circuit* c, uint32_t* input_bits, uint32_t* output_bits) {
    GCB_TRY
    if (!c) return fail(GCB_E_ARG, "null argument");
    if (input_bits && !c->inputs.empty()) memcpy(input_bits, c->inputs.data(), c->inputs.size() * 4);
    if (output_bits && !c->outputs.empty()) memcpy(output_bits, c->outputs.data(), c->outputs.size() * 4);
    return GCB_OK;
    GCB_CATCH
}
int gcb_circuit_compute(const gcb_circuit* c, uint32_t batch, const uint8_t* in_bits, uint8_t* out_bits) {
    GCB_TRY
    if (!c || (batch && ((!in_bits && c->num_inputs) || (!out_bits && c->num_outputs)))) return fail(GCB_E_ARG, "null argument");
    std::vector<uint8_t> w(c->num_wires);
    for (uint32_t b = 0; b < batch; b++) {
        const uint8_t* in = in_bits + (size_t)b * c->num_inputs;
        for (uint32_t i = 0; i < c->num_inputs; i++) w[i] = in[i] & 1;
        for (const gcb_gate& g : c->gates) {         // computer.go:44-76
            uint8_t r;
            switch (g.op) {
                case OP_XOR: r = w[g.in0] ^ w[g.in1]; break;
                case OP_XNOR: r = (uint8_t)((w[g.in0] ^ w[g.in1]) == 0); break;
                case OP_AND: r = w[g.in0] & w[g.in1]; break;
                case OP_OR: r = w[g.in0] | w[g.in1]; break;
                default: r = (uint8_t)(w[g.in0] == 0); break;      // INV
            }
            w[g.out] = r;
        }
        if (c->num_outputs) memcpy(out_bits + (size_t)b * c->num_outputs, w.data() + (c->num_wires - c->num_outputs), c->num_outputs);
    }
    return GCB_OK;
    GCB_CATCH
}
int gcb_circuit_plan(const gcb_circuit* c, gcb_plan** out) {
    GCB_TRY
    if (!c) return fail(GCB_E_ARG, "null argument");
    return gcb_plan_create(c->gates.data(), (uint32_t)c->gates.size(), c->num_wires, c->num_inputs, c->num_outputs, out);
    GCB_CATCH
}

}  // extern "C"
