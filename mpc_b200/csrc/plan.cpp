// plan.cpp -- host-side circuit plan compiler (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <cstdio>
#include <functional>
#include <limits>
#include <queue>

namespace gcb {

namespace {
constexpr uint32_t kCipherSub = 0xffffffffu;   // sub-level of the cipher gates of a level
constexpr int64_t kForever = std::numeric_limits<int64_t>::max();

inline int op_class(uint8_t op) { return op <= OP_XNOR ? 0 : (op == OP_INV ? 2 : 1); }
}  // namespace

int build_plan(const PlanSpec& spec, Plan& plan, std::string& err) {
    char msg[160];
    const uint32_t ng = spec.num_gates, nw = spec.num_wires;
    const bool identity = spec.loc.empty();
    const uint32_t nloc = identity ? nw : spec.num_locs;
    if (!identity && spec.loc.size() != nw) { err = "location map size mismatch"; return GCB_E_ARG; }
    auto loc_of = [&](uint32_t w) { return identity ? w : spec.loc[w]; };

    const size_t ninit = spec.live_in.size();
    const size_t ndefs = ninit + ng;
    // A definition is one value of one location: the initial value of a live-in
    // location, or the output of one gate.  Re-assigned locations (legal in the
    // file formats, and the norm for aliased streaming wires) get a new
    // definition each time, which makes every hazard a plain RAW dependency.
    std::vector<int64_t> cur_def(nloc, -1);
    std::vector<uint32_t> cd(ndefs, 0), xd(ndefs, 0);       // cipher depth, free sub-depth
    std::vector<int64_t> born(ndefs, -1), last(ndefs, -1);
    for (size_t k = 0; k < ninit; k++) {
        const uint32_t l = spec.live_in[k];
        if (l >= nloc) { err = "live-in location out of range"; return GCB_E_WIRE; }
        if (cur_def[l] >= 0) { err = "duplicate live-in location"; return GCB_E_ARG; }
        cur_def[l] = (int64_t)k;
    }

    std::vector<uint64_t> key(ng);
    std::vector<int64_t> def_a(ng), def_b(ng);
    uint32_t n_and = 0, n_or = 0, n_inv = 0, n_free = 0;
    plan.row_off.assign(ng + 1, 0);
    plan.ops.resize(ng);
    std::vector<uint32_t> tweak_of(ng);
    uint32_t tweak = 0, row = 0;
    for (uint32_t i = 0; i < ng; i++) {
        const gcb_gate& g = spec.gates[i];
        if (g.op > OP_INV) {
            snprintf(msg, sizeof msg, "invalid gate type %u (gate %u)", g.op, i);
            err = msg;
            return GCB_E_BADOP;
        }
        const bool unary = g.op == OP_INV;
        if (g.in0 >= nw || (!unary && g.in1 >= nw) || g.out >= nw) {
            snprintf(msg, sizeof msg, "invalid wire in gate %u [0...%u[", i, nw);
            err = msg;
            return GCB_E_WIRE;
        }
        const int64_t da = cur_def[loc_of(g.in0)];
        const int64_t db = unary ? da : cur_def[loc_of(g.in1)];
        if (da < 0 || db < 0) {
            snprintf(msg, sizeof msg, "input %u of gate %u not set", da < 0 ? g.in0 : g.in1, i);
            err = msg;
            return GCB_E_WIRE;
        }
        uint32_t d, x;
        if (cd[da] > cd[db]) { d = cd[da]; x = xd[da]; }
        else if (cd[db] > cd[da]) { d = cd[db]; x = xd[db]; }
        else { d = cd[da]; x = std::max(xd[da], xd[db]); }
        const size_t dout = ninit + i;
        if (g.op >= OP_AND) {
            key[i] = ((uint64_t)d << 32) | kCipherSub;
            cd[dout] = d + 1; xd[dout] = 0;
        } else {
            key[i] = ((uint64_t)d << 32) | x;
            cd[dout] = d; xd[dout] = x + 1;
        }
        def_a[i] = da; def_b[i] = db;
        cur_def[loc_of(g.out)] = (int64_t)dout;
        plan.ops[i] = g.op;
        plan.row_off[i] = row;
        tweak_of[i] = tweak;
        switch (g.op) {
            case OP_AND: tweak += 2; row += 2; n_and++; break;
            case OP_OR: tweak += 1; row += 3; n_or++; break;
            case OP_INV: tweak += 1; row += 1; n_inv++; break;
            default: n_free++; break;
        }
    }
    plan.row_off[ng] = row;

    // ---- steps: sort gates by (level key, class, original index)
    std::vector<uint32_t> order(ng);
    for (uint32_t i = 0; i < ng; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t p, uint32_t q) {
        if (key[p] != key[q]) return key[p] < key[q];
        const int cp = op_class(spec.gates[p].op), cq = op_class(spec.gates[q].op);
        if (cp != cq) return cp < cq;
        return p < q;
    });
    std::vector<int64_t> step_of(ng);
    plan.steps.clear();
    for (uint32_t pos = 0; pos < ng;) {
        uint32_t end = pos;
        StepRec st{pos, 0, 0, 0};
        while (end < ng && key[order[end]] == key[order[pos]]) {
            const int c = op_class(spec.gates[order[end]].op);
            (c == 0 ? st.n_free : c == 1 ? st.n_quad : st.n_inv)++;
            step_of[order[end]] = (int64_t)plan.steps.size();
            end++;
        }
        plan.steps.push_back(st);
        pos = end;
    }
    const int64_t nsteps = (int64_t)plan.steps.size();

    // ---- liveness
    for (uint32_t i = 0; i < ng; i++) {
        const int64_t s = step_of[i];
        born[ninit + i] = s;
        last[def_a[i]] = std::max(last[def_a[i]], s);
        last[def_b[i]] = std::max(last[def_b[i]], s);
    }
    plan.live_out.clear();
    for (size_t k = 0; k < spec.live_out.size(); k++) {
        const uint32_t l = spec.live_out[k];
        if (l >= nloc || cur_def[l] < 0) {
            snprintf(msg, sizeof msg, "wire %u not assigned", l);
            err = msg;
            return GCB_E_WIRE;
        }
        last[cur_def[l]] = kForever;
    }
    for (size_t d = 0; d < ndefs; d++) last[d] = std::max(last[d], born[d]);

    // ---- slots: lowest free index first; a slot whose value was last read in
    // step s may be rewritten from step s+1 on (reads and writes of one step are
    // unordered).
    std::vector<std::vector<size_t>> expire((size_t)nsteps + 1);
    std::vector<uint32_t> slot(ndefs, 0);
    std::priority_queue<uint32_t, std::vector<uint32_t>, std::greater<uint32_t>> free_slots;
    uint32_t next_slot = 0;
    plan.live_in.clear();
    for (size_t k = 0; k < ninit; k++) {
        slot[k] = next_slot++;
        plan.live_in.push_back(SlotRef{slot[k], (uint32_t)k});
        if (last[k] < 0) free_slots.push(slot[k]);                 // never read, not live-out
        else if (last[k] != kForever) expire[(size_t)last[k]].push_back(k);
    }
    for (int64_t s = 0; s < nsteps; s++) {
        if (s > 0) {
            for (size_t d : expire[(size_t)s - 1]) free_slots.push(slot[d]);
            expire[(size_t)s - 1].clear();
        }
        const StepRec& st = plan.steps[(size_t)s];
        const uint32_t n = st.n_free + st.n_quad + st.n_inv;
        for (uint32_t j = 0; j < n; j++) {
            const size_t d = ninit + order[st.first + j];
            if (free_slots.empty()) slot[d] = next_slot++;
            else { slot[d] = free_slots.top(); free_slots.pop(); }
            if (last[d] != kForever) expire[(size_t)last[d]].push_back(d);
        }
    }
    if (next_slot > 65535) {
        snprintf(msg, sizeof msg, "circuit needs %u live wire slots (limit 65535)", next_slot);
        err = msg;
        return GCB_E_TOO_LARGE;
    }
    for (size_t k = 0; k < spec.live_out.size(); k++)
        plan.live_out.push_back(SlotRef{slot[(size_t)cur_def[spec.live_out[k]]], (uint32_t)k});

    // ---- gate records in schedule order, grouped into phases
    plan.phases.clear(); plan.frecs.clear(); plan.crecs.clear();
    plan.fout_wire.clear(); plan.cout_wire.clear();
    {
        PhaseRec ph{};
        bool open = false;
        uint32_t run_start = 0;                       // index in frecs of the phase's first free gate
        std::vector<uint32_t> rank_of;                // sub-level rank of every free gate of the open phase
        auto close_phase = [&]() {
            // waves: rank of the gate's sub-level minus the rank of its chunk's first gate
            for (uint32_t i = 0; i < ph.n_free; i++) {
                const uint32_t chunk_first = i & ~31u;
                plan.frecs[run_start + i].wave = (uint8_t)(rank_of[i] - rank_of[chunk_first]);
            }
            while (plan.frecs.size() % 32) {          // whole chunks
                plan.frecs.push_back(FreeRec{0, 0, 0, FREE_PAD, 0});
                plan.fout_wire.push_back(0);
            }
            ph.n_chunks = (ph.n_free + 31) / 32;
            plan.phases.push_back(ph);
            ph = PhaseRec{};
            rank_of.clear();
            open = false;
        };
        uint32_t rank = 0;
        for (size_t si = 0; si < plan.steps.size(); si++) {
            const StepRec& st = plan.steps[si];
            if (!open) {
                run_start = (uint32_t)plan.frecs.size();
                ph.free_chunk = run_start / 32;
                ph.cipher_first = (uint32_t)plan.crecs.size();
                rank = 0;
                open = true;
            }
            const uint32_t n = st.n_free + st.n_quad + st.n_inv;
            for (uint32_t j = 0; j < n; j++) {
                const uint32_t i = order[st.first + j];
                const gcb_gate& g = spec.gates[i];
                const uint16_t sa = (uint16_t)slot[(size_t)def_a[i]], sb = (uint16_t)slot[(size_t)def_b[i]];
                const uint16_t sc = (uint16_t)slot[ninit + i];
                if (g.op <= OP_XNOR) {
                    plan.frecs.push_back(FreeRec{sa, sb, sc, g.op, 0});
                    plan.fout_wire.push_back(g.out);
                    rank_of.push_back(rank);
                    ph.n_free++;
                } else {
                    plan.crecs.push_back(GateRec{sa, sb, sc, g.op, 0, tweak_of[i], plan.row_off[i]});
                    plan.cout_wire.push_back(g.out);
                    (g.op == OP_INV ? ph.n_inv : ph.n_quad)++;
                }
            }
            if (st.n_free) rank++;
            if (st.n_quad + st.n_inv) { ph.cipher_first = (uint32_t)plan.crecs.size() - st.n_quad - st.n_inv; close_phase(); }
        }
        if (open) close_phase();
    }

    gcb_plan_info& in = plan.info;
    in = gcb_plan_info{};
    in.num_gates = ng;
    in.num_wires = nw;
    in.num_inputs = (uint32_t)ninit;
    in.num_outputs = (uint32_t)spec.live_out.size();
    in.num_rows = row;
    in.num_tweaks = tweak;
    in.num_steps = (uint32_t)nsteps;
    in.num_slots = next_slot;
    in.num_and = n_and; in.num_or = n_or; in.num_inv = n_inv; in.num_free = n_free;
    in.garble_hashes = 4 * n_and + 4 * n_or + 2 * n_inv;
    in.eval_hashes = 2 * n_and + n_or + n_inv;
    return GCB_OK;
}

}  // namespace gcb

namespace gcb {

int build_stream_layout(const std::vector<gcb_gate>& gates, uint32_t num_wires, const uint32_t* in, uint32_t nin,
                        const uint32_t* out, uint32_t nout, StreamLayout& lay, std::string& err) {
    if (nout > num_wires) { err = "more output ids than wires"; return GCB_E_ARG; }
    const uint32_t first_tmp = nin, first_out = num_wires - nout;          // stream_garble.go:112-113
    lay.tmpl.clear();
    lay.row_pos.clear();
    lay.tmpl.reserve(gates.size() * 10);
    // Streaming.Get / Set index resolution (stream_garble.go:131-157)
    auto resolve = [&](uint32_t w, uint32_t& index, bool& tmp) {
        if (w < first_tmp) { index = in[w]; tmp = false; }
        else if (w >= first_out) { index = out[w - first_out]; tmp = false; }
        else { index = w; tmp = true; }
    };
    auto put16 = [&](uint32_t v) { lay.tmpl.push_back((uint8_t)(v >> 8)); lay.tmpl.push_back((uint8_t)v); };
    auto put32 = [&](uint32_t v) { put16(v >> 16); put16(v & 0xffff); };
    for (size_t i = 0; i < gates.size(); i++) {
        const gcb_gate& g = gates[i];
        if (g.op > OP_INV) { err = "invalid gate type"; return GCB_E_BADOP; }
        uint32_t ai = 0, bi = 0, ci = 0;
        bool at = false, bt = false, ct = false;
        const bool unary = g.op == OP_INV;
        if (!unary) resolve(g.in1, bi, bt);
        resolve(g.in0, ai, at);
        resolve(g.out, ci, ct);
        uint8_t op = g.op;
        if (at) op |= 0x80;
        if (bt) op |= 0x40;
        if (ct) op |= 0x20;
        const bool shrt = ai <= 0xffff && bi <= 0xffff && ci <= 0xffff;
        if (shrt) op |= 0x10;
        lay.tmpl.push_back(op);
        if (shrt) { put16(ai); if (!unary) put16(bi); put16(ci); }
        else { put32(ai); if (!unary) put32(bi); put32(ci); }
        const int rows = g.op == OP_AND ? 2 : g.op == OP_OR ? 3 : g.op == OP_INV ? 1 : 0;
        for (int k = 0; k < rows; k++) {
            if (lay.tmpl.size() > 0xfffffff0u) { err = "record stream exceeds 4 GiB"; return GCB_E_TOO_LARGE; }
            lay.row_pos.push_back((uint32_t)lay.tmpl.size());
            lay.tmpl.insert(lay.tmpl.end(), 16, 0);
        }
    }
    return GCB_OK;
}

}  // namespace gcb
