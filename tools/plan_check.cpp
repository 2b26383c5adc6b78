// plan_check.cpp -- host-only interpreter of compiled plans: the plan compiler checked without a GPU.
//
// Builds plans of a circuit with the product's compiler (plan.cpp) under many settings -- every schedule policy,
// fan-in limits, hot caps (whole-life hot / cold and split live ranges) -- and EXECUTES each plan symbolically the
// way the gate kernels do (gc_kernels.cuh): live-in loads, per phase the evict / reload lists, the node waves, the
// cipher level, live-out reads.  Every wire carries a 64-bit fingerprint (inputs and ciphered gate outputs: random;
// XOR: the XOR of its inputs; XNOR: that XOR a fingerprint of R), so a plan is right exactly when
//   * every ciphered gate finds the fingerprints of its two input wires in the slots it reads, carries the static
//     tweak id and slab row of its ORIGINAL position (circuit/garble.go:357-359,419-420,451-452) and the right op;
//   * every node's leaves XOR (plus R for odd XNOR parity) to the fingerprint of the wire it materialises;
//   * the caller's outputs are where live_out says;
//   * no slot is read and written within one unordered step (a wave, a cipher level), nothing reads the target of a
//     reload before the phase that queued it has ended, and nothing a reload reads is evicted over in the same phase;
//   * slots stay below the plan's bounds (hot slots below num_hot_slots when live ranges are split).
//
//   g++ -O2 -std=c++17 -pthread -I include -o tools/_build/plan_check tools/plan_check.cpp mpc_b200/csrc/plan.cpp
//   tools/_build/plan_check file.gates [file.gates ...]      (written by tools/dump_gates.py)
//   tools/_build/plan_check --random N                       N random circuits with all five gate types
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../mpc_b200/csrc/plan.hpp"

using namespace gcb;

namespace gcb { int fail(int code, const char*, ...) { return code; } }   // plan.cpp reports through this in the library

static const uint64_t kPoison = 0xdeadbeefdeadbeefull;

struct Circuit {
    std::vector<gcb_gate> gates;
    uint32_t num_wires = 0, num_in = 0, num_out = 0;
    std::string name;
};

static uint64_t mix(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

static int check_plan(const Circuit& c, const Plan& plan, const char* what) {
    const gcb_plan_info& in = plan.info;
    const uint32_t ng = (uint32_t)c.gates.size();
    const uint64_t RFP = mix(0x5151);
    int errors = 0;
    auto err = [&](const char* fmt, auto... a) {
        if (errors++ < 8) { fprintf(stderr, "%s [%s]: ", c.name.c_str(), what); fprintf(stderr, fmt, a...); fprintf(stderr, "\n"); }
    };
    // expected fingerprints, tweaks and rows in original order
    std::vector<uint64_t> fp(c.num_wires, 0);
    std::vector<int64_t> def_gate(c.num_wires, -1);
    std::vector<uint32_t> tweak(ng, 0);
    for (uint32_t w = 0; w < c.num_in; w++) fp[w] = mix(w + 1);
    uint32_t id = 0;
    for (uint32_t i = 0; i < ng; i++) {
        const gcb_gate& g = c.gates[i];
        tweak[i] = id;
        switch (g.op) {
        case OP_XOR: fp[g.out] = fp[g.in0] ^ fp[g.in1]; break;
        case OP_XNOR: fp[g.out] = fp[g.in0] ^ fp[g.in1] ^ RFP; break;
        case OP_AND: fp[g.out] = mix(0x1000000ull + i); id += 2; break;
        case OP_OR: fp[g.out] = mix(0x1000000ull + i); id += 1; break;
        default: fp[g.out] = mix(0x1000000ull + i); id += 1; break;
        }
        def_gate[g.out] = i;
    }
    const bool split = !plan.phase_copy.empty();
    const uint32_t n_hot = in.num_hot_slots, n_slots = in.num_slots;
    if (n_hot > n_slots) err("hot slots %u > slots %u", n_hot, n_slots);
    std::vector<uint64_t> S(n_slots, kPoison), G(split ? n_slots - n_hot : 0, kPoison);
    // a plan that is not split addresses one slot space (the kernels put slots >= n_smem in the scratch in place)
    const uint32_t direct = split ? n_hot : n_slots;
    auto rd = [&](uint32_t s, const char* who) -> uint64_t {
        if (s >= direct) { err("%s reads slot %u beyond %u", who, s, direct); return kPoison; }
        if (S[s] == kPoison) err("%s reads slot %u that holds nothing", who, s);
        return S[s];
    };
    // live-in
    for (const SlotRef& r : plan.live_in) {
        const uint32_t src = r.index & 0x7fffffffu;
        if (src >= c.num_in) { err("live-in index %u", src); continue; }
        if (split && r.slot >= n_hot) {
            if (r.slot - n_hot >= G.size()) err("live-in scratch %u", r.slot); else G[r.slot - n_hot] = fp[src];
        } else if (r.slot >= direct) err("live-in slot %u", r.slot);
        else S[r.slot] = fp[src];
    }
    if (split && plan.phase_copy.size() < plan.phases.size()) err("copy lists for %zu of %zu phases", plan.phase_copy.size(), plan.phases.size());
    std::vector<uint8_t> mark(n_slots, 0);
    std::vector<std::pair<uint32_t, uint64_t>> pending;
    std::vector<uint32_t> evict_src;                              // hot slots the evicts of this phase read: no barrier lies
    std::vector<uint8_t> seen(ng, 0);                             // between them and the first step of the phase
    for (size_t pi = 0; pi < plan.phases.size(); pi++) {
        const PhaseRec& ph = plan.phases[pi];
        pending.clear();
        if (split && pi < plan.phase_copy.size()) {
            const auto& pc = plan.phase_copy[pi];
            const uint32_t first = pc[0], n_ev = pc[1], n_rl = pc[2];
            if ((size_t)2 * (first + n_ev + n_rl) > plan.copies.size()) { err("phase %zu: copy list out of range", pi); break; }
            std::vector<uint32_t> evicted;
            evict_src.clear();
            for (uint32_t k = 0; k < n_ev; k++) {
                const uint32_t src = plan.copies[2 * (first + k)], dst = plan.copies[2 * (first + k) + 1];
                if (src >= n_hot || dst >= G.size()) { err("phase %zu: evict %u -> %u out of range", pi, src, dst); continue; }
                if (S[src] == kPoison) err("phase %zu: evict of empty slot %u", pi, src);
                G[dst] = S[src];
                evicted.push_back(dst);
                evict_src.push_back(src);
            }
            for (uint32_t k = 0; k < n_rl; k++) {
                const uint32_t src = plan.copies[2 * (first + n_ev + k)], dst = plan.copies[2 * (first + n_ev + k) + 1];
                if (dst >= n_hot || src >= G.size()) { err("phase %zu: reload %u -> %u out of range", pi, src, dst); continue; }
                for (uint32_t e : evicted) if (e == src) err("phase %zu: reload reads scratch %u that the same phase evicts to", pi, src);
                if (G[src] == kPoison) err("phase %zu: reload of empty scratch %u", pi, src);
                pending.push_back({dst, G[src]});
                S[dst] = kPoison;                              // in flight until the phase ends
            }
        }
        if (!split) evict_src.clear();
        auto is_pending = [&](uint32_t s) { for (auto& q : pending) if (q.first == s) return true; return false; };
        bool first_step = true;
        auto evict_reads = [&](uint32_t s) { if (first_step) for (uint32_t e : evict_src) if (e == s) return true; return false; };
        const uint32_t nw = ph.n_waves & 0x7fffffffu;
        for (uint32_t w = 0; w < nw; w++) {
            const WaveRec& wr = plan.waves[ph.wave_first + w];
            std::vector<std::pair<uint32_t, uint64_t>> writes;
            for (uint32_t j = 0; j < wr.count; j++) {
                const NodeRec& nd = plan.nodes[wr.first + j];
                if (nd.k > NODE_MAX_FANIN) { err("node fan-in %u", nd.k); continue; }
                uint64_t acc = (nd.parity & NODE_PARITY) ? RFP : 0;
                for (uint32_t k = 0; k < nd.k; k++) { acc ^= rd(nd.leaf[k], "node"); mark[nd.leaf[k] < n_slots ? nd.leaf[k] : 0] = 1; }
                const uint32_t wire = plan.nout_wire[wr.first + j];
                if (wire >= c.num_wires) { err("node wire %u", wire); continue; }
                if (acc != fp[wire]) err("phase %zu wave %u: node %u does not produce wire %u", pi, w, j, wire);
                if (nd.dst >= direct) { err("node dst %u", nd.dst); continue; }
                if (is_pending(nd.dst)) err("phase %zu: node writes slot %u with a reload in flight", pi, nd.dst);
                if (evict_reads(nd.dst)) err("phase %zu: node of the first wave writes slot %u that an evict of the phase reads", pi, nd.dst);
                writes.push_back({nd.dst, acc});
            }
            for (auto& q : writes) if (mark[q.first]) err("phase %zu wave %u: slot %u read and written in one wave", pi, w, q.first);
            for (uint32_t j = 0; j < wr.count; j++)
                for (uint32_t k = 0; k < plan.nodes[wr.first + j].k; k++) { const uint32_t l = plan.nodes[wr.first + j].leaf[k]; if (l < n_slots) mark[l] = 0; }
            for (auto& q : writes) S[q.first] = q.second;
            if (wr.count) first_step = false;
        }
        std::vector<std::pair<uint32_t, uint64_t>> writes;
        std::vector<uint32_t> reads;
        for (uint32_t gi = 0; gi < ph.n_quad + ph.n_inv; gi++) {
            const GateRec& g = plan.crecs[ph.cipher_first + gi];
            const uint32_t wire = plan.cout_wire[ph.cipher_first + gi];
            if (wire >= c.num_wires || def_gate[wire] < 0) { err("gate wire %u", wire); continue; }
            const uint32_t o = (uint32_t)def_gate[wire];
            const gcb_gate& og = c.gates[o];
            if (seen[o]++) err("gate %u scheduled twice", o);
            if (g.op != og.op) err("gate %u: op %u, want %u", o, g.op, og.op);
            if ((gi < ph.n_quad) != (og.op == OP_AND || og.op == OP_OR)) err("gate %u in the wrong part of its level", o);
            if (g.tweak != tweak[o]) err("gate %u: tweak %u, want %u", o, g.tweak, tweak[o]);
            if (g.row != plan.row_off[o]) err("gate %u: row %u, want %u", o, g.row, plan.row_off[o]);
            if (rd(g.a, "gate") != fp[og.in0]) err("phase %zu: gate %u input 0 is not wire %u", pi, o, og.in0);
            reads.push_back(g.a);
            if (og.op != OP_INV) { if (rd(g.b, "gate") != fp[og.in1]) err("phase %zu: gate %u input 1 is not wire %u", pi, o, og.in1); reads.push_back(g.b); }
            else if (g.b >= direct) err("INV gate %u names slot %u", o, g.b);      // the kernel loads it (unused)
            if (g.c >= direct) { err("gate dst %u", g.c); continue; }
            if (is_pending(g.c)) err("phase %zu: gate writes slot %u with a reload in flight", pi, g.c);
            if (evict_reads(g.c)) err("phase %zu: gate writes slot %u that an evict of the phase reads", pi, g.c);
            writes.push_back({g.c, fp[wire]});
        }
        for (uint32_t r : reads) if (r < n_slots) mark[r] = 1;
        for (auto& q : writes) if (mark[q.first]) err("phase %zu: slot %u read and written in one cipher level", pi, q.first);
        for (uint32_t r : reads) if (r < n_slots) mark[r] = 0;
        for (auto& q : writes) S[q.first] = q.second;
        for (auto& q : pending) S[q.first] = q.second;           // cp.async.wait_all before the phase's last barrier
    }
    for (uint32_t i = 0; i < ng; i++) if (c.gates[i].op >= OP_AND && !seen[i]) err("ciphered gate %u never scheduled", i);
    if (plan.live_out.size() != c.num_out) err("%zu live-out entries for %u outputs", plan.live_out.size(), c.num_out);
    for (const SlotRef& r : plan.live_out) {
        if (r.index >= c.num_out) { err("live-out index %u", r.index); continue; }
        const uint32_t wire = c.num_wires - c.num_out + r.index;
        uint64_t v = kPoison;
        if (split && r.slot >= n_hot) { if (r.slot - n_hot < G.size()) v = G[r.slot - n_hot]; }
        else if (r.slot < direct) v = S[r.slot];
        if (v != fp[wire]) err("output %u (wire %u) is not in slot %u", r.index, wire, r.slot);
    }
    return errors;
}

static bool load_gates(const char* path, Circuit& c) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return false; }
    uint32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) { fclose(f); return false; }
    c.gates.resize(hdr[0]);
    const bool ok = fread(c.gates.data(), sizeof(gcb_gate), hdr[0], f) == hdr[0];
    fclose(f);
    c.num_wires = hdr[1]; c.num_in = hdr[2]; c.num_out = hdr[3];
    c.name = path;
    return ok;
}

static Circuit random_circuit(uint32_t seed) {
    std::mt19937 rng(seed);
    Circuit c;
    c.num_in = 8 + rng() % 120;
    const uint32_t ng = 200 + rng() % 6000;
    c.num_out = 1 + rng() % 40;
    const uint32_t window = 4 + rng() % 300;                 // how far back a gate reaches: narrow and deep or wide and shallow
    const uint32_t p_cipher = 5 + rng() % 60;                // per cent ciphered gates
    uint32_t nw = c.num_in;
    for (uint32_t i = 0; i < ng; i++) {
        gcb_gate g{};
        const uint32_t lo = nw > window ? nw - window : 0;
        g.in0 = lo + rng() % (nw - lo);
        g.in1 = (rng() % 4 == 0) ? rng() % nw : lo + rng() % (nw - lo);
        g.out = nw++;
        const uint32_t r = rng() % 100;
        if (r < p_cipher) g.op = (rng() % 8 == 0) ? OP_OR : (rng() % 5 == 0) ? OP_INV : OP_AND;
        else g.op = (rng() % 4 == 0) ? OP_XNOR : OP_XOR;
        if (g.op == OP_INV) g.in1 = 0;
        c.gates.push_back(g);
    }
    c.num_wires = nw;
    if (c.num_out > ng) c.num_out = ng;
    c.name = "random" + std::to_string(seed);
    return c;
}

static int check_circuit(const Circuit& c, uint64_t* plans) {
    PlanSpec spec;
    spec.gates = c.gates.data(); spec.num_gates = (uint32_t)c.gates.size(); spec.num_wires = c.num_wires;
    for (uint32_t i = 0; i < c.num_in; i++) spec.live_in.push_back(i);
    for (uint32_t i = 0; i < c.num_out; i++) spec.live_out.push_back(c.num_wires - c.num_out + i);
    int errors = 0;
    auto run = [&](int fanin, int balance, uint32_t hot_cap, int policy, const char* tag) {
        Plan plan;
        std::string e;
        const int rc = build_plan(spec, plan, e, fanin, balance, hot_cap, policy);
        char what[96];
        snprintf(what, sizeof what, "%s fanin %d balance %d hot_cap %u policy %d", tag, fanin, balance, hot_cap, policy);
        if (rc == GCB_E_TOO_LARGE && hot_cap) return 0u;       // the hot set does not fit this cap: a legitimate refusal
        if (rc) { fprintf(stderr, "%s [%s]: build_plan failed: %d %s\n", c.name.c_str(), what, rc, e.c_str()); errors++; return 0u; }
        (*plans)++;
        errors += check_plan(c, plan, what);
        if (hot_cap && plan.info.num_hot_slots > hot_cap) { fprintf(stderr, "%s [%s]: %u hot slots\n", c.name.c_str(), what, plan.info.num_hot_slots); errors++; }
        return plan.info.num_slots;
    };
    const uint32_t slots = run(NODE_MAX_FANIN, 1, 0, -1, "default");
    for (int policy = 0; policy <= 4; policy++) run(NODE_MAX_FANIN, 1, 0, policy, "policy");
    run(2, 1, 0, -1, "full-wire");
    for (int fanin : {3, 5, 9}) run(fanin, 1, 0, -1, "fan-in");
    run(NODE_MAX_FANIN, 0, 0, -1, "unbalanced");
    if (slots > 64)
        for (uint32_t cap : {slots - 1, slots * 3 / 4, slots / 2, slots / 3, slots / 5}) {
            if (cap < 32) continue;
            run(NODE_MAX_FANIN, 1, cap, -1, "hot cap");
            run(NODE_MAX_FANIN, 1, cap, 3, "hot cap, balanced schedule");
        }
    {   // what the library itself picks for a batch that overflows the all-hot plan (plan.cpp: build_best_plan)
        Plan plan;
        std::string e;
        const int rc = build_best_plan(spec, plan, e, NODE_MAX_FANIN, ~0ull);
        if (rc) { fprintf(stderr, "%s: build_best_plan failed: %d %s\n", c.name.c_str(), rc, e.c_str()); errors++; }
        else { (*plans)++; errors += check_plan(c, plan, "the library's plan for large batches"); }
    }
    return errors;
}

int main(int argc, char** argv) {
    int errors = 0;
    uint64_t plans = 0, circuits = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--random") && i + 1 < argc) {
            const int n = atoi(argv[++i]);
            for (int s = 1; s <= n; s++) { const Circuit c = random_circuit((uint32_t)s); errors += check_circuit(c, &plans); circuits++; }
            continue;
        }
        Circuit c;
        if (!load_gates(argv[i], c)) return 2;
        errors += check_circuit(c, &plans);
        circuits++;
    }
    printf("plan_check: %llu plans of %llu circuits interpreted (mode %s), %d errors\n", (unsigned long long)plans,
           (unsigned long long)circuits, getenv("GCB_HOT_MODE") && atoi(getenv("GCB_HOT_MODE")) == 1 ? "whole-life hot / cold" : "split live ranges", errors);
    return errors ? 1 : 0;
}
