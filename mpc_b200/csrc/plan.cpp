// plan.cpp -- host-side circuit plan compiler (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstdio>
#include <functional>
#include <iterator>
#include <limits>
#include <queue>
#include <set>
#include <thread>

namespace gcb {

namespace {
constexpr int64_t kForever = std::numeric_limits<int64_t>::max();

inline int op_class(uint8_t op) { return op <= OP_XNOR ? 0 : (op == OP_INV ? 2 : 1); }
}  // namespace

int build_plan(const PlanSpec& spec, Plan& plan, std::string& err, int max_fanin, int balance, uint32_t hot_cap, int only_policy) {
    char msg[160];
    const uint32_t ng = spec.num_gates, nw = spec.num_wires;
    const bool identity = spec.loc.empty();
    const uint32_t nloc = identity ? nw : spec.num_locs;
    if (!identity && spec.loc.size() != nw) { err = "location map size mismatch"; return GCB_E_ARG; }
    if (max_fanin < 2 || max_fanin > NODE_MAX_FANIN) { err = "bad fan-in limit"; return GCB_E_ARG; }
    const bool keep_all = max_fanin == 2;
    const size_t K = (size_t)max_fanin;
    auto loc_of = [&](uint32_t w) { return identity ? w : spec.loc[w]; };

    const size_t ninit = spec.live_in.size();
    const size_t ndefs = ninit + ng;
    // A definition is one value of one location: the initial value of a live-in
    // location, or the output of one gate.  Re-assigned locations (legal in the
    // file formats, and the norm for aliased streaming wires) get a new
    // definition each time, which makes every hazard a plain RAW dependency.
    std::vector<int64_t> cur_def(nloc, -1);
    std::vector<uint32_t> cd(ndefs, 0);                     // cipher depth of a definition
    for (size_t k = 0; k < ninit; k++) {
        const uint32_t l = spec.live_in[k];
        if (l >= nloc) { err = "live-in location out of range"; return GCB_E_WIRE; }
        if (cur_def[l] >= 0) { err = "duplicate live-in location"; return GCB_E_ARG; }
        cur_def[l] = (int64_t)k;
    }

    // ---- pass 1: definitions, phases, static tweak ids and slab rows (original order)
    std::vector<int64_t> def_a(ng), def_b(ng);
    std::vector<uint32_t> phase_of(ng);
    uint32_t n_and = 0, n_or = 0, n_inv = 0, n_free = 0, n_phases = 0;
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
        const uint32_t d = std::max(cd[da], cd[db]);
        phase_of[i] = d;
        n_phases = std::max(n_phases, d + 1);
        cd[ninit + i] = g.op >= OP_AND ? d + 1 : d;
        def_a[i] = da; def_b[i] = db;
        cur_def[loc_of(g.out)] = (int64_t)(ninit + i);
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
    auto is_free_def = [&](int64_t d) { return d >= (int64_t)ninit && spec.gates[(size_t)d - ninit].op <= OP_XNOR; };

    // ---- the latest phase at which each gate may run without lengthening the schedule
    // (ALAP): a ciphered gate must run one level before its earliest consumer, a free gate
    // no later than it.  Gates whose output nothing reads stay where ASAP put them.
    const std::vector<uint32_t> asap(phase_of);
    std::vector<uint8_t> is_out_def(ndefs, 0);
    for (size_t k = 0; k < spec.live_out.size(); k++) {
        const uint32_t l = spec.live_out[k];
        if (l >= nloc || cur_def[l] < 0) {
            snprintf(msg, sizeof msg, "wire %u not assigned", l);
            err = msg;
            return GCB_E_WIRE;
        }
        is_out_def[(size_t)cur_def[l]] = 1;
    }
    // GCB_HOT_MODE = 1: hot / cold by whole-life classification with synchronous scratch reads (the first form, kept for
    // comparison); default: live-range splitting with asynchronous reloads
    static const bool split_mode = [] { const char* e = getenv("GCB_HOT_MODE"); return !(e && atoi(e) == 1); }();
    // fixed: cipher levels chosen by the caller (the balanced schedule); the free gates then go as late as their
    // consumers allow, but never before their own inputs exist under those cipher levels.
    auto alap_levels = [&](bool cipher_asap, const std::vector<uint32_t>* fixed = nullptr) {
        constexpr uint32_t kNone = 0xffffffffu;
        std::vector<uint32_t> need(ndefs, kNone);          // latest phase in which the definition may appear
        const uint32_t last_phase = n_phases ? n_phases - 1 : 0;
        // Values the caller reads afterwards may wait until the last phase -- unless live ranges are split (hot_cap): then
        // sha256's 256 outputs and the 769 values they are made of would all be hot at the very end (the peak of the whole
        // schedule); computed when their inputs are there, they are evicted and read from the scratch at the end.
        if (!(hot_cap && split_mode))
            for (size_t d = 0; d < ndefs; d++) if (is_out_def[d]) need[d] = last_phase;
        std::vector<uint32_t> earliest;
        if (fixed) {                                       // when each value exists, front to back
            std::vector<uint32_t> avail(ndefs, 0);
            earliest.resize(ng);
            for (uint32_t i = 0; i < ng; i++) {
                earliest[i] = std::max(avail[(size_t)def_a[i]], avail[(size_t)def_b[i]]);
                avail[ninit + i] = spec.gates[i].op >= OP_AND ? (*fixed)[i] + 1 : earliest[i];
            }
        }
        std::vector<uint32_t> lv(ng);
        for (uint32_t i = ng; i-- > 0;) {
            const bool cipher = spec.gates[i].op >= OP_AND;
            const uint32_t lo = fixed ? earliest[i] : asap[i];
            uint32_t l = need[ninit + i];
            if (cipher) l = fixed ? (*fixed)[i] : (l == kNone || cipher_asap) ? asap[i] : (l == 0 ? 0 : l - 1);
            else if (l == kNone) l = lo;
            if (l < lo) l = lo;
            lv[i] = l;
            // what this gate demands of its inputs: available in phase lv (free producers) --
            // the kernel runs a phase's free wires before its cipher level
            for (const int64_t d : {def_a[i], def_b[i]}) need[(size_t)d] = std::min(need[(size_t)d], lv[i]);
        }
        return lv;
    };

    // ---- balanced cipher levels.  A cipher level costs one warp pass per 32 tasks STARTED (4 tasks per AND / OR
    // and 2 per INV in the garbler, 2 and 1 in the evaluator): a level of 39 tasks pays for 64.  Deep, narrow
    // circuits (sha256: 2,387 levels of ~39 garbler tasks) have slack -- most gates are not on the critical path --
    // so the levels are filled front to back: a level takes the gates that can wait no longer (ALAP level reached)
    // and tops its last pass up with ready gates that could still wait, least slack first.  The depth is unchanged.
    auto balanced_levels = [&](int mode) {
        const uint32_t wq = mode == 1 ? 4u : 2u, wi = mode == 1 ? 2u : 1u;       // tasks per AND/OR, per INV
        const std::vector<uint32_t> alap = alap_levels(false);
        std::vector<std::vector<uint32_t>> users(ndefs);
        for (uint32_t i = 0; i < ng; i++) {
            users[(size_t)def_a[i]].push_back(i);
            if (def_b[i] != def_a[i]) users[(size_t)def_b[i]].push_back(i);
        }
        std::vector<uint32_t> avail(ndefs, 0), lv(ng, 0);
        std::vector<uint8_t> waiting(ng);
        for (uint32_t i = 0; i < ng; i++) waiting[i] = def_a[i] == def_b[i] ? 1 : 2;
        std::vector<std::vector<uint32_t>> ready_at(n_phases + 1);
        std::vector<size_t> stack;
        auto resolve = [&](size_t d0) {                        // definition d0 now has its availability level
            stack.assign(1, d0);
            while (!stack.empty()) {
                const size_t d = stack.back();
                stack.pop_back();
                for (uint32_t u : users[d]) {
                    if (--waiting[u]) continue;
                    const uint32_t e = std::max(avail[(size_t)def_a[u]], avail[(size_t)def_b[u]]);
                    if (spec.gates[u].op >= OP_AND) ready_at[std::min<uint32_t>(e, n_phases)].push_back(u);
                    else { avail[ninit + u] = e; stack.push_back(ninit + u); }
                }
            }
        };
        for (size_t k = 0; k < ninit; k++) resolve(k);
        using Item = std::pair<uint32_t, uint32_t>;             // (ALAP level, gate): least slack first
        std::priority_queue<Item, std::vector<Item>, std::greater<Item>> pool;
        std::vector<uint32_t> chosen;
        for (uint32_t L = 0; L < n_phases; L++) {
            for (uint32_t g : ready_at[L]) pool.push(Item{alap[g], g});
            chosen.clear();
            uint32_t tasks = 0;
            while (!pool.empty() && pool.top().first <= L) {       // cannot wait
                const uint32_t g = pool.top().second;
                pool.pop();
                chosen.push_back(g);
                tasks += spec.gates[g].op == OP_INV ? wi : wq;
            }
            // top the last pass up (never open a new one for gates that can wait)
            const uint32_t cap = (tasks + 31) / 32 * 32;
            std::vector<Item> skipped;
            while (!pool.empty() && tasks < cap) {
                const Item it = pool.top();
                pool.pop();
                const uint32_t w = spec.gates[it.second].op == OP_INV ? wi : wq;
                if (tasks + w > cap) { skipped.push_back(it); continue; }
                chosen.push_back(it.second);
                tasks += w;
            }
            for (const Item& it : skipped) pool.push(it);
            for (uint32_t g : chosen) {
                lv[g] = L;
                avail[ninit + g] = L + 1;
            }
            for (uint32_t g : chosen) resolve(ninit + g);
        }
        return alap_levels(false, &lv);
    };

    // emit = false: only the figures the policies are compared on (slots, steps); the records (and the
    // leaf ordering, the expensive part) are produced once, for the schedule that won
    auto schedule = [&](const std::vector<uint32_t>& phase_of, Plan& out, bool emit, std::string& err, uint32_t hot_cap) -> int {
        char msg[160];
        size_t ndefs = ninit + ng;                          // grows when live ranges are split (reload segments)
        // ---- pass 2: which free wires must exist in a slot: read by a ciphered gate, by a
        // free gate of a later phase, or by the caller afterwards
        std::vector<uint8_t> required(ng, keep_all ? 1 : 0);
        for (uint32_t i = 0; i < ng; i++) {
            const bool cipher = spec.gates[i].op >= OP_AND;
            for (const int64_t d : {def_a[i], def_b[i]})
                if (is_free_def(d) && (cipher || phase_of[(size_t)d - ninit] != phase_of[i])) required[(size_t)d - ninit] = 1;
        }
        std::vector<int64_t> out_def(spec.live_out.size());
        for (size_t k = 0; k < spec.live_out.size(); k++) {
            const uint32_t l = spec.live_out[k];
            if (l >= nloc || cur_def[l] < 0) {
                snprintf(msg, sizeof msg, "wire %u not assigned", l);
                err = msg;
                return GCB_E_WIRE;
            }
            out_def[k] = cur_def[l];
            if (is_free_def(cur_def[l])) required[(size_t)cur_def[l] - ninit] = 1;
        }

        // ---- pass 3: flatten the free gates of each phase into nodes of bounded fan-in.
        // flat[i] = the leaves whose XOR is gate i's output: definitions that exist when the
        // phase starts, or nodes of the same phase that had to be cut off to respect the fan-in
        // limit (a "forced" node).  Same-phase free inputs are otherwise inlined, which keeps
        // the dependency depth (waves) low.
        struct Node { uint32_t gate; uint32_t wave; uint8_t parity; std::vector<uint32_t> leaves; };
        std::vector<std::vector<uint32_t>> flat(ng);
        std::vector<uint8_t> parity(ng, 0), forced(ng, 0), emitted(ng, 0);
        std::vector<uint32_t> wave_of(ng, 0);
        std::vector<std::vector<Node>> phase_nodes(n_phases);
        auto sym_diff = [](const std::vector<uint32_t>& x, const std::vector<uint32_t>& y, std::vector<uint32_t>& o) {
            o.clear();
            std::set_symmetric_difference(x.begin(), x.end(), y.begin(), y.end(), std::back_inserter(o));
        };
        auto emit_node = [&](uint32_t i) {
            if (emitted[i]) return;
            emitted[i] = 1;
            uint32_t w = 0;
            for (uint32_t l : flat[i])
                if (is_free_def(l) && phase_of[l - ninit] == phase_of[i]) w = std::max(w, wave_of[l - ninit] + 1);
            wave_of[i] = w;
            phase_nodes[phase_of[i]].push_back(Node{i, w, parity[i], flat[i]});
        };
        std::vector<uint32_t> la, lb, tmp;
        for (uint32_t i = 0; i < ng; i++) {
            const gcb_gate& g = spec.gates[i];
            if (g.op > OP_XNOR) continue;
            for (int attempt = 0;; attempt++) {
                uint8_t par = g.op == OP_XNOR;
                auto leafset = [&](int64_t d, std::vector<uint32_t>& o) {
                    if (is_free_def(d) && phase_of[(size_t)d - ninit] == phase_of[i] && !forced[(size_t)d - ninit] && !keep_all) {
                        o = flat[(size_t)d - ninit];
                        par ^= parity[(size_t)d - ninit];
                    } else {
                        o.assign(1, (uint32_t)d);
                    }
                };
                leafset(def_a[i], la);
                leafset(def_b[i], lb);
                sym_diff(la, lb, tmp);
                if (tmp.size() <= K || attempt == 2) { flat[i] = tmp; parity[i] = par; break; }
                // too wide: cut off the wider inlined input as a node of its own and refer to it
                const int64_t cand = (la.size() >= lb.size() && la.size() > 1) ? def_a[i] : (lb.size() > 1 ? def_b[i] : def_a[i]);
                const uint32_t cg = (uint32_t)((size_t)cand - ninit);
                emit_node(cg);
                forced[cg] = 1;
            }
            if (required[i]) emit_node(i);
        }

        // ---- pass 4: the time axis (one step per wave, one per cipher level), liveness
        std::vector<std::vector<uint32_t>> phase_cipher(n_phases);
        for (uint32_t i = 0; i < ng; i++)
            if (spec.gates[i].op >= OP_AND) phase_cipher[phase_of[i]].push_back(i);
        std::vector<int64_t> born(ndefs, -1), last(ndefs, -1);
        std::vector<uint32_t> wave_step0(n_phases, 0), cipher_step(n_phases, 0), n_waves(n_phases, 0);
        int64_t nsteps = 0;
        for (uint32_t p = 0; p < n_phases; p++) {
            for (const Node& nd : phase_nodes[p]) n_waves[p] = std::max(n_waves[p], nd.wave + 1);
            wave_step0[p] = (uint32_t)nsteps;
            nsteps += n_waves[p];
            cipher_step[p] = (uint32_t)nsteps;
            if (!phase_cipher[p].empty()) nsteps++;
        }
        for (uint32_t p = 0; p < n_phases; p++) {
            for (const Node& nd : phase_nodes[p]) {
                const int64_t s = wave_step0[p] + nd.wave;
                born[ninit + nd.gate] = s;
                for (uint32_t l : nd.leaves) last[l] = std::max(last[l], s);
            }
            for (uint32_t i : phase_cipher[p]) {
                const int64_t s = cipher_step[p];
                born[ninit + i] = s;
                last[def_a[i]] = std::max(last[def_a[i]], s);
                last[def_b[i]] = std::max(last[def_b[i]], s);
            }
        }
        for (const int64_t d : out_def) last[(size_t)d] = kForever;
        for (size_t d = 0; d < ndefs; d++) last[d] = std::max(last[d], born[d]);

        // ---- hot / cold by LIVE-RANGE SPLITTING (hot_cap > 0, the default mode).  Deep circuits keep most labels idle
        // for hundreds of levels between bursts of use (sha256: its 768 inputs, the message schedule, the state words):
        // 1,272 labels are live at the peak but fewer than 330 are within 16 phases of a use.  A value therefore
        // leaves shared memory after a cluster of uses (EVICT: one copy to the instance's L2 scratch, queued a phase
        // after it was produced) and comes back before the next cluster (RELOAD: an asynchronous copy queued one phase
        // ahead, gc_kernels.cuh: cp.async) when more than `gap` phases lie between them.  Every read and write of the
        // gate kernels stays a plain shared-memory access; the scratch traffic is off the dependency chain.  `gap` is
        // the largest value whose hot set fits hot_cap (fewest copies).  Consumers name the hot SEGMENT they read: every
        // reload cluster is a new definition with its own slot.
        const size_t ndefs0 = ndefs;
        std::vector<std::vector<uint32_t>> reload_at(n_phases), evict_at(n_phases);
        std::vector<uint32_t> seg_parent;                        // reload segment (def id - ndefs0) -> the value it reloads
        std::vector<std::vector<std::pair<int64_t, uint32_t>>> seg_map;   // value -> (last step of cluster, def id), ascending
        std::vector<uint8_t> hot_input(ninit, 1), cold_input(ninit, 0), out_cold(ndefs0, 0), is_out(ndefs0, 0);
        std::vector<uint32_t> cold_until(ndefs0, 0);             // last phase in which the cold copy is read
        std::vector<uint32_t> ga(ng), gb(ng);                    // the definitions (segments) the ciphered gates read
        for (uint32_t i = 0; i < ng; i++) { ga[i] = (uint32_t)def_a[i]; gb[i] = (uint32_t)def_b[i]; }
        for (const int64_t d : out_def) is_out[(size_t)d] = 1;
        const bool use_split = hot_cap != 0 && split_mode;
        uint64_t n_reloads = 0, n_evicts = 0;
        if (use_split) {
            std::vector<uint32_t> step_phase((size_t)nsteps + 1, n_phases ? n_phases - 1 : 0);
            std::vector<int64_t> phase_start(n_phases + 1, nsteps);
            std::vector<uint8_t> phase_has_steps(n_phases, 0);
            for (uint32_t p = 0; p < n_phases; p++) {
                const int64_t s0 = wave_step0[p], s1 = (int64_t)cipher_step[p] + (phase_cipher[p].empty() ? 0 : 1);
                for (int64_t q = s0; q < s1; q++) step_phase[(size_t)q] = p;
                phase_start[p] = s0;
                phase_has_steps[p] = s1 > s0;
            }
            std::vector<std::vector<int64_t>> ev(ndefs0);
            for (uint32_t p = 0; p < n_phases; p++) {
                for (const Node& nd : phase_nodes[p]) for (uint32_t l : nd.leaves) ev[l].push_back((int64_t)wave_step0[p] + nd.wave);
                for (uint32_t i : phase_cipher[p]) { ev[(size_t)def_a[i]].push_back(cipher_step[p]); ev[(size_t)def_b[i]].push_back(cipher_step[p]); }
            }
            for (auto& e : ev) { std::sort(e.begin(), e.end()); e.erase(std::unique(e.begin(), e.end()), e.end()); }
            struct Cluster { int64_t lo, hi; uint32_t first_phase; };
            // the clusters of one value for a gap; c0 = true: the first cluster starts at the value's birth (or load)
            auto clusters_of = [&](size_t d, uint32_t gap, std::vector<Cluster>& cl, bool& c0, bool& hot_end, bool& needs_cold) {
                cl.clear();
                const bool input = d < ninit;
                const int64_t b = std::max<int64_t>(born[d], 0);
                const uint32_t pb = input ? 0u : step_phase[(size_t)b];
                uint32_t last_phase = pb;
                c0 = !input;
                if (!input) cl.push_back(Cluster{b, b, pb});
                // A reload is queued at the top of the last phase with steps before its cluster; the value reaches the scratch
                // at the top of the phase after its birth.  Both run unordered within their phase, so the reload must be queued
                // in a LATER phase than the evict: a cluster that would start sooner (tiny gaps, empty phases in between) stays
                // part of the previous one.  (Found by tools/plan_check.cpp: add64 with a cap just below its all-hot need
                // reloaded values before they were born.)  Inputs are in the scratch before phase 0.
                const uint32_t evict_phase = std::min(pb + 1, n_phases - 1);
                auto reload_after_evict = [&](uint32_t first_phase) {
                    if (input) return true;
                    uint32_t p = first_phase - 1;
                    while (p > 0 && !phase_has_steps[p]) p--;
                    return p > evict_phase;
                };
                for (int64_t st : ev[d]) {
                    const uint32_t ph = step_phase[(size_t)st];
                    if (cl.empty()) {
                        if (ph <= gap) { c0 = true; cl.push_back(Cluster{0, st, 0}); }        // an input used early stays hot from the load
                        else cl.push_back(Cluster{st, st, ph});
                    } else if (ph > last_phase + gap && ph >= 1 && reload_after_evict(ph)) cl.push_back(Cluster{st, st, ph});
                    else cl.back().hi = st;
                    last_phase = ph;
                }
                // a live-out value is read by the caller after the last phase: from its hot slot if its last cluster is
                // recent, else from the scratch
                hot_end = is_out[d] && !cl.empty() && n_phases - 1 - std::min(last_phase, n_phases - 1) <= gap;
                const size_t n_reload_clusters = cl.size() - (c0 ? 1 : 0);
                needs_cold = n_reload_clusters > 0 || (is_out[d] && !hot_end);
                if (!input && needs_cold) {
                    // the evict runs at the start of the phase after the birth: the value must still be hot there.  A value born
                    // in the last phase is hot at the end by construction (gap >= 1).
                    const uint32_t pe = std::min(pb + 1, n_phases - 1);
                    cl[0].hi = std::max(cl[0].hi, phase_start[pe]);
                }
                if (hot_end) cl.back().hi = nsteps;
            };
            auto reload_phase = [&](uint32_t first_phase) {                 // the phase at whose start the reload is queued
                uint32_t p = first_phase - 1;
                while (p > 0 && !phase_has_steps[p]) p--;
                return p;
            };
            std::vector<Cluster> cl;
            auto peak_for = [&](uint32_t gap) {
                std::vector<int32_t> diff((size_t)nsteps + 3, 0);
                for (size_t d = 0; d < ndefs0; d++) {
                    if (!(d < ninit || born[d] >= 0)) continue;
                    bool c0, hot_end, needs_cold;
                    clusters_of(d, gap, cl, c0, hot_end, needs_cold);
                    for (size_t k = 0; k < cl.size(); k++) {
                        const int64_t lo = (k == 0 && c0) ? cl[k].lo : phase_start[reload_phase(cl[k].first_phase)];
                        diff[(size_t)lo]++; diff[(size_t)std::min<int64_t>(cl[k].hi, nsteps) + 1]--;
                    }
                }
                int32_t cur = 0, peak = 0;
                size_t at = 0;
                for (size_t q = 0; q < diff.size(); q++) { cur += diff[q]; if (cur > peak) { peak = cur; at = q; } }
                if (getenv("GCB_PLAN_DEBUG") && gap == 64) {
                    size_t n_in = 0, n_c0 = 0, n_rel = 0, n_end = 0;
                    for (size_t d = 0; d < ndefs0; d++) {
                        if (!(d < ninit || born[d] >= 0)) continue;
                        bool c0, hot_end, needs_cold;
                        clusters_of(d, gap, cl, c0, hot_end, needs_cold);
                        for (size_t k = 0; k < cl.size(); k++) {
                            const int64_t lo = (k == 0 && c0) ? cl[k].lo : phase_start[reload_phase(cl[k].first_phase)];
                            if (lo <= (int64_t)at && (int64_t)at <= std::min<int64_t>(cl[k].hi, nsteps)) {
                                if (d < ninit) n_in++; else if (k == 0) n_c0++; else n_rel++;
                                if (hot_end && k + 1 == cl.size()) n_end++;
                            }
                        }
                    }
                    fprintf(stderr, "split peak at step %zu of %lld: inputs %zu, birth clusters %zu, reload clusters %zu, of them hot to the end %zu\n",
                            at, (long long)nsteps, n_in, n_c0, n_rel, n_end);
                }
                return (uint32_t)peak;
            };
            static const uint32_t kGaps[] = {0xffffffffu, 512, 256, 128, 96, 64, 48, 32, 24, 16, 12, 8, 6, 4, 3, 2};
            uint32_t gap = 0;
            for (uint32_t g : kGaps) {
                const uint32_t pk = peak_for(g);
                if (getenv("GCB_PLAN_DEBUG")) fprintf(stderr, "split: gap %u -> hot peak %u (cap %u)\n", g, pk, hot_cap);
                if (pk <= hot_cap) { gap = g; break; }
            }
            if (gap == 0) { err = "hot set does not fit"; return GCB_E_TOO_LARGE; }
            seg_map.assign(ndefs0, {});
            for (size_t d = 0; d < ndefs0; d++) {
                if (!(d < ninit || born[d] >= 0)) continue;
                bool c0, hot_end, needs_cold;
                clusters_of(d, gap, cl, c0, hot_end, needs_cold);
                const bool input = d < ninit;
                if (input) { hot_input[d] = c0; cold_input[d] = needs_cold; }
                if (!needs_cold && cl.size() <= 1 && c0) continue;                    // an ordinary value: one hot segment, as before
                uint32_t last_reload = 0;
                for (size_t k = 0; k < cl.size(); k++) {
                    uint32_t id = (uint32_t)d;
                    if (k == 0 && c0) last[d] = cl[k].hi >= nsteps ? kForever : cl[k].hi;
                    else {
                        id = (uint32_t)(ndefs0 + seg_parent.size());
                        seg_parent.push_back((uint32_t)d);
                        const uint32_t pr = reload_phase(cl[k].first_phase);
                        born.push_back(phase_start[pr]);
                        last.push_back(cl[k].hi >= nsteps ? kForever : cl[k].hi);
                        reload_at[pr].push_back(id);
                        last_reload = std::max(last_reload, pr);
                        n_reloads++;
                    }
                    seg_map[d].push_back({cl[k].hi, id});
                }
                if (!c0 && input) last[d] = -1;                                       // no hot slot of its own
                if (needs_cold) {
                    out_cold[d] = is_out[d] && !hot_end;
                    cold_until[d] = out_cold[d] ? n_phases + 1 : last_reload;
                    if (!input) { evict_at[std::min(step_phase[(size_t)std::max<int64_t>(born[d], 0)] + 1, n_phases - 1)].push_back((uint32_t)d); n_evicts++; }
                }
            }
            ndefs = born.size();
            // consumers read the segment that is hot at their step
            auto seg_for = [&](uint32_t d, int64_t step) {
                if (d >= seg_map.size() || seg_map[d].empty()) return d;
                for (const auto& s : seg_map[d]) if (step <= s.first) return s.second;
                return seg_map[d].back().second;
            };
            for (uint32_t p = 0; p < n_phases; p++) {
                for (Node& nd : phase_nodes[p]) for (uint32_t& l : nd.leaves) l = seg_for(l, (int64_t)wave_step0[p] + nd.wave);
                for (uint32_t i : phase_cipher[p]) { ga[i] = seg_for(ga[i], cipher_step[p]); gb[i] = seg_for(gb[i], cipher_step[p]); }
            }
        }

        // ---- hot / cold.  With hot_cap > 0 at most hot_cap labels may be live in shared memory at any step; the
        // others are COLD: they live in the team's scratch in global memory (L2-resident) for their whole life and
        // are read from there by the nodes and gates that use them (gc_kernels.cuh: SlotsSpill).  Deep, narrow
        // circuits are bound by the latency of a level, so the instances an SM holds count more than an L2 round trip
        // now and then: sha256 keeps its 768 inputs and its message schedule live for hundreds of levels between
        // uses.  A value is cold when it idles long per access (lifetime / (reads + 1) above a threshold); the
        // threshold is the largest one whose hot set fits.
        std::vector<uint8_t> cold(ndefs, 0);
        uint64_t cold_accesses = 0;
        if (hot_cap && !use_split) {
            std::vector<uint32_t> nreads(ndefs, 0);
            for (uint32_t p = 0; p < n_phases; p++) {
                for (const Node& nd : phase_nodes[p]) for (uint32_t l : nd.leaves) nreads[l]++;
                for (uint32_t i : phase_cipher[p]) { nreads[(size_t)def_a[i]]++; nreads[(size_t)def_b[i]]++; }
            }
            std::vector<int64_t> b0(ndefs), l0(ndefs);
            std::vector<double> idle(ndefs, 0.0);
            std::vector<uint8_t> stored(ndefs, 0);                 // values that get a slot at all (not the inlined XORs)
            for (size_t d = 0; d < ndefs; d++) {
                stored[d] = d < ninit || born[d] >= 0;
                if (!stored[d]) continue;
                b0[d] = std::max<int64_t>(born[d], 0);
                l0[d] = last[d] == kForever ? nsteps : std::max<int64_t>(last[d], b0[d]);
                idle[d] = (double)(l0[d] - b0[d] + 1) / (nreads[d] + 1);
            }
            // live labels per step, as a max segment tree with lazy range add
            size_t nleaf = 1;
            while (nleaf < (size_t)nsteps + 2) nleaf <<= 1;
            std::vector<int32_t> mx(2 * nleaf, 0), lz(2 * nleaf, 0);
            std::function<void(size_t, size_t, size_t, size_t, size_t, int32_t)> add = [&](size_t n, size_t l, size_t r, size_t ql, size_t qr, int32_t v) {
                if (qr < l || r < ql) return;
                if (ql <= l && r <= qr) { mx[n] += v; lz[n] += v; return; }
                const size_t m = (l + r) / 2;
                add(2 * n, l, m, ql, qr, v); add(2 * n + 1, m + 1, r, ql, qr, v);
                mx[n] = std::max(mx[2 * n], mx[2 * n + 1]) + lz[n];
            };
            std::function<int32_t(size_t, size_t, size_t, size_t, size_t)> qmax = [&](size_t n, size_t l, size_t r, size_t ql, size_t qr) -> int32_t {
                if (qr < l || r < ql) return INT32_MIN / 2;
                if (ql <= l && r <= qr) return mx[n];
                const size_t m = (l + r) / 2;
                return std::max(qmax(2 * n, l, m, ql, qr), qmax(2 * n + 1, m + 1, r, ql, qr)) + lz[n];
            };
            std::vector<size_t> order;
            for (size_t d = 0; d < ndefs; d++)
                if (stored[d]) { order.push_back(d); add(1, 0, nleaf - 1, (size_t)b0[d], (size_t)l0[d], 1); }
            // longest idle per access first; a value goes cold only if it is live across a step that is still over the cap
            std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return idle[x] != idle[y] ? idle[x] > idle[y] : x < y; });
            for (size_t d : order) {
                if (mx[1] <= (int32_t)hot_cap) break;
                if (qmax(1, 0, nleaf - 1, (size_t)b0[d], (size_t)l0[d]) <= (int32_t)hot_cap) continue;
                cold[d] = 1;
                cold_accesses += nreads[d] + 1;
                add(1, 0, nleaf - 1, (size_t)b0[d], (size_t)l0[d], -1);
            }
            if (mx[1] > (int32_t)hot_cap) { err = "hot set does not fit"; return GCB_E_TOO_LARGE; }
        }

        // ---- slots: lowest free index first; a slot whose value was last read in step s may
        // be rewritten from step s+1 on (reads and writes of one step are unordered).  Cold values take slots
        // from their own pool, numbered from kColdBase while the hot pool grows and renumbered behind it at the end.
        constexpr uint32_t kColdBase = 0x40000000u;
        std::set<uint32_t> free_cold;
        uint32_t next_cold = 0;
        std::vector<std::vector<size_t>> expire((size_t)nsteps + 1);
        std::vector<uint32_t> slot(ndefs, 0);
        // Free slots by bank group (slot & 7: a 16-byte label covers four of the 32 banks).  A
        // 128-bit shared-memory access is served one quarter warp at a time, and the eight lanes of
        // a quarter warp are conflict-free when their labels sit in eight different bank groups:
        // the value written by lane position `pos` of a step prefers bank group pos & 7.  Taking a
        // free slot of the wanted group instead of the lowest one never grows the slot count (a new
        // slot is opened only when no slot is free).
        std::set<uint32_t> free_slots[8];
        uint32_t next_slot = 0;
        auto release = [&](uint32_t s) { if (s >= kColdBase) free_cold.insert(s); else free_slots[s & 7].insert(s); };
        auto take_cold = [&]() {
            if (free_cold.empty()) return kColdBase + next_cold++;
            const uint32_t c = *free_cold.begin();
            free_cold.erase(free_cold.begin());
            return c;
        };
        // scratch slots of evicted values (live-range splitting): by phase, freed two phases after the last reload
        std::vector<uint32_t> cold_slot(ndefs0, 0);
        std::vector<std::vector<uint32_t>> cold_expire(n_phases + 4);
        struct CopyOp { uint32_t src, dst; };
        std::vector<CopyOp> copies;
        std::vector<std::array<uint32_t, 3>> phase_copy(n_phases, std::array<uint32_t, 3>{0, 0, 0});   // first, evicts, reloads
        out.live_in.clear();
        for (size_t k = 0; k < ninit; k++) {
            if (use_split) {
                if (cold_input[k]) {
                    cold_slot[k] = take_cold();
                    cold_expire[std::min<size_t>(cold_until[k] + 2, n_phases + 3)].push_back((uint32_t)k);
                    out.live_in.push_back(SlotRef{cold_slot[k], (uint32_t)k | (hot_input[k] ? 0x80000000u : 0u)});
                }
                if (!hot_input[k]) continue;
            }
            slot[k] = cold[k] ? take_cold() : next_slot++;
            out.live_in.push_back(SlotRef{slot[k], (uint32_t)k});
            if (last[k] < 0) release(slot[k]);                         // never read, not live-out
            else if (last[k] != kForever) expire[(size_t)last[k]].push_back(k);
        }
        // Readers: the eight nodes of a quarter warp read leaf j in one instruction, and (XOR commutes)
        // each node's leaves can be put in any order; an order without bank conflicts exists exactly
        // when no bank group holds more of the quarter's leaves than its longest node has leaves
        // (edge colouring of the node x bank-group multigraph).  So the bank group of a value is chosen
        // to keep the leaves of the quarters that will read it balanced: quarter_load[q][b] counts the
        // leaves of quarter q already placed in group b.
        std::vector<std::vector<uint32_t>> readers(ndefs);
        std::vector<std::array<uint16_t, 8>> quarter_load;
        for (uint32_t p = 0; p < n_phases; p++) {
            // nodes sorted by (wave, fan-in descending) so that the lanes of a warp do similar work
            std::stable_sort(phase_nodes[p].begin(), phase_nodes[p].end(), [](const Node& x, const Node& y) {
                if (x.wave != y.wave) return x.wave < y.wave;
                return x.leaves.size() > y.leaves.size();
            });
            uint32_t pos = 0, wave = 0;
            for (const Node& nd : phase_nodes[p]) {
                if (nd.wave != wave) { wave = nd.wave; pos = 0; }
                if ((pos & 7) == 0) quarter_load.push_back(std::array<uint16_t, 8>{});
                const uint32_t q = (uint32_t)quarter_load.size() - 1;
                for (uint32_t l : nd.leaves) readers[l].push_back(q);
                pos++;
            }
        }
        auto place = [&](size_t d) { for (uint32_t q : readers[d]) quarter_load[q][slot[d] & 7]++; };
        for (size_t k = 0; k < ninit; k++) place(k);
        // The store of the value is one lane of a quarter warp too: store_used counts the bank groups
        // the lanes before it in the same quarter (`lane` & 7 == 0 starts a new one) have taken.
        uint16_t store_used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto take_slot = [&](size_t d, uint32_t lane) {
            if ((lane & 7) == 0) std::fill(store_used, store_used + 8, (uint16_t)0);
            if (cold[d]) {
                slot[d] = take_cold();
                if (last[d] != kForever) expire[(size_t)last[d]].push_back(d);
                return;
            }
            int from = -1;
            uint32_t best = 0xffffffffu;
            for (int b = 0; b < 8; b++) {
                if (free_slots[b].empty()) continue;
                uint32_t cost = 4u * store_used[b];
                for (uint32_t q : readers[d]) cost += 2u * quarter_load[q][b];
                if (cost < best) { best = cost; from = b; }
            }
            if (from < 0) slot[d] = next_slot++;
            else { slot[d] = *free_slots[from].begin(); free_slots[from].erase(free_slots[from].begin()); }
            if (last[d] != kForever) expire[(size_t)last[d]].push_back(d);
            store_used[slot[d] & 7]++;
            place(d);
        };
        // Order of the ciphered gates of a level: AND/OR first, then INV (the kernels give 4 / 2
        // tasks to each); inside a class, gates are grouped so that the lanes of a quarter warp
        // (four AND gates in eval, two in garble; eight / four INV gates) read their first inputs
        // from different bank groups, and their second inputs too.
        auto group_gates = [&](std::vector<uint32_t>& gates_of_level) {
            std::stable_sort(gates_of_level.begin(), gates_of_level.end(), [&](uint32_t x, uint32_t y) {
                return op_class(spec.gates[x].op) < op_class(spec.gates[y].op);
            });
            size_t lo = 0;
            while (lo < gates_of_level.size()) {
                size_t hi = lo;
                const int cls = op_class(spec.gates[gates_of_level[lo]].op);
                while (hi < gates_of_level.size() && op_class(spec.gates[gates_of_level[hi]].op) == cls) hi++;
                const size_t group = cls == 2 ? 8 : 4;
                std::vector<uint32_t> rest(gates_of_level.begin() + lo, gates_of_level.begin() + hi), done;
                done.reserve(rest.size());
                while (!rest.empty()) {
                    uint32_t sa[8], sb[8];
                    size_t n = 0;
                    while (n < group && !rest.empty()) {
                        size_t pick = 0;
                        int best = 1 << 30;
                        for (size_t c = 0; c < rest.size() && best > 0; c++) {
                            const uint32_t a = slot[ga[rest[c]]], b = slot[gb[rest[c]]];
                            int cost = 0;
                            for (size_t q = 0; q < n; q++) {
                                cost += (sa[q] != a && (sa[q] & 7) == (a & 7));
                                cost += (sb[q] != b && (sb[q] & 7) == (b & 7));
                            }
                            if (cost < best) { best = cost; pick = c; }
                        }
                        sa[n] = slot[ga[rest[pick]]]; sb[n] = slot[gb[rest[pick]]];
                        n++;
                        done.push_back(rest[pick]);
                        rest.erase(rest.begin() + (long)pick);
                    }
                }
                std::copy(done.begin(), done.end(), gates_of_level.begin() + lo);
                lo = hi;
            }
        };
        {
            int64_t s = 0;
            auto advance = [&]() {
                if (s > 0) {
                    for (size_t d : expire[(size_t)s - 1]) release(slot[d]);
                    expire[(size_t)s - 1].clear();
                }
            };
            for (uint32_t p = 0; p < n_phases; p++) {
                if (use_split) {
                    // the copies queued at the top of this phase: evicts of values produced in the previous phase (their hot slot
                    // stays theirs through this step), reloads into hot slots taken now for the clusters that start next phase
                    advance();
                    for (uint32_t d : cold_expire[p]) release(cold_slot[d]);
                    phase_copy[p][0] = (uint32_t)copies.size();
                    for (uint32_t d : evict_at[p]) {
                        cold_slot[d] = take_cold();
                        cold_expire[std::min<size_t>(cold_until[d] + 2, n_phases + 3)].push_back(d);
                        copies.push_back(CopyOp{slot[d], cold_slot[d]});
                    }
                    phase_copy[p][1] = (uint32_t)evict_at[p].size();
                    uint32_t lane = 0;
                    for (uint32_t id : reload_at[p]) {
                        take_slot(id, lane++);
                        copies.push_back(CopyOp{cold_slot[seg_parent[id - ndefs0]], slot[id]});
                    }
                    phase_copy[p][2] = (uint32_t)reload_at[p].size();
                }
                size_t pos = 0;
                for (uint32_t w = 0; w < n_waves[p]; w++, s++) {
                    advance();
                    for (uint32_t lane = 0; pos < phase_nodes[p].size() && phase_nodes[p][pos].wave == w; pos++, lane++)
                        take_slot(ninit + phase_nodes[p][pos].gate, lane);
                }
                if (!phase_cipher[p].empty()) {
                    advance();
                    group_gates(phase_cipher[p]);
                    uint32_t lane = 0;
                    for (uint32_t i : phase_cipher[p]) take_slot(ninit + i, lane++);
                    s++;
                }
            }
        }
        if (next_slot + next_cold > 65535) {
            snprintf(msg, sizeof msg, "circuit needs %u live wire slots (limit 65535)", next_slot + next_cold);
            err = msg;
            return GCB_E_TOO_LARGE;
        }
        const uint32_t num_hot = next_slot;
        for (size_t d = 0; d < slot.size(); d++) if (slot[d] >= kColdBase) slot[d] = num_hot + (slot[d] - kColdBase);
        for (uint32_t& c : cold_slot) if (c >= kColdBase) c = num_hot + (c - kColdBase);
        for (SlotRef& r : out.live_in) if (r.slot >= kColdBase) r.slot = num_hot + (r.slot - kColdBase);
        for (CopyOp& c : copies) {                                   // scratch side: index into the instance's scratch
            if (c.src >= kColdBase) c.src -= kColdBase;
            if (c.dst >= kColdBase) c.dst -= kColdBase;
        }
        next_slot += next_cold;
        uint32_t g_passes = 0, e_passes = 0;
        for (uint32_t p = 0; p < n_phases; p++) {
            uint32_t nq = 0, ni = 0;
            for (uint32_t i : phase_cipher[p]) (spec.gates[i].op == OP_INV ? ni : nq)++;
            g_passes += (4 * nq + 2 * ni + 31) / 32;
            e_passes += (2 * nq + ni + 31) / 32;
        }
        if (!emit) {
            out.info = gcb_plan_info{};
            out.info.num_slots = next_slot;
            out.info.num_hot_slots = num_hot;
            out.info.num_steps = (uint32_t)nsteps;
            out.cold_accesses = use_split ? n_reloads + n_evicts : cold_accesses;
            out.info.garble_passes = g_passes;
            out.info.eval_passes = e_passes;
            return GCB_OK;
        }
        out.live_out.clear();
        for (size_t k = 0; k < spec.live_out.size(); k++)
        {
            const size_t d = (size_t)out_def[k];
            uint32_t from = slot[d];
            if (use_split && d < ndefs0 && !seg_map[d].empty()) from = out_cold[d] ? cold_slot[d] : slot[seg_map[d].back().second];
            else if (use_split && d < ndefs0 && out_cold[d]) from = cold_slot[d];
            out.live_out.push_back(SlotRef{from, (uint32_t)k});
        }

        // ---- leaf order.  The eight nodes of a quarter warp read leaf j in the same instruction; XOR
        // commutes, so each node's leaves are permuted to put labels of different bank groups side by
        // side (equal slots are a broadcast and cost nothing).  Greedy per leaf position, a few seeded
        // restarts, the cheapest kept.
        auto order_leaves = [&](NodeRec* q, size_t n) {
            auto cost_of = [&](const NodeRec* r) {
                int total = 0;
                for (int j = 0; j < NODE_MAX_FANIN; j++) {
                    uint32_t seen[8]; int ns = 0, cnt[8] = {0}, worst = 0;
                    for (size_t i = 0; i < n; i++) {
                        if (j >= r[i].k) continue;
                        bool dup = false;
                        for (int x = 0; x < ns; x++) dup |= seen[x] == r[i].leaf[j];
                        if (dup) continue;
                        seen[ns++] = r[i].leaf[j];
                        worst = std::max(worst, ++cnt[r[i].leaf[j] & 7]);
                    }
                    total += worst;
                }
                return total;
            };
            NodeRec best[8], cand[8];
            std::copy(q, q + n, best);
            int best_cost = cost_of(best), floor_cost = 0;       // one wavefront per leaf position at least
            for (size_t i = 0; i < n; i++) floor_cost = std::max<int>(floor_cost, q[i].k);
            uint32_t rng = 0x9e3779b9u;
            for (int trial = 0; trial < 6 && best_cost > floor_cost; trial++) {
                std::copy(q, q + n, cand);
                uint16_t pool[8][NODE_MAX_FANIN];
                int left[8];
                for (size_t i = 0; i < n; i++) {
                    left[i] = cand[i].k;
                    std::copy(cand[i].leaf, cand[i].leaf + cand[i].k, pool[i]);
                    if (trial) for (int x = left[i] - 1; x > 0; x--) {       // seeded shuffle
                        rng = rng * 1664525u + 1013904223u;
                        std::swap(pool[i][x], pool[i][(rng >> 8) % (uint32_t)(x + 1)]);
                    }
                }
                for (int j = 0; j < NODE_MAX_FANIN; j++) {
                    uint32_t seen[8]; int ns = 0, cnt[8] = {0};
                    // most constrained first: nodes with the fewest leaves left choose first
                    size_t ord[8];
                    for (size_t i = 0; i < n; i++) ord[i] = i;
                    std::stable_sort(ord, ord + n, [&](size_t x, size_t y) { return left[x] < left[y]; });
                    for (size_t oi = 0; oi < n; oi++) {
                        const size_t i = ord[oi];
                        if (left[i] == 0) continue;
                        int pick = 0, pick_cost = 1 << 30;
                        for (int x = 0; x < left[i]; x++) {
                            bool dup = false;
                            for (int y = 0; y < ns; y++) dup |= seen[y] == pool[i][x];
                            const int c = dup ? -1 : cnt[pool[i][x] & 7];
                            if (c < pick_cost) { pick_cost = c; pick = x; }
                        }
                        const uint16_t leaf = pool[i][pick];
                        pool[i][pick] = pool[i][--left[i]];
                        cand[i].leaf[j] = leaf;
                        if (pick_cost >= 0) { seen[ns++] = leaf; cnt[leaf & 7]++; }
                    }
                }
                int c = cost_of(cand);
                // local search: swap two leaf positions of one node while that lowers the cost of the two
                // positions it touches
                auto pos_cost = [&](int j) {
                    uint32_t seen[8]; int ns = 0, cnt[8] = {0}, worst = 0;
                    for (size_t i = 0; i < n; i++) {
                        if (j >= cand[i].k) continue;
                        bool dup = false;
                        for (int x = 0; x < ns; x++) dup |= seen[x] == cand[i].leaf[j];
                        if (dup) continue;
                        seen[ns++] = cand[i].leaf[j];
                        worst = std::max(worst, ++cnt[cand[i].leaf[j] & 7]);
                    }
                    return worst;
                };
                for (int pass = 0; pass < 4 && c > floor_cost; pass++) {
                    bool improved = false;
                    for (size_t i = 0; i < n; i++)
                        for (int j1 = 0; j1 < cand[i].k; j1++)
                            for (int j2 = j1 + 1; j2 < cand[i].k; j2++) {
                                if ((cand[i].leaf[j1] & 7) == (cand[i].leaf[j2] & 7)) continue;
                                const int before = pos_cost(j1) + pos_cost(j2);
                                if (before <= 2) continue;
                                std::swap(cand[i].leaf[j1], cand[i].leaf[j2]);
                                const int after = pos_cost(j1) + pos_cost(j2);
                                if (after < before) { c += after - before; improved = true; }
                                else std::swap(cand[i].leaf[j1], cand[i].leaf[j2]);
                            }
                    if (!improved) break;
                }
                if (c < best_cost) { best_cost = c; std::copy(cand, cand + n, best); }
            }
            std::copy(best, best + n, q);
        };

        // ---- records in schedule order
        out.phases.clear(); out.waves.clear(); out.nodes.clear(); out.crecs.clear();
        out.nout_wire.clear(); out.cout_wire.clear();
        out.node_loads = 0;
        for (uint32_t p = 0; p < n_phases; p++) {
            PhaseRec ph{};
            ph.wave_first = (uint32_t)out.waves.size();
            ph.n_waves = n_waves[p];
            size_t pos = 0;
            for (uint32_t w = 0; w < n_waves[p]; w++) {
                WaveRec wr{(uint32_t)out.nodes.size(), 0};
                for (; pos < phase_nodes[p].size() && phase_nodes[p][pos].wave == w; pos++) {
                    const Node& nd = phase_nodes[p][pos];
                    NodeRec r{};
                    r.dst = (uint16_t)slot[ninit + nd.gate];
                    r.k = (uint8_t)nd.leaves.size();
                    r.parity = nd.parity;
                    for (size_t j = 0; j < nd.leaves.size(); j++) r.leaf[j] = (uint16_t)slot[nd.leaves[j]];
                    out.nodes.push_back(r);
                    if ((wr.count & 7) == 7) order_leaves(out.nodes.data() + out.nodes.size() - 8, 8);
                    out.nout_wire.push_back(spec.gates[nd.gate].out);
                    out.node_loads += r.k;
                    wr.count++;
                }
                if (wr.count & 7) order_leaves(out.nodes.data() + out.nodes.size() - (wr.count & 7), wr.count & 7);
                out.waves.push_back(wr);
                if (w == 0) ph.w0_first = wr.first;
                if (w < 4) ph.wave_count[w] = (uint16_t)(wr.count < 0xffffu ? wr.count : 0xffffu);
            }
            // ciphered gates, in the order group_gates chose: AND/OR first, then INV
            ph.cipher_first = (uint32_t)out.crecs.size();
            for (uint32_t i : phase_cipher[p]) {
                const gcb_gate& g = spec.gates[i];
                out.crecs.push_back(GateRec{(uint16_t)slot[ga[i]], (uint16_t)slot[gb[i]],
                                             (uint16_t)slot[ninit + i], g.op, 0, tweak_of[i], plan.row_off[i]});
                out.cout_wire.push_back(g.out);
                (g.op == OP_INV ? ph.n_inv : ph.n_quad)++;
                if (g.op == OP_OR) ph.n_waves |= 0x80000000u;       // flag: the level has OR gates (no AND-only fast path)
            }
            out.phases.push_back(ph);
        }

        gcb_plan_info& in = out.info;
        in = gcb_plan_info{};
        in.num_gates = ng;
        in.num_wires = nw;
        in.num_inputs = (uint32_t)ninit;
        in.num_outputs = (uint32_t)spec.live_out.size();
        in.num_rows = row;
        in.num_tweaks = tweak;
        in.num_steps = (uint32_t)nsteps;
        in.num_slots = next_slot;
        in.num_hot_slots = num_hot;
        out.cold_accesses = use_split ? n_reloads + n_evicts : cold_accesses;
        out.copies.clear();
        out.phase_copy.clear();
        if (use_split) {
            for (const CopyOp& c : copies) { out.copies.push_back(c.src); out.copies.push_back(c.dst); }
            out.phase_copy = phase_copy;
        }
        in.num_and = n_and; in.num_or = n_or; in.num_inv = n_inv; in.num_free = n_free;
        in.garble_hashes = 4 * n_and + 4 * n_or + 2 * n_inv;
        in.eval_hashes = 2 * n_and + n_or + n_inv;
        in.garble_passes = g_passes;
        in.eval_passes = e_passes;
        return GCB_OK;
    };

    // ---- try the schedules and keep the best (the trials are independent and run on their own threads: the
    // compiler is on the critical path of the first use of a circuit, e.g. a streaming evaluator that meets a new
    // sub-circuit).  Policies 0-2 are compared on wire slots.  Policies 3 and 4, the schedules balanced for the
    // garbler's and for the evaluator's pass size, replace the winner when they lower the warp passes (garbler +
    // evaluator: the plan serves both) per instance resident beside two T-tables by at least 2 %.
    {
        auto levels_of = [&](int policy) {
            return policy == 0 ? asap : policy >= 3 ? balanced_levels(policy - 2) : alap_levels(policy == 2);
        };
        int best_policy = -1, best_rc = GCB_OK;
        uint32_t best_slots = 0, best_steps = 0, best_cap = hot_cap;
        std::string first_err;
        const int n_policies = keep_all ? 1 : (balance ? 5 : 3);          // the full-wire plan keeps the simple schedule
        Plan cand[5];
        int rcs[5] = {GCB_OK, GCB_OK, GCB_OK, GCB_OK, GCB_OK};
        std::string errs[5];
        auto trial = [&](int policy) {
            rcs[policy] = schedule(levels_of(policy), cand[policy], false, errs[policy], hot_cap);
        };
        if (only_policy >= 0 && only_policy < n_policies) best_policy = only_policy;      // the caller knows which schedule it wants
        else {
            // joined on every path out of this scope (an exception in trial(0) or in thread creation must not
            // leave joinable threads behind: that would terminate the process)
            struct Joiner {
                std::vector<std::thread> v;
                ~Joiner() { for (std::thread& t : v) if (t.joinable()) t.join(); }
            } workers;
            auto guarded = [&](int policy) {
                try { trial(policy); }
                catch (const std::exception& e) { rcs[policy] = GCB_E_TOO_LARGE; errs[policy] = e.what(); }
            };
            for (int policy = 1; policy < n_policies; policy++) workers.v.emplace_back(guarded, policy);
            guarded(0);
        }
        for (int policy = 0; only_policy < 0 && policy < std::min(n_policies, 3); policy++) {
            if (rcs[policy] != GCB_OK) { if (best_policy < 0 && first_err.empty()) { first_err = errs[policy]; best_rc = rcs[policy]; } continue; }
            if (best_policy < 0 || cand[policy].info.num_slots < best_slots ||
                (cand[policy].info.num_slots == best_slots && cand[policy].info.num_steps < best_steps)) {
                best_policy = policy; best_slots = cand[policy].info.num_slots; best_steps = cand[policy].info.num_steps;
            }
        }
        if (best_policy < 0) { err = first_err; return best_rc; }
        if (getenv("GCB_PLAN_DEBUG"))
            for (int policy = 0; policy < n_policies; policy++)
                fprintf(stderr, "plan policy %d: rc %d slots %u steps %u passes %u + %u\n", policy, rcs[policy], cand[policy].info.num_slots,
                        cand[policy].info.num_steps, cand[policy].info.garble_passes, cand[policy].info.eval_passes);
        if (n_policies == 5 && only_policy < 0) {
            auto resident = [](uint32_t slots) {                  // instances beside two T-tables
                const size_t n = teams_that_fit(slots, kAssumedSmemBase, 2);
                return n > 16 ? (n >= 32 ? (size_t)32 : (size_t)16) : n;
            };
            // cost of a schedule: warp passes per resident instance (deep circuits are bound by the latency of a level,
            // so an instance more per SM is worth as much as proportionally fewer passes)
            auto cost = [&](const gcb_plan_info& i) {
                const size_t r = resident(hot_cap ? i.num_hot_slots : i.num_slots);
                return r ? ((double)i.garble_passes + i.eval_passes) / (double)r : 1e30;
            };
            double best_cost = cost(cand[best_policy].info) * 0.98;                             // worth it from 2 % on
            for (int policy = 3; policy < 5; policy++) {
                if (rcs[policy] != GCB_OK) continue;
                const double c = cost(cand[policy].info);
                if (c <= best_cost) { best_policy = policy; best_cost = c; }
            }
        }
        Plan best;
        best.row_off = plan.row_off; best.ops = plan.ops;
        const int rc = schedule(levels_of(best_policy), best, true, err, best_cap);
        if (rc != GCB_OK) return rc;
        plan.info = best.info;
        plan.policy = best_policy;
        plan.phases.swap(best.phases); plan.waves.swap(best.waves); plan.nodes.swap(best.nodes);
        plan.crecs.swap(best.crecs); plan.nout_wire.swap(best.nout_wire); plan.cout_wire.swap(best.cout_wire);
        plan.live_in.swap(best.live_in); plan.live_out.swap(best.live_out);
        plan.node_loads = best.node_loads;
        plan.cold_accesses = best.cold_accesses;
        plan.copies.swap(best.copies); plan.phase_copy.swap(best.phase_copy);
    }
    return GCB_OK;
}

}  // namespace gcb

namespace gcb {

int build_best_plan(const PlanSpec& spec, Plan& plan, std::string& err, int max_fanin, uint64_t batch_hint) {
    int rc = build_plan(spec, plan, err, max_fanin);
    if (rc != GCB_OK || max_fanin == 2) return rc;
    // The second plan trades the latency of one instance (narrower teams, L2 round trips) for resident instances:
    // it only pays when the batch does not fit the all-hot plan in one wave (148 SMs assumed).
    const char* forced = getenv("GCB_HOT_TEAMS");
    if (!(forced && atoi(forced) > 0) && batch_hint <= teams_that_fit(plan.info.num_slots, kAssumedSmemBase, 2) * 148) return GCB_OK;
    // Measured on B200 (profiles/r02_hot_cold.txt, profiles/r02_split.txt), M AND/s garble + eval on batches of 16 x 148:
    //                                    all-hot          whole-life hot / cold, synchronous     live-range splitting,
    //                                                     scratch reads (GCB_HOT_MODE=1)         asynchronous reloads (default)
    //   sha512     3 instances per SM      870            16 instances: 1,718                    16 instances: 2,497
    //   sha256     8                     1,788            16: 1,705                              16: 2,440
    //   sha256xor  9                     1,825            16: 1,718                              16: 2,462
    //   chacha20  10                     1,661            16: 1,506                              16: 2,021
    // With whole-life classification every cold leaf is an L2 round trip on the dependency chain of its level; with split
    // live ranges a value is evicted once after its birth and reloaded (cp.async, one phase ahead) before each cluster of
    // uses, so every gate and node access stays in shared memory and the extra instances pay.  The second plan is built
    // when fewer than 16 instances fit, with 16 as the target (halved until the hot set fits and the evicts + reloads
    // stay below a fifth of the label accesses).  GCB_HOT_TEAMS = 0 switches it off, N forces a target.
    int force = -1;
    if (const char* e = getenv("GCB_HOT_TEAMS")) force = atoi(e);
    if (force == 0) return GCB_OK;
    const gcb_plan_info& in = plan.info;
    const size_t np = plan.phases.size();
    const uint32_t width = np ? (uint32_t)(in.garble_hashes / np) : 0u;
    const size_t have = teams_that_fit(in.num_slots, kAssumedSmemBase, 2);
    if (force < 0 && (width >= 128 || have >= 16 || in.num_slots < 256)) return GCB_OK;    // wide (pipe-bound) or enough resident
    // the target: 16 resident instances, or what the caller forces (halved until the hot set fits)
    for (size_t target = force > 0 ? (size_t)force : 16; target > have; target /= 2) {
        uint32_t lo = 32, hi = in.num_slots;               // largest hot set with which `target` teams fit
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) / 2;
            if (teams_that_fit(mid, kAssumedSmemBase, 2) >= target) lo = mid; else hi = mid - 1;
        }
        if (teams_that_fit(lo, kAssumedSmemBase, 2) < target) continue;
        Plan cand;
        std::string e2;
        if (build_plan(spec, cand, e2, max_fanin, 1, lo, plan.policy) != GCB_OK) continue;
        // label accesses of an instance: node leaves + node results + three per ciphered gate
        const uint64_t accesses = (uint64_t)cand.node_loads + cand.nodes.size() + 3ull * cand.crecs.size();
        if (force <= 0 && cand.cold_accesses * 100 > accesses * 20) continue;           // more than 20 % would go to L2
        plan.info = cand.info; plan.policy = cand.policy;
        plan.phases.swap(cand.phases); plan.waves.swap(cand.waves); plan.nodes.swap(cand.nodes); plan.crecs.swap(cand.crecs);
        plan.nout_wire.swap(cand.nout_wire); plan.cout_wire.swap(cand.cout_wire);
        plan.live_in.swap(cand.live_in); plan.live_out.swap(cand.live_out);
        plan.row_off.swap(cand.row_off); plan.ops.swap(cand.ops);
        plan.node_loads = cand.node_loads; plan.cold_accesses = cand.cold_accesses;
        plan.copies.swap(cand.copies); plan.phase_copy.swap(cand.phase_copy);
        break;
    }
    return GCB_OK;
}

int parse_stream(const uint8_t* buf, size_t len, uint32_t ngates, std::vector<StreamGate>& gates,
                 std::vector<uint32_t>& row_pos, size_t* consumed, std::string& err) {
    gates.clear(); row_pos.clear();
    // the count comes from the peer: a record is at least 5 bytes (INV with 16-bit ids), so a count the
    // buffer cannot hold is rejected before anything is sized by it
    if ((uint64_t)ngates * 5 > (uint64_t)len) { err = "record stream truncated"; return GCB_E_BUFFER; }
    gates.reserve(ngates);
    size_t pos = 0;
    auto need = [&](size_t n) { return pos + n <= len; };
    for (uint32_t i = 0; i < ngates; i++) {
        if (!need(1)) { err = "record stream truncated"; return GCB_E_BUFFER; }
        uint8_t gop = buf[pos++];
        StreamGate g{};
        g.a_tmp = (gop & 0x80) != 0; g.b_tmp = (gop & 0x40) != 0; g.c_tmp = (gop & 0x20) != 0;
        const bool shrt = (gop & 0x10) != 0;
        gop &= 0x0f;
        if (gop > OP_INV) { err = "invalid operation in record stream"; return GCB_E_BADOP; }
        g.op = gop;
        const int nidx = gop == OP_INV ? 2 : 3;
        const size_t w = shrt ? 2 : 4;
        if (!need(nidx * w)) { err = "record stream truncated"; return GCB_E_BUFFER; }
        uint32_t idx[3] = {0, 0, 0};
        for (int k = 0; k < nidx; k++) {
            uint32_t v = 0;
            for (size_t q = 0; q < w; q++) v = (v << 8) | buf[pos++];
            idx[k] = v;
        }
        g.a = idx[0];
        if (nidx == 3) { g.b = idx[1]; g.c = idx[2]; } else { g.b = 0; g.c = idx[1]; g.b_tmp = 0; }
        const int rows = gop == OP_AND ? 2 : gop == OP_OR ? 3 : gop == OP_INV ? 1 : 0;
        if (!need((size_t)rows * 16)) { err = "record stream truncated"; return GCB_E_BUFFER; }
        for (int k = 0; k < rows; k++) {
            if (pos > 0xfffffff0u) { err = "record stream exceeds 4 GiB"; return GCB_E_TOO_LARGE; }
            row_pos.push_back((uint32_t)pos);
            pos += 16;
        }
        gates.push_back(g);
    }
    if (consumed) *consumed = pos;
    return GCB_OK;
}

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
