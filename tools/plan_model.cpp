// plan_model.cpp -- host-only model of the shared-memory traffic of a plan.
//
// Builds the plan of a circuit (tools/_build/<name>.gates, written by tools/dump_gates.py)
// with the product's plan compiler and counts, for a given team width, the shared-memory
// wavefronts the gate kernels spend on wire labels: a 128-bit access is served one quarter
// warp at a time, and a quarter warp takes as many wavefronts as the largest number of
// distinct 16-byte words that fall into the same group of four banks (word index mod 8).
//
//   g++ -O2 -std=c++17 -pthread -I include -o tools/_build/plan_model tools/plan_model.cpp mpc_b200/csrc/plan.cpp
//   tools/_build/plan_model tools/_build/aes_128.gates 96
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../mpc_b200/csrc/plan.hpp"

using namespace gcb;

struct Acc {
    uint64_t instr = 0, wavefronts = 0, ideal8 = 0;   // ideal8: distinct words (8 per wavefront at best)
    void add(const std::vector<int>& lane_slot) {     // -1 = inactive lane; 32 entries
        bool any = false;
        for (int q = 0; q < 4; q++) {
            int cnt[8] = {0};
            int seen[8], ns = 0;
            for (int l = 0; l < 8; l++) {
                const int s = lane_slot[q * 8 + l];
                if (s < 0) continue;
                bool dup = false;
                for (int i = 0; i < ns; i++) dup |= seen[i] == s;
                if (dup) continue;
                seen[ns++] = s;
                cnt[s & 7]++;
            }
            int m = 0;
            for (int b = 0; b < 8; b++) m = cnt[b] > m ? cnt[b] : m;
            wavefronts += m;
            ideal8 += ns;
            any |= ns > 0;
        }
        instr += any;
    }
};

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: plan_model file.gates team_threads [balance]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    uint32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return 1;
    std::vector<gcb_gate> gates(hdr[0]);
    if (fread(gates.data(), sizeof(gcb_gate), hdr[0], f) != hdr[0]) return 1;
    fclose(f);
    const uint32_t TT = (uint32_t)atoi(argv[2]);
    PlanSpec spec;
    spec.gates = gates.data(); spec.num_gates = hdr[0]; spec.num_wires = hdr[1];
    for (uint32_t i = 0; i < hdr[2]; i++) spec.live_in.push_back(i);
    for (uint32_t i = 0; i < hdr[3]; i++) spec.live_out.push_back(hdr[1] - hdr[3] + i);
    Plan plan;
    std::string err;
    const int balance = argc > 3 ? atoi(argv[3]) : 0;      // 0 none, 1 garbler, 2 evaluator (plan.cpp: balanced_levels)
    const uint32_t hot_cap = argc > 4 ? (uint32_t)atoi(argv[4]) : 0;
    const int max_fanin_arg = argc > 5 ? atoi(argv[5]) : NODE_MAX_FANIN;   // 0: all labels in shared memory
    if (int rc = build_plan(spec, plan, err, max_fanin_arg, balance, hot_cap)) { fprintf(stderr, "plan: %d %s\n", rc, err.c_str()); return 1; }
    printf("balance %d  policy %d  garble_passes %u  eval_passes %u  hot slots %u of %u  cold accesses %llu (node loads %u, gates %u)\n", balance, plan.policy,
           plan.info.garble_passes, plan.info.eval_passes, plan.info.num_hot_slots, plan.info.num_slots, (unsigned long long)plan.cold_accesses,
           plan.node_loads, plan.info.num_and + plan.info.num_inv + plan.info.num_or);
    if (!plan.phase_copy.empty()) {
        uint64_t ev = 0, rl = 0;
        for (const auto& c : plan.phase_copy) { ev += c[1]; rl += c[2]; }
        printf("live-range splitting: %llu evicts, %llu reloads per instance, scratch %u labels\n", (unsigned long long)ev, (unsigned long long)rl,
               plan.info.num_slots - plan.info.num_hot_slots);
    }
    const gcb_plan_info& in = plan.info;
    printf("gates %u  and %u inv %u or %u free %u  slots %u  steps %u  phases %zu  waves %zu  nodes %zu  node_loads %u\n",
           in.num_gates, in.num_and, in.num_inv, in.num_or, in.num_free, in.num_slots, in.num_steps, plan.phases.size(),
           plan.waves.size(), plan.nodes.size(), plan.node_loads);

    // what lane splitting would cost: a wave whose nodes fill at most TT / f lanes gives each node f adjacent lanes
    // (leaves dealt round robin), combined with log2(f) shuffle steps (counted as one loop iteration each)
    {
        uint64_t now = 0, split = 0, rows = 0;
        for (const WaveRec& wr : plan.waves) {
            const uint32_t nrow = (wr.count + TT - 1) / TT;
            uint32_t f = 1;
            while (f < 4 && (wr.count * f * 2 + TT - 1) / TT == nrow) f *= 2;
            for (uint32_t base = 0; base < wr.count; base += 32 / f) {
                uint32_t kmax = 0;
                for (uint32_t l = 0; l < 32 / f && base + l < wr.count; l++) kmax = std::max<uint32_t>(kmax, plan.nodes[wr.first + base + l].k);
                split += (kmax + f - 1) / f + (f == 4 ? 2 : f == 2 ? 1 : 0);
                rows++;
            }
            for (uint32_t base = 0; base < wr.count; base += 32) {
                uint32_t kmax = 0;
                for (uint32_t l = 0; l < 32 && base + l < wr.count; l++) kmax = std::max<uint32_t>(kmax, plan.nodes[wr.first + base + l].k);
                now += kmax;
            }
        }
        printf("node row loop iterations per instance: now %llu, with lane splitting %llu (%llu warp rows)\n", (unsigned long long)now,
               (unsigned long long)split, (unsigned long long)rows);
    }
    // warp-level hash calls of the two-block kernels (gc_kernels.cuh: the cipher pass loop) and what they would be with other team widths
    for (uint32_t tt : {32u, 64u, 96u, 128u}) {
        for (int garble = 1; garble >= 0; garble--) {
            uint64_t dbl = 0, sgl = 0, rounds = 0;
            std::vector<uint32_t> hist(9, 0);
            for (const PhaseRec& ph : plan.phases) {
                const uint32_t ntask = garble ? 4 * ph.n_quad + 2 * ph.n_inv : 2 * ph.n_quad + ph.n_inv;
                if (!ntask) continue;
                hist[std::min<uint32_t>(8, (ntask + 31) / 32)]++;
                const uint32_t per_thread = (ntask + tt - 1) / tt;
                for (uint32_t k0 = 0; k0 < per_thread;) {
                    const uint32_t uu = per_thread - k0 >= 2 ? 2 : 1;
                    bool any2 = false;
                    for (uint32_t w = 0; w < tt / 32; w++) {
                        if (uu == 2 && (k0 + 1) * tt + 32 * w < ntask) { dbl++; any2 = true; }
                        else if (k0 * tt + 32 * w < ntask) sgl++;
                    }
                    rounds += 1;   // serial steps of the team (a double call costs about 1.3 single calls of latency)
                    (void)any2;
                    k0 += uu;
                }
            }
            printf("%s TT=%u: double calls %llu, single calls %llu, serial call rounds %llu", garble ? "garble" : "eval  ", tt,
                   (unsigned long long)dbl, (unsigned long long)sgl, (unsigned long long)rounds);
            if (tt == 32) { printf("  | phases by 32-block passes 1..8+:"); for (int i = 1; i <= 8; i++) printf(" %u", hist[i]); }
            printf("\n");
        }
    }
    Acc node_ld, node_st, g_ld, g_st, e_ld, e_st;
    uint64_t barriers = 0, garble_passes = 0, eval_passes = 0, node_iters = 0, node_floor = 0;
    std::vector<int> ls(32);
    for (const PhaseRec& ph : plan.phases) {
        const uint32_t nw = ph.n_waves & 0x7fffffffu;
        for (uint32_t w = 0; w < nw; w++) {
            const WaveRec& wr = plan.waves[ph.wave_first + w];
            barriers++;
            for (uint32_t base = 0; base < wr.count; base += TT) {
                for (uint32_t wb = 0; wb < TT && base + wb < wr.count; wb += 32) {
                    node_iters++;
                    uint32_t kmax = 0;
                    for (int l = 0; l < 32; l++) {
                        const uint32_t j = base + wb + l;
                        if (j < wr.count) kmax = std::max<uint32_t>(kmax, plan.nodes[wr.first + j].k);
                    }
                    for (uint32_t k = 0; k < kmax; k++) {
                        for (int l = 0; l < 32; l++) {
                            const uint32_t j = base + wb + l;
                            ls[l] = -1;
                            if (j < wr.count && k < plan.nodes[wr.first + j].k) ls[l] = plan.nodes[wr.first + j].leaf[k];
                        }
                        node_ld.add(ls);
                        for (int q = 0; q < 4; q++) {          // floor: one wavefront per quarter warp with an active lane
                            bool any = false;
                            for (int l = 0; l < 8; l++) any |= ls[q * 8 + l] >= 0;
                            node_floor += any;
                        }
                    }
                    for (int l = 0; l < 32; l++) {
                        const uint32_t j = base + wb + l;
                        ls[l] = j < wr.count ? (int)plan.nodes[wr.first + j].dst : -1;
                    }
                    node_st.add(ls);
                }
            }
        }
        const uint32_t ngate = ph.n_quad + ph.n_inv;
        if (ngate) barriers++;
        for (int garble = 0; garble < 2; garble++) {
            const uint32_t qs = garble ? 4 : 2, is = garble ? 2 : 1;
            const uint32_t ntask = qs * ph.n_quad + is * ph.n_inv;
            Acc& ld = garble ? g_ld : e_ld;
            Acc& st = garble ? g_st : e_st;
            for (uint32_t t0 = 0; t0 < ntask; t0 += 32) {
                (garble ? garble_passes : eval_passes)++;
                std::vector<int> la(32, -1), lb(32, -1), lc(32, -1);
                for (int l = 0; l < 32; l++) {
                    const uint32_t t = t0 + l;
                    if (t >= ntask) continue;
                    uint32_t gi, k;
                    if (t < qs * ph.n_quad) { gi = t / qs; k = t % qs; }
                    else { gi = ph.n_quad + (t - qs * ph.n_quad) / is; k = (t - qs * ph.n_quad) % is; }
                    const GateRec& g = plan.crecs[ph.cipher_first + gi];
                    la[l] = g.a; lb[l] = g.b;
                    if (k == 0) lc[l] = g.c;
                }
                ld.add(la); ld.add(lb); st.add(lc);
            }
        }
    }
    // invariant the kernels rely on: the reads and writes of one step (a wave, a cipher level) are unordered,
    // so no slot may be read and written in the same step
    {
        uint64_t bad = 0;
        for (size_t pi = 0; pi < plan.phases.size(); pi++) {
            const PhaseRec& ph = plan.phases[pi];
            const uint32_t nw = ph.n_waves & 0x7fffffffu;
            for (uint32_t w = 0; w < nw; w++) {
                const WaveRec& wr = plan.waves[ph.wave_first + w];
                std::vector<uint8_t> rd(65536, 0);
                for (uint32_t j = 0; j < wr.count; j++)
                    for (uint32_t k = 0; k < plan.nodes[wr.first + j].k; k++) rd[plan.nodes[wr.first + j].leaf[k]] = 1;
                for (uint32_t j = 0; j < wr.count; j++)
                    if (rd[plan.nodes[wr.first + j].dst]) { bad++; printf("phase %zu wave %u: node %u writes slot %u that the wave reads\n", pi, w, j, plan.nodes[wr.first + j].dst); }
            }
            std::vector<uint8_t> rd(65536, 0);
            for (uint32_t g = 0; g < ph.n_quad + ph.n_inv; g++) { rd[plan.crecs[ph.cipher_first + g].a] = 1; rd[plan.crecs[ph.cipher_first + g].b] = 1; }
            for (uint32_t g = 0; g < ph.n_quad + ph.n_inv; g++)
                if (rd[plan.crecs[ph.cipher_first + g].c]) { bad++; if (bad < 20) printf("phase %zu cipher level: gate %u writes slot %u that the level reads\n", pi, g, plan.crecs[ph.cipher_first + g].c); }
        }
        printf("same-step read/write overlaps: %llu\n", (unsigned long long)bad);
    }
    // AES warp-blocks (160 wavefronts each) per instance: the kernels' pass policy against a packing in
    // which each warp takes a contiguous block of tasks and runs only the block slots it needs
    {
        uint64_t cur[2] = {0, 0}, packed[2] = {0, 0}, ideal[2] = {0, 0};
        const uint32_t W = TT / 32;
        for (const PhaseRec& ph : plan.phases)
            for (int garble = 0; garble < 2; garble++) {
                const uint32_t ntask = (garble ? 4 : 2) * ph.n_quad + (garble ? 2 : 1) * ph.n_inv;
                if (!ntask) continue;
                const uint32_t per_thread = (ntask + TT - 1) / TT;
                for (uint32_t k0 = 0; k0 < per_thread;) {
                    const uint32_t left = per_thread - k0, uu = left >= 2 ? 2 : 1;
                    if (uu == 2) cur[garble] += 2 * W;
                    else for (uint32_t w = 0; w < W; w++) cur[garble] += (k0 * TT + 32 * w < ntask);
                    k0 += uu;
                }
                // packed: rounds of up to 64 tasks per warp (two slots), W warps
                uint32_t rest = ntask;
                while (rest) {
                    for (uint32_t w = 0; w < W && rest; w++) {
                        const uint32_t take = rest < 64 ? rest : 64;
                        packed[garble] += (take + 31) / 32;
                        rest -= take;
                    }
                }
                ideal[garble] += (ntask + 31) / 32;
            }
        printf("AES warp-blocks per instance: garble now %llu packed %llu floor %llu ; eval now %llu packed %llu floor %llu\n",
               (unsigned long long)cur[1], (unsigned long long)packed[1], (unsigned long long)ideal[1],
               (unsigned long long)cur[0], (unsigned long long)packed[0], (unsigned long long)ideal[0]);
    }
    auto pr = [](const char* n, const Acc& a) {
        printf("%-14s instr %8llu  wavefronts %8llu  ideal %8.0f  (x%.2f)\n", n, (unsigned long long)a.instr,
               (unsigned long long)a.wavefronts, a.ideal8 / 8.0, a.wavefronts / (a.ideal8 / 8.0 + 1e-9));
    };
    pr("node loads", node_ld); pr("node stores", node_st);
    printf("node loads floor (one wavefront per active quarter warp and leaf position): %llu\n", (unsigned long long)node_floor);
    pr("garble loads", g_ld); pr("garble stores", g_st);
    pr("eval loads", e_ld); pr("eval stores", e_st);
    printf("barriers %llu  node warp-iterations %llu  cipher warp-passes garble %llu eval %llu\n",
           (unsigned long long)barriers, (unsigned long long)node_iters, (unsigned long long)garble_passes,
           (unsigned long long)eval_passes);
    const uint64_t aes_g = (uint64_t)in.garble_hashes * 160 / 32, aes_e = (uint64_t)in.eval_hashes * 160 / 32;
    printf("per instance wavefronts: garble AES %llu + labels %llu ; eval AES %llu + labels %llu\n",
           (unsigned long long)aes_g, (unsigned long long)(node_ld.wavefronts + node_st.wavefronts + g_ld.wavefronts + g_st.wavefronts),
           (unsigned long long)aes_e, (unsigned long long)(node_ld.wavefronts + node_st.wavefronts + e_ld.wavefronts + e_st.wavefronts));
    return 0;
}
