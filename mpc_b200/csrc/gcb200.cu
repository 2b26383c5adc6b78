// gcb200.cu -- the C ABI of include/gcb200.h: argument checking, device
// selection, plan upload, kernel launches and the host-buffer staging paths.
// There is no CPU fallback in this file: without a CUDA device every compute
// entry point fails with GCB_E_CUDA.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gcb200.h"
#include "gc_launch.hpp"
#include "ot_kernels.cuh"
#include "stream_kernels.cuh"
#include "plan.hpp"
#include "async.hpp"

namespace gcb {

// ------------------------------------------------------------------ errors ----
static thread_local std::string tl_err;
static thread_local int tl_device = -1;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    tl_err = buf;
    return code;
}
// Kernels this library has launched (every launch site checks cudaGetLastError through launched()).
static std::atomic<uint64_t> g_launches{0};
static cudaError_t launched() { g_launches++; return cudaGetLastError(); }
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(GCB_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)

// The library keeps eight streams per device busy beside the caller's own; the default of 8 hardware work queues
// per context makes streams share queues (false dependencies between jobs).  Takes effect when the library is
// loaded before the CUDA context is created and the variable is not set by the user.
__attribute__((constructor)) static void more_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

// ----------------------------------------------------------------- devices ----
constexpr uint32_t kCounterRing = 1024;

struct DeviceInfo {
    int sm_count = 0;
    uint32_t smem_base = 1024;                  // shared-window address at which dynamic shared memory starts
    uint32_t* counters = nullptr;               // ring of work-claim counters
    std::atomic<uint32_t> next{0};
};
static std::mutex g_dev_mu;
static std::map<int, std::unique_ptr<DeviceInfo>> g_devs;

template <class K>
static cudaError_t opt_in(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemOptin);
}

// Where dynamic shared memory starts in the shared window (the AES tables are placed at
// the next multiple of 64 KiB, see aes_core.cuh).
__global__ void smem_base_probe(uint32_t* out) {
    extern __shared__ __align__(16) uint8_t smem[];
    if (threadIdx.x == 0) *out = smem_addr(smem);
}

// Which device(s) a call runs on.  A thread that called gcb_set_device(d >= 0) uses d.  Otherwise the
// process-wide list of gcb_set_devices applies: calls that can split their work (host-pointer garble /
// eval / IKNP / MiTCCRH, and the _dev forms with GCB_FLAG_FANOUT) spread it over the whole list, all
// others run on its first entry.  With neither set: LOCAL_RANK, else device 0.
static std::mutex g_devlist_mu;
static std::shared_ptr<const std::vector<int>> g_devlist;
static int default_device() {
    static const int d = [] { const char* lr = getenv("LOCAL_RANK"); return lr ? atoi(lr) : 0; }();
    return d;
}
static std::vector<int> call_devices() {
    if (tl_device >= 0) return {tl_device};
    std::shared_ptr<const std::vector<int>> l;
    { std::lock_guard<std::mutex> lk(g_devlist_mu); l = g_devlist; }
    if (l && !l->empty()) return *l;
    return {default_device()};
}
static int current_device() { return call_devices()[0]; }

// Makes `device` current for the calling thread and returns its per-device state (created on first use).
int use_device(int device, DeviceInfo** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(GCB_E_CUDA, "no CUDA device available (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(GCB_E_CUDA, "device %d out of range (%d devices)", device, count);
    CK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto& slot = g_devs[device];
    if (!slot) {
        auto di = std::make_unique<DeviceInfo>();
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(GCB_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                        prop.major, prop.minor);
        di->sm_count = prop.multiProcessorCount;
        CK(cudaMalloc(&di->counters, kCounterRing * sizeof(uint32_t)));
        CK(cudaMemset(di->counters, 0, kCounterRing * sizeof(uint32_t)));
        smem_base_probe<<<1, 32, 1024>>>(di->counters);
        CK(cudaMemcpy(&di->smem_base, di->counters, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemset(di->counters, 0, sizeof(uint32_t)));
        CK(gc_opt_in_garble((int)kSmemOptin));
        CK(gc_opt_in_eval((int)kSmemOptin));
        CK(opt_in(hash_half_kernel<10>)); CK(opt_in(hash_half_kernel<12>)); CK(opt_in(hash_half_kernel<14>));
        CK(opt_in(mitccrh_kernel));
        CK(opt_in(iknp_kernel<false, false>)); CK(opt_in(iknp_kernel<true, false>));
        CK(opt_in(iknp_kernel<false, true>)); CK(opt_in(iknp_kernel<true, true>));
        CK(opt_in(cot_kernel<COT_SEND>)); CK(opt_in(cot_kernel<COT_RECEIVE>));
        CK(opt_in(cot_kernel<ROT_SEND>)); CK(opt_in(cot_kernel<ROT_RECEIVE>));
        CK(opt_in(iknp_check_kernel));
        slot = std::move(di);
    }
    if (out) *out = slot.get();
    return GCB_OK;
}
// The calling thread's device (see call_devices); remembered in tl_cur for the code that follows.
static thread_local int tl_cur = 0;
int select_device(DeviceInfo** out) {
    tl_cur = current_device();
    return use_device(tl_cur, out);
}

// A zeroed claim counter for one launch on `stream`.
static int fresh_counter(DeviceInfo* di, cudaStream_t stream, uint32_t** out) {
    uint32_t* c = di->counters + (di->next.fetch_add(1) % kCounterRing);
    CK(cudaMemsetAsync(c, 0, sizeof(uint32_t), stream));
    *out = c;
    return GCB_OK;
}

// Developer hook (not in the public header): device buffer of 4 timestamps per phase
// written by block 0 / team 0 of the next garble launches; nullptr switches it off.
static long long* g_trace = nullptr;
extern "C" void gcb_debug_set_trace(long long* dev_buf) { g_trace = dev_buf; }
// Developer hook (not in the public header): the staging copier of async.hpp, so that it can be tested without a device.
extern "C" void gcb_debug_host_copy(void* dst, const void* src, size_t n) { host_copy(dst, src, n); }

// ------------------------------------------------------------- plan upload ----
DevicePlan::~DevicePlan() {
    // best effort: the context may already be gone at process exit
    if (phases) cudaFree(phases);
    if (nodes) cudaFree(nodes);
    if (crecs) cudaFree(crecs);
    if (nout_wire) cudaFree(nout_wire);
    if (cout_wire) cudaFree(cout_wire);
    if (live_in) cudaFree(live_in);
    if (live_out) cudaFree(live_out);
    if (copies) cudaFree(copies);
}

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& v) {
    const size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

int plan_on_device(const Plan& plan, int device, uint32_t team_threads, std::shared_ptr<DevicePlan>* out) {
    std::lock_guard<std::mutex> lk(plan.mu);
    const auto key = std::make_pair(device, team_threads);
    auto it = plan.dev.find(key);
    if (it != plan.dev.end()) { *out = it->second; return GCB_OK; }
    auto dp = std::make_shared<DevicePlan>();
    dp->device = device;
    dp->team_threads = team_threads;
    // node rows: every wave padded to whole rows of team_threads records (plan.hpp)
    const uint32_t TT = team_threads;
    std::vector<DevPhaseRec> ph;
    std::vector<NodeRec> rows;
    std::vector<uint32_t> row_wire;
    ph.reserve(plan.phases.size() + 2);
    for (const PhaseRec& src : plan.phases) {
        DevPhaseRec d{};
        d.cipher_first = src.cipher_first; d.n_quad = src.n_quad; d.n_inv = src.n_inv;
        d.row_first = (uint32_t)(rows.size() / TT);
        const uint32_t nw = src.n_waves & 0x7fffffffu;
        for (uint32_t w = 0; w < nw; w++) {
            const WaveRec& wr = plan.waves[src.wave_first + w];
            const uint32_t nrow = (wr.count + TT - 1) / TT;
            for (uint32_t r = 0; r < nrow; r++) {
                const uint8_t end = r + 1 == nrow ? NODE_WAVE_END : 0;
                for (uint32_t t = 0; t < TT; t++) {
                    const uint32_t j = r * TT + t;
                    NodeRec rec{};
                    uint32_t wire = 0;
                    if (j < wr.count) {
                        rec = plan.nodes[wr.first + j];
                        rec.parity = (uint8_t)((rec.parity & NODE_PARITY) | NODE_ACTIVE);
                        wire = plan.nout_wire[wr.first + j];
                    }
                    rec.parity |= end;
                    rows.push_back(rec);
                    row_wire.push_back(wire);
                }
            }
        }
        d.n_rows = (uint32_t)(rows.size() / TT) - d.row_first;
        if (src.n_waves & 0x80000000u) d.n_rows |= 0x80000000u;
        if (!plan.phase_copy.empty()) {                     // live-range splitting: this phase's evict / reload lists
            const auto& c = plan.phase_copy[ph.size()];
            d.copy_first = c[0];
            d.n_evict = c[1] | (c[2] << 16);                // the kernels read (copy_first, evicts | reloads << 16)
            d.n_reload = 0;
        }
        ph.push_back(d);
    }
    ph.push_back(DevPhaseRec{});                        // the kernels read two records ahead
    ph.push_back(DevPhaseRec{});
    rows.resize(rows.size() + (size_t)GC_NODE_PIPE_MAX * TT, NodeRec{});   // and GC_NODE_PIPE_MAX rows ahead
    CK(upload(&dp->phases, ph));
    CK(upload(&dp->nodes, rows));
    CK(upload(&dp->crecs, plan.crecs));
    CK(upload(&dp->nout_wire, row_wire));
    CK(upload(&dp->cout_wire, plan.cout_wire));
    CK(upload(&dp->live_in, plan.live_in));
    CK(upload(&dp->live_out, plan.live_out));
    {
        std::vector<uint32_t> packed;                       // (src, dst) pairs -> src | dst << 16, plus one spare (read-ahead)
        for (size_t i = 0; i + 1 < plan.copies.size(); i += 2) packed.push_back(plan.copies[i] | (plan.copies[i + 1] << 16));
        packed.push_back(0);
        CK(upload(&dp->copies, packed));
    }
    plan.dev[key] = dp;
    *out = dp;
    return GCB_OK;
}

// Team geometry for a plan: how many instances one SM keeps resident, how many threads work on
// each, which kernel variant runs (AES blocks a thread interleaves, resident T-tables).
// GCB_NT / GCB_TEAMS / GCB_ILP / GCB_TEAM_THREADS / GCB_STAGGER override the choice (tuning experiments).
struct Geometry { uint32_t n_teams = 0, team_threads = 0, ilp = 1, nt = 4, stagger = 0, n_smem = 0; int spill = 0; bool split = false, twin = false; };
// width = AES blocks per cipher level of the garbler (4 per AND / OR, 2 per INV), averaged.
// num_hot < num_slots: a hot / cold plan (plan.cpp) -- only the hot labels need shared memory.
// copies: the hot / cold plan moves its values with copy lists (live-range splitting) instead of in-place scratch access.
static Geometry compute_geometry(uint32_t num_slots, uint32_t width, uint32_t smem_base, uint32_t num_hot = 0, bool copies = false) {
    Geometry g;
    if (num_hot && num_hot < num_slots) {
        size_t n = teams_that_fit(num_hot, smem_base, 2, &g.split);
        if (n > 16) n = 16;
        if (n == 0) return g;
        if (const char* e = getenv("GCB_TEAMS")) { const int v = atoi(e); if (v >= 1 && (size_t)v <= n) n = (size_t)v; }
        g.n_teams = (uint32_t)n; g.ilp = 1; g.nt = 2; g.spill = copies ? 2 : 1; g.n_smem = num_hot;
        g.team_threads = n >= 8 ? 32u : 32u * (uint32_t)(512 / 32 / n > 3 ? 3 : 512 / 32 / n);
        g.stagger = n > 1 ? 100000 : 0;
        if (const char* e = getenv("GCB_STAGGER")) g.stagger = (uint32_t)atoi(e);
        // 16 one-warp teams run as 8 lock-step pairs (gc_kernels.cuh: team_ctx)
        g.twin = g.team_threads == 32 && (n == 16 || n == 8) && n > 8;
        if (const char* e = getenv("GCB_TWIN")) g.twin = atoi(e) != 0 && g.team_threads == 32 && n % 8 == 0;
        return g;
    }
    bool split4 = false, split2 = false;
    const size_t n4 = teams_that_fit(num_slots, smem_base, 4, &split4), n2 = teams_that_fit(num_slots, smem_base, 2, &split2);
    // Wide levels keep the shared-memory pipe busy with a few resident instances: four tables.
    // Deep, narrow circuits (sha256: 39 blocks per level) are bound by the latency of each level,
    // so the number of resident instances is what counts: two tables, twice the label space.
    // Wide levels: what counts is warps per SM (the 512-thread variants hold 16), then teams -- smaller teams wait less
    // at their barriers, one-warp teams have none -- and two resident tables only cost eight PRMTs per round, which fit
    // the ALU pipe's slack.  For each table count the team count n <= fit that gives the most warps (n teams of
    // min(3, 16 / n) warps), ties to the larger n; two tables when that is strictly better.  Measured
    // (profiles/r02_geometry_p.txt, r02_geometry_z.txt): aes_128 5 x 96 on four tables 4,126 M AND/s, 8 x 64 on two 4,222;
    // mul64 10 x 32 on four 3,257, 8 x 64 on four 3,307, 16 x 32 on two 3,756.
    auto best_teams = [](size_t fit, uint32_t* warps) {
        size_t best = 0;
        *warps = 0;
        for (size_t k = 1; k <= fit && k <= 16; k++) {
            const uint32_t w = (uint32_t)(k * std::min<size_t>(3, 16 / k));
            if (w >= *warps) { *warps = w; best = k; }
        }
        return best;
    };
    uint32_t w4 = 0, w2 = 0;
    const size_t b4 = best_teams(n4, &w4), b2 = best_teams(n2, &w2);
    const bool tuned = width >= 128 && n4 < 16 && !getenv("GCB_TEAMS") && !getenv("GCB_TEAM_THREADS");
    const bool wide2 = tuned && (w2 > w4 || (w2 == w4 && b2 > b4));
    uint32_t nt = (n4 == 0 || wide2 || (width < 128 && n2 > n4 && n4 < 16)) ? 2u : 4u;
    if (const char* e = getenv("GCB_NT")) { const int v = atoi(e); if (v == 2 || v == 4) nt = (uint32_t)v; }
    size_t n = nt == 2 ? n2 : n4;
    if (tuned && !getenv("GCB_NT")) n = nt == 2 ? b2 : b4;
    g.split = nt == 2 ? split2 : split4;
    g.n_smem = num_slots;
    if (n == 0) {
        // More live labels than one team's share of shared memory holds even beside two tables: one team
        // per SM keeps the low (hot) slots in the region above the tables and spills the rest to a
        // global-memory scratch (gc_kernels.cuh: SlotsSpill).
        const size_t pad = table_pad(smem_base);
        const size_t above = kSmemOptin - pad - (size_t)aes_table_bytes(2), below = pad - kTeamHeaderBytes;
        g.n_teams = 1; g.team_threads = 256; g.ilp = 2; g.nt = 2; g.stagger = 0; g.spill = 1; g.split = true;
        g.n_smem = (uint32_t)((above > below ? above : below) / 16);
        return g;
    }
    if (n >= 32) n = 32; else if (n > 16) n = 16;
    if (const char* e = getenv("GCB_TEAMS")) { const int v = atoi(e); if (v >= 1 && (size_t)v <= n) n = (size_t)v; }
    // measured on B200 (tools/tune_geometry.py): two interleaved AES blocks per thread and
    // three-warp teams for wide levels; one block per thread for narrow ones and one-warp teams
    uint32_t ilp = (n > 16 || (nt == 2 && width < 128)) ? 1 : 2;
    if (const char* e = getenv("GCB_ILP")) { const int v = atoi(e); if (v == 1 || v == 2) ilp = (uint32_t)v; }
    // threads per CTA: the 512-thread variants have 128 registers per thread and no spills
    const uint32_t maxt = (ilp == 1 && nt == 4) ? 1024 : 512;
    while (n * 32 > maxt) n--;                      // at least one warp per team
    uint32_t tt = 32u * (uint32_t)(maxt / 32 / n);
    if (tt > 96) tt = 96;
    if (nt == 2 && ilp == 1 && width < 64 && n >= 8) tt = 32;   // very narrow levels, many teams: one-warp teams (no named barriers) measured best
    if (n > 16) tt = 32;                            // named barriers: at most 16 multi-warp teams
    if (const char* e = getenv("GCB_TEAM_THREADS")) {
        const int v = atoi(e);
        if (v >= 32 && v % 32 == 0 && (size_t)v * n <= maxt && (v == 32 || n <= 16)) tt = (uint32_t)v;
    }
    g.n_teams = (uint32_t)n; g.team_threads = tt; g.ilp = ilp; g.nt = nt;
    // one-warp teams start ~50 us apart (in lock step they collide in the same phases: sha256 x 16 teams 13.7 instead of
    // 11.2-12.4 ms); multi-warp teams drift apart on their own, a start offset only idles the SM (aes_128: +2.7 % without)
    g.stagger = n > 1 && tt == 32 ? 100000 : 0;
    if (const char* e = getenv("GCB_STAGGER")) g.stagger = (uint32_t)atoi(e);
    if (const char* e = getenv("GCB_TWIN")) g.twin = atoi(e) != 0 && tt == 32 && n % 8 == 0 && n <= 16;
    return g;
}
static uint32_t plan_width(const Plan& plan) {
    const size_t np = plan.phases.size();
    return np ? (uint32_t)(plan.info.garble_hashes / np) : 0u;
}
void team_geometry(Plan& plan) {                    // what gcb_plan_get_info reports (typical device)
    const Geometry g = compute_geometry(plan.info.num_slots, plan_width(plan), kAssumedSmemBase, plan.info.num_hot_slots, !plan.phase_copy.empty());
    plan.info.teams_per_sm = g.n_teams; plan.info.team_threads = g.team_threads;
    plan.ilp = g.ilp; plan.stagger = g.stagger;
}
size_t gc_smem_bytes(uint32_t num_slots, uint32_t n_teams, uint32_t nt, uint32_t smem_base, bool split) {
    const size_t per_team = team_block_bytes(num_slots, split);
    const size_t pad = table_pad(smem_base), in_a = teams_below(num_slots, n_teams, smem_base, split);
    return pad + (size_t)aes_table_bytes((int)nt) + (n_teams > in_a ? (n_teams - in_a) * per_team : 0);
}

// The plan a call runs on: the flattened one, or -- when the caller wants every wire
// label (Garbled.Wires / Eval's in-place wires) -- one that keeps every XOR gate.
static int plan_for(const gcb_plan* plan, bool full, const Plan** out, uint64_t batch = 0) {
    if (!full) {
        *out = &plan->p;
        const Plan& base = plan->p;
        // a batch that overflows the resident instances of the all-hot plan may run faster on a plan that holds more
        // instances with only a hot subset of the labels in shared memory (plan.cpp: build_best_plan)
        // (narrow circuits only: wide cipher levels are bound by the shared-memory pipe, not by resident instances)
        if (batch > (uint64_t)base.info.teams_per_sm * 148 && base.info.teams_per_sm < 16 && plan_width(base) < 128 &&
            base.info.num_slots >= 256) {
            std::lock_guard<std::mutex> lk(plan->many_mu);
            if (!plan->many_tried) {
                plan->many_tried = true;
                PlanSpec spec;
                spec.gates = base.gates.data(); spec.num_gates = (uint32_t)base.gates.size(); spec.num_wires = base.info.num_wires;
                for (uint32_t i = 0; i < base.info.num_inputs; i++) spec.live_in.push_back(i);
                for (uint32_t i = 0; i < base.info.num_outputs; i++) spec.live_out.push_back(base.info.num_wires - base.info.num_outputs + i);
                auto mp = std::make_unique<Plan>();
                std::string err;
                if (build_best_plan(spec, *mp, err) == GCB_OK && mp->info.num_hot_slots < mp->info.num_slots) {
                    team_geometry(*mp);
                    if (mp->info.teams_per_sm > base.info.teams_per_sm) plan->many = std::move(mp);
                }
            }
            if (plan->many) *out = plan->many.get();
        }
        return GCB_OK;
    }
    std::lock_guard<std::mutex> lk(plan->full_mu);
    if (!plan->full) {
        const Plan& base = plan->p;
        PlanSpec spec;
        spec.gates = base.gates.data(); spec.num_gates = (uint32_t)base.gates.size(); spec.num_wires = base.info.num_wires;
        for (uint32_t i = 0; i < base.info.num_inputs; i++) spec.live_in.push_back(i);
        for (uint32_t i = 0; i < base.info.num_outputs; i++) spec.live_out.push_back(base.info.num_wires - base.info.num_outputs + i);
        auto fp = std::make_unique<Plan>();
        std::string err;
        int rc = build_plan(spec, *fp, err, 2);
        if (rc) return fail(rc, "%s", err.c_str());
        team_geometry(*fp);
        if (fp->info.teams_per_sm == 0)
            return fail(GCB_E_TOO_LARGE, "circuit keeps %u wire labels live with every wire materialised", fp->info.num_slots);
        plan->full = std::move(fp);
    }
    *out = plan->full.get();
    return GCB_OK;
}

static int check_keylen(uint32_t keylen) {
    if (keylen != 16 && keylen != 24 && keylen != 32)
        return fail(GCB_E_KEYLEN, "crypto/aes: invalid key size %u", keylen);
    return GCB_OK;
}

static int launch_gc(bool garble, const Plan& plan, DeviceInfo* di, int device, const uint8_t* keys,
                     uint32_t keylen, uint32_t key_stride, uint32_t batch, const gcb_label* r,
                     const gcb_label* in_labels, gcb_label* tables, void* io, void* wires_full,
                     cudaStream_t stream, const uint32_t* in_ids = nullptr, const uint32_t* out_ids = nullptr,
                     uint4* const* pages = nullptr) {
    const gcb_plan_info& in = plan.info;
    Geometry geo = compute_geometry(in.num_slots, plan_width(plan), di->smem_base, in.num_hot_slots, !plan.phase_copy.empty());
    if (geo.n_teams == 0) return fail(GCB_E_TOO_LARGE, "circuit keeps %u wire labels live; they do not fit on chip", in.num_slots);
    // A batch smaller than one wave is spread over the SMs instead of filling a few of them: a CTA then carries only as
    // many teams as it needs (256 aes_128 instances: 2 teams on each of 128 SMs instead of 8 on each of 32), and every
    // instance gets a larger share of its SM's shared-memory pipe.
    const char* spread_env = getenv("GCB_SPREAD");                       // GCB_SPREAD=0: tests that want full CTAs for tiny batches
    const bool spread = !(spread_env && atoi(spread_env) == 0);
    if (spread && (uint64_t)batch < (uint64_t)geo.n_teams * di->sm_count) {
        const uint32_t had = geo.n_teams;
        geo.n_teams = (batch + di->sm_count - 1) / di->sm_count;
        if (geo.n_teams == 0) geo.n_teams = 1;
        geo.twin = geo.twin && geo.n_teams % 8 == 0;
        if (geo.n_teams == 1) geo.stagger = 0;
        // ... and the threads the absent teams would have had go to the present ones: an instance alone on its SM is
        // bound by the latency of its levels.  Measured (profiles/r02_small_batch.txt): aes_128 x 148, garble + eval
        // 1.70 ms in full CTAs on 19 SMs, 0.95 ms spread with 64-thread teams, 0.54 ms with one 256-thread team per SM
        // (480 threads: no better); sha256 / sha512 x 148: 64-thread teams 5 % / 12 % faster than one warp, 96 no better.
        if (!geo.spill && !getenv("GCB_TEAM_THREADS")) {
            const uint32_t maxt = (geo.ilp == 1 && geo.nt == 4) ? 1024u : 512u;
            uint32_t tt = geo.team_threads;
            if (plan_width(plan) >= 128) tt = 256;
            else if (geo.n_teams * 2 <= had && tt < 64) tt = 64;
            while (tt > 32 && tt * geo.n_teams > maxt) tt -= 32;
            if (tt > geo.team_threads) { geo.team_threads = tt; geo.stagger = 0; geo.twin = false; }
        }
    }
    std::shared_ptr<DevicePlan> dp;
    int rc = plan_on_device(plan, device, geo.team_threads, &dp);
    if (rc) return rc;
    GcParams p{};
    p.phases = reinterpret_cast<const uint4*>(dp->phases);
    p.nodes = reinterpret_cast<const uint4*>(dp->nodes);
    p.crecs = reinterpret_cast<const uint4*>(dp->crecs);
    p.nout_wire = dp->nout_wire;
    p.cout_wire = dp->cout_wire;
    p.live_in = reinterpret_cast<const uint2*>(dp->live_in);
    p.live_out = reinterpret_cast<const uint2*>(dp->live_out);
    p.copies = dp->copies;
    p.n_live_in = (uint32_t)plan.live_in.size();
    p.n_phases = (uint32_t)plan.phases.size(); p.n_in = in.num_inputs; p.n_out = in.num_outputs;
    p.n_slots = in.num_slots; p.n_rows = in.num_rows; p.n_wires = in.num_wires;
    p.keys = keys; p.keylen = keylen; p.key_stride = key_stride; p.batch = batch;
    p.r = reinterpret_cast<const uint4*>(r);
    p.in_labels = reinterpret_cast<const uint4*>(in_labels);
    p.tables = reinterpret_cast<uint4*>(tables);
    p.io = reinterpret_cast<uint4*>(io);
    p.wires_full = reinterpret_cast<uint4*>(wires_full);
    p.team_threads = geo.team_threads; p.n_teams = geo.n_teams;
    p.in_ids = in_ids; p.out_ids = out_ids; p.pages = pages;
    p.stagger = geo.stagger;
    p.trace = g_trace;
    rc = fresh_counter(di, stream, &p.counter);
    if (rc) return rc;
    const uint32_t want = (batch + p.n_teams - 1) / p.n_teams;
    const dim3 grid(want < (uint32_t)di->sm_count ? want : (uint32_t)di->sm_count);
    const dim3 block(p.n_teams * p.team_threads);
    const size_t smem = gc_smem_bytes(geo.n_smem, p.n_teams, geo.nt, di->smem_base, geo.split);
    p.hdr_split = geo.split ? 1u : 0u;
    p.twin = geo.twin ? 1u : 0u;
    const bool full = wires_full != nullptr;
    const int mode = pages ? GC_STREAM : full ? GC_FULL : GC_PLAIN;
    const GcVariant var{geo.ilp, geo.nt, geo.spill};
    p.n_smem = geo.n_smem;
    if (geo.spill) {                                   // stream-ordered scratch: concurrent launches never share it
        const size_t bytes = (size_t)grid.x * p.n_teams * (p.n_slots - p.n_smem) * 16;
        CK(cudaMallocAsync(reinterpret_cast<void**>(&p.spill), bytes, stream));
    }
    if (garble) gc_launch_garble(mode, keylen, var, grid, block, smem, stream, p);
    else gc_launch_eval(mode, keylen, var, grid, block, smem, stream, p);
    if (geo.spill) CK(cudaFreeAsync(p.spill, stream));
    CK(launched());
    return GCB_OK;
}

// ---- small label-plumbing kernels (circuit/helpers.go:10-27) -------------------
__global__ void select_labels_kernel(const uint4* wires, size_t wire_stride, const uint8_t* bits, uint4* out,
                                     uint32_t batch, uint32_t n) {
    const size_t total = (size_t)batch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t inst = i / n, k = i % n;
        const uint4* w = wires + (inst * wire_stride + k) * 2;
        out[i] = __ldg(w + (bits[i] ? 1 : 0));            // LabelForBit
    }
}
__global__ void decode_bits_kernel(const uint4* wires, size_t wire_stride, const uint4* labels, uint8_t* bits,
                                   uint32_t batch, uint32_t n) {
    const size_t total = (size_t)batch * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t inst = i / n, k = i % n;
        const uint4* w = wires + (inst * wire_stride + k) * 2;
        const uint4 l = __ldg(labels + i), l0 = __ldg(w), l1 = __ldg(w + 1);
        const bool is0 = l.x == l0.x && l.y == l0.y && l.z == l0.z && l.w == l0.w;
        const bool is1 = l.x == l1.x && l.y == l1.y && l.z == l1.z && l.w == l1.w;
        bits[i] = is0 ? 0 : is1 ? 1 : 2;                  // BitFromLabel: 2 = "unknown label"
    }
}

// ---- host staging helpers -------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
struct StreamGuard {
    cudaStream_t s = nullptr;
    ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};

// ------------------------------------------------------- streaming garbler ----
}  // namespace gcb

struct gcb_stream {
    int device = 0;
    uint32_t batch = 0, keylen = 0, key_stride = 0;
    gcb::DevBuf keys, r;
    std::vector<uint4*> pages;              // host copy of the page table
    gcb::DevBuf page_table;
    size_t page_table_cap = 0;
    cudaStream_t cs = nullptr;
    gcb::DevBuf slab, ser, ids, tmpl, row_pos, wires;
    size_t slab_cap = 0, ser_cap = 0, ids_cap = 0, tmpl_cap = 0, row_pos_cap = 0, wires_cap = 0;
    std::mutex mu;
    // evaluator side: plans recovered from record streams, keyed by a hash of the gate headers
    // `canon`: the canonical gate list (op | tmp flags, three canonical locations per gate) the plan was built
    // from.  The headers come from the peer and the hash is not collision resistant, so a hit is only used
    // after the whole list compared equal.
    struct EvalPlan { gcb::Plan plan; std::vector<uint32_t> in_ids, out_ids, canon; uint32_t ntmp = 0, np = 0; };
    std::unordered_map<uint64_t, std::shared_ptr<EvalPlan>> eval_plans;
    // garbler side: plans built for aliased in / out ids, keyed by the plan and the aliasing pattern
    struct AliasPlan { uint64_t plan_uid; std::vector<uint32_t> loc; std::unique_ptr<gcb::Plan> plan; std::vector<uint32_t> in_locs, out_locs; };
    std::vector<std::shared_ptr<AliasPlan>> alias_plans;
    // permanent id -> canonical location of the call in progress: a flat array with an epoch stamp per
    // entry (no clearing, no hashing: a sub-circuit touches its ids a few hundred thousand times)
    std::vector<uint32_t> perm_loc, perm_epoch;
    uint32_t epoch = 0;
    // evaluator side: the record bytes of call k+1 are copied in (stream ds, second staging buffer) while the
    // kernel of call k runs (stream cs)
    gcb::DevBuf ser2;
    size_t ser2_cap = 0;
    cudaStream_t ds = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_free[2] = {nullptr, nullptr};
    bool ev_valid[2] = {false, false};
    uint32_t ser_turn = 0;
    uint8_t* stage[2] = {nullptr, nullptr};    // garbler side: page-locked staging of header templates / row offsets
    size_t stage_cap[2] = {0, 0};
    ~gcb_stream() {
        for (uint4* p : pages) if (p) cudaFree(p);
        for (uint8_t* h : stage) if (h) cudaFreeHost(h);
        if (ev_h2d) cudaEventDestroy(ev_h2d);
        for (cudaEvent_t e : ev_free) if (e) cudaEventDestroy(e);
        if (ds) cudaStreamDestroy(ds);
        if (cs) cudaStreamDestroy(cs);
    }
};

namespace gcb {

static int grow(DevBuf& b, size_t& cap, size_t bytes) {
    if (bytes <= cap) return GCB_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; cap = 0; }
    bytes += bytes / 4;
    CK(b.alloc(bytes));
    cap = bytes;
    return GCB_OK;
}

// ensureWires (stream_garble.go:78-100): pages up to max_id exist, zero-filled.
constexpr uint32_t kMaxStreamWires = 1u << 28;
static int ensure_wires(gcb_stream* s, uint32_t max_id) {
    const size_t need = ((size_t)max_id >> WF_PAGE_SHIFT) + 1;
    if (need <= s->pages.size()) return GCB_OK;
    const size_t page_bytes = (size_t)s->batch * WF_PAGE_IDS * 16;
    {   // the highest wire id can come from the peer (the OpCircuit header of a record stream): a wire file that cannot
        // fit is refused before the first page is allocated instead of after a hundred thousand of them
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const size_t want = (need - s->pages.size()) * page_bytes;
        if (want > free_b)
            return fail(GCB_E_TOO_LARGE, "wire file up to id %u needs %zu more bytes of device memory for %u instances; %zu are free",
                        max_id, want, s->batch, free_b);
    }
    while (s->pages.size() < need) {
        uint4* p = nullptr;
        CK(cudaMalloc(&p, page_bytes));
        s->pages.push_back(p);
        CK(cudaMemsetAsync(p, 0, page_bytes, s->cs));
    }
    if (s->pages.size() > s->page_table_cap) {
        CK(cudaStreamSynchronize(s->cs));               // kernels may still read the old table
        int rc = grow(s->page_table, s->page_table_cap, s->pages.size() * 2 * sizeof(uint4*));
        if (rc) return rc;
        s->page_table_cap /= sizeof(uint4*);
    }
    CK(cudaMemcpyAsync(s->page_table.p, s->pages.data(), s->pages.size() * sizeof(uint4*), cudaMemcpyHostToDevice,
                       s->cs));
    return GCB_OK;
}

// ------------------------------------------------------------------ jobs -------
// Host-pointer garble / eval as asynchronous jobs (async.hpp).
static ResPool g_res_pool;

static int lease_res(int device, std::unique_ptr<JobRes>* out, int cls = 0, size_t want = 0, size_t want_pin = 0) {
    cudaError_t e = cudaSuccess;
    *out = g_res_pool.lease(device, cls, want, &e, want_pin);
    if (!*out) return cuda_fail(e, "stream creation");
    return GCB_OK;
}
// Grows the part's device arena: to the largest request any part of this device has made so far, not just to this one.
static cudaError_t reserve_dev(JobRes& r) {
    if (r.dev.used <= r.dev.cap) return cudaSuccess;
    const size_t hint = g_res_pool.alloc_hint(r.device);
    return r.dev.reserve(hint > r.dev.used ? hint : r.dev.used);
}

// One operand of a staged call: `per` bytes per instance at host address `host` (already offset to the
// part's first instance); per == 0 marks an operand that is absent.
struct Operand {
    uint8_t* host = nullptr;
    size_t per = 0;
    bool pinned = false;
    size_t dev_off = 0, pin_off = 0;
    uint8_t* dev(const JobRes& r, size_t inst) const { return r.dev.base + dev_off + inst * per; }
};
static Operand operand(const void* host, size_t per) {
    Operand o;
    o.host = static_cast<uint8_t*>(const_cast<void*>(host));
    o.per = host ? per : 0;
    return o;
}

// Instances per slice: whole kernel waves (`wave` = instances the device holds at once), about 64 MB of
// traffic, so that the copies of one slice hide behind the kernel and copies of its neighbours.
static uint32_t slice_instances(size_t per_inst, uint32_t batch, uint32_t wave) {
    constexpr size_t kSliceBytes = 64u << 20;
    if (wave == 0) wave = 1;
    size_t n = kSliceBytes / (per_inst ? per_inst : 1) / wave * wave;
    if (n < wave) n = wave;
    return n >= batch ? batch : (uint32_t)n;
}

// Queue one part of a garble (garble = true) or eval call on `device`: `batch` instances whose operands
// start at the given host addresses.  ins / outs: per-instance operands; `key` is the shared key (key_stride
// == 0) or null.  Returns with everything queued on the part's streams.
// peer.home >= 0: the operands are DEVICE buffers on device peer.home (the fan-out of a _dev call): copies become
// peer copies over NVLink, start once `after` (recorded on the caller's stream) has completed, and `done` is
// recorded behind the last one.
struct PeerMode { int home = -1; cudaEvent_t after = nullptr; cudaEvent_t* done = nullptr; };
static int begin_part(bool garble, const Plan& plan, int device, const uint8_t* key, uint32_t keylen, uint32_t key_stride,
                      uint32_t batch, std::vector<Operand>& ins, std::vector<Operand>& outs, JobPart* part,
                      const PeerMode& peer = PeerMode()) {
    DeviceInfo* di;
    int rc = use_device(device, &di);
    if (rc) return rc;
    size_t per_inst = 0, want = 512, want_pin = 512;
    for (std::vector<Operand>* v : {&ins, &outs})
        for (Operand& o : *v) {
            if (!o.per) continue;
            o.pinned = peer.home >= 0 || is_pinned(o.host);
            want += (((size_t)batch * o.per + 255) & ~(size_t)255);
            if (!o.pinned) want_pin += (((size_t)batch * o.per + 255) & ~(size_t)255);
        }
    if ((rc = lease_res(device, &part->res, garble ? 0 : 1, want, want_pin))) return rc;
    JobRes& r = *part->res;
    r.dev.used = r.pin.used = 0; r.ev_used = 0; r.sealed = false;
    const size_t key_pin = r.pin.take(64), key_dev = r.dev.take(64);      // the shared key always goes through pinned staging
    for (std::vector<Operand>* v : {&ins, &outs})
        for (Operand& o : *v) {
            if (!o.per) continue;
            o.dev_off = r.dev.take((size_t)batch * o.per);
            if (!o.pinned) o.pin_off = r.pin.take((size_t)batch * o.per);
            per_inst += o.per;
        }
    cudaError_t e;
    if ((e = reserve_dev(r)) != cudaSuccess) return cuda_fail(e, "device staging allocation");
    if ((e = r.pin.reserve(r.pin.used)) != cudaSuccess) return cuda_fail(e, "pinned staging allocation");
    auto copy_in = [&](void* dst, const void* src, size_t n) {
        return peer.home >= 0 ? cudaMemcpyPeerAsync(dst, device, src, peer.home, n, r.h2d)
                              : cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, r.h2d);
    };
    auto copy_out = [&](void* dst, const void* src, size_t n) {
        return peer.home >= 0 ? cudaMemcpyPeerAsync(dst, peer.home, src, device, n, r.d2h)
                              : cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, r.d2h);
    };
    if (peer.after) CK(cudaStreamWaitEvent(r.h2d, peer.after, 0));
    if (key && peer.home >= 0) CK(copy_in(r.dev.base + key_dev, key, keylen));
    else if (key) {
        memcpy(r.pin.base + key_pin, key, keylen);
        CK(copy_in(r.dev.base + key_dev, r.pin.base + key_pin, keylen));
    }
    const gcb_plan_info& in = plan.info;
    const uint32_t wave = (in.teams_per_sm ? in.teams_per_sm : 1) * (uint32_t)di->sm_count;
    const uint32_t slice = slice_instances(per_inst, batch, wave);
    for (uint32_t b0 = 0; b0 < batch; b0 += slice) {
        const uint32_t nb = batch - b0 < slice ? batch - b0 : slice;
        for (const Operand& o : ins) {
            if (!o.per) continue;
            const size_t off = (size_t)b0 * o.per, n = (size_t)nb * o.per;
            const uint8_t* src = o.host + off;
            if (!o.pinned) { host_copy(r.pin.base + o.pin_off + off, src, n); src = r.pin.base + o.pin_off + off; }
            CK(copy_in(r.dev.base + o.dev_off + off, src, n));
        }
        cudaEvent_t ev_in, ev_k;
        CK(r.event(&ev_in)); CK(r.event(&ev_k));
        CK(cudaEventRecord(ev_in, r.h2d));
        CK(cudaStreamWaitEvent(r.k, ev_in, 0));
        // operand order (see gcb_garble_begin / gcb_eval_begin): garble ins = {keys, r, l0}, outs = {tables, io, full};
        // eval ins = {keys, tables, in_labels}, outs = {out_labels, full}
        const uint8_t* dk = key ? r.dev.base + key_dev : ins[0].dev(r, b0);
        if (garble)
            rc = launch_gc(true, plan, di, device, dk, keylen, key_stride, nb, (const gcb_label*)ins[1].dev(r, b0),
                           ins[2].per ? (const gcb_label*)ins[2].dev(r, b0) : nullptr,
                           outs[0].per ? (gcb_label*)outs[0].dev(r, b0) : nullptr, outs[1].per ? outs[1].dev(r, b0) : nullptr,
                           outs[2].per ? outs[2].dev(r, b0) : nullptr, r.k);
        else
            rc = launch_gc(false, plan, di, device, dk, keylen, key_stride, nb, nullptr,
                           ins[2].per ? (const gcb_label*)ins[2].dev(r, b0) : nullptr,
                           ins[1].per ? (gcb_label*)ins[1].dev(r, b0) : nullptr, outs[0].per ? outs[0].dev(r, b0) : nullptr,
                           outs[1].per ? outs[1].dev(r, b0) : nullptr, r.k);
        if (rc) return rc;
        CK(cudaEventRecord(ev_k, r.k));
        CK(cudaStreamWaitEvent(r.d2h, ev_k, 0));
        for (const Operand& o : outs) {
            if (!o.per) continue;
            const size_t off = (size_t)b0 * o.per, n = (size_t)nb * o.per;
            uint8_t* dst = o.pinned ? o.host + off : r.pin.base + o.pin_off + off;
            CK(copy_out(dst, r.dev.base + o.dev_off + off, n));
            if (!o.pinned) {
                cudaEvent_t ready;
                CK(r.event(&ready));
                CK(cudaEventRecord(ready, r.d2h));
                part->late.push_back(LateCopy{o.host + off, dst, n, ready});
            }
        }
    }
    if (peer.done) {
        CK(r.event(peer.done));
        CK(cudaEventRecord(*peer.done, r.d2h));
    }
    CK(r.seal());
    return GCB_OK;
}

// Waits for every part, finishes the copies into pageable memory, returns the resources.  The first error wins,
// but every part is drained: afterwards nothing refers to the caller's buffers.
static int finish_job(gcb_job* job) {
    int rc = GCB_OK;
    for (JobPart& part : job->parts) {
        if (!part.res) continue;
        JobRes& r = *part.res;
        cudaSetDevice(r.device);
        for (const LateCopy& c : part.late) {
            cudaError_t e = cudaEventSynchronize(c.ready);
            if (e != cudaSuccess) { if (!rc) rc = cuda_fail(e, "result copy"); break; }
            host_copy(c.dst, c.src, c.bytes);
        }
        cudaError_t e = r.quiesce();
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "job completion");
        g_res_pool.give_back(std::move(part.res));
    }
    job->parts.clear();
    return rc;
}

// Per-device share of a host-pointer call above which the blocking entry points go through the device in
// several rounds (the staging arena of one round stays below this).
constexpr size_t kMaxPartBytes = (size_t)24 << 30;

static int garble_begin_impl(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                             const gcb_label* r, const gcb_label* in_l0, gcb_label* tables, gcb_wire* io_wires,
                             gcb_wire* wires_full, gcb_job* job) {
    const Plan* use;
    std::vector<int> devs = call_devices();
    int rc = plan_for(plan, wires_full != nullptr, &use, batch / (devs.empty() ? 1 : devs.size()));
    if (rc) return rc;
    const gcb_plan_info& in = use->info;
    const size_t nin = in.num_inputs, nout = in.num_outputs, rows = in.num_rows, nw = in.num_wires;
    const uint32_t nd = (uint32_t)std::min<size_t>(devs.size(), batch);
    job->parts.resize(nd);
    for (uint32_t d = 0; d < nd; d++) {
        uint64_t lo, hi;
        share_range(batch, d, nd, &lo, &hi);
        std::vector<Operand> ins = {operand(key_stride ? keys + lo * key_stride : nullptr, key_stride),
                                    operand(r + lo, 16), operand(nin ? in_l0 + lo * nin : nullptr, nin * 16)};
        std::vector<Operand> outs = {operand(rows ? tables + lo * rows : nullptr, rows * 16),
                                     operand(io_wires ? io_wires + lo * (nin + nout) : nullptr, (nin + nout) * 32),
                                     operand(wires_full ? wires_full + lo * nw : nullptr, nw * 32)};
        if ((rc = begin_part(true, *use, devs[d], key_stride ? nullptr : keys, keylen, key_stride, (uint32_t)(hi - lo), ins, outs,
                             &job->parts[d])))
            return rc;
    }
    return GCB_OK;
}

static int eval_begin_impl(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                           const gcb_label* tables, const gcb_label* in_labels, gcb_label* out_labels, gcb_label* wires_full,
                           gcb_job* job) {
    const Plan* use;
    std::vector<int> devs = call_devices();
    int rc = plan_for(plan, wires_full != nullptr, &use, batch / (devs.empty() ? 1 : devs.size()));
    if (rc) return rc;
    const gcb_plan_info& in = use->info;
    const size_t nin = in.num_inputs, nout = in.num_outputs, rows = in.num_rows, nw = in.num_wires;
    const uint32_t nd = (uint32_t)std::min<size_t>(devs.size(), batch);
    job->parts.resize(nd);
    for (uint32_t d = 0; d < nd; d++) {
        uint64_t lo, hi;
        share_range(batch, d, nd, &lo, &hi);
        std::vector<Operand> ins = {operand(key_stride ? keys + lo * key_stride : nullptr, key_stride),
                                    operand(rows ? tables + lo * rows : nullptr, rows * 16),
                                    operand(nin ? in_labels + lo * nin : nullptr, nin * 16)};
        std::vector<Operand> outs = {operand(nout ? out_labels + lo * nout : nullptr, nout * 16),
                                     operand(wires_full ? wires_full + lo * nw : nullptr, nw * 16)};
        if ((rc = begin_part(false, *use, devs[d], key_stride ? nullptr : keys, keylen, key_stride, (uint32_t)(hi - lo), ins, outs,
                             &job->parts[d])))
            return rc;
    }
    return GCB_OK;
}

// ---- fan-out of a device-resident call (GCB_FLAG_FANOUT) ------------------------------------
// The operands live on the calling thread's device (`home`).  Every other device of the gcb_set_devices list
// takes a contiguous block of instances: its inputs are scattered to it and its results gathered back with peer
// copies over NVLink, slice by slice behind the kernels, all ordered against the caller's stream by events -- the
// call only enqueues.  This is the "scatter input labels, gather garbled tables" exchange of the batch split
// (SURVEY.md section 8e); there is no collective because no rank needs another rank's data.
static std::mutex g_peer_mu;
static std::map<std::pair<int, int>, bool> g_peer_on;
static void enable_peer(int a, int b) {                      // best effort: without it the copies are staged by the driver
    std::lock_guard<std::mutex> lk(g_peer_mu);
    for (const auto& pr : {std::make_pair(a, b), std::make_pair(b, a)}) {
        if (g_peer_on.count(pr)) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, pr.first, pr.second);
        if (can) {
            cudaSetDevice(pr.first);
            const cudaError_t e = cudaDeviceEnablePeerAccess(pr.second, 0);
            if (e != cudaSuccess) cudaGetLastError();            // already enabled, or refused
        }
        g_peer_on[pr] = can != 0;
    }
}
struct FanoutCleanup { std::vector<std::unique_ptr<JobRes>> res; };
static void CUDART_CB fanout_done(void* p) {
    std::unique_ptr<FanoutCleanup> c(static_cast<FanoutCleanup*>(p));
    for (auto& r : c->res) g_res_pool.retire(std::move(r));
}
// run_home(lo, hi): queue the home device's own block on the caller's stream.
template <class RunHome>
static int fanout_dev(bool garble, const Plan& plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                      const std::vector<Operand>& ins0, const std::vector<Operand>& outs0, cudaStream_t stream, RunHome run_home) {
    const int home = tl_cur;
    std::vector<int> devs;
    { std::lock_guard<std::mutex> lk(g_devlist_mu); if (g_devlist) devs = *g_devlist; }
    devs.erase(std::remove(devs.begin(), devs.end(), home), devs.end());
    devs.insert(devs.begin(), home);
    const uint32_t nd = (uint32_t)std::min<size_t>(devs.size(), batch);
    if (nd <= 1) return run_home(0, batch);
    DeviceInfo* di;
    int rc = use_device(home, &di);
    if (rc) return rc;
    auto cleanup = std::make_unique<FanoutCleanup>();
    std::unique_ptr<JobRes> home_res;                          // only for its events
    if ((rc = lease_res(home, &home_res))) return rc;
    home_res->ev_used = 0;
    cudaEvent_t ready;
    CK(home_res->event(&ready));
    CK(cudaEventRecord(ready, stream));
    std::vector<cudaEvent_t> done(nd, nullptr);
    for (uint32_t d = 1; d < nd && !rc; d++) {
        uint64_t lo, hi;
        share_range(batch, d, nd, &lo, &hi);
        enable_peer(home, devs[d]);
        std::vector<Operand> ins = ins0, outs = outs0;
        for (std::vector<Operand>* v : {&ins, &outs})
            for (Operand& o : *v) if (o.per) o.host += lo * o.per;
        JobPart part;
        PeerMode pm;
        pm.home = home; pm.after = ready; pm.done = &done[d];
        rc = begin_part(garble, plan, devs[d], key_stride ? nullptr : keys, keylen, key_stride, (uint32_t)(hi - lo), ins, outs, &part, pm);
        if (part.res) cleanup->res.push_back(std::move(part.res));
    }
    if (!rc) rc = use_device(home, &di);
    uint64_t lo, hi;
    share_range(batch, 0, nd, &lo, &hi);
    if (!rc) rc = run_home((uint32_t)lo, (uint32_t)hi);
    cudaSetDevice(home);
    for (uint32_t d = 1; d < nd; d++)
        if (done[d]) cudaStreamWaitEvent(stream, done[d], 0);
    cleanup->res.push_back(std::move(home_res));
    if (rc) {                                                  // drain what was queued, then give the parts back
        for (auto& r : cleanup->res) { cudaSetDevice(r->device); r->quiesce(); g_res_pool.give_back(std::move(r)); }
        cudaSetDevice(home);
        return rc;
    }
    CK(cudaLaunchHostFunc(stream, fanout_done, cleanup.release()));
    return GCB_OK;
}

// ---- simple staged calls (IKNP, MiTCCRH, COT / ROT, hashes, wire format) ---------------------
// One part per device on the part's kernel stream: inputs up, the _dev entry point, results down.  The parts
// of all devices are queued before any is waited for.
struct DeviceScope {                                   // the _dev entry points run on the calling thread's device
    int saved;
    explicit DeviceScope(int d) : saved(tl_device) { tl_device = d; }
    ~DeviceScope() { tl_device = saved; }
};
struct HostOp {
    const void* src;          // copied to the device before the kernel (null: none)
    void* dst;                // copied back after it (null: none)
    size_t bytes;
    void* dptr = nullptr;
    size_t dev_off = 0, pin_off = 0;
    bool pinned = false;
    const void* host() const { return src ? src : dst; }
};
static int stage_begin(int device, std::vector<HostOp*> ops, JobPart* part) {
    int rc = use_device(device, nullptr);
    if (rc) return rc;
    size_t want = 0;
    for (HostOp* o : ops) want += ((o->bytes ? o->bytes : 16) + 255) & ~(size_t)255;
    if ((rc = lease_res(device, &part->res, 0, want))) return rc;
    JobRes& r = *part->res;
    r.dev.used = r.pin.used = 0; r.ev_used = 0; r.sealed = false;
    for (HostOp* o : ops) {
        o->pinned = o->host() && is_pinned(o->host());
        o->dev_off = r.dev.take(o->bytes ? o->bytes : 16);
        if (!o->pinned) o->pin_off = r.pin.take(o->bytes ? o->bytes : 16);
    }
    cudaError_t e;
    if ((e = reserve_dev(r)) != cudaSuccess) return cuda_fail(e, "device staging allocation");
    if ((e = r.pin.reserve(r.pin.used ? r.pin.used : 16)) != cudaSuccess) return cuda_fail(e, "pinned staging allocation");
    for (HostOp* o : ops) {
        o->dptr = r.dev.base + o->dev_off;
        if (!o->src || !o->bytes) continue;
        const void* h = o->src;
        if (!o->pinned) { host_copy(r.pin.base + o->pin_off, o->src, o->bytes); h = r.pin.base + o->pin_off; }
        CK(cudaMemcpyAsync(o->dptr, h, o->bytes, cudaMemcpyHostToDevice, r.k));
    }
    return GCB_OK;
}
static int stage_end(std::vector<HostOp*> ops, JobPart* part) {
    JobRes& r = *part->res;
    for (HostOp* o : ops) {
        if (!o->dst || !o->bytes) continue;
        void* h = o->pinned ? o->dst : (void*)(r.pin.base + o->pin_off);
        CK(cudaMemcpyAsync(h, o->dptr, o->bytes, cudaMemcpyDeviceToHost, r.k));
        if (!o->pinned) {
            cudaEvent_t ready;
            CK(r.event(&ready));
            CK(cudaEventRecord(ready, r.k));
            part->late.push_back(LateCopy{o->dst, static_cast<const uint8_t*>(h), o->bytes, ready});
        }
    }
    CK(r.seal());
    return GCB_OK;
}
// Runs body(device, lo, hi, part) for every device's block of `units`, then waits for all parts.
template <class Body>
static int fan_out(uint64_t units, Body body) {
    const std::vector<int> devs = call_devices();
    const uint32_t nd = (uint32_t)std::min<uint64_t>(devs.size(), units ? units : 1);
    gcb_job job;
    job.parts.resize(nd);
    int rc = GCB_OK;
    for (uint32_t d = 0; d < nd && !rc; d++) {
        uint64_t lo, hi;
        share_range(units, d, nd, &lo, &hi);
        DeviceScope scope(devs[d]);
        rc = body(devs[d], lo, hi, &job.parts[d]);
    }
    const std::string msg = tl_err;
    const int rc2 = finish_job(&job);
    if (rc) { tl_err = msg; return rc; }
    return rc2;
}

// Instances one round of a blocking call may take so that no device's staging exceeds kMaxPartBytes.
static uint32_t round_instances(size_t per_inst, uint32_t batch) {
    const size_t nd = call_devices().size();
    const size_t n = kMaxPartBytes / (per_inst ? per_inst : 1) * (nd ? nd : 1);
    return n >= batch ? batch : (uint32_t)(n ? n : 1);
}


}  // namespace gcb

using namespace gcb;

// ======================================================================= C ABI ==
extern "C" {

const char* gcb_last_error(void) { return tl_err.c_str(); }
const char* gcb_version(void) { return "gcb200 0.2 (sm_100a)"; }
uint64_t gcb_launch_count(void) { return g_launches.load(); }

int gcb_set_device(int device) {
    GCB_TRY
    tl_device = device < 0 ? -1 : device;                   // -1: back to the process-wide choice
    return GCB_OK;
    GCB_CATCH
}
int gcb_set_devices(const int* ids, int n) {
    GCB_TRY
    if (n < 0 || (n > 0 && !ids)) return fail(GCB_E_ARG, "bad device list");
    int count = 0;
    if (n > 0 && (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)) {
        cudaGetLastError();
        return fail(GCB_E_CUDA, "no CUDA device available; this library has no CPU path");
    }
    auto l = std::make_shared<std::vector<int>>();
    for (int i = 0; i < n; i++) {
        if (ids[i] < 0 || ids[i] >= count) return fail(GCB_E_ARG, "device %d out of range (%d devices)", ids[i], count);
        for (int v : *l) if (v == ids[i]) return fail(GCB_E_ARG, "device %d listed twice", ids[i]);
        l->push_back(ids[i]);
    }
    std::lock_guard<std::mutex> lk(g_devlist_mu);
    g_devlist = n ? l : nullptr;
    return GCB_OK;
    GCB_CATCH
}
int gcb_get_devices(int* ids, int cap) {
    GCB_TRY
    const std::vector<int> d = call_devices();
    for (int i = 0; i < cap && i < (int)d.size(); i++) if (ids) ids[i] = d[i];
    return (int)d.size();
    GCB_CATCH
}
int gcb_device_count(void) {
    GCB_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
    GCB_CATCH
}

// Page-locked host memory for the slabs the Go side pools per circuit
// (garbleScratchPool, circuit/garble.go:193-224): buffers from here are DMA'd
// in place by the host entry points instead of being staged.
void* gcb_host_alloc(size_t bytes) {
    if (select_device(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        fail(GCB_E_CUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void gcb_host_free(void* p) { if (p) cudaFreeHost(p); }

void* gcb_dev_alloc(size_t bytes) {
    if (select_device(nullptr)) return nullptr;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
        fail(GCB_E_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void gcb_dev_free(void* p) { if (p) cudaFree(p); }
int gcb_dev_upload(void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    GCB_TRY
    if (bytes && (!dst_dev || !src_host)) return fail(GCB_E_ARG, "null argument");
    int rc = select_device(nullptr);
    if (rc) return rc;
    if (bytes) CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return GCB_OK;
    GCB_CATCH
}
int gcb_dev_download(void* dst_host, const void* src_dev, size_t bytes, void* stream) {
    GCB_TRY
    if (bytes && (!dst_host || !src_dev)) return fail(GCB_E_ARG, "null argument");
    int rc = select_device(nullptr);
    if (rc) return rc;
    if (bytes) CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return GCB_OK;
    GCB_CATCH
}
int gcb_dev_stream_create(void** stream) {
    GCB_TRY
    if (!stream) return fail(GCB_E_ARG, "null argument");
    int rc = select_device(nullptr);
    if (rc) return rc;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = st;
    return GCB_OK;
    GCB_CATCH
}
void gcb_dev_stream_destroy(void* stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }
int gcb_dev_sync(void* stream) {
    GCB_TRY
    int rc = select_device(nullptr);
    if (rc) return rc;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return GCB_OK;
    GCB_CATCH
}

// ------------------------------------------------------------------ plans -------
int gcb_plan_create(const gcb_gate* gates, uint32_t num_gates, uint32_t num_wires, uint32_t num_inputs,
                    uint32_t num_outputs, gcb_plan** out) {
    GCB_TRY
    if (!out) return fail(GCB_E_ARG, "null plan output pointer");
    *out = nullptr;
    if (!gates && num_gates) return fail(GCB_E_ARG, "null gate array");
    if (num_inputs > num_wires || num_outputs > num_wires) return fail(GCB_E_ARG, "more I/O wires than wires");
    // every wire is an input or is assigned by a gate (circuit/parser.go:199 rejects files with unassigned wires): a wire
    // count beyond that is refused before anything is sized by it (0xffffffff used to cost 32 GB and a minute)
    if ((uint64_t)num_wires > (uint64_t)num_inputs + num_gates)
        return fail(GCB_E_WIRE, "corrupted circuit: wire %llu not assigned", (unsigned long long)num_inputs + num_gates);
    PlanSpec spec;
    spec.gates = gates; spec.num_gates = num_gates; spec.num_wires = num_wires;
    for (uint32_t i = 0; i < num_inputs; i++) spec.live_in.push_back(i);
    for (uint32_t i = 0; i < num_outputs; i++) spec.live_out.push_back(num_wires - num_outputs + i);
    auto pl = std::make_unique<gcb_plan>();
    static std::atomic<uint64_t> next_uid{1};
    pl->uid = next_uid.fetch_add(1);
    std::string err;
    pl->p.gates.assign(gates, gates + num_gates);
    // GCB_HOT_TEAMS = N (experiments, tests): the plan itself is the hot / cold one; otherwise that plan is built on
    // demand for batches that overflow this one (plan_for)
    const char* forced = getenv("GCB_HOT_TEAMS");
    int rc = (forced && atoi(forced) > 0) ? build_best_plan(spec, pl->p, err) : build_plan(spec, pl->p, err);
    if (rc) return fail(rc, "%s", err.c_str());
    team_geometry(pl->p);
    if (pl->p.info.teams_per_sm == 0)
        return fail(GCB_E_TOO_LARGE, "circuit keeps %u wire labels live; at most %zu fit on chip",
                    pl->p.info.num_slots, (kSmemOptin - table_pad(kAssumedSmemBase) - (size_t)aes_table_bytes(2)) / 16);
    pl->p.gates.assign(gates, gates + num_gates);
    *out = pl.release();
    return GCB_OK;
    GCB_CATCH
}
void gcb_plan_destroy(gcb_plan* plan) { delete plan; }
int gcb_plan_get_info(const gcb_plan* plan, gcb_plan_info* info) {
    GCB_TRY
    if (!plan || !info) return fail(GCB_E_ARG, "null argument");
    *info = plan->p.info;
    return GCB_OK;
    GCB_CATCH
}
int gcb_plan_get_info_for_batch(const gcb_plan* plan, uint64_t batch, gcb_plan_info* info) {
    GCB_TRY
    if (!plan || !info) return fail(GCB_E_ARG, "null argument");
    const Plan* use;
    int rc = plan_for(plan, false, &use, batch);
    if (rc) return rc;
    *info = use->info;
    return GCB_OK;
    GCB_CATCH
}
int gcb_plan_row_offsets(const gcb_plan* plan, uint32_t* row_off) {
    GCB_TRY
    if (!plan || !row_off) return fail(GCB_E_ARG, "null argument");
    memcpy(row_off, plan->p.row_off.data(), plan->p.row_off.size() * sizeof(uint32_t));
    return GCB_OK;
    GCB_CATCH
}

// ----------------------------------------------------------- garble / eval ------
int gcb_garble_dev(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride,
                   uint32_t batch, const gcb_label* r, const gcb_label* in_l0, gcb_label* tables,
                   gcb_wire* io_wires, gcb_wire* wires_full, uint32_t flags, void* stream) {
    GCB_TRY
    if (!plan || !keys || !r || (!in_l0 && plan->p.info.num_inputs) || (!tables && plan->p.info.num_rows))
        return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    if (batch == 0) return GCB_OK;
    DeviceInfo* di;
    if ((rc = select_device(&di))) return rc;
    const Plan* use;
    if ((rc = plan_for(plan, wires_full != nullptr, &use, batch))) return rc;
    if (flags & GCB_FLAG_FANOUT) {
        if (wires_full) return fail(GCB_E_ARG, "GCB_FLAG_FANOUT does not produce wires_full");
        const size_t nin = use->info.num_inputs, nout = use->info.num_outputs, rows = use->info.num_rows;
        const std::vector<Operand> ins = {operand(key_stride ? keys : nullptr, key_stride), operand(r, 16), operand(nin ? in_l0 : nullptr, nin * 16)};
        const std::vector<Operand> outs = {operand(rows ? tables : nullptr, rows * 16), operand(io_wires, (nin + nout) * 32), operand(nullptr, 0)};
        return fanout_dev(true, *use, keys, keylen, key_stride, batch, ins, outs, (cudaStream_t)stream, [&](uint32_t lo, uint32_t hi) -> int {
            return launch_gc(true, *use, di, tl_cur, key_stride ? keys + (size_t)lo * key_stride : keys, keylen, key_stride, hi - lo, r + lo,
                             in_l0 ? in_l0 + (size_t)lo * nin : nullptr, tables ? tables + (size_t)lo * rows : nullptr,
                             io_wires ? io_wires + (size_t)lo * (nin + nout) : nullptr, nullptr, (cudaStream_t)stream);
        });
    }
    return launch_gc(true, *use, di, tl_cur, keys, keylen, key_stride, batch, r, in_l0, tables, io_wires,
                     wires_full, (cudaStream_t)stream);
    GCB_CATCH
}

int gcb_eval_dev(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                 const gcb_label* tables, const gcb_label* in_labels, gcb_label* out_labels,
                 gcb_label* wires_full, uint32_t flags, void* stream) {
    GCB_TRY
    if (!plan || !keys || (!in_labels && plan->p.info.num_inputs) || (!tables && plan->p.info.num_rows) ||
        (!out_labels && plan->p.info.num_outputs))
        return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    if (batch == 0) return GCB_OK;
    DeviceInfo* di;
    if ((rc = select_device(&di))) return rc;
    const Plan* use;
    if ((rc = plan_for(plan, wires_full != nullptr, &use, batch))) return rc;
    if (flags & GCB_FLAG_FANOUT) {
        if (wires_full) return fail(GCB_E_ARG, "GCB_FLAG_FANOUT does not produce wires_full");
        const size_t nin = use->info.num_inputs, nout = use->info.num_outputs, rows = use->info.num_rows;
        const std::vector<Operand> ins = {operand(key_stride ? keys : nullptr, key_stride), operand(rows ? tables : nullptr, rows * 16),
                                          operand(nin ? in_labels : nullptr, nin * 16)};
        const std::vector<Operand> outs = {operand(nout ? out_labels : nullptr, nout * 16), operand(nullptr, 0)};
        return fanout_dev(false, *use, keys, keylen, key_stride, batch, ins, outs, (cudaStream_t)stream, [&](uint32_t lo, uint32_t hi) -> int {
            return launch_gc(false, *use, di, tl_cur, key_stride ? keys + (size_t)lo * key_stride : keys, keylen, key_stride, hi - lo, nullptr,
                             in_labels ? in_labels + (size_t)lo * nin : nullptr, const_cast<gcb_label*>(tables ? tables + (size_t)lo * rows : nullptr),
                             out_labels ? out_labels + (size_t)lo * nout : nullptr, nullptr, (cudaStream_t)stream);
        });
    }
    return launch_gc(false, *use, di, tl_cur, keys, keylen, key_stride, batch, nullptr, in_labels,
                     const_cast<gcb_label*>(tables), out_labels, wires_full, (cudaStream_t)stream);
    GCB_CATCH
}

// Host-buffer variants.  gcb_garble / gcb_eval = _begin + gcb_job_wait; a batch whose staging would not fit one
// arena per device goes through in rounds.
static int check_garble_args(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride,
                             const gcb_label* r, const gcb_label* in_l0, const gcb_label* tables) {
    if (!plan || !keys || !r || (!in_l0 && plan->p.info.num_inputs) || (!tables && plan->p.info.num_rows))
        return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    return GCB_OK;
}
static int check_eval_args(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride,
                           const gcb_label* tables, const gcb_label* in_labels, const gcb_label* out_labels) {
    if (!plan || !keys || (!in_labels && plan->p.info.num_inputs) || (!tables && plan->p.info.num_rows) ||
        (!out_labels && plan->p.info.num_outputs))
        return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    return GCB_OK;
}

int gcb_garble_begin(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                     const gcb_label* r, const gcb_label* in_l0, gcb_label* tables, gcb_wire* io_wires,
                     gcb_wire* wires_full, uint32_t flags, gcb_job** job) {
    GCB_TRY
    (void)flags;
    if (!job) return fail(GCB_E_ARG, "null job output pointer");
    *job = nullptr;
    int rc = check_garble_args(plan, keys, keylen, key_stride, r, in_l0, tables);
    if (rc) return rc;
    auto j = std::make_unique<gcb_job>();
    if (batch && (rc = garble_begin_impl(plan, keys, keylen, key_stride, batch, r, in_l0, tables, io_wires, wires_full, j.get()))) {
        const std::string msg = tl_err;                    // keep the first error over whatever the drain reports
        finish_job(j.get());
        tl_err = msg;
        return rc;
    }
    *job = j.release();
    return GCB_OK;
    GCB_CATCH
}

int gcb_eval_begin(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                   const gcb_label* tables, const gcb_label* in_labels, gcb_label* out_labels, gcb_label* wires_full,
                   uint32_t flags, gcb_job** job) {
    GCB_TRY
    (void)flags;
    if (!job) return fail(GCB_E_ARG, "null job output pointer");
    *job = nullptr;
    int rc = check_eval_args(plan, keys, keylen, key_stride, tables, in_labels, out_labels);
    if (rc) return rc;
    auto j = std::make_unique<gcb_job>();
    if (batch && (rc = eval_begin_impl(plan, keys, keylen, key_stride, batch, tables, in_labels, out_labels, wires_full, j.get()))) {
        const std::string msg = tl_err;
        finish_job(j.get());
        tl_err = msg;
        return rc;
    }
    *job = j.release();
    return GCB_OK;
    GCB_CATCH
}

int gcb_job_wait(gcb_job* job) {
    GCB_TRY
    if (!job) return fail(GCB_E_ARG, "null job");
    std::unique_ptr<gcb_job> j(job);
    return finish_job(j.get());
    GCB_CATCH
}

int gcb_job_done(gcb_job* job) {
    GCB_TRY
    if (!job) return 1;
    for (JobPart& part : job->parts) {
        if (!part.res) continue;
        cudaSetDevice(part.res->device);
        if (!part.res->sealed) continue;
        for (cudaEvent_t t : part.res->tail) {
            const cudaError_t e = cudaEventQuery(t);
            if (e == cudaErrorNotReady) return 0;
            if (e != cudaSuccess) return 1;              // gcb_job_wait reports it
        }
    }
    return 1;
    GCB_CATCH
}

int gcb_garble(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
               const gcb_label* r, const gcb_label* in_l0, gcb_label* tables, gcb_wire* io_wires,
               gcb_wire* wires_full, uint32_t flags) {
    GCB_TRY
    int rc = check_garble_args(plan, keys, keylen, key_stride, r, in_l0, tables);
    if (rc || batch == 0) return rc;
    const gcb_plan_info& in = plan->p.info;
    const size_t nin = in.num_inputs, nout = in.num_outputs, rows = in.num_rows, nw = in.num_wires;
    const size_t per_inst = 16 + nin * 16 + rows * 16 + (io_wires ? (nin + nout) * 32 : 0) + (wires_full ? nw * 32 : 0) + key_stride;
    const uint32_t round = round_instances(per_inst, batch);
    for (uint32_t b0 = 0; b0 < batch; b0 += round) {
        const uint32_t nb = batch - b0 < round ? batch - b0 : round;
        gcb_job* job = nullptr;
        rc = gcb_garble_begin(plan, key_stride ? keys + (size_t)b0 * key_stride : keys, keylen, key_stride, nb, r + b0,
                              in_l0 ? in_l0 + (size_t)b0 * nin : nullptr, tables ? tables + (size_t)b0 * rows : nullptr,
                              io_wires ? io_wires + (size_t)b0 * (nin + nout) : nullptr,
                              wires_full ? wires_full + (size_t)b0 * nw : nullptr, flags, &job);
        if (rc) return rc;
        if ((rc = gcb_job_wait(job))) return rc;
    }
    return GCB_OK;
    GCB_CATCH
}

int gcb_eval(const gcb_plan* plan, const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
             const gcb_label* tables, const gcb_label* in_labels, gcb_label* out_labels, gcb_label* wires_full,
             uint32_t flags) {
    GCB_TRY
    int rc = check_eval_args(plan, keys, keylen, key_stride, tables, in_labels, out_labels);
    if (rc || batch == 0) return rc;
    const gcb_plan_info& in = plan->p.info;
    const size_t nin = in.num_inputs, nout = in.num_outputs, rows = in.num_rows, nw = in.num_wires;
    const size_t per_inst = nin * 16 + rows * 16 + nout * 16 + (wires_full ? nw * 16 : 0) + key_stride;
    const uint32_t round = round_instances(per_inst, batch);
    for (uint32_t b0 = 0; b0 < batch; b0 += round) {
        const uint32_t nb = batch - b0 < round ? batch - b0 : round;
        gcb_job* job = nullptr;
        rc = gcb_eval_begin(plan, key_stride ? keys + (size_t)b0 * key_stride : keys, keylen, key_stride, nb,
                            tables ? tables + (size_t)b0 * rows : nullptr, in_labels ? in_labels + (size_t)b0 * nin : nullptr,
                            out_labels ? out_labels + (size_t)b0 * nout : nullptr,
                            wires_full ? wires_full + (size_t)b0 * nw : nullptr, flags, &job);
        if (rc) return rc;
        if ((rc = gcb_job_wait(job))) return rc;
    }
    return GCB_OK;
    GCB_CATCH
}

int gcb_select_labels_dev(const gcb_wire* wires, size_t wire_stride, const uint8_t* bits, gcb_label* out,
                          uint32_t batch, uint32_t n, void* stream) {
    GCB_TRY
    if (!wires || !bits || !out) return fail(GCB_E_ARG, "null argument");
    if (!batch || !n) return GCB_OK;
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    select_labels_kernel<<<di->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(wires), wire_stride, bits, reinterpret_cast<uint4*>(out), batch, n);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}
int gcb_decode_bits_dev(const gcb_wire* wires, size_t wire_stride, const gcb_label* labels, uint8_t* bits,
                        uint32_t batch, uint32_t n, void* stream) {
    GCB_TRY
    if (!wires || !bits || !labels) return fail(GCB_E_ARG, "null argument");
    if (!batch || !n) return GCB_OK;
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    decode_bits_kernel<<<di->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(wires), wire_stride, reinterpret_cast<const uint4*>(labels), bits, batch, n);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}

// ------------------------------------------------------------- gate hashes ------
// key: DEVICE pointer in the _dev variant.
int gcb_hash_half_dev(const uint8_t* key, uint32_t keylen, const gcb_label* x, uint32_t tweak0, gcb_label* out,
                      uint64_t n, void* stream) {
    GCB_TRY
    if (!key || (n && (!x || !out))) return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (n == 0) return GCB_OK;
    DeviceInfo* di;
    if ((rc = select_device(&di))) return rc;
    HashParams p{key, keylen, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), tweak0, n};
    const uint64_t want = (n + 1023) / 1024;
    const dim3 grid((unsigned)(want < (uint64_t)di->sm_count ? want : (uint64_t)di->sm_count));
    const size_t smem = table_pad(di->smem_base) + AES_TABLE_BYTES + 256;
    cudaStream_t s = (cudaStream_t)stream;
    if (keylen == 16) hash_half_kernel<10><<<grid, 1024, smem, s>>>(p);
    else if (keylen == 24) hash_half_kernel<12><<<grid, 1024, smem, s>>>(p);
    else hash_half_kernel<14><<<grid, 1024, smem, s>>>(p);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}
int gcb_hash_half(const uint8_t* key, uint32_t keylen, const gcb_label* x, uint32_t tweak0, gcb_label* out,
                  uint64_t n) {
    GCB_TRY
    if (!key || (n && (!x || !out))) return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (n == 0) return GCB_OK;
    return fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        HostOp k{key, nullptr, keylen}, xi{x + lo, nullptr, (size_t)(hi - lo) * 16}, o{nullptr, out + lo, (size_t)(hi - lo) * 16};
        int rc = stage_begin(dev, {&k, &xi, &o}, part);
        if (rc) return rc;
        if ((rc = gcb_hash_half_dev((const uint8_t*)k.dptr, keylen, (const gcb_label*)xi.dptr, tweak0 + (uint32_t)lo, (gcb_label*)o.dptr,
                                    hi - lo, part->res->k)))
            return rc;
        return stage_end({&o}, part);
    });
    GCB_CATCH
}

// --------------------------------------------------------------- streaming ------
int gcb_stream_create(const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch, const gcb_label* r,
                      const uint32_t* input_ids, uint32_t ninputs, const gcb_label* in_l0, gcb_stream** out) {
    GCB_TRY
    if (!out) return fail(GCB_E_ARG, "null stream output pointer");
    *out = nullptr;
    if (!keys || !r || (ninputs && (!input_ids || !in_l0)) || batch == 0) return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    DeviceInfo* di;
    if ((rc = select_device(&di))) return rc;
    auto s = std::make_unique<gcb_stream>();
    s->device = tl_cur; s->batch = batch; s->keylen = keylen; s->key_stride = key_stride;
    CK(cudaStreamCreateWithFlags(&s->cs, cudaStreamNonBlocking));
    const size_t kb = key_stride ? (size_t)key_stride * batch : keylen;
    CK(s->keys.alloc(kb));
    CK(cudaMemcpyAsync(s->keys.p, keys, kb, cudaMemcpyHostToDevice, s->cs));
    CK(s->r.alloc((size_t)batch * 16));
    CK(cudaMemcpyAsync(s->r.p, r, (size_t)batch * 16, cudaMemcpyHostToDevice, s->cs));
    force_s_kernel<<<(batch + 255) / 256, 256, 0, s->cs>>>(s->r.as<uint4>(), batch);
    CK(launched());
    if (ninputs) {
        uint32_t mx = 0;
        for (uint32_t i = 0; i < ninputs; i++) mx = input_ids[i] > mx ? input_ids[i] : mx;
        if ((rc = ensure_wires(s.get(), mx))) return rc;
        DevBuf dl0;
        CK(dl0.alloc((size_t)batch * ninputs * 16));
        if ((rc = grow(s->ids, s->ids_cap, (size_t)ninputs * 4))) return rc;
        CK(cudaMemcpyAsync(s->ids.p, input_ids, (size_t)ninputs * 4, cudaMemcpyHostToDevice, s->cs));
        CK(cudaMemcpyAsync(dl0.p, in_l0, (size_t)batch * ninputs * 16, cudaMemcpyHostToDevice, s->cs));
        wf_set_kernel<<<di->sm_count * 4, 256, 0, s->cs>>>(reinterpret_cast<uint4* const*>(s->page_table.p),
                                                           s->ids.as<uint32_t>(), ninputs, dl0.as<uint4>(), batch);
        CK(launched());
        CK(cudaStreamSynchronize(s->cs));
    }
    CK(cudaStreamSynchronize(s->cs));
    *out = s.release();
    return GCB_OK;
    GCB_CATCH
}
void gcb_stream_destroy(gcb_stream* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->ds) cudaStreamSynchronize(s->ds);
    if (s->cs) cudaStreamSynchronize(s->cs);
    delete s;
}

int gcb_stream_get_wires(gcb_stream* s, const uint32_t* ids, uint32_t n, gcb_wire* wires) {
    GCB_TRY
    if (!s || (n && (!ids || !wires))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceInfo* di;
    int rc = use_device(s->device, &di);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++)
        if (((size_t)ids[i] >> WF_PAGE_SHIFT) >= s->pages.size())
            return fail(GCB_E_WIRE, "wire %u not allocated", ids[i]);
    if ((rc = grow(s->ids, s->ids_cap, (size_t)n * 4))) return rc;
    if ((rc = grow(s->wires, s->wires_cap, (size_t)s->batch * n * 32))) return rc;
    CK(cudaMemcpyAsync(s->ids.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s->cs));
    wf_get_kernel<<<di->sm_count * 4, 256, 0, s->cs>>>(reinterpret_cast<uint4* const*>(s->page_table.p),
                                                       s->ids.as<uint32_t>(), n, s->r.as<uint4>(),
                                                       s->wires.as<uint4>(), s->batch);
    CK(launched());
    CK(cudaMemcpyAsync(wires, s->wires.p, (size_t)s->batch * n * 32, cudaMemcpyDeviceToHost, s->cs));
    CK(cudaStreamSynchronize(s->cs));
    return GCB_OK;
    GCB_CATCH
}

int gcb_stream_step_size(gcb_stream* s, const gcb_plan* plan, const uint32_t* in, uint32_t nin, const uint32_t* out,
                         uint32_t nout, size_t* bytes) {
    GCB_TRY
    if (!s || !plan || !bytes || (nin && !in) || (nout && !out)) return fail(GCB_E_ARG, "null argument");
    StreamLayout lay;
    std::string err;
    int rc = build_stream_layout(plan->p.gates, plan->p.info.num_wires, in, nin, out, nout, lay, err);
    if (rc) return fail(rc, "%s", err.c_str());
    *bytes = lay.tmpl.size();
    return GCB_OK;
    GCB_CATCH
}

// Streaming.Garble for one sub-circuit.  wait = false (gcb_stream_garble_begin): everything is queued -- the gate
// kernel and the serialiser on stream cs, the copy to `dst` on stream ds from one of two staging buffers -- and
// the call returns; the next step's kernel then runs while this step's bytes are still crossing PCIe.
static int stream_garble_impl(gcb_stream* s, const gcb_plan* plan, const uint32_t* in, uint32_t nin, const uint32_t* out,
                              uint32_t nout, uint8_t* dst, size_t dst_stride, size_t* written, uint64_t* ns_init,
                              uint64_t* ns_garble, bool wait) {
    using clk = std::chrono::steady_clock;
    const auto t_start = clk::now();
    if (!s || !plan || (nin && !in) || (nout && !out)) return fail(GCB_E_ARG, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceInfo* di;
    int rc = use_device(s->device, &di);
    if (rc) return rc;
    const Plan& base = plan->p;
    const uint32_t nw = base.info.num_wires;
    if (nin > nw || nout > nw) return fail(GCB_E_ARG, "more wire ids than wires");
    // initCircuit (stream_garble.go:102-114).  The byte layout of the record stream (every header,
    // the offset of every row) is host work proportional to the gate count; when the caller's buffer
    // is large enough for ANY id assignment (13 header bytes per gate at most) the gate kernel is
    // launched first and the layout is built while it runs.
    StreamLayout lay;
    std::string err;
    const size_t bound = 13 * base.gates.size() + 16 * (size_t)base.info.num_rows;
    const bool late_layout = dst && dst_stride >= bound;
    size_t total = 0;
    if (!late_layout) {
        if ((rc = build_stream_layout(base.gates, nw, in, nin, out, nout, lay, err))) return fail(rc, "%s", err.c_str());
        total = lay.tmpl.size();
        if (written) *written = total;
        if (total && (!dst || dst_stride < total)) return fail(GCB_E_BUFFER, "stream buffer too small: need %zu bytes per instance", total);
    }
    uint32_t mx = 0;
    for (uint32_t i = 0; i < nin; i++) mx = in[i] > mx ? in[i] : mx;
    for (uint32_t i = 0; i < nout; i++) mx = out[i] > mx ? out[i] : mx;
    if ((rc = ensure_wires(s, mx))) return rc;

    // The compiled plan can be used as is when the in / out id sets are disjoint,
    // the out ids are distinct, and no gate assigns an input wire; otherwise
    // build a plan in which aliased wires share one location, so that the result
    // equals the reference's sequential reads and writes of the wire file.
    bool alias = nin != base.info.num_inputs || nout != base.info.num_outputs;
    std::unordered_map<uint32_t, uint32_t> loc_of_id;
    if (!alias) {
        for (uint32_t i = 0; i < nin; i++) loc_of_id.emplace(in[i], 0);
        for (uint32_t i = 0; i < nout && !alias; i++) alias = !loc_of_id.emplace(out[i], 1).second;
        const uint32_t first_out = nw - nout;
        for (size_t g = 0; g < base.gates.size() && !alias; g++) {
            const gcb_gate& gt = base.gates[g];
            alias = gt.out < nin || (first_out < nin);
        }
    }
    std::shared_ptr<gcb_stream::AliasPlan> aliased;
    std::vector<uint32_t> in_eff(in, in + nin), out_eff(out, out + nout);
    const Plan* use = &base;
    if (alias) {
        loc_of_id.clear();
        std::vector<uint32_t> ids;
        auto loc_for = [&](uint32_t id) {
            auto it = loc_of_id.find(id);
            if (it != loc_of_id.end()) return it->second;
            const uint32_t l = (uint32_t)ids.size();
            loc_of_id.emplace(id, l);
            ids.push_back(id);
            return l;
        };
        PlanSpec spec;
        spec.gates = base.gates.data(); spec.num_gates = (uint32_t)base.gates.size(); spec.num_wires = nw;
        spec.loc.resize(nw);
        const uint32_t first_out = nw - nout;
        for (uint32_t w = 0; w < nw; w++) {
            if (w < nin) spec.loc[w] = loc_for(in[w]);
            else if (w >= first_out) spec.loc[w] = loc_for(out[w - first_out]);
        }
        const uint32_t np = (uint32_t)ids.size();
        for (uint32_t w = nin; w < first_out && w < nw; w++) spec.loc[w] = np + w;
        spec.num_locs = np + nw;
        for (uint32_t l = 0; l < np; l++) spec.live_in.push_back(l);
        std::vector<uint8_t> written_loc(np, 0);
        for (const gcb_gate& gt : base.gates) if (spec.loc[gt.out] < np) written_loc[spec.loc[gt.out]] = 1;
        in_eff = ids;
        out_eff.clear();
        for (uint32_t l = 0; l < np; l++) if (written_loc[l]) { spec.live_out.push_back(l); out_eff.push_back(ids[l]); }
        // The plan depends on the aliasing PATTERN (spec.loc numbers locations by first appearance), not on the ids:
        // the steps of a program that reuse a sub-circuit with the same pattern share one compiled plan, which also
        // stays alive in the cache while its kernel runs (no implicit device synchronisation on return).
        for (const auto& c : s->alias_plans)
            if (c->plan_uid == plan->uid && c->loc == spec.loc) { aliased = c; break; }
        if (!aliased) {
            auto ap = std::make_shared<gcb_stream::AliasPlan>();
            ap->plan_uid = plan->uid;
            ap->plan = std::make_unique<Plan>();
            if ((rc = build_best_plan(spec, *ap->plan, err, NODE_MAX_FANIN, s->batch))) return fail(rc, "%s", err.c_str());
            team_geometry(*ap->plan);
            if (ap->plan->info.teams_per_sm == 0) return fail(GCB_E_TOO_LARGE, "sub-circuit keeps %u wire labels live", ap->plan->info.num_slots);
            ap->loc = spec.loc;
            if (s->alias_plans.size() >= 16) s->alias_plans.erase(s->alias_plans.begin());
            s->alias_plans.push_back(ap);
            aliased = ap;
        }
        use = aliased->plan.get();
    }
    const size_t n_rows = use->info.num_rows;
    const size_t nids = in_eff.size() + out_eff.size();
    if ((rc = grow(s->ids, s->ids_cap, (nids ? nids : 1) * 4))) return rc;
    if ((rc = grow(s->slab, s->slab_cap, (size_t)s->batch * n_rows * 16))) return rc;
    uint32_t* d_in = s->ids.as<uint32_t>();
    uint32_t* d_out = d_in + in_eff.size();
    if (!in_eff.empty()) CK(cudaMemcpyAsync(d_in, in_eff.data(), in_eff.size() * 4, cudaMemcpyHostToDevice, s->cs));
    if (!out_eff.empty()) CK(cudaMemcpyAsync(d_out, out_eff.data(), out_eff.size() * 4, cudaMemcpyHostToDevice, s->cs));
    const auto t_mid = clk::now();

    // the gate loop (stream_garble.go:179-190)
    rc = launch_gc(true, *use, di, s->device, s->keys.as<uint8_t>(), s->keylen, s->key_stride, s->batch,
                   s->r.as<gcb_label>(), nullptr, s->slab.as<gcb_label>(), nullptr, nullptr, s->cs, d_in, d_out,
                   reinterpret_cast<uint4* const*>(s->page_table.p));
    if (rc) return rc;
    if (late_layout) {
        rc = build_stream_layout(base.gates, nw, in, nin, out, nout, lay, err);
        if (rc) { cudaStreamSynchronize(s->cs); return fail(rc, "%s", err.c_str()); }
        total = lay.tmpl.size();
        if (written) *written = total;
    }
    const size_t stride16 = (total + 15) & ~(size_t)15;
    if (!s->ds) {
        CK(cudaStreamCreateWithFlags(&s->ds, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
        for (cudaEvent_t& e : s->ev_free) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint32_t cur = s->ser_turn++ & 1u;                 // staging buffers used in turn
    DevBuf& ser = cur ? s->ser2 : s->ser;
    size_t& ser_cap = cur ? s->ser2_cap : s->ser_cap;
    if ((rc = grow(ser, ser_cap, (size_t)s->batch * stride16))) return rc;
    if ((rc = grow(s->tmpl, s->tmpl_cap, stride16 + 16))) return rc;
    if ((rc = grow(s->row_pos, s->row_pos_cap, (n_rows ? n_rows : 1) * 4))) return rc;
    // header template and row offsets go through page-locked staging (two sets, like the buffers above): a
    // copy from pageable memory would hold this thread until the stream has drained up to it
    const size_t stage_bytes = stride16 + (n_rows ? n_rows : 1) * 4;
    if (s->stage_cap[cur] < stage_bytes) {
        if (s->stage[cur]) { CK(cudaStreamSynchronize(s->cs)); cudaFreeHost(s->stage[cur]); s->stage[cur] = nullptr; s->stage_cap[cur] = 0; }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&s->stage[cur]), stage_bytes + stage_bytes / 4, cudaHostAllocDefault));
        s->stage_cap[cur] = stage_bytes + stage_bytes / 4;
    }
    if (s->ev_valid[cur]) CK(cudaEventSynchronize(s->ev_free[cur]));   // the copy that last used this set (two calls ago) is done
    uint8_t* h_tmpl = s->stage[cur];
    uint32_t* h_rows = reinterpret_cast<uint32_t*>(s->stage[cur] + stride16);
    memcpy(h_tmpl, lay.tmpl.data(), total);
    memset(h_tmpl + total, 0, stride16 - total);
    if (n_rows) memcpy(h_rows, lay.row_pos.data(), n_rows * 4);
    if (total) CK(cudaMemcpyAsync(s->tmpl.p, h_tmpl, stride16, cudaMemcpyHostToDevice, s->cs));
    if (n_rows) CK(cudaMemcpyAsync(s->row_pos.p, h_rows, n_rows * 4, cudaMemcpyHostToDevice, s->cs));
    if (total) {
        SerParams sp{s->tmpl.as<uint8_t>(), (uint32_t)total, s->row_pos.as<uint32_t>(), (uint32_t)n_rows,
                     s->slab.as<uint4>(), ser.as<uint8_t>(), stride16};
        const dim3 grid((unsigned)((total + SER_TILE - 1) / SER_TILE), s->batch);
        serialize_kernel<<<grid, SER_THREADS, 0, s->cs>>>(sp);
        CK(launched());
        CK(cudaEventRecord(s->ev_h2d, s->cs));
        CK(cudaStreamWaitEvent(s->ds, s->ev_h2d, 0));
        CK(cudaMemcpy2DAsync(dst, dst_stride, ser.p, stride16, total, s->batch, cudaMemcpyDeviceToHost, s->ds));
    }
    CK(cudaEventRecord(s->ev_free[cur], s->ds));
    s->ev_valid[cur] = true;
    if (wait) { CK(cudaStreamSynchronize(s->ds)); CK(cudaStreamSynchronize(s->cs)); }
    const auto t_end = clk::now();
    if (ns_init) *ns_init = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_mid - t_start).count();
    if (ns_garble) *ns_garble = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_end - t_mid).count();
    return GCB_OK;
}

int gcb_stream_garble(gcb_stream* s, const gcb_plan* plan, const uint32_t* in, uint32_t nin, const uint32_t* out,
                      uint32_t nout, uint8_t* dst, size_t dst_stride, size_t* written, uint64_t* ns_init,
                      uint64_t* ns_garble) {
    GCB_TRY
    return stream_garble_impl(s, plan, in, nin, out, nout, dst, dst_stride, written, ns_init, ns_garble, true);
    GCB_CATCH
}
int gcb_stream_garble_begin(gcb_stream* s, const gcb_plan* plan, const uint32_t* in, uint32_t nin, const uint32_t* out,
                            uint32_t nout, uint8_t* dst, size_t dst_stride, size_t* written) {
    GCB_TRY
    return stream_garble_impl(s, plan, in, nin, out, nout, dst, dst_stride, written, nullptr, nullptr, false);
    GCB_CATCH
}
int gcb_stream_garble_wait(gcb_stream* s, uint32_t leave_in_flight) {
    GCB_TRY
    if (!s) return fail(GCB_E_ARG, "null argument");
    if (leave_in_flight > 1) return fail(GCB_E_ARG, "at most one step can stay in flight");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = use_device(s->device, nullptr);
    if (rc) return rc;
    if (leave_in_flight == 1) {                       // everything but the step begun last
        const uint32_t prev = s->ser_turn & 1u;       // the set the step before the last one used
        if (s->ser_turn >= 2 && s->ev_valid[prev]) CK(cudaEventSynchronize(s->ev_free[prev]));
        return GCB_OK;
    }
    if (s->ds) CK(cudaStreamSynchronize(s->ds));
    CK(cudaStreamSynchronize(s->cs));
    return GCB_OK;
    GCB_CATCH
}

// ------------------------------------------------------- streaming evaluator ----
}  // extern "C"  (helpers below are C++)

struct gcb_seval : gcb_stream {};

extern "C" {

int gcb_seval_create(const uint8_t* keys, uint32_t keylen, uint32_t key_stride, uint32_t batch, gcb_seval** out) {
    GCB_TRY
    if (!out) return fail(GCB_E_ARG, "null output pointer");
    *out = nullptr;
    if (!keys || batch == 0) return fail(GCB_E_ARG, "null argument");
    int rc = check_keylen(keylen);
    if (rc) return rc;
    if (key_stride && key_stride < keylen) return fail(GCB_E_ARG, "key_stride smaller than keylen");
    if ((rc = select_device(nullptr))) return rc;
    auto s = std::make_unique<gcb_seval>();
    s->device = tl_cur; s->batch = batch; s->keylen = keylen; s->key_stride = key_stride;
    CK(cudaStreamCreateWithFlags(&s->cs, cudaStreamNonBlocking));
    const size_t kb = key_stride ? (size_t)key_stride * batch : keylen;
    CK(s->keys.alloc(kb));
    CK(cudaMemcpyAsync(s->keys.p, keys, kb, cudaMemcpyHostToDevice, s->cs));
    CK(cudaStreamSynchronize(s->cs));
    *out = s.release();
    return GCB_OK;
    GCB_CATCH
}
void gcb_seval_destroy(gcb_seval* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->ds) cudaStreamSynchronize(s->ds);
    if (s->cs) cudaStreamSynchronize(s->cs);
    delete s;
}

// StreamEval.Set / SetInputs (stream_evaluator.go:68-96): labels [batch][n].
int gcb_seval_set_wires(gcb_seval* s, const uint32_t* ids, uint32_t n, const gcb_label* labels) {
    GCB_TRY
    if (!s || (n && (!ids || !labels))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceInfo* di;
    int rc = use_device(s->device, &di);
    if (rc) return rc;
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n; i++) mx = ids[i] > mx ? ids[i] : mx;
    if ((rc = ensure_wires(s, mx))) return rc;
    if ((rc = grow(s->ids, s->ids_cap, (size_t)n * 4))) return rc;
    if ((rc = grow(s->wires, s->wires_cap, (size_t)s->batch * n * 16))) return rc;
    CK(cudaMemcpyAsync(s->ids.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s->cs));
    CK(cudaMemcpyAsync(s->wires.p, labels, (size_t)s->batch * n * 16, cudaMemcpyHostToDevice, s->cs));
    wf_set_kernel<<<di->sm_count * 4, 256, 0, s->cs>>>(reinterpret_cast<uint4* const*>(s->page_table.p),
                                                       s->ids.as<uint32_t>(), n, s->wires.as<uint4>(), s->batch);
    CK(launched());
    CK(cudaStreamSynchronize(s->cs));
    return GCB_OK;
    GCB_CATCH
}
// StreamEval.Get (stream_evaluator.go:57-66): labels [batch][n].
int gcb_seval_get_wires(gcb_seval* s, const uint32_t* ids, uint32_t n, gcb_label* labels) {
    GCB_TRY
    if (!s || (n && (!ids || !labels))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceInfo* di;
    int rc = use_device(s->device, &di);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++)
        if (((size_t)ids[i] >> WF_PAGE_SHIFT) >= s->pages.size()) return fail(GCB_E_WIRE, "wire %u not allocated", ids[i]);
    if ((rc = grow(s->ids, s->ids_cap, (size_t)n * 4))) return rc;
    if ((rc = grow(s->wires, s->wires_cap, (size_t)s->batch * n * 16))) return rc;
    CK(cudaMemcpyAsync(s->ids.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s->cs));
    wf_get_labels_kernel<<<di->sm_count * 4, 256, 0, s->cs>>>(reinterpret_cast<uint4* const*>(s->page_table.p),
                                                              s->ids.as<uint32_t>(), n, s->wires.as<uint4>(), s->batch);
    CK(launched());
    CK(cudaMemcpyAsync(labels, s->wires.p, (size_t)s->batch * n * 16, cudaMemcpyDeviceToHost, s->cs));
    CK(cudaStreamSynchronize(s->cs));
    return GCB_OK;
    GCB_CATCH
}

// The gate loop of StreamEvaluator for one OpCircuit body (stream_evaluator.go:270-432).
// src: [batch][src_stride] host bytes, each instance's record stream for this circuit (all
// instances carry the same gate headers: same circuit, same wire ids; only the rows differ).
int gcb_seval_circuit(gcb_seval* s, const uint8_t* src, size_t src_stride, size_t len, uint32_t ngates,
                      uint32_t ntmp, uint32_t nwires, size_t* consumed) {
    GCB_TRY
    if (!s || (ngates && !src)) return fail(GCB_E_ARG, "null argument");
    if (src_stride < len && s->batch > 1) return fail(GCB_E_ARG, "src_stride smaller than len");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceInfo* di;
    int rc = use_device(s->device, &di);
    if (rc) return rc;
    // the record bytes start moving to the device first: parsing the headers and recovering the plan
    // on the host overlap the copy (2 GB per step in the config-5 stand-in)
    const size_t stride16 = ((len + 15) & ~(size_t)15) + 16;
    if (!s->ds) {
        CK(cudaStreamCreateWithFlags(&s->ds, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
        for (cudaEvent_t& e : s->ev_free) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // Two staging buffers used in turn, the copy on its own stream: the call returns when ITS bytes are on
    // the device (the caller may reuse `src`) while its kernel is still queued or running, so the copy of
    // the next call overlaps it.  Readers of the wire file (gcb_seval_get_wires) are ordered on stream cs.
    const uint32_t cur = s->ser_turn++ & 1u;
    DevBuf& ser = cur ? s->ser2 : s->ser;
    size_t& ser_cap = cur ? s->ser2_cap : s->ser_cap;
    if (ngates && len) {
        if ((rc = grow(ser, ser_cap, (size_t)s->batch * stride16))) return rc;
        if (s->ev_valid[cur]) CK(cudaStreamWaitEvent(s->ds, s->ev_free[cur], 0));   // its previous reader has run
        CK(cudaMemcpy2DAsync(ser.p, stride16, src, s->batch > 1 ? src_stride : len, len, s->batch, cudaMemcpyHostToDevice, s->ds));
        CK(cudaEventRecord(s->ev_h2d, s->ds));
    }
    // on every path out of this call the copy has finished before the caller may reuse `src`
    struct SyncOnExit { cudaStream_t st; ~SyncOnExit() { cudaStreamSynchronize(st); } } sync_on_exit{s->ds};
    // parse instance 0's headers
    std::vector<StreamGate> sg;
    std::vector<uint32_t> row_pos;
    std::string err;
    size_t used = 0;
    if ((rc = parse_stream(src, len, ngates, sg, row_pos, &used, err))) return fail(rc, "%s", err.c_str());
    if (consumed) *consumed = used;
    if (ngates == 0) return GCB_OK;
    // InitCircuit (stream_evaluator.go:86-96): wires up to nwires exist.  The count comes from the peer; the reference
    // would size a Go slice by it.  Here: at most 2^28 wire ids per evaluator (4 GiB of labels per instance).
    if (nwires > kMaxStreamWires)
        return fail(GCB_E_TOO_LARGE, "record stream declares %u wires; at most %u are supported", nwires, kMaxStreamWires);
    if (nwires) { if ((rc = ensure_wires(s, nwires - 1))) return rc; }
    // wire space of the recovered circuit: [permanent ids in order of first use | tmp wires].  The
    // plan depends only on this canonical form, not on the actual ids, so the steps of a program
    // that reuse a sub-circuit with fresh ids share one cached plan.
    if (s->perm_loc.size() < nwires) { s->perm_loc.resize(nwires, 0); s->perm_epoch.resize(nwires, 0); }
    if (++s->epoch == 0) { std::fill(s->perm_epoch.begin(), s->perm_epoch.end(), 0u); s->epoch = 1; }
    const uint32_t epoch = s->epoch;
    std::vector<uint32_t> perm_ids;
    std::vector<uint8_t> first_is_read, written;
    auto perm_loc = [&](uint32_t id, bool is_read) {
        if (s->perm_epoch[id] == epoch) { const uint32_t l = s->perm_loc[id]; if (!is_read) written[l] = 1; return l; }
        const uint32_t l = (uint32_t)perm_ids.size();
        s->perm_epoch[id] = epoch; s->perm_loc[id] = l; perm_ids.push_back(id);
        first_is_read.push_back(is_read); written.push_back(!is_read);
        return l;
    };
    struct Ref { uint32_t loc; bool tmp; };
    std::vector<std::array<Ref, 3>> refs(sg.size());
    std::vector<uint32_t> canon;
    canon.reserve(sg.size() * 4);
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix(ntmp);
    for (size_t i = 0; i < sg.size(); i++) {
        const StreamGate& g = sg[i];
        for (const auto& t : {std::pair<uint32_t, bool>{g.a, g.a_tmp}, {g.b, g.b_tmp}, {g.c, g.c_tmp}})
            if (t.second && t.first >= ntmp) return fail(GCB_E_WIRE, "tmp wire %u out of range (gate %zu)", t.first, i);
        if (!g.a_tmp && g.a >= nwires) return fail(GCB_E_WIRE, "wire %u out of range (gate %zu)", g.a, i);
        if (g.op != OP_INV && !g.b_tmp && g.b >= nwires) return fail(GCB_E_WIRE, "wire %u out of range (gate %zu)", g.b, i);
        if (!g.c_tmp && g.c >= nwires) return fail(GCB_E_WIRE, "wire %u out of range (gate %zu)", g.c, i);
        refs[i][0] = Ref{g.a_tmp ? g.a : perm_loc(g.a, true), (bool)g.a_tmp};
        refs[i][1] = g.op == OP_INV ? refs[i][0] : Ref{g.b_tmp ? g.b : perm_loc(g.b, true), (bool)g.b_tmp};
        refs[i][2] = Ref{g.c_tmp ? g.c : perm_loc(g.c, false), (bool)g.c_tmp};
        const uint32_t head = g.op | (g.a_tmp << 8) | (g.b_tmp << 9) | (g.c_tmp << 10);
        mix(head);
        mix(refs[i][0].loc); mix(refs[i][1].loc); mix(refs[i][2].loc);
        canon.push_back(head); canon.push_back(refs[i][0].loc); canon.push_back(refs[i][1].loc); canon.push_back(refs[i][2].loc);
    }
    const uint32_t np = (uint32_t)perm_ids.size();
    std::shared_ptr<gcb_stream::EvalPlan> ep;
    auto it = s->eval_plans.find(h);
    if (it != s->eval_plans.end() && it->second->ntmp == ntmp && it->second->np == np && it->second->canon == canon)
        ep = it->second;
    else {
        std::vector<gcb_gate> gates(sg.size());
        for (size_t i = 0; i < sg.size(); i++) {
            auto w = [&](const Ref& r) { return r.tmp ? np + r.loc : r.loc; };
            gates[i] = gcb_gate{w(refs[i][0]), w(refs[i][1]), w(refs[i][2]), sg[i].op, {0, 0, 0}, 0};
        }
        ep = std::make_shared<gcb_stream::EvalPlan>();
        PlanSpec spec;
        spec.gates = gates.data(); spec.num_gates = (uint32_t)gates.size(); spec.num_wires = np + ntmp;
        for (uint32_t l = 0; l < np; l++) {
            if (first_is_read[l]) { spec.live_in.push_back(l); ep->in_ids.push_back(l); }     // canonical locations;
            if (written[l]) { spec.live_out.push_back(l); ep->out_ids.push_back(l); }         // mapped to ids per call
        }
        if ((rc = build_best_plan(spec, ep->plan, err, NODE_MAX_FANIN, s->batch))) return fail(rc == GCB_E_WIRE ? GCB_E_CORRUPT : rc, "corrupted circuit: %s", err.c_str());
        team_geometry(ep->plan);
        if (ep->plan.info.teams_per_sm == 0) return fail(GCB_E_TOO_LARGE, "sub-circuit keeps %u wire labels live", ep->plan.info.num_slots);
        ep->canon.swap(canon);
        ep->ntmp = ntmp; ep->np = np;
        if (s->eval_plans.size() > 32) s->eval_plans.clear();
        s->eval_plans[h] = ep;                          // replaces a colliding entry
    }
    std::vector<uint32_t> in_ids(ep->in_ids.size()), out_ids(ep->out_ids.size());
    for (size_t k = 0; k < in_ids.size(); k++) in_ids[k] = perm_ids[ep->in_ids[k]];
    for (size_t k = 0; k < out_ids.size(); k++) out_ids[k] = perm_ids[ep->out_ids[k]];
    const size_t n_rows = row_pos.size();
    const size_t nids = in_ids.size() + out_ids.size();
    if ((rc = grow(s->ids, s->ids_cap, (nids ? nids : 1) * 4))) return rc;
    if ((rc = grow(s->slab, s->slab_cap, (size_t)s->batch * (n_rows ? n_rows : 1) * 16))) return rc;
    if ((rc = grow(s->row_pos, s->row_pos_cap, (n_rows ? n_rows : 1) * 4))) return rc;
    uint32_t* d_in = s->ids.as<uint32_t>();
    uint32_t* d_out = d_in + in_ids.size();
    if (!in_ids.empty()) CK(cudaMemcpyAsync(d_in, in_ids.data(), in_ids.size() * 4, cudaMemcpyHostToDevice, s->cs));
    if (!out_ids.empty()) CK(cudaMemcpyAsync(d_out, out_ids.data(), out_ids.size() * 4, cudaMemcpyHostToDevice, s->cs));
    if (n_rows) {
        CK(cudaMemcpyAsync(s->row_pos.p, row_pos.data(), n_rows * 4, cudaMemcpyHostToDevice, s->cs));
        CK(cudaStreamWaitEvent(s->cs, s->ev_h2d, 0));
        DeserParams dp{ser.as<uint8_t>(), stride16, s->row_pos.as<uint32_t>(), (uint32_t)n_rows, s->slab.as<uint4>(), s->batch};
        deserialize_kernel<<<di->sm_count * 8, 256, 0, s->cs>>>(dp);
        CK(launched());
        CK(cudaEventRecord(s->ev_free[cur], s->cs));
        s->ev_valid[cur] = true;
    }
    rc = launch_gc(false, ep->plan, di, s->device, s->keys.as<uint8_t>(), s->keylen, s->key_stride, s->batch, nullptr,
                   nullptr, s->slab.as<gcb_label>(), nullptr, nullptr, s->cs, d_in, d_out,
                   reinterpret_cast<uint4* const*>(s->page_table.p));
    if (rc) return rc;
    return GCB_OK;                          // sync_on_exit waits for the copy; the kernel completes on stream cs
    GCB_CATCH
}

// ------------------------------------------------------------------- IKNP -------
size_t gcb_iknp_u_size(uint64_t n) {
    const uint64_t full = n / IKNP_CHUNK_ROWS, rem = n % IKNP_CHUNK_ROWS;
    return (size_t)(full * 64 * 128 + ((rem + 7) / 8) * 128);
}
uint64_t gcb_iknp_stream_advance(uint64_t n) {
    const uint64_t full = n / IKNP_CHUNK_ROWS, rem = n % IKNP_CHUNK_ROWS;
    return full * 64 + (rem + 7) / 8;
}

static int launch_iknp(bool receiver, bool bits, const IknpParams& p0, void* stream) {
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    IknpParams p = p0;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = fresh_counter(di, s, &p.counter))) return rc;
    const uint64_t nchunks = (p.n + IKNP_CHUNK_ROWS - 1) / IKNP_CHUNK_ROWS;
    const uint32_t groups = receiver ? 2 : 4;
    const uint64_t want = (nchunks + groups - 1) / groups;
    const dim3 grid((unsigned)(want < (uint64_t)di->sm_count ? want : (uint64_t)di->sm_count));
    const size_t smem = table_pad(di->smem_base) + AES_TABLE_BYTES + groups * (IKNP_STAGE_BYTES + (receiver ? 8192 : 0) + 128);
    if (receiver && bits) iknp_kernel<true, true><<<grid, IKNP_CTA_THREADS, smem, s>>>(p);
    else if (receiver) iknp_kernel<true, false><<<grid, IKNP_CTA_THREADS, smem, s>>>(p);
    else if (bits) iknp_kernel<false, true><<<grid, IKNP_CTA_THREADS, smem, s>>>(p);
    else iknp_kernel<false, false><<<grid, IKNP_CTA_THREADS, smem, s>>>(p);
    CK(launched());
    return GCB_OK;
}

int gcb_iknp_receiver_expand_dev(const gcb_label* k0, const gcb_label* k1, uint64_t stream_pos,
                                 const uint8_t* choice, uint64_t n, uint8_t* u_out, gcb_label* labels,
                                 void* stream) {
    GCB_TRY
    if (!k0 || !k1 || (n && (!choice || !u_out || !labels))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    if (reinterpret_cast<uintptr_t>(u_out) & 15) return fail(GCB_E_ARG, "u_out must be 16-byte aligned");
    IknpParams p{};
    p.k0 = reinterpret_cast<const uint4*>(k0); p.k1 = reinterpret_cast<const uint4*>(k1);
    p.stream_pos = stream_pos; p.choice = choice; p.u_out = u_out;
    p.labels = reinterpret_cast<uint4*>(labels); p.n = n;
    return launch_iknp(true, false, p, stream);
    GCB_CATCH
}
int gcb_iknp_sender_expand_dev(const gcb_label* k, const gcb_label* delta, uint64_t stream_pos, const uint8_t* u,
                               size_t u_len, uint64_t n, gcb_label* labels, void* stream) {
    GCB_TRY
    if (!k || !delta || (n && (!u || !labels))) return fail(GCB_E_ARG, "null argument");
    if (u_len != gcb_iknp_u_size(n)) return fail(GCB_E_CHUNK, "invalid chunk size: %zu bytes for %llu OTs", u_len,
                                                 (unsigned long long)n);
    if (n == 0) return GCB_OK;
    if (reinterpret_cast<uintptr_t>(u) & 15) return fail(GCB_E_ARG, "u must be 16-byte aligned");
    IknpParams p{};
    p.k0 = reinterpret_cast<const uint4*>(k); p.delta = reinterpret_cast<const uint4*>(delta);
    p.stream_pos = stream_pos; p.u_in = u; p.labels = reinterpret_cast<uint4*>(labels); p.n = n;
    return launch_iknp(false, false, p, stream);
    GCB_CATCH
}

// Host-pointer variants: rows split by 512-row chunk ranges over the selected devices (the CTR
// keystream is random access, so a range only needs its byte offset into the column streams).
static uint64_t iknp_chunks(uint64_t n) { return (n + IKNP_CHUNK_ROWS - 1) / IKNP_CHUNK_ROWS; }
int gcb_iknp_receiver_expand(const gcb_label k0[128], const gcb_label k1[128], uint64_t stream_pos,
                             const uint8_t* choice, uint64_t n, uint8_t* u_out, gcb_label* labels) {
    GCB_TRY
    if (!k0 || !k1 || (n && (!choice || !u_out || !labels))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    return fan_out(iknp_chunks(n), [&](int dev, uint64_t clo, uint64_t chi, JobPart* part) -> int {
        const uint64_t lo = clo * IKNP_CHUNK_ROWS, hi = std::min<uint64_t>(n, chi * IKNP_CHUNK_ROWS), m = hi - lo;
        HostOp a{k0, nullptr, 128 * 16}, b{k1, nullptr, 128 * 16}, c{choice + lo, nullptr, (size_t)m};
        HostOp u{nullptr, u_out + clo * 64 * 128, gcb_iknp_u_size(m)}, l{nullptr, labels + lo, (size_t)m * 16};
        int rc = stage_begin(dev, {&a, &b, &c, &u, &l}, part);
        if (rc) return rc;
        if ((rc = gcb_iknp_receiver_expand_dev((const gcb_label*)a.dptr, (const gcb_label*)b.dptr, stream_pos + clo * 64,
                                               (const uint8_t*)c.dptr, m, (uint8_t*)u.dptr, (gcb_label*)l.dptr, part->res->k)))
            return rc;
        return stage_end({&u, &l}, part);
    });
    GCB_CATCH
}
int gcb_iknp_sender_expand(const gcb_label k[128], const gcb_label* delta, uint64_t stream_pos, const uint8_t* u,
                           size_t u_len, uint64_t n, gcb_label* labels) {
    GCB_TRY
    if (!k || !delta || (n && (!u || !labels))) return fail(GCB_E_ARG, "null argument");
    if (u_len != gcb_iknp_u_size(n)) return fail(GCB_E_CHUNK, "invalid chunk size: %zu bytes for %llu OTs", u_len,
                                                 (unsigned long long)n);
    if (n == 0) return GCB_OK;
    return fan_out(iknp_chunks(n), [&](int dev, uint64_t clo, uint64_t chi, JobPart* part) -> int {
        const uint64_t lo = clo * IKNP_CHUNK_ROWS, hi = std::min<uint64_t>(n, chi * IKNP_CHUNK_ROWS), m = hi - lo;
        HostOp a{k, nullptr, 128 * 16}, d{delta, nullptr, 16}, ui{u + clo * 64 * 128, nullptr, gcb_iknp_u_size(m)};
        HostOp l{nullptr, labels + lo, (size_t)m * 16};
        int rc = stage_begin(dev, {&a, &d, &ui, &l}, part);
        if (rc) return rc;
        if ((rc = gcb_iknp_sender_expand_dev((const gcb_label*)a.dptr, (const gcb_label*)d.dptr, stream_pos + clo * 64,
                                             (const uint8_t*)ui.dptr, ui.bytes, m, (gcb_label*)l.dptr, part->res->k)))
            return rc;
        return stage_end({&l}, part);
    });
    GCB_CATCH
}

// Bit-COT variants: ReceiveBits / SendBits (ot/iknp.go:554-620, 259-310).  choices / result are
// the packed LSB-first []uint64 of the Go API ((n+63)/64 words; result is overwritten).
int gcb_iknp_receiver_expand_bits_dev(const gcb_label* k0, const gcb_label* k1, uint64_t stream_pos,
                                      const uint64_t* choices, uint64_t n, uint8_t* u_out, uint64_t* result,
                                      void* stream) {
    GCB_TRY
    if (!k0 || !k1 || (n && (!choices || !u_out || !result))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    if (reinterpret_cast<uintptr_t>(u_out) & 15) return fail(GCB_E_ARG, "u_out must be 16-byte aligned");
    CK(cudaMemsetAsync(result, 0, ((n + 63) / 64) * 8, (cudaStream_t)stream));
    IknpParams p{};
    p.k0 = reinterpret_cast<const uint4*>(k0); p.k1 = reinterpret_cast<const uint4*>(k1);
    p.stream_pos = stream_pos; p.choice_bits = reinterpret_cast<const uint32_t*>(choices); p.u_out = u_out;
    p.result_bits = reinterpret_cast<uint32_t*>(result); p.n = n;
    return launch_iknp(true, true, p, stream);
    GCB_CATCH
}
int gcb_iknp_sender_expand_bits_dev(const gcb_label* k, const gcb_label* delta, uint64_t stream_pos, const uint8_t* u,
                                    size_t u_len, uint64_t n, uint64_t* result, void* stream) {
    GCB_TRY
    if (!k || !delta || (n && (!u || !result))) return fail(GCB_E_ARG, "null argument");
    if (u_len != gcb_iknp_u_size(n)) return fail(GCB_E_CHUNK, "invalid chunk size: %zu bytes for %llu OTs", u_len,
                                                 (unsigned long long)n);
    if (n == 0) return GCB_OK;
    if (reinterpret_cast<uintptr_t>(u) & 15) return fail(GCB_E_ARG, "u must be 16-byte aligned");
    CK(cudaMemsetAsync(result, 0, ((n + 63) / 64) * 8, (cudaStream_t)stream));
    IknpParams p{};
    p.k0 = reinterpret_cast<const uint4*>(k); p.delta = reinterpret_cast<const uint4*>(delta);
    p.stream_pos = stream_pos; p.u_in = u; p.result_bits = reinterpret_cast<uint32_t*>(result); p.n = n;
    return launch_iknp(false, true, p, stream);
    GCB_CATCH
}
int gcb_iknp_receiver_expand_bits(const gcb_label k0[128], const gcb_label k1[128], uint64_t stream_pos,
                                  const uint64_t* choices, uint64_t n, uint8_t* u_out, uint64_t* result) {
    GCB_TRY
    if (!k0 || !k1 || (n && (!choices || !u_out || !result))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    const uint64_t words = (n + 63) / 64;
    return fan_out(iknp_chunks(n), [&](int dev, uint64_t clo, uint64_t chi, JobPart* part) -> int {
        const uint64_t lo = clo * IKNP_CHUNK_ROWS, hi = std::min<uint64_t>(n, chi * IKNP_CHUNK_ROWS), m = hi - lo;
        // the reference reads whole 64-bit choice words per full 8-byte row group: the range's words, zero-padded
        // to the end of its last chunk
        const uint64_t w0 = lo / 64, cwords = (chi - clo) * 8, have = std::min<uint64_t>(cwords, words - w0);
        std::vector<uint64_t> cw(cwords, 0);
        memcpy(cw.data(), choices + w0, have * 8);
        HostOp a{k0, nullptr, 128 * 16}, b{k1, nullptr, 128 * 16}, c{cw.data(), nullptr, (size_t)cwords * 8};
        HostOp u{nullptr, u_out + clo * 64 * 128, gcb_iknp_u_size(m)}, r{nullptr, result + w0, (size_t)((m + 63) / 64) * 8};
        int rc = stage_begin(dev, {&a, &b, &c, &u, &r}, part);      // cw is pageable: copied into staging right here
        if (rc) return rc;
        if ((rc = gcb_iknp_receiver_expand_bits_dev((const gcb_label*)a.dptr, (const gcb_label*)b.dptr, stream_pos + clo * 64,
                                                    (const uint64_t*)c.dptr, m, (uint8_t*)u.dptr, (uint64_t*)r.dptr, part->res->k)))
            return rc;
        return stage_end({&u, &r}, part);
    });
    GCB_CATCH
}
int gcb_iknp_sender_expand_bits(const gcb_label k[128], const gcb_label* delta, uint64_t stream_pos, const uint8_t* u,
                                size_t u_len, uint64_t n, uint64_t* result) {
    GCB_TRY
    if (!k || !delta || (n && (!u || !result))) return fail(GCB_E_ARG, "null argument");
    if (u_len != gcb_iknp_u_size(n)) return fail(GCB_E_CHUNK, "invalid chunk size: %zu bytes for %llu OTs", u_len,
                                                 (unsigned long long)n);
    if (n == 0) return GCB_OK;
    return fan_out(iknp_chunks(n), [&](int dev, uint64_t clo, uint64_t chi, JobPart* part) -> int {
        const uint64_t lo = clo * IKNP_CHUNK_ROWS, hi = std::min<uint64_t>(n, chi * IKNP_CHUNK_ROWS), m = hi - lo;
        HostOp a{k, nullptr, 128 * 16}, d{delta, nullptr, 16}, ui{u + clo * 64 * 128, nullptr, gcb_iknp_u_size(m)};
        HostOp r{nullptr, result + lo / 64, (size_t)((m + 63) / 64) * 8};
        int rc = stage_begin(dev, {&a, &d, &ui, &r}, part);
        if (rc) return rc;
        if ((rc = gcb_iknp_sender_expand_bits_dev((const gcb_label*)a.dptr, (const gcb_label*)d.dptr, stream_pos + clo * 64,
                                                  (const uint8_t*)ui.dptr, ui.bytes, m, (uint64_t*)r.dptr, part->res->k)))
            return rc;
        return stage_end({&r}, part);
    });
    GCB_CATCH
}

// ----------------------------------------------------------------- MiTCCRH ------
int gcb_mitccrh_hash_dev(const gcb_label* seed_host, uint64_t gid_start, gcb_label* blks, uint64_t nkeys, uint32_t h,
                         void* stream) {
    GCB_TRY
    if (!seed_host || (nkeys && !blks)) return fail(GCB_E_ARG, "null argument");
    if (h == 0 || h > 64) return fail(GCB_E_ARG, "MITCCRH.Hash: invalid H %u", h);   // mitccrh.go:94-102 panics
    if (nkeys == 0) return GCB_OK;
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    MitccrhParams p{seed_host->d0, seed_host->d1, gid_start, reinterpret_cast<uint4*>(blks), nkeys, h};
    const uint64_t want = (nkeys + 511) / 512;
    const dim3 grid((unsigned)(want < (uint64_t)di->sm_count ? want : (uint64_t)di->sm_count));
    mitccrh_kernel<<<grid, 512, table_pad(di->smem_base) + AES_TABLE_BYTES, (cudaStream_t)stream>>>(p);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}
int gcb_mitccrh_hash(const gcb_label* seed, uint64_t gid_start, gcb_label* blks, uint64_t nkeys, uint32_t h) {
    GCB_TRY
    if (!seed || (nkeys && !blks)) return fail(GCB_E_ARG, "null argument");
    if (h == 0 || h > 64) return fail(GCB_E_ARG, "MITCCRH.Hash: invalid H %u", h);
    if (nkeys == 0) return GCB_OK;
    return fan_out(nkeys, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        HostOp b{blks + lo * h, blks + lo * h, (size_t)(hi - lo) * h * 16};       // in place
        int rc = stage_begin(dev, {&b}, part);
        if (rc) return rc;
        if ((rc = gcb_mitccrh_hash_dev(seed, gid_start + lo, (gcb_label*)b.dptr, hi - lo, h, part->res->k))) return rc;
        return stage_end({&b}, part);
    });
    GCB_CATCH
}

}  // extern "C"

// ------------------------------------------------- COT / ROT post-processing ------
template <int MODE>
static int launch_cot(const gcb_label* seed, const gcb_label* delta, const void* data, const void* wires,
                      const uint8_t* choice, const void* msgs_in, void* out, uint64_t n, uint32_t flags, void* stream,
                      uint64_t first = 0) {
    if (!seed || (n && (!data || !out))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    CotParams p{};
    p.seed_d0 = seed->d0; p.seed_d1 = seed->d1;
    if (delta) { p.delta_d0 = delta->d0; p.delta_d1 = delta->d1; }
    p.data = reinterpret_cast<const uint4*>(data);
    p.wires = reinterpret_cast<const uint4*>(wires);
    p.flags = choice;
    p.msgs_in = reinterpret_cast<const uint4*>(msgs_in);
    p.out = reinterpret_cast<uint4*>(out);
    p.n = n;
    p.first = first;
    p.wire_bytes = (flags & GCB_COT_WIRE_BYTES) ? 1u : 0u;
    const uint64_t want = (n + 511) / 512;
    const dim3 grid((unsigned)(want < (uint64_t)di->sm_count ? want : (uint64_t)di->sm_count));
    cot_kernel<MODE><<<grid, 512, table_pad(di->smem_base) + AES_TABLE_BYTES, (cudaStream_t)stream>>>(p);
    CK(launched());
    return GCB_OK;
}

extern "C" {

int gcb_cot_send_dev(const gcb_label* seed, const gcb_label* delta, const gcb_label* q, const gcb_wire* wires,
                     uint64_t n, gcb_label* msgs, uint32_t flags, void* stream) {
    GCB_TRY
    if (!delta || (n && !wires)) return fail(GCB_E_ARG, "null argument");
    return launch_cot<COT_SEND>(seed, delta, q, wires, nullptr, nullptr, msgs, n, flags, stream);
    GCB_CATCH
}
int gcb_cot_receive_dev(const gcb_label* seed, const uint8_t* choice, const gcb_label* msgs, const gcb_label* t,
                        uint64_t n, gcb_label* result, uint32_t flags, void* stream) {
    GCB_TRY
    if (n && (!choice || !msgs)) return fail(GCB_E_ARG, "null argument");
    return launch_cot<COT_RECEIVE>(seed, nullptr, t, nullptr, choice, msgs, result, n, flags, stream);
    GCB_CATCH
}
int gcb_rot_send_dev(const gcb_label* seed, const gcb_label* delta, const gcb_label* q, uint64_t n, gcb_wire* wires,
                     void* stream) {
    GCB_TRY
    if (!delta) return fail(GCB_E_ARG, "null argument");
    return launch_cot<ROT_SEND>(seed, delta, q, nullptr, nullptr, nullptr, wires, n, 0, stream);
    GCB_CATCH
}
int gcb_rot_receive_dev(const gcb_label* seed, const gcb_label* t, uint64_t n, gcb_label* result, void* stream) {
    GCB_TRY
    return launch_cot<ROT_RECEIVE>(seed, nullptr, t, nullptr, nullptr, nullptr, result, n, 0, stream);
    GCB_CATCH
}

// Host-pointer variants: OT ranges split over the selected devices (each OT is independent: MiTCCRH key number =
// OT number); the bulk path keeps the labels on the device with _dev.
int gcb_cot_send(const gcb_label* seed, const gcb_label* delta, const gcb_label* q, const gcb_wire* wires, uint64_t n,
                 gcb_label* msgs, uint32_t flags) {
    GCB_TRY
    if (!seed || !delta || (n && (!q || !wires || !msgs))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    return fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dq{q + lo, nullptr, m * 16}, dw{wires + lo, nullptr, m * 32}, dm{nullptr, msgs + 2 * lo, m * 32};
        int rc = stage_begin(dev, {&dq, &dw, &dm}, part);
        if (rc) return rc;
        if ((rc = launch_cot<COT_SEND>(seed, delta, dq.dptr, dw.dptr, nullptr, nullptr, dm.dptr, m, flags, part->res->k, lo))) return rc;
        return stage_end({&dm}, part);
    });
    GCB_CATCH
}
int gcb_cot_receive(const gcb_label* seed, const uint8_t* choice, const gcb_label* msgs, const gcb_label* t, uint64_t n,
                    gcb_label* result, uint32_t flags) {
    GCB_TRY
    if (!seed || (n && (!choice || !msgs || !t || !result))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    return fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dc{choice + lo, nullptr, m}, dm{msgs + 2 * lo, nullptr, m * 32}, dt{t + lo, result + lo, m * 16};
        int rc = stage_begin(dev, {&dc, &dm, &dt}, part);
        if (rc) return rc;
        if ((rc = launch_cot<COT_RECEIVE>(seed, nullptr, dt.dptr, nullptr, (const uint8_t*)dc.dptr, dm.dptr, dt.dptr, m, flags, part->res->k, lo)))
            return rc;
        return stage_end({&dt}, part);
    });
    GCB_CATCH
}
int gcb_rot_send(const gcb_label* seed, const gcb_label* delta, const gcb_label* q, uint64_t n, gcb_wire* wires) {
    GCB_TRY
    if (!seed || !delta || (n && (!q || !wires))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    return fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dq{q + lo, nullptr, m * 16}, dw{nullptr, wires + lo, m * 32};
        int rc = stage_begin(dev, {&dq, &dw}, part);
        if (rc) return rc;
        if ((rc = launch_cot<ROT_SEND>(seed, delta, dq.dptr, nullptr, nullptr, nullptr, dw.dptr, m, 0, part->res->k, lo))) return rc;
        return stage_end({&dw}, part);
    });
    GCB_CATCH
}
int gcb_rot_receive(const gcb_label* seed, const gcb_label* t, uint64_t n, gcb_label* result) {
    GCB_TRY
    if (!seed || (n && (!t || !result))) return fail(GCB_E_ARG, "null argument");
    if (n == 0) return GCB_OK;
    return fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dt{t + lo, result + lo, m * 16};
        int rc = stage_begin(dev, {&dt}, part);
        if (rc) return rc;
        if ((rc = launch_cot<ROT_RECEIVE>(seed, nullptr, dt.dptr, nullptr, nullptr, nullptr, dt.dptr, m, 0, part->res->k, lo))) return rc;
        return stage_end({&dt}, part);
    });
    GCB_CATCH
}

// ------------------------------------------- IKNP malicious-mode consistency sums --
// acc: 12 zeroed device words, XOR-accumulated
static int check_sums_enqueue(const gcb_label* seed2, uint64_t chi_start, const gcb_label* labels, const uint8_t* choice,
                              uint64_t n, uint32_t* acc, cudaStream_t stream) {
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    CheckParams p{seed2->d0, seed2->d1, chi_start, reinterpret_cast<const uint4*>(labels), choice, n, acc};
    const uint64_t want = (n + 511) / 512;
    const dim3 grid((unsigned)(want < (uint64_t)di->sm_count ? want : (uint64_t)di->sm_count));
    iknp_check_kernel<<<grid, 512, table_pad(di->smem_base) + AES_TABLE_BYTES + 512, stream>>>(p);
    CK(launched());
    return GCB_OK;
}
static void sums_from_words(const uint32_t* w, gcb_label out[3]) {
    for (int k = 0; k < 3; k++) {
        out[k].d0 ^= (uint64_t)w[4 * k] | ((uint64_t)w[4 * k + 1] << 32);
        out[k].d1 ^= (uint64_t)w[4 * k + 2] | ((uint64_t)w[4 * k + 3] << 32);
    }
}
int gcb_iknp_check_sums_dev(const gcb_label* seed2, uint64_t chi_start, const gcb_label* labels, const uint8_t* choice,
                            uint64_t n, gcb_label out[3], void* stream) {
    GCB_TRY
    if (!seed2 || !out || (n && !labels)) return fail(GCB_E_ARG, "null argument");
    memset(out, 0, 3 * sizeof(gcb_label));
    if (n == 0) return GCB_OK;
    int rc = select_device(nullptr);
    if (rc) return rc;
    uint32_t* acc = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&acc), 12 * sizeof(uint32_t), (cudaStream_t)stream));
    CK(cudaMemsetAsync(acc, 0, 12 * sizeof(uint32_t), (cudaStream_t)stream));
    rc = check_sums_enqueue(seed2, chi_start, labels, choice, n, acc, (cudaStream_t)stream);
    uint32_t w[12] = {0};
    if (!rc) CK(cudaMemcpyAsync(w, acc, sizeof w, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaFreeAsync(acc, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (rc) return rc;
    sums_from_words(w, out);
    return GCB_OK;
    GCB_CATCH
}
// Host-pointer variant: OT ranges over the selected devices; the partial sums are XORed on the host (SURVEY.md 8e:
// "gather 32 B per rank and fold on host" -- NCCL has no XOR reduction).
int gcb_iknp_check_sums(const gcb_label* seed2, uint64_t chi_start, const gcb_label* labels, const uint8_t* choice,
                        uint64_t n, gcb_label out[3]) {
    GCB_TRY
    if (!seed2 || !out || (n && !labels)) return fail(GCB_E_ARG, "null argument");
    memset(out, 0, 3 * sizeof(gcb_label));
    if (n == 0) return GCB_OK;
    const size_t nd = call_devices().size();
    std::vector<std::array<uint32_t, 12>> partial(nd ? nd : 1);
    for (auto& w : partial) w.fill(0);
    uint32_t next = 0;
    int rc = fan_out(n, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dl{labels + lo, nullptr, m * 16}, dc{choice ? choice + lo : nullptr, nullptr, choice ? m : 0};
        HostOp acc{nullptr, partial[next++].data(), 12 * sizeof(uint32_t)};
        int rc = stage_begin(dev, {&dl, &dc, &acc}, part);
        if (rc) return rc;
        CK(cudaMemsetAsync(acc.dptr, 0, 12 * sizeof(uint32_t), part->res->k));
        if ((rc = check_sums_enqueue(seed2, chi_start + lo, (const gcb_label*)dl.dptr, choice ? (const uint8_t*)dc.dptr : nullptr, m,
                                     (uint32_t*)acc.dptr, part->res->k)))
            return rc;
        return stage_end({&acc}, part);
    });
    if (rc) return rc;
    for (const auto& w : partial) sums_from_words(w.data(), out);
    return GCB_OK;
    GCB_CATCH
}

}  // extern "C"

// ------------------------------------------- garbled tables in the wire format ------
gcb::DevWireLayout::~DevWireLayout() {
    if (tmpl) cudaFree(tmpl);
    if (row_pos) cudaFree(row_pos);
}
static size_t tables_wire_bytes(const gcb_plan* plan) {
    return 4 + 4 * (size_t)plan->p.info.num_gates + 16 * (size_t)plan->p.info.num_rows;
}
static void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static int wire_layout_on_device(const gcb_plan* plan, int device, std::shared_ptr<DevWireLayout>* out) {
    std::lock_guard<std::mutex> lk(plan->wire_mu);
    auto it = plan->wire_dev.find(device);
    if (it != plan->wire_dev.end()) { *out = it->second; return GCB_OK; }
    const Plan& pl = plan->p;
    const size_t total = tables_wire_bytes(plan);
    if (total > 0xfffffff0u) return fail(GCB_E_TOO_LARGE, "garbled tables exceed 4 GiB per instance");
    std::vector<uint8_t> tmpl(((total + 15) & ~(size_t)15) + 16, 0);
    std::vector<uint32_t> row_pos(pl.info.num_rows ? pl.info.num_rows : 1);
    put_be32(tmpl.data(), pl.info.num_gates);                      // garbler.go:69
    size_t pos = 4;
    for (uint32_t g = 0; g < pl.info.num_gates; g++) {
        const uint32_t r0 = pl.row_off[g], cnt = pl.row_off[g + 1] - r0;
        put_be32(tmpl.data() + pos, cnt);                          // garbler.go:74
        pos += 4;
        for (uint32_t k = 0; k < cnt; k++, pos += 16) row_pos[r0 + k] = (uint32_t)pos;
    }
    auto dl = std::make_shared<DevWireLayout>();
    CK(upload(&dl->tmpl, tmpl));
    CK(upload(&dl->row_pos, row_pos));
    plan->wire_dev[device] = dl;
    *out = dl;
    return GCB_OK;
}

extern "C" {

int gcb_tables_wire_size(const gcb_plan* plan, size_t* bytes) {
    GCB_TRY
    if (!plan || !bytes) return fail(GCB_E_ARG, "null argument");
    *bytes = tables_wire_bytes(plan);
    return GCB_OK;
    GCB_CATCH
}
int gcb_tables_to_wire_dev(const gcb_plan* plan, uint32_t batch, const gcb_label* tables, uint8_t* dst, size_t stride,
                           void* stream) {
    GCB_TRY
    if (!plan || (batch && (!dst || (!tables && plan->p.info.num_rows)))) return fail(GCB_E_ARG, "null argument");
    const size_t total = tables_wire_bytes(plan);
    if ((stride & 15) || stride < ((total + 15) & ~(size_t)15)) return fail(GCB_E_BUFFER, "wire stride must be a multiple of 16 and at least %zu", (total + 15) & ~(size_t)15);
    if (batch == 0) return GCB_OK;
    int rc = select_device(nullptr);
    if (rc) return rc;
    std::shared_ptr<DevWireLayout> dl;
    if ((rc = wire_layout_on_device(plan, tl_cur, &dl))) return rc;
    SerParams sp{dl->tmpl, (uint32_t)total, dl->row_pos, plan->p.info.num_rows, reinterpret_cast<const uint4*>(tables), dst, stride};
    const dim3 grid((unsigned)((total + SER_TILE - 1) / SER_TILE), batch);
    serialize_kernel<<<grid, SER_THREADS, 0, (cudaStream_t)stream>>>(sp);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}
int gcb_tables_from_wire_dev(const gcb_plan* plan, uint32_t batch, const uint8_t* src, size_t stride, gcb_label* tables,
                             void* stream) {
    GCB_TRY
    if (!plan || (batch && (!src || (!tables && plan->p.info.num_rows)))) return fail(GCB_E_ARG, "null argument");
    const size_t total = tables_wire_bytes(plan);
    if ((stride & 15) || stride < ((total + 15) & ~(size_t)15) + 16) return fail(GCB_E_BUFFER, "wire stride must be a multiple of 16 and at least %zu", ((total + 15) & ~(size_t)15) + 16);
    if (batch == 0 || plan->p.info.num_rows == 0) return GCB_OK;
    DeviceInfo* di;
    int rc = select_device(&di);
    if (rc) return rc;
    std::shared_ptr<DevWireLayout> dl;
    if ((rc = wire_layout_on_device(plan, tl_cur, &dl))) return rc;
    DeserParams dp{src, stride, dl->row_pos, plan->p.info.num_rows, reinterpret_cast<uint4*>(tables), batch};
    const size_t work = (size_t)batch * plan->p.info.num_rows;
    const unsigned blocks = (unsigned)std::min<size_t>((work + 255) / 256, (size_t)di->sm_count * 8);
    deserialize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dp);
    CK(launched());
    return GCB_OK;
    GCB_CATCH
}
int gcb_tables_to_wire(const gcb_plan* plan, uint32_t batch, const gcb_label* tables, uint8_t* dst, size_t stride) {
    GCB_TRY
    if (!plan || (batch && (!dst || (!tables && plan->p.info.num_rows)))) return fail(GCB_E_ARG, "null argument");
    const size_t total = tables_wire_bytes(plan);
    if (stride < total) return fail(GCB_E_BUFFER, "wire buffer too small: need %zu bytes per instance", total);
    if (batch == 0) return GCB_OK;
    const size_t rows = plan->p.info.num_rows, s16 = (total + 15) & ~(size_t)15;
    return fan_out(batch, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dt{rows ? tables + lo * rows : nullptr, nullptr, m * rows * 16}, dw{nullptr, nullptr, m * s16};
        int rc = stage_begin(dev, {&dt, &dw}, part);
        if (rc) return rc;
        if ((rc = gcb_tables_to_wire_dev(plan, (uint32_t)m, (const gcb_label*)dt.dptr, (uint8_t*)dw.dptr, s16, part->res->k))) return rc;
        // rows of `total` bytes at the caller's stride (pageable or pinned: the 2-D copy handles both)
        CK(cudaMemcpy2DAsync(dst + lo * stride, stride, dw.dptr, s16, total, m, cudaMemcpyDeviceToHost, part->res->k));
        CK(part->res->seal());
        return GCB_OK;
    });
    GCB_CATCH
}
int gcb_tables_from_wire(const gcb_plan* plan, uint32_t batch, const uint8_t* src, size_t stride, gcb_label* tables) {
    GCB_TRY
    if (!plan || (batch && (!src || (!tables && plan->p.info.num_rows)))) return fail(GCB_E_ARG, "null argument");
    const size_t total = tables_wire_bytes(plan);
    if (stride < total) return fail(GCB_E_BUFFER, "wire buffer too small: need %zu bytes per instance", total);
    // the counts the evaluator checks while it reads (evaluator.go:44-52, eval.go:55,87,103)
    const Plan& pl = plan->p;
    for (uint32_t b = 0; b < batch; b++) {
        const uint8_t* q = src + (size_t)b * stride;
        auto be32 = [](const uint8_t* x) { return ((uint32_t)x[0] << 24) | ((uint32_t)x[1] << 16) | ((uint32_t)x[2] << 8) | x[3]; };
        if (be32(q) != pl.info.num_gates)
            return fail(GCB_E_CORRUPT, "wrong number of gates: got %u, expected %u", be32(q), pl.info.num_gates);
        size_t pos = 4;
        for (uint32_t g = 0; g < pl.info.num_gates; g++) {
            const uint32_t cnt = pl.row_off[g + 1] - pl.row_off[g];
            if (be32(q + pos) != cnt)
                return fail(GCB_E_CORRUPT, "corrupted circuit: gate %u has %u garbled rows, expected %u", g, be32(q + pos), cnt);
            pos += 4 + (size_t)cnt * 16;
        }
    }
    if (batch == 0 || pl.info.num_rows == 0) return GCB_OK;
    const size_t rows = pl.info.num_rows, s16 = ((total + 15) & ~(size_t)15) + 16;
    return fan_out(batch, [&](int dev, uint64_t lo, uint64_t hi, JobPart* part) -> int {
        const size_t m = (size_t)(hi - lo);
        HostOp dw{nullptr, nullptr, m * s16}, dt{nullptr, tables + lo * rows, m * rows * 16};
        int rc = stage_begin(dev, {&dw, &dt}, part);
        if (rc) return rc;
        CK(cudaMemcpy2DAsync(dw.dptr, s16, src + lo * stride, stride, total, m, cudaMemcpyHostToDevice, part->res->k));
        if ((rc = gcb_tables_from_wire_dev(plan, (uint32_t)m, (const uint8_t*)dw.dptr, s16, (gcb_label*)dt.dptr, part->res->k))) return rc;
        return stage_end({&dt}, part);
    });
    GCB_CATCH
}

}  // extern "C"
