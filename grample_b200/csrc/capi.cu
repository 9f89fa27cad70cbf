// C ABI of grample_b200 (include/grample_b200.h): host orchestration around the kernels in
// kernels.cuh.  There is no CPU fallback: every compute entry needs a CUDA device and fails
// loudly without one.  Host-only models (device = -1) exist for schedule / blanket inspection.
#include "../../include/grample_b200.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "comm.hpp"
#include "host_model.hpp"
#include "kernels.cuh"
#if !defined(GB_PART) || GB_PART == 0 || GB_PART == 1
#include "bits.cuh"
#endif

// Build partition (see __graft_entry__.build): the log-sum-exp sweep kernels are ~70 heavy template instantiations, so the
// library is compiled as three translation units in parallel from this one source — GB_PART 1 = everything but the LSE
// kernels, 2 = the float32 LSE kernels, 3 = the float64 ones (0 / undefined = all in one unit).  The LSE launches sit
// behind four plain functions (gbh::lse_*), defined in the unit that instantiates their kernels.
#ifndef GB_PART
#define GB_PART 0
#endif
#define GB_MAIN (GB_PART == 0 || GB_PART == 1)
#define GB_LSE_F32 (GB_PART == 0 || GB_PART == 2)
#define GB_LSE_F64 (GB_PART == 0 || GB_PART == 3)

namespace {

thread_local std::string g_err;

#define CUDA_CHECK(expr)                                                                             \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            throw gb::Err(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);      \
    } while (0)

#define GB_TRY try {
#define GB_END                        \
    }                                 \
    catch (const std::exception& e) { \
        g_err = e.what();             \
        return 1;                     \
    }                                 \
    return 0;

template <typename T>
T* dev_upload(const std::vector<T>& h) {
    T* d = nullptr;
    size_t n = h.empty() ? 1 : h.size();
    CUDA_CHECK(cudaMalloc(&d, n * sizeof(T)));
    if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

constexpr int kMaxDevices = 64;  // per-device caches of launch configuration (indexed by device, checked in require_device)

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1)
        throw gb::Err("grample_b200: no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= n || device >= kMaxDevices) throw gb::Err("grample_b200: invalid device index " + std::to_string(device));
    CUDA_CHECK(cudaSetDevice(device));
}

int grid_for(int64_t items, int threads) {
    int64_t blocks = (items + threads - 1) / threads;
    const int64_t cap = 148 * 8;  // B200: 148 SMs x 8 resident 256-thread CTAs
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

std::mutex g_cfg_mu;             // guards those caches: handles on different threads may configure the same kernel

}  // namespace

// Thread safety ("every entry thread-safe per handle"): every entry point that takes a gb_chains* holds the
// handle's mutex for its whole duration, so goroutines / pthreads may call into ONE handle concurrently and the calls
// serialise in arrival order (stream order on the device follows).  A model's lazily built tables have their own lock.
#define GB_LOCK(c) std::lock_guard<std::recursive_mutex> _gb_lock((c)->mu)

struct gb_model {
    gb::HostModel h;
    int device = -1;
    std::vector<void*> allocs;
    gb::DevModel dev{};
    int32_t* d_order = nullptr;
    int32_t* d_colour_off = nullptr;
    gb::DevTab tab{};
    bool tab_built = false;
    std::recursive_mutex mu;  // lazily built threshold tables: a model is shared by the groups of several handles

    ~gb_model() {
        if (device >= 0) {
            cudaSetDevice(device);
            for (void* p : allocs) cudaFree(p);
        }
    }
    template <typename T>
    T* up(const std::vector<T>& v) {
        T* d = dev_upload(v);
        allocs.push_back(d);
        return d;
    }
    void upload(int dev_index) {
        device = dev_index;
        if (device < 0) return;
        require_device(device);
        std::vector<int32_t> entry_var(h.total_card);
        for (int v = 0; v < h.n_vars; v++)
            for (int k = 0; k < h.card[v]; k++) entry_var[h.card_off[v] + k] = v;
        // device copies of the tables are padded to a multiple of 4 entries (16-byte TMA bulk-copy granules)
        std::vector<double> tab64(h.log_tab);
        tab64.resize((tab64.size() + 3) & ~(size_t)3, 0.0);
        std::vector<float> tab32(tab64.begin(), tab64.end());
        dev.n_tab = (int32_t)tab64.size();
        dev.n_vars = h.n_vars;
        dev.total_card = h.total_card;
        dev.max_card = h.max_card;
        dev.card = up(h.card);
        dev.card_off = up(h.card_off);
        dev.fixed = up(h.fixed);
        dev.prog_off = up(h.prog_off);
        dev.prog = up(h.prog);
        dev.pw_off = up(h.pw_off);
        dev.pw_rec = reinterpret_cast<const int4*>(up(h.pw_rec));
        dev.tab64 = up(tab64);
        dev.tab32 = up(tab32);
        dev.entry_var = up(entry_var);
        d_order = up(h.order);
        std::vector<int32_t> pos_rec(std::max<size_t>(1, h.order.size()) * 4, 0);
        for (size_t j = 0; j < h.order.size(); j++) {
            const int v = h.order[j];
            pos_rec[4 * j] = v;
            pos_rec[4 * j + 1] = h.card[v];
            pos_rec[4 * j + 2] = h.pw_off[v];
            pos_rec[4 * j + 3] = h.prog[h.prog_off[v]];
        }
        dev.pos_rec = reinterpret_cast<const int4*>(up(pos_rec));
        d_colour_off = up(h.colour_off);
    }
    // thresholds: evaluate every (tabulated variable, neighbour configuration) conditional once on the device
    bool thr_built = false;
    void ensure_thresholds() {
        std::lock_guard<std::recursive_mutex> lk(mu);
        if (thr_built) return;
        if (device < 0) throw gb::Err("table / hybrid mode needs a device-resident model (there is no CPU fallback)");
        require_device(device);
        tab.tp_off = up(h.tp_off);
        tab.tprog = up(h.tprog);
        tab.n_order = (int32_t)h.order.size();
        tab.n_thr = (int32_t)h.n_thresholds;
        uint32_t* thr = nullptr;
        CUDA_CHECK(cudaMalloc(&thr, (size_t)std::max<int64_t>(h.n_thresholds, 1) * sizeof(uint32_t)));
        allocs.push_back(thr);
        tab.thr = thr;
        const int n = (int)h.order.size();
        if (h.n_tab_vars > 0) {
            gb::k_build_thresholds<<<std::max(1, std::min((n + 127) / 128, 148 * 16)), 128>>>(dev, tab, d_order, n);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaDeviceSynchronize());
        }
        thr_built = true;
    }
    // table mode proper: every sampled variable tabulated with <= 256 configurations, fixed-size records
    void ensure_bits() {
        if (!h.bits_ok) throw gb::Err("bit-sliced table mode does not apply to this model: " + h.bits_why);
        ensure_tab();
    }
    void ensure_tab() {
        std::lock_guard<std::recursive_mutex> lk(mu);
        if (!h.tab_ok) throw gb::Err("table mode does not apply to this model: " + h.tab_why);
        if (tab_built) return;
        ensure_thresholds();
        tab.trec = up(h.trec);
        tab_built = true;
    }
    // hybrid mode tabulates what qualifies (cardinality <= 4, <= 65536 configurations) in models whose cardinalities are <= 4
    bool hybrid_tables() const { return h.max_card <= 4 && h.n_tab_vars > 0; }
    // hybrid mode on a model where EVERY sampled variable got a table (plain binary models and their
    // single-collapsed variants): the resident table kernel runs it, wide variables through their tprog entry
    bool hybrid_all_tables() const { return hybrid_tables() && h.tab_all; }
    void ensure_hybrid() {
        std::lock_guard<std::recursive_mutex> lk(mu);
        if (!hybrid_tables()) return;
        ensure_thresholds();
        if (h.tab_all && !tab_built) {
            tab.trec = up(h.trec);
            tab_built = true;
        }
    }
};

namespace gbh {  // host-side types shared by the translation units

struct Group {
    gb_model* model = nullptr;
    bool owns_model = false;
    int32_t n_chains = 0, n_pad = 0;
    uint64_t first_chain = 0;
    uint32_t sweep = 0;  // next Philox sweep index
    int64_t total_samples = 0;  // Chain.TotalSampleCount summed over the group's chains
    uint64_t scan_step = 0;     // next random-scan step index
    bool window_filled = false; // the group has been through an AdvanceChain round (its half-window histograms are valid)
    uint8_t* d_state = nullptr;
    uint32_t* d_bits = nullptr;  // GB_TABLE_BITS: [n_vars][n_words] one bit per chain (d_state stays null)
    unsigned int* d_tile_ctr = nullptr;  // GB_TABLE_BITS: {next tile, CTAs done} of the dynamically scheduled sweep launches
    int32_t n_words = 0;
    unsigned long long* d_counts = nullptr;
    uint16_t* d_hist = nullptr;
    gb::DevGroup dev{};
};

// Resident-path launch plan: chains per CTA, dynamic shared memory, and whether the log-space tables are
// staged in shared memory by TMA bulk copies (ts).  ch == 0: the model does not qualify.
struct ResidentPlan {
    int ch = 0;
    size_t smem = 0;
    bool ts = false;
    int32_t n_stage = 0;  // table entries staged in shared memory: all of them (ts) or a prefix ending on a factor boundary
};

}  // namespace gbh
using gbh::Group;
using gbh::ResidentPlan;

struct gb_chains {
    std::recursive_mutex mu;  // GB_LOCK: one call at a time per handle
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t seed = 0;
    int precision = GB_F64;
    uint32_t flags = 0;
    std::vector<Group> groups;
    int64_t total_samples = 0;
    double* d_merge = nullptr;  // [total_card]
    double* d_wb = nullptr;     // [2*n_vars]
    uint8_t* d_skip = nullptr;  // [n_vars]
    double* d_merged_in = nullptr;
    int32_t last_cw = -1;       // ConvergenceWindow of the last gb_chains_advance
    int64_t launches = 0;       // kernels launched on behalf of this handle
    // merge path caches (rebuilt when a group is added)
    std::vector<uint8_t> col_any;      // collapsed in any group
    std::vector<int> col_first_group;  // first group (list order) in which a variable is collapsed
    std::vector<int> col_vars;         // indices of the collapsed-in-any variables
    bool skip_uploaded = false;
    std::vector<uint8_t> skip_bits;    // host copy of d_skip
    std::vector<int32_t> col_any32;    // col_any widened for the ABI's int32 output
    double* h_merge = nullptr;         // pinned staging buffer for the device -> host copy
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // interval path (cmd/root.go:498-539) off the sweep stream: integer count sums -> NCCL all-reduce -> finalize -> D2H
    // run on `merge_stream` behind a snapshot event, so the sweeps enqueued after gb_chains_merge_begin overlap with it
    gb_comm* comm = nullptr;                 // borrowed; nullptr or world == 1: no collective
    cudaStream_t merge_stream = nullptr;
    // Up to kMergeSlots merges are in flight at once (begin, begin, end, begin, end, ...): the host can enqueue the next
    // round's sweeps AND its snapshot before it waits for the previous round's result, so the device never runs dry even
    // when a merge takes about as long as a round (its NCCL kernel only gets SMs at a sweep kernel's boundary).
    static constexpr int kMergeSlots = 2;
    struct MergeSlot {
        cudaEvent_t ev_snap = nullptr, ev_done = nullptr;
        cudaEvent_t ev_t[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // timing marks (gb_chains_merge_timing)
        unsigned long long* d_cnt = nullptr;   // [total_card + 2]: summed counts, chain count, TotalSampleCount
        double* d_out = nullptr;               // [total_card] marginals on the device
        unsigned long long* h_tail = nullptr;  // pinned [2]: the (all-reduced) tail of d_cnt
        double* h_stage = nullptr;             // pinned staging buffer when the destination is pageable
        bool pending = false, ever = false, staged = false;
        bool needs_finalize = false;           // reduced counts are (or will be) in d_cnt; conversion + host copy not enqueued yet
        double* out = nullptr;                 // destination of the pending merge
        int32_t* col_out = nullptr;
    } slots[kMergeSlots];
    int slot_head = 0, slot_tail = 0, slots_pending = 0;  // next slot to begin / to end
    int last_done_slot = -1;
    bool merge_ever = false;
    int64_t global_chains = -1, global_samples = -1;  // tail of the last completed merge

    // groups are independent between monitor intervals (like the reference's goroutine per chain,
    // chain.go:197-215): their launches fan out over side streams and join back on `stream`
    std::vector<cudaStream_t> side;
    std::vector<cudaEvent_t> ev_join;
    cudaEvent_t ev_fork = nullptr;

    ~gb_chains() {
        cudaSetDevice(device);
        for (auto st : side) cudaStreamDestroy(st);
        for (auto e : ev_join) cudaEventDestroy(e);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& g : groups) {
            cudaFree(g.d_state);
            cudaFree(g.d_bits);
            cudaFree(g.d_tile_ctr);
            cudaFree(g.d_counts);
            cudaFree(g.d_hist);
            if (g.owns_model) delete g.model;
        }
        cudaFree(d_merge);
        cudaFree(d_wb);
        cudaFree(d_skip);
        cudaFree(d_merged_in);
        for (auto& sl : slots) {
            cudaFree(sl.d_cnt);
            cudaFree(sl.d_out);
            if (sl.h_tail) cudaFreeHost(sl.h_tail);
            if (sl.h_stage) cudaFreeHost(sl.h_stage);
            if (sl.ev_snap) cudaEventDestroy(sl.ev_snap);
            if (sl.ev_done) cudaEventDestroy(sl.ev_done);
            for (auto e : sl.ev_t)
                if (e) cudaEventDestroy(e);
        }
        if (merge_stream) cudaStreamDestroy(merge_stream);
        if (h_merge) cudaFreeHost(h_merge);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    const gb::HostModel& base() const { return groups[0].model->h; }
};

// The L1 / shared-memory split of a launch is the driver's choice; left alone it sizes the carve-out for the MOST CTAs that
// could fit by shared memory (233 KB for a resident table kernel with histograms) even when two CTAs per SM exist, and the
// records and thresholds then miss in what is left of L1 (ncu: L1 hit rate 99 % -> 69 %, long-scoreboard stall 0.6 -> 3.1
// per issue).  Ask for the carve-out that holds the CTAs each SM really hosts — counted over all groups of the handle,
// whose kernels run concurrently — and leave the rest to L1.  Table kernel only: the log-sum-exp resident kernels keep their
// tables in shared memory and measured slower with the hint (ObjectDetection_11 f64 33.8 -> 39.5 us/sweep).
template <typename K>
void prefer_carveout(K kernel, const gb_chains* c, int ch, size_t smem_per_cta) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    int64_t ctas = 0;
    for (const auto& gg : c->groups) ctas += (gg.n_pad + ch - 1) / ch;
    const int64_t per_sm = (ctas + sms - 1) / sms + 1;  // one spare: block scheduling is not perfectly even
    const int64_t pct = std::min<int64_t>(100, (100 * per_sm * (int64_t)(smem_per_cta + 1024) + 228 * 1024 - 1) / (228 * 1024));
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::max<int64_t>(pct, 8));
}

namespace gbh {
// one colour of one group / n_sweeps sweeps of one group on the resident path, log-sum-exp kernels of one precision
void lse_colour_f32(gb_chains* c, Group& g, const int32_t* d_vars, int32_t n, int record, int hist_half);
void lse_colour_f64(gb_chains* c, Group& g, const int32_t* d_vars, int32_t n, int record, int hist_half);
void lse_resident_f32(gb_chains* c, Group& g, const ResidentPlan& p, int32_t n_sweeps, int record, int32_t n_pre, int32_t n_half);
void lse_resident_f64(gb_chains* c, Group& g, const ResidentPlan& p, int32_t n_sweeps, int record, int32_t n_pre, int32_t n_half);
}  // namespace gbh

namespace {

#if GB_MAIN
void add_group(gb_chains* c, gb_model* model, int32_t n_chains, uint64_t first_chain, bool owns) {
    if (!model) throw gb::Err("No model supplied");
    if (model->device != c->device) throw gb::Err("model lives on a different device than the chains");
    if (n_chains < 1) throw gb::Err("a chain group needs at least 1 chain");
    if (first_chain % 8) throw gb::Err("first_chain_id must be a multiple of 8 (chains share Philox calls in blocks of 8)");
    const bool bits = c->precision == GB_TABLE_BITS;
    if (bits && first_chain % 32) throw gb::Err("first_chain_id must be a multiple of 32 under GB_TABLE_BITS (one state word holds 32 chains)");
    if (bits && (c->flags & GB_CHAINS_HISTORY)) throw gb::Err("GB_TABLE_BITS keeps no per-chain histories (no GB_CHAINS_HISTORY): use GB_TABLE");
    if (c->precision == GB_TABLE) model->ensure_tab();
    if (bits) model->ensure_bits();
    if (c->precision == GB_HYBRID) model->ensure_hybrid();
    if (!c->groups.empty() && (model->h.n_vars != c->base().n_vars || model->h.card != c->base().card))
        throw gb::Err("Cannot merge chain with different variables");
    if (model->h.order.empty()) throw gb::Err("No Variables to select");
    Group g;
    g.model = model;
    g.owns_model = owns;
    g.n_chains = n_chains;
    g.n_pad = (n_chains + 7) / 8 * 8;
    g.first_chain = first_chain;
    const gb::HostModel& h = model->h;
    g.n_words = (g.n_chains + 31) / 32;
    try {
        if (bits) {
            CUDA_CHECK(cudaMalloc(&g.d_bits, (size_t)h.n_vars * g.n_words * sizeof(uint32_t)));
            CUDA_CHECK(cudaMalloc(&g.d_tile_ctr, 2 * sizeof(unsigned int)));
            CUDA_CHECK(cudaMemsetAsync(g.d_tile_ctr, 0, 2 * sizeof(unsigned int), c->stream));
        } else {
            CUDA_CHECK(cudaMalloc(&g.d_state, (size_t)h.n_vars * g.n_pad));
        }
        CUDA_CHECK(cudaMalloc(&g.d_counts, (size_t)h.total_card * sizeof(unsigned long long)));
        CUDA_CHECK(cudaMemsetAsync(g.d_counts, 0, (size_t)h.total_card * sizeof(unsigned long long), c->stream));
        if (c->flags & GB_CHAINS_HISTORY) {
            size_t hb = (size_t)2 * h.total_card * g.n_pad * sizeof(uint16_t);
            CUDA_CHECK(cudaMalloc(&g.d_hist, hb));
            CUDA_CHECK(cudaMemsetAsync(g.d_hist, 0, hb, c->stream));
        }
    } catch (...) {  // a failed allocation must not leak the ones before it
        cudaFree(g.d_state);
        cudaFree(g.d_bits);
        cudaFree(g.d_tile_ctr);
        cudaFree(g.d_counts);
        cudaFree(g.d_hist);
        throw;
    }
    g.dev.state = g.d_state;
    g.dev.counts = g.d_counts;
    g.dev.hist = g.d_hist;
    g.dev.n_chains = g.n_chains;
    g.dev.n_pad = g.n_pad;
    g.dev.first_chain = first_chain;
    g.dev.seed_lo = (uint32_t)c->seed;
    g.dev.seed_hi = (uint32_t)(c->seed >> 32);
    g.dev.rb = (c->flags & GB_CHAINS_RAO_BLACKWELL) ? 1 : 0;
    if (bits) {
        gb::k_init_bits<<<grid_for((int64_t)h.n_vars * g.n_words, 256), 256, 0, c->stream>>>(model->dev, g.dev, g.d_bits, g.n_words);
    } else {
        const int64_t items = (int64_t)h.n_vars * (g.n_pad / 4);
        gb::k_init_state<<<grid_for(items, 256), 256, 0, c->stream>>>(model->dev, g.dev);
    }
    c->launches++;
    CUDA_CHECK(cudaGetLastError());
    c->groups.push_back(g);
    c->col_any.clear();  // invalidate the merge caches
    c->skip_uploaded = false;
}

#endif  // GB_MAIN

template <typename Real, int MAXC, int CW>
void launch_colour(gb_chains* c, Group& g, const int32_t* d_vars, int32_t n, int record, int hist_half) {
    const int64_t items = (int64_t)n * (g.n_pad / 4);
    const int hybrid = c->precision == GB_HYBRID && g.model->hybrid_tables();
    if (g.dev.rb)
        gb::k_sweep_colour<Real, MAXC, CW, true><<<grid_for(items, 256), 256, 0, c->stream>>>(g.model->dev, g.dev, d_vars, n, g.sweep,
                                                                                             record, hist_half, g.model->tab, hybrid);
    else
        gb::k_sweep_colour<Real, MAXC, CW, false><<<grid_for(items, 256), 256, 0, c->stream>>>(g.model->dev, g.dev, d_vars, n, g.sweep,
                                                                                              record, hist_half, g.model->tab, hybrid);
    c->launches++;
}

#if GB_MAIN
template <int VB, int NN, bool HIST, int PF>
void launch_tab_variant(gb_chains* c, Group& g, int col, int32_t n, int record, int hist_half) {
    static int resident_dev[kMaxDevices] = {};  // per device: CTAs that fit at once (persistent tile loop)
    std::lock_guard<std::mutex> cfg_lock(g_cfg_mu);
    int& resident = resident_dev[c->device];
    constexpr size_t ring = (size_t)(PF > 0 ? (PF + 1) * NN * 256 * 8 : 0);  // cp.async prefetch ring
    if (!resident) {
        int per_sm = 0, sms = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (ring > 0)
            CUDA_CHECK(cudaFuncSetAttribute(gb::k_sweep_tab<VB, NN, HIST, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gb::k_sweep_tab<VB, NN, HIST, PF>, 256, ring));
        resident = std::max(1, per_sm * sms);
    }
    const gb::HostModel& h = g.model->h;
    const int64_t tiles = (int64_t)((g.n_pad / 8 + 255) / 256) * ((n + VB - 1) / VB);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, resident));
    gb::k_sweep_tab<VB, NN, HIST, PF><<<grid, 256, ring, c->stream>>>(g.model->dev, g.model->tab, g.dev, h.colour_off[col], n,
                                                                     g.sweep, record, hist_half);
    c->launches++;
}

void launch_tab(gb_chains* c, Group& g, int col, int32_t n, int record, int hist_half) {
    if (g.dev.rb)
        throw gb::Err("GB_CHAINS_RAO_BLACKWELL under GB_TABLE needs a model small enough for the shared-memory-resident table kernel "
                      "(the per-colour table kernel keeps only threshold halves): use GB_HYBRID");
    const bool hist = g.d_hist != nullptr && hist_half >= 0;
    // GB_TAB_PF=0 selects the register-prefetch variant (kept for A/B measurements; default: cp.async ring, depth 3)
    static const bool reg_prefetch = std::getenv("GB_TAB_PF") && std::atoi(std::getenv("GB_TAB_PF")) == 0;
    constexpr int VB = gb::kTabTile;
    if (g.model->h.tab_max_nbr > 4) {
        if (hist) launch_tab_variant<32, 8, true, 3>(c, g, col, n, record, hist_half);
        else launch_tab_variant<32, 8, false, 3>(c, g, col, n, record, hist_half);
    } else if (hist) {
        launch_tab_variant<VB, 4, true, 3>(c, g, col, n, record, hist_half);
    } else if (reg_prefetch) {
        launch_tab_variant<VB, 4, false, 0>(c, g, col, n, record, hist_half);
    } else {
        launch_tab_variant<VB, 4, false, 3>(c, g, col, n, record, hist_half);
    }
}

// GB_TABLE_BITS: one colour of one group on bit-packed state; W = state words per thread, NT = threads per CTA
template <int W, int NT, int S = 1>
void launch_bits_w(gb_chains* c, Group& g, int col, int32_t n, int record) {
    static int resident_dev[kMaxDevices] = {};  // per device: CTAs that fit at once (persistent CTAs, tiles handed out by an atomic counter)
    int resident;
    {
        std::lock_guard<std::mutex> cfg_lock(g_cfg_mu);
        int& r = resident_dev[c->device];
        if (!r) {
            int per_sm = 0, sms = 0;
            CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gb::k_sweep_bits<W, NT, S>, NT, 0));
            r = std::max(1, per_sm * sms);
        }
        resident = r;
    }
    // the group's counter pair {next tile, CTAs done}: zero at every launch because the last CTA of the previous launch
    // re-armed it, and a group's launches are ordered (one stream at a time, joined to the handle's stream in between)
    const gb::HostModel& h = g.model->h;
    constexpr int chunk_words = NT * W / S;
    const int64_t tiles = (int64_t)((g.n_words + chunk_words - 1) / chunk_words) * ((n + gb::kBitsVB - 1) / gb::kBitsVB);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, resident));
    gb::k_sweep_bits<W, NT, S><<<grid, NT, 0, c->stream>>>(g.model->dev, g.model->tab, g.dev, g.d_bits, g.n_words, h.colour_off[col], n, g.sweep,
                                                         record, g.d_tile_ctr, gb::philox_keys(g.dev.seed_lo, g.dev.seed_hi));
    c->launches++;
}
// CTA shape by population: two words per thread amortise the warp-uniform coefficient reads; a chunk must not exceed the
// row, or threads idle — 8192 chains per GPU (the 8-way split of 65536) are 256 words = one chunk of the split shape
// (256 threads x 2 words, the CTA's halves taking half of the tile's positions each)
void launch_bits(gb_chains* c, Group& g, int col, int32_t n, int record) {
    const char* env = std::getenv("GB_BITS_SHAPE");  // A/B and test knob: "W,NT" or "W,NT,S"
    int w = g.n_words >= 256 ? 2 : 1, nt = g.n_words >= 256 ? 256 : 128, sp = (g.n_words >= 256 && g.n_words < 512) ? 2 : 1;
    if (env && std::strlen(env) >= 3) {
        w = env[0] - '0';
        nt = std::atoi(env + 2);
        sp = std::strlen(env) >= 7 ? env[6] - '0' : 1;
    }
    if (w == 2 && nt == 256 && sp == 2) launch_bits_w<2, 256, 2>(c, g, col, n, record);
    else if (w == 2 && nt == 256) launch_bits_w<2, 256>(c, g, col, n, record);
    else if (w == 2) launch_bits_w<2, 128>(c, g, col, n, record);
    else if (nt == 256) launch_bits_w<1, 256>(c, g, col, n, record);
    else launch_bits_w<1, 128>(c, g, col, n, record);
}

void flush_finalizes(gb_chains* c);  // (merge path, below)

// one sweep of one group: one launch per colour
void sweep_group(gb_chains* c, Group& g, int record, int hist_half) {
    const gb::HostModel& h = g.model->h;
    const int n_col = (int)h.colour_off.size() - 1;
    for (int col = 0; col < n_col; col++) {
        const int32_t* dv = g.model->d_order + h.colour_off[col];
        const int32_t n = h.colour_off[col + 1] - h.colour_off[col];
        if (n == 0) continue;
        if (c->precision == GB_TABLE) {
            launch_tab(c, g, col, n, record, hist_half);
        } else if (c->precision == GB_TABLE_BITS) {
            launch_bits(c, g, col, n, record);
        } else if (c->precision == GB_F32) {
            gbh::lse_colour_f32(c, g, dv, n, record, hist_half);
        } else {
            gbh::lse_colour_f64(c, g, dv, n, record, hist_half);
        }
        flush_finalizes(c);  // a pending merge's conversion goes behind this kernel; its NCCL kernel overlaps with it
    }
    g.sweep++;
    if (record) {
        c->total_samples += (int64_t)h.order.size() * g.n_chains;
        g.total_samples += (int64_t)h.order.size() * g.n_chains;
    }
}

#endif  // GB_MAIN

// Shared-memory-resident path for small models: all sweeps of one group in ONE launch.
template <typename Real, int MAXC, int CW, bool TS, bool RB>
void launch_resident_rb(gb_chains* c, Group& g, int ch, size_t smem, int32_t hist_off, int32_t n_sweeps, int record, int32_t n_pre,
                        int32_t n_half, int32_t n_stage) {
    // block size >= work items of the largest colour of one CTA (one item per thread keeps the
    // per-colour critical path at a single update), capped at 256
    const gb::HostModel& hm = g.model->h;
    int max_col = 1;
    for (size_t i = 0; i + 1 < hm.colour_off.size(); i++) max_col = std::max(max_col, hm.colour_off[i + 1] - hm.colour_off[i]);
    // the largest colour's items are spread evenly over the fewest passes a CTA of <= cap threads needs (whole warps).
    // One thread per chain (CW == 0, cardinality >= 8): a warp holds 32 chains of one variable, and these models have
    // few variables per colour, so the cap is 512 threads to keep a colour at one pass.
    // (Quad kernels keep 256 threads and a short last pass: balancing their passes measured slower.)
    const int64_t items = (int64_t)max_col * (CW == 0 ? ch : ch / 4);
    const int64_t passes = (items + 511) / 512;
    const int threads = CW == 0 ? (int)(((items + passes - 1) / passes + 31) / 32 * 32) : (int)std::min<int64_t>(256, (items + 31) / 32 * 32);
    static size_t configured_dev[kMaxDevices] = {};  // function attributes are per device
    std::lock_guard<std::mutex> cfg_lock(g_cfg_mu);
    size_t& configured = configured_dev[c->device];
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gb::k_sweep_resident<Real, MAXC, CW, TS, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const gb::HostModel& h = g.model->h;
    gb::k_sweep_resident<Real, MAXC, CW, TS, RB><<<g.n_pad / ch, threads, smem, c->stream>>>(
        g.model->dev, g.dev, g.model->d_order, g.model->d_colour_off, (int32_t)h.colour_off.size() - 1, ch, g.sweep, n_sweeps,
        record, n_pre, n_half, g.model->tab, (int)(c->precision == GB_HYBRID && g.model->hybrid_tables()), hist_off, n_stage);
    c->launches++;
}

template <typename Real, int MAXC, int CW, bool TS>
void launch_resident_ts(gb_chains* c, Group& g, int ch, size_t smem, int32_t hist_off, int32_t n_sweeps, int record, int32_t n_pre,
                        int32_t n_half, int32_t n_stage) {
    if (g.dev.rb) launch_resident_rb<Real, MAXC, CW, TS, true>(c, g, ch, smem, hist_off, n_sweeps, record, n_pre, n_half, n_stage);
    else launch_resident_rb<Real, MAXC, CW, TS, false>(c, g, ch, smem, hist_off, n_sweeps, record, n_pre, n_half, n_stage);
}

bool tab_resident(const gb_chains* c, const Group& g);

// Where the CTA's per-chain half-window histograms ([2][total_card][ch] u16) go for a launch that records
// them: appended to the resident layout when that still fits (returns the byte offset and grows *smem),
// else -1 = updated in global memory.
int32_t place_histograms(const gb_chains* c, const Group& g, const ResidentPlan& p, int32_t n_half, size_t* smem) {
    *smem = p.smem;
    if (!(c->flags & GB_CHAINS_HISTORY) || n_half < 0 || !g.d_hist) return -1;
    if (std::getenv("GB_HIST_GLOBAL")) return -1;  // A/B knob: histograms updated in global memory
    const size_t off = (p.smem + 15) & ~(size_t)15;
    // the resident table kernel keeps only the ones of its (binary) variables, by sweep position: [2][n_order][ch];
    // the others [2][total_card][ch]
    const size_t rows = tab_resident(c, g) ? g.model->h.order.size() : (size_t)g.model->h.total_card;
    const size_t bytes = (size_t)2 * rows * p.ch * sizeof(uint16_t);
    if (off + bytes > (p.ts ? 160 : 100) * 1024) return -1;
    *smem = off + bytes;
    return (int32_t)off;
}

// the resident table kernel runs table mode, and hybrid mode when every sampled variable has a table
bool tab_resident(const gb_chains* c, const Group& g) {
    const bool no_hy = std::getenv("GB_HYBRID_NO_TAB_KERNEL") != nullptr;  // A/B and test knob: hybrid stays on the LSE kernels
    return c->precision == GB_TABLE || (c->precision == GB_HYBRID && !no_hy && g.model->hybrid_all_tables());
}

#if GB_MAIN
ResidentPlan resident_plan(const gb_chains* c, const Group& g) {
    static const int disabled = std::getenv("GB_NO_RESIDENT") ? 1 : 0;
    static const int no_ts = std::getenv("GB_NO_SMEM_TABLES") ? 1 : 0;  // A/B knob (function-local statics initialise thread-safely)
    ResidentPlan p;
    if (disabled || (c->flags & GB_CHAINS_PER_COLOUR) || c->precision == GB_TABLE_BITS) return p;
    const gb::HostModel& h = g.model->h;
    if (h.n_vars > 4096) return p;
    auto base = [&](int ch) { return (((size_t)h.n_vars * ch + 15) & ~(size_t)15) + (size_t)h.total_card * 4; };
    if (tab_resident(c, g)) {  // units of 8 chains; thresholds and records stay in L1
        int sms_t = 148;
        cudaDeviceGetAttribute(&sms_t, cudaDevAttrMultiProcessorCount, c->device);
        for (int ch : {8, 16, 32, 64}) {
            if (g.n_pad % ch) continue;
            if (ch < 64 && g.n_pad / ch > sms_t * 16) continue;
            if (base(ch) > 100 * 1024) break;
            p.ch = ch; p.smem = base(ch);
            return p;
        }
        return p;
    }
    const size_t real_bytes = c->precision == GB_F32 ? 4 : 8;
    const size_t tab_bytes = 16 + (((size_t)h.log_tab.size() + 3) & ~(size_t)3) * real_bytes;
    // the groups of a handle run concurrently (side streams), so co-residency is judged on the CTAs of ALL of them
    auto ctas = [&](int ch) {
        int64_t n = 0;
        for (const auto& gg : c->groups) n += (gg.n_pad + ch - 1) / ch;
        return std::max<int64_t>(n, g.n_pad / ch);
    };
    constexpr size_t kSmemPerSm = 220 * 1024, kSmemPerCta = 200 * 1024;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    // 1) tables staged in shared memory.  Preferably with every CTA of the launch co-resident (these models expose little
    //    parallelism per colour, so chains per CTA grow only as far as that needs); a population too large for that runs
    //    64 chains per CTA in several waves — a CTA carries its chains through the whole round on its own, so waves cost
    //    only the partly filled last one, far less than reading the tables through L1 (BASELINE configs[3] at 32768
    //    chains per GPU: 217 -> us per sweep)
    if (!no_ts) {
        ResidentPlan waves;
        for (int ch : {8, 16, 32, 64}) {
            if (g.n_pad % ch) continue;
            const size_t smem = base(ch) + tab_bytes;
            if (smem > kSmemPerCta) break;
            const int64_t per_sm = std::min<int64_t>(8, (int64_t)(kSmemPerSm / (smem + 1024)));
            waves.ch = ch; waves.smem = smem; waves.ts = true;
            waves.n_stage = (int32_t)(((size_t)h.log_tab.size() + 3) & ~(size_t)3);
            if (ch < 64 && ctas(ch) > per_sm * sms) continue;
            if ((int64_t)(g.n_pad / ch) > per_sm * sms) break;
            return waves;
        }
        if (waves.ch && ctas(waves.ch) >= 2 * sms) return waves;  // (at least a couple of CTAs per SM: not a tiny launch that merely failed to fit)
    }
    // 1b) one-thread-per-chain kernels (cardinality >= 8): stage the longest PREFIX of the tables that fits and ends on
    //     a factor boundary.  A collapsed variant keeps the model's small factors first and appends the large factor
    //     over the collapsed variable's blanket (up to 11^6 entries on ObjectDetection_11), so every small factor is
    //     served from shared memory and only the large one goes through L1/L2.
    if (!no_ts && h.max_card >= 8) {
        ResidentPlan waves;
        for (int ch : {8, 16, 32, 64}) {
            if (g.n_pad % ch) continue;
            if (base(ch) + 4096 > kSmemPerCta) break;
            const size_t budget = kSmemPerCta - base(ch) - 32;
            int64_t n = 0;
            for (const auto& f : h.funcs) {  // factors are laid out in this order
                if ((size_t)((f.off + f.size + 3) & ~(int64_t)3) * real_bytes > budget) break;
                n = f.off + f.size;
            }
            if (n == 0) break;
            const size_t smem = base(ch) + 16 + (size_t)((n + 3) & ~(int64_t)3) * real_bytes;
            const int64_t per_sm = std::min<int64_t>(8, (int64_t)(kSmemPerSm / (smem + 1024)));
            waves.ch = ch; waves.smem = smem; waves.ts = false; waves.n_stage = (int32_t)n;
            if (ch < 64 && ctas(ch) > per_sm * sms) continue;
            if ((int64_t)(g.n_pad / ch) > per_sm * sms) break;
            return waves;
        }
        if (waves.ch && ctas(waves.ch) >= 2 * sms) return waves;
    }
    // 2) tables through L1: few chains per CTA = many CTAs = better SM fill and latency hiding; grow the
    //    CTA's chain count only when that would exceed ~16 CTAs per SM
    for (int ch : {8, 16, 32}) {
        if (g.n_pad % ch) continue;
        if (ch < 32 && g.n_pad / ch > sms * 16) continue;
        const size_t smem = base(ch);
        if (smem > 100 * 1024) continue;
        p.ch = ch; p.smem = smem;
        return p;
    }
    return p;
}

#endif  // GB_MAIN

template <typename Real, int MAXC, int CW>
void launch_resident(gb_chains* c, Group& g, const ResidentPlan& p, int32_t n_sweeps, int record, int32_t n_pre, int32_t n_half) {
    size_t smem = 0;
    const int32_t hist_off = place_histograms(c, g, p, n_half, &smem);
    if (p.ts) launch_resident_ts<Real, MAXC, CW, true>(c, g, p.ch, smem, hist_off, n_sweeps, record, n_pre, n_half, p.n_stage);
    else launch_resident_ts<Real, MAXC, CW, false>(c, g, p.ch, smem, hist_off, n_sweeps, record, n_pre, n_half, p.n_stage);
}

}  // namespace

namespace gbh {
#if GB_LSE_F32
void lse_colour_f32(gb_chains* c, Group& g, const int32_t* dv, int32_t n, int record, int hist_half) {
    const int mc = g.model->h.max_card;
    if (mc <= 2) launch_colour<float, 2, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 4) launch_colour<float, 4, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 8) launch_colour<float, 8, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 16) launch_colour<float, 16, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 32) launch_colour<float, 32, 1>(c, g, dv, n, record, hist_half);
    else launch_colour<float, 64, 1>(c, g, dv, n, record, hist_half);
}
void lse_resident_f32(gb_chains* c, Group& g, const ResidentPlan& plan, int32_t ns, int record, int32_t pre, int32_t n_half) {
    const int mc = g.model->h.max_card;
    if (mc <= 2) launch_resident<float, 2, 4>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 4) launch_resident<float, 4, 4>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 8) launch_resident<float, 8, 0>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 16) launch_resident<float, 16, 0>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 32) launch_resident<float, 32, 0>(c, g, plan, ns, record, pre, n_half);
    else launch_resident<float, 64, 0>(c, g, plan, ns, record, pre, n_half);
}
#endif
#if GB_LSE_F64
void lse_colour_f64(gb_chains* c, Group& g, const int32_t* dv, int32_t n, int record, int hist_half) {
    const int mc = g.model->h.max_card;
    if (mc <= 2) launch_colour<double, 2, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 4) launch_colour<double, 4, 4>(c, g, dv, n, record, hist_half);
    else if (mc <= 8) launch_colour<double, 8, 1>(c, g, dv, n, record, hist_half);
    else if (mc <= 16) launch_colour<double, 16, 1>(c, g, dv, n, record, hist_half);
    else if (mc <= 32) launch_colour<double, 32, 1>(c, g, dv, n, record, hist_half);
    else launch_colour<double, 64, 1>(c, g, dv, n, record, hist_half);
}
void lse_resident_f64(gb_chains* c, Group& g, const ResidentPlan& plan, int32_t ns, int record, int32_t pre, int32_t n_half) {
    const int mc = g.model->h.max_card;
    if (mc <= 2) launch_resident<double, 2, 4>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 4) launch_resident<double, 4, 4>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 8) launch_resident<double, 8, 0>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 16) launch_resident<double, 16, 0>(c, g, plan, ns, record, pre, n_half);
    else if (mc <= 32) launch_resident<double, 32, 0>(c, g, plan, ns, record, pre, n_half);
    else launch_resident<double, 64, 0>(c, g, plan, ns, record, pre, n_half);
}
#endif
}  // namespace gbh

#if GB_MAIN
namespace {

void launch_tab_resident(gb_chains* c, Group& g, const ResidentPlan& p, int32_t n_sweeps, int record, int32_t n_pre, int32_t n_half) {
    const gb::HostModel& h = g.model->h;
    int max_col = 1;
    for (size_t i = 0; i + 1 < h.colour_off.size(); i++) max_col = std::max(max_col, h.colour_off[i + 1] - h.colour_off[i]);
    const int64_t items = (int64_t)max_col * (p.ch / 8);
    const int threads = (int)std::min<int64_t>(256, (items + 31) / 32 * 32);  // whole warps, no idle ones at the colour barrier
    size_t smem = 0;
    const int32_t hist_off = place_histograms(c, g, p, n_half, &smem);
    static size_t configured_dev[8][kMaxDevices] = {};  // function attributes are per device
    const bool wide = h.tab_max_nbr > 8, multi = !h.tab_all_binary, rb = g.dev.rb != 0;
    std::lock_guard<std::mutex> cfg_lock(g_cfg_mu);
    size_t& configured = configured_dev[4 * rb + 2 * multi + wide][c->device];
    using Kernel = void (*)(const gb::DevModel, const gb::DevTab, const gb::DevGroup, const int32_t*, const int32_t, const int32_t,
                            const uint32_t, const int32_t, const int, const int32_t, const int32_t, const int32_t);
    static const Kernel kernels[8] = {gb::k_sweep_tab_resident<false, false, false>, gb::k_sweep_tab_resident<true, false, false>,
                                      gb::k_sweep_tab_resident<false, true, false>,  gb::k_sweep_tab_resident<true, true, false>,
                                      gb::k_sweep_tab_resident<false, false, true>,  gb::k_sweep_tab_resident<true, false, true>,
                                      gb::k_sweep_tab_resident<false, true, true>,   gb::k_sweep_tab_resident<true, true, true>};
    const Kernel kernel = kernels[4 * rb + 2 * multi + wide];
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    prefer_carveout(kernel, c, p.ch, smem);
    kernel<<<g.n_pad / p.ch, threads, smem, c->stream>>>(g.model->dev, g.model->tab, g.dev, g.model->d_colour_off,
                                                         (int32_t)h.colour_off.size() - 1, p.ch, g.sweep, n_sweeps, record, n_pre, n_half,
                                                         hist_off);
    c->launches++;
}

// n_sweeps sweeps of one group with the AdvanceChain window schedule (n_half < 0: no histograms)
void run_group(gb_chains* c, Group& g, int64_t n_sweeps, int record, int32_t n_pre, int32_t n_half) {
    if (n_sweeps <= 0) return;
    if (!(c->flags & GB_CHAINS_HISTORY)) n_half = -1;
    const ResidentPlan plan = resident_plan(c, g);
    const gb::HostModel& h = g.model->h;
    if (plan.ch) {
        flush_finalizes(c);  // one launch runs the whole round: a pending merge's conversion goes in front of it
        // One launch runs at most kMaxSweepsPerLaunch sweeps: the CTA's shared-memory counters are 32-bit
        // (sweeps x chains per CTA must stay below 2^31) and a single kernel should not run for minutes.
        // A later launch continues the window schedule: its n_pre is shifted (it may go negative).
        int64_t kMaxSweepsPerLaunch = std::min<int64_t>(1 << 20, ((int64_t)1 << 31) / plan.ch - 1);
        if (const char* e = std::getenv("GB_MAX_SWEEPS_PER_LAUNCH")) kMaxSweepsPerLaunch = std::max(1, std::atoi(e));  // test knob
        for (int64_t s0 = 0; s0 < n_sweeps; s0 += kMaxSweepsPerLaunch) {
            const int32_t ns = (int32_t)std::min<int64_t>(kMaxSweepsPerLaunch, n_sweeps - s0);
            const int32_t pre = (int32_t)std::max<int64_t>((int64_t)n_pre - s0, -(1ll << 30));
            if (tab_resident(c, g)) {
                launch_tab_resident(c, g, plan, ns, record, pre, n_half);
            } else if (c->precision == GB_F32) {
                gbh::lse_resident_f32(c, g, plan, ns, record, pre, n_half);
            } else {
                gbh::lse_resident_f64(c, g, plan, ns, record, pre, n_half);
            }
            g.sweep += (uint32_t)ns;
        }
        if (record) {
            c->total_samples += n_sweeps * (int64_t)h.order.size() * g.n_chains;
            g.total_samples += n_sweeps * (int64_t)h.order.size() * g.n_chains;
        }
        return;
    }
    for (int64_t s = 0; s < n_sweeps; s++) {
        const int hist_half = (n_half < 0 || s < n_pre) ? -1 : (s < (int64_t)n_pre + n_half ? 0 : 1);
        sweep_group(c, g, record, hist_half);
    }
}

// Run f(group) for every group with the groups' kernels on different streams (up to 128 at once): a
// variant's group is a few hundred chains, far too few to fill the device alone.  Everything a group's
// kernels touch (state, counts, histograms) is private to the group.  Fork: the side streams wait for
// what is already queued on the main stream; join: the main stream waits for every side stream.
static const size_t kSideStreams = std::getenv("GB_SIDE_STREAMS") ? (size_t)std::atoi(std::getenv("GB_SIDE_STREAMS")) : 128;  // = the device's limit of concurrently resident kernels
template <typename F>
void for_each_group_concurrent(gb_chains* c, F&& f) {
    static const bool serial = std::getenv("GB_SERIAL_GROUPS") != nullptr;  // A/B knob
    if (c->groups.size() < 2 || serial) {
        for (auto& g : c->groups) f(g);
        return;
    }
    const size_t K = std::min(kSideStreams, c->groups.size());
    while (c->side.size() < K) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        c->side.push_back(st);
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->ev_join.push_back(ev);
    }
    if (!c->ev_fork) CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    cudaStream_t main_stream = c->stream;
    CUDA_CHECK(cudaEventRecord(c->ev_fork, main_stream));
    for (size_t k = 0; k < K; k++) CUDA_CHECK(cudaStreamWaitEvent(c->side[k], c->ev_fork, 0));
    try {
        for (size_t i = 0; i < c->groups.size(); i++) {
            c->stream = c->side[i % K];
            f(c->groups[i]);
        }
    } catch (...) {
        c->stream = main_stream;
        throw;
    }
    c->stream = main_stream;
    for (size_t k = 0; k < K; k++) {
        CUDA_CHECK(cudaEventRecord(c->ev_join[k], c->side[k]));
        CUDA_CHECK(cudaStreamWaitEvent(main_stream, c->ev_join[k], 0));
    }
}

void sweeps(gb_chains* c, int64_t n, int record) {
    CUDA_CHECK(cudaSetDevice(c->device));
    for_each_group_concurrent(c, [&](Group& g) { run_group(c, g, n, record, 0, -1); });
    CUDA_CHECK(cudaGetLastError());
}

// (*Chain).AdvanceChain for one group: cw + 1 recorded sweeps.  buffer/circular.go: the window keeps the last
// 2 * (cw / 2) values (NewCircularInt rounds the size down to even, circular.go:16-27) and FirstHalf / SecondHalf split
// exactly those — for an odd cw the newest cw - 1 samples, i.e. the round's first 2 sweeps stay out of the histograms
// (1 for an even cw).
void advance_group(gb_chains* c, Group& g, int32_t cw) {
    const int32_t half = cw / 2;
    if (c->flags & GB_CHAINS_HISTORY) {
        if (half > 65535) throw gb::Err("convergence window too large for 16-bit half-window histograms");
        CUDA_CHECK(cudaMemsetAsync(g.d_hist, 0, (size_t)2 * g.model->h.total_card * g.n_pad * sizeof(uint16_t), c->stream));
    }
    run_group(c, g, (int64_t)cw + 1, 1, cw + 1 - 2 * half, half);
    g.window_filled = true;
}

// skip flags: bit0 = collapsed in any group, bit1 = fixed
std::vector<uint8_t> collapsed_any(const gb_chains* c, std::vector<int>* first_group = nullptr) {
    const int nv = c->base().n_vars;
    std::vector<uint8_t> col(nv, 0);
    if (first_group) first_group->assign(nv, -1);
    for (size_t gi = 0; gi < c->groups.size(); gi++)
        for (int v = 0; v < nv; v++)
            if (c->groups[gi].model->h.collapsed[v] && !col[v]) {
                col[v] = 1;
                if (first_group) (*first_group)[v] = (int)gi;
            }
    return col;
}

void ensure_scratch(gb_chains* c) {
    const gb::HostModel& h = c->base();
    if (!c->d_merge) CUDA_CHECK(cudaMalloc(&c->d_merge, (size_t)h.total_card * sizeof(double)));
    if (!c->d_merged_in) CUDA_CHECK(cudaMalloc(&c->d_merged_in, (size_t)h.total_card * sizeof(double)));
    if (!c->d_wb) CUDA_CHECK(cudaMalloc(&c->d_wb, ((size_t)2 * h.n_vars + 1) * sizeof(double)));  // + the chain count
    if (!c->d_skip) CUDA_CHECK(cudaMalloc(&c->d_skip, (size_t)h.n_vars));
    if (!c->merge_stream) {
        int lo = 0, hi = 0;  // highest priority: its small kernels and the NCCL kernel take the first free SM slots
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_CHECK(cudaStreamCreateWithPriority(&c->merge_stream, cudaStreamNonBlocking, hi));
        for (auto& sl : c->slots) {
            CUDA_CHECK(cudaMalloc(&sl.d_cnt, ((size_t)h.total_card + 2) * sizeof(unsigned long long)));
            CUDA_CHECK(cudaMalloc(&sl.d_out, (size_t)h.total_card * sizeof(double)));
            CUDA_CHECK(cudaMallocHost(&sl.h_tail, 2 * sizeof(unsigned long long)));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.ev_snap, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
            for (auto& e : sl.ev_t) CUDA_CHECK(cudaEventCreate(&e));
        }
    }
}

void refresh_collapsed_cache(gb_chains* c) {
    if (!c->col_any.empty()) return;
    c->col_any = collapsed_any(c, &c->col_first_group);
    c->col_vars.clear();
    for (int v = 0; v < (int)c->col_any.size(); v++)
        if (c->col_any[v]) c->col_vars.push_back(v);
}

// d_skip[v]: bit0 = collapsed in any group (skipped by the merge), bit1 = fixed (skipped, with
// bit0, by the convergence pass); uploaded once per change of the group list
void upload_skip(gb_chains* c) {
    refresh_collapsed_cache(c);
    if (c->skip_uploaded) return;
    ensure_scratch(c);
    const gb::HostModel& h = c->base();
    c->skip_bits.resize(h.n_vars);
    for (int v = 0; v < h.n_vars; v++) c->skip_bits[v] = (uint8_t)(c->col_any[v] | (h.fixed[v] >= 0 ? 2 : 0));
    if (c->merge_ever) CUDA_CHECK(cudaStreamSynchronize(c->merge_stream));  // an earlier merge may still read the old flags
    CUDA_CHECK(cudaMemcpyAsync(c->d_skip, c->skip_bits.data(), c->skip_bits.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->col_any32.assign(c->col_any.begin(), c->col_any.end());
    c->skip_uploaded = true;
}

int64_t local_chains(const gb_chains* c) {
    int64_t n = 0;
    for (const auto& g : c->groups) n += g.n_chains;
    return n;
}
bool has_peers(const gb_chains* c) { return c->comm && c->comm->world > 1; }

// ---- MergeChains (chain.go:96-148) in three phases, so that a fleet of handles can put phase 2 of all its devices
// ---- inside one NCCL group and so that phases 2-3 overlap with the sweeps enqueued after phase 1.
// phase 1, sweep stream: snapshot = integer sums of the groups' counts (+ chain count, TotalSampleCount)
void merge_snapshot(gb_chains* c) {
    CUDA_CHECK(cudaSetDevice(c->device));
    if (c->slots_pending >= gb_chains::kMergeSlots)
        throw gb::Err("too many merges are already pending on this handle (" + std::to_string(gb_chains::kMergeSlots) +
                      "): call gb_chains_merge_end first");
    ensure_scratch(c);
    const gb::HostModel& h = c->base();
    upload_skip(c);
    auto& sl = c->slots[c->slot_head];
    if (sl.ever) CUDA_CHECK(cudaStreamWaitEvent(c->stream, sl.ev_done, 0));  // the slot's previous merge has been consumed
    CUDA_CHECK(cudaEventRecord(sl.ev_t[0], c->stream));
    CUDA_CHECK(cudaMemsetAsync(sl.d_cnt, 0, ((size_t)h.total_card + 2) * sizeof(unsigned long long), c->stream));
    bool first = true;
    for (auto& g : c->groups) {
        gb::k_merge_counts<<<(h.total_card + 255) / 256, 256, 0, c->stream>>>(
            g.model->dev, g.d_counts, c->d_skip, sl.d_cnt, first ? (unsigned long long)local_chains(c) : 0ull,
            first ? (unsigned long long)c->total_samples : 0ull);
        first = false;
    }
    c->launches += (int64_t)c->groups.size();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(sl.ev_snap, c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(c->merge_stream, sl.ev_snap, 0));
    CUDA_CHECK(cudaEventRecord(sl.ev_t[1], c->merge_stream));
}
// phase 2, merge stream: sum over the ranks of the communicator (collective: every rank calls it; a fleet brackets
// the calls of its devices with ncclGroupStart / ncclGroupEnd)
void merge_reduce(gb_chains* c) {
    if (!has_peers(c)) return;
    CUDA_CHECK(cudaSetDevice(c->device));
    auto& sl = c->slots[c->slot_head];
    NCCL_CHECK(gbn::api().AllReduce(sl.d_cnt, sl.d_cnt, (size_t)c->base().total_card + 2, gbn::kNcclUint64, gbn::kNcclSum,
                                    c->comm->nccl, c->merge_stream));
}
// phase 3: counts -> marginals (every chain starts at uniform 1/card, model/variable.go:45) -> host.
// The conversion kernel runs on the SWEEP stream, behind the reduction's event: a kernel on the side stream would only get
// SMs when a (persistent, device-filling) sweep kernel ends, i.e. one more kernel boundary later.  Without peers it is
// enqueued at once (right behind the count sums); with peers it is deferred to just after the NEXT sweep kernel launch
// (flush_finalizes), so that the NCCL kernel — which starts the moment the count sums are done, all SMs being free then —
// overlaps with that sweep kernel instead of holding the sweep stream up.  The copy to the host follows on the side stream.
// A page-locked destination (cudaHostAlloc / cudaHostRegister / a pinned torch tensor) receives the DMA directly; a
// pageable one goes through the slot's pinned staging buffer.  out == nullptr: no host copy of the marginals (a rank
// that only takes part in the reduction; the totals still arrive).
void finalize_slot(gb_chains* c, gb_chains::MergeSlot& sl) {
    if (!sl.needs_finalize) return;
    sl.needs_finalize = false;
    const gb::HostModel& h = c->base();
    const size_t bytes = (size_t)h.total_card * sizeof(double);
    const double unit = (c->flags & GB_CHAINS_RAO_BLACKWELL) ? 1.0 / gb::kRbScale : 1.0;
    CUDA_CHECK(cudaStreamWaitEvent(c->stream, sl.ev_t[2], 0));  // the reduction (or, without peers, the snapshot) is done
    gb::k_merge_finalize<<<(h.total_card + 255) / 256, 256, 0, c->stream>>>(c->groups[0].model->dev, sl.d_cnt, c->d_skip, sl.d_out, unit);
    c->launches++;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(sl.ev_t[3], c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(c->merge_stream, sl.ev_t[3], 0));
    sl.staged = false;
    if (sl.out) {
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, sl.out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();  // (older drivers report an unregistered host pointer as an error)
        if (pinned) {
            CUDA_CHECK(cudaMemcpyAsync(sl.out, sl.d_out, bytes, cudaMemcpyDeviceToHost, c->merge_stream));
        } else {
            if (!sl.h_stage) CUDA_CHECK(cudaMallocHost(&sl.h_stage, bytes));
            CUDA_CHECK(cudaMemcpyAsync(sl.h_stage, sl.d_out, bytes, cudaMemcpyDeviceToHost, c->merge_stream));
            sl.staged = true;
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(sl.h_tail, sl.d_cnt + h.total_card, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               c->merge_stream));
    CUDA_CHECK(cudaEventRecord(sl.ev_t[4], c->merge_stream));
    CUDA_CHECK(cudaEventRecord(sl.ev_done, c->merge_stream));
}
void flush_finalizes(gb_chains* c) {
    if (!c->slots_pending) return;
    for (int i = 0; i < gb_chains::kMergeSlots; i++) finalize_slot(c, c->slots[(c->slot_tail + i) % gb_chains::kMergeSlots]);
}
void merge_collect(gb_chains* c, double* out, int32_t* collapsed_out) {
    CUDA_CHECK(cudaSetDevice(c->device));
    auto& sl = c->slots[c->slot_head];
    CUDA_CHECK(cudaEventRecord(sl.ev_t[2], c->merge_stream));  // behind the NCCL kernel (or right behind the snapshot)
    sl.out = out;
    sl.col_out = collapsed_out;
    sl.pending = true;
    sl.ever = true;
    sl.needs_finalize = true;
    c->merge_ever = true;
    c->slot_head = (c->slot_head + 1) % gb_chains::kMergeSlots;
    c->slots_pending++;
    if (!has_peers(c)) finalize_slot(c, sl);
}
// host side of the merge on an already-reduced vector: collapsed-in-any variables take the first such chain's local marginal
void merge_overrides(gb_chains* c, double* out, int32_t* collapsed_out) {
    const gb::HostModel& h = c->base();
    refresh_collapsed_cache(c);
    if (collapsed_out) {
        upload_skip(c);
        std::memcpy(collapsed_out, c->col_any32.data(), (size_t)h.n_vars * sizeof(int32_t));
    }
    if (!out) return;
    for (int v : c->col_vars) {
        const auto& m = c->groups[c->col_first_group[v]].model->h.coll_marg[v];  // chain.go:113-129: first chain found
        for (int k = 0; k < h.card[v]; k++) out[h.card_off[v] + k] = m[k];
    }
}
// completes the OLDEST pending merge
void merge_wait(gb_chains* c) {
    if (c->slots_pending == 0) throw gb::Err("no merge is pending on this handle: call gb_chains_merge_begin first");
    CUDA_CHECK(cudaSetDevice(c->device));
    auto& sl = c->slots[c->slot_tail];
    finalize_slot(c, sl);  // (no sweep followed the merge: nothing has flushed it yet)
    CUDA_CHECK(cudaEventSynchronize(sl.ev_done));
    sl.pending = false;
    c->last_done_slot = c->slot_tail;
    c->slot_tail = (c->slot_tail + 1) % gb_chains::kMergeSlots;
    c->slots_pending--;
    const gb::HostModel& h = c->base();
    if (sl.staged) std::memcpy(sl.out, sl.h_stage, (size_t)h.total_card * sizeof(double));
    c->global_chains = (int64_t)sl.h_tail[0];
    c->global_samples = (int64_t)sl.h_tail[1];
    merge_overrides(c, sl.out, sl.col_out);
}
void merged_marginals(gb_chains* c, double* out, int32_t* collapsed_out) {
    while (c->slots_pending) merge_wait(c);  // (a blocking merge drains the asynchronous ones first: results arrive in order)
    merge_snapshot(c);
    merge_reduce(c);
    merge_collect(c, out, collapsed_out);
    merge_wait(c);
}

// Legacy multi-device form (the caller all-reduces a float64 device buffer itself): this device's contribution
// n_local / card + counts, collapsed variables zero
void merge_partial(gb_chains* c) {
    CUDA_CHECK(cudaSetDevice(c->device));
    ensure_scratch(c);
    const gb::HostModel& h = c->base();
    upload_skip(c);
    CUDA_CHECK(cudaMemsetAsync(c->d_merge, 0, (size_t)h.total_card * sizeof(double), c->stream));
    for (auto& g : c->groups)
        gb::k_merge_partial<<<(h.total_card + 255) / 256, 256, 0, c->stream>>>(g.model->dev, g.d_counts,
                                                                              (double)g.n_chains, c->d_skip, c->d_merge,
                                                                              g.dev.rb ? 1.0 / gb::kRbScale : 1.0);
    c->launches += (int64_t)c->groups.size();
    CUDA_CHECK(cudaGetLastError());
}

void merge_finalize(gb_chains* c, double* out, int32_t* collapsed_out) {
    CUDA_CHECK(cudaSetDevice(c->device));
    const gb::HostModel& h = c->base();
    const size_t bytes = (size_t)h.total_card * sizeof(double);
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    if (pinned) {
        CUDA_CHECK(cudaMemcpyAsync(out, c->d_merge, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
        if (!c->h_merge) CUDA_CHECK(cudaMallocHost(&c->h_merge, bytes));
        CUDA_CHECK(cudaMemcpyAsync(c->h_merge, c->d_merge, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::memcpy(out, c->h_merge, bytes);
    }
    merge_overrides(c, out, collapsed_out);
}

// ---- ChainConvergence (chain.go:32-92) in phases, like the merge.
// phase 1, sweep stream: per-variable sums of within / between distances over this device's chains, chain count in the tail
void convergence_enqueue(gb_chains* c, int measure, const double* merged) {
    CUDA_CHECK(cudaSetDevice(c->device));
    if (!(c->flags & GB_CHAINS_HISTORY)) throw gb::Err("chains were created without GB_CHAINS_HISTORY");
    if (measure < 0 || measure > 3) throw gb::Err("unknown measure");
    ensure_scratch(c);
    const gb::HostModel& h = c->base();
    upload_skip(c);
    CUDA_CHECK(cudaMemcpyAsync(c->d_merged_in, merged, (size_t)h.total_card * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->d_wb, 0, ((size_t)2 * h.n_vars + 1) * sizeof(double), c->stream));
    for (auto& g : c->groups) {
        const int64_t items = (int64_t)h.n_vars * g.n_chains;
        gb::k_chain_dist<<<grid_for(items, 256), 256, 0, c->stream>>>(g.model->dev, g.dev, c->d_merged_in, c->d_skip,
                                                                     measure, c->d_wb, (double)g.n_chains);
        c->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
}
// phase 2: sum over the ranks (collective)
void convergence_reduce(gb_chains* c) {
    if (!has_peers(c)) return;
    CUDA_CHECK(cudaSetDevice(c->device));
    NCCL_CHECK(gbn::api().AllReduce(c->d_wb, c->d_wb, (size_t)2 * c->base().n_vars + 1, gbn::kNcclFloat64, gbn::kNcclSum,
                                    c->comm->nccl, c->stream));
}
// phase 3: scores on the host (chain.go:46-59, 69-88) from the reduced sums and the global chain count
void convergence_finalize(const gb::HostModel& h, const double* wb, int32_t cw, int64_t total_chains, const uint8_t* collapsed,
                          double* out);
void convergence_collect(gb_chains* c, double* out) {
    CUDA_CHECK(cudaSetDevice(c->device));
    const gb::HostModel& h = c->base();
    std::vector<double> wb((size_t)2 * h.n_vars + 1);
    CUDA_CHECK(cudaMemcpyAsync(wb.data(), c->d_wb, wb.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const int64_t total = (int64_t)(wb.back() + 0.5);
    if (total < 2) throw gb::Err("Convergence requires at least 2 chains");
    refresh_collapsed_cache(c);
    convergence_finalize(h, wb.data(), c->last_cw, total, c->col_any.data(), out);
}
// the three phases on one handle; merged == nullptr: MergeChains first (collective when a communicator is attached)
void convergence_scores(gb_chains* c, int measure, const double* merged, double* out) {
    if (c->last_cw < 2) throw gb::Err("Total seen < Convergence Window: run gb_chains_advance first");
    for (const auto& g : c->groups)
        if (!g.window_filled) throw gb::Err("Total seen < Convergence Window: a chain group has not advanced since it was added");
    std::vector<double> tmp;
    if (!merged) {
        tmp.resize(c->base().total_card);
        merged_marginals(c, tmp.data(), nullptr);
        merged = tmp.data();
    }
    convergence_enqueue(c, measure, merged);
    convergence_reduce(c);
    convergence_collect(c, out);
}
// legacy: this device's sums only, left in d_wb for the caller's own all-reduce
void convergence_partial(gb_chains* c, int measure, const double* merged) {
    std::vector<double> tmp;
    if (!merged) {
        tmp.resize(c->base().total_card);
        merge_partial(c);
        merge_finalize(c, tmp.data(), nullptr);
        merged = tmp.data();
    }
    convergence_enqueue(c, measure, merged);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// chain.go:46-59, 69-88
void convergence_finalize(const gb::HostModel& h, const double* wb, int32_t cw, int64_t total_chains,
                          const uint8_t* collapsed, double* out) {
    if (total_chains < 2) throw gb::Err("Convergence requires at least 2 chains");
    const double n = (double)cw, m = (double)total_chains;
    const double b_norm = n / (m - 1), w_factor = (n - 1) / n, b_factor = (m + 1) / (m * n);
    for (int v = 0; v < h.n_vars; v++) {
        if (collapsed[v] || h.fixed[v] >= 0) {
            out[v] = 1.0;
            continue;
        }
        double W = (1e-8 + wb[v]) / m;
        double B = (1e-8 + wb[h.n_vars + v]) * b_norm;
        double vhat = w_factor * W + b_factor * B;
        out[v] = std::sqrt((4.0 * vhat) / (2.0 * W));
    }
}

// gibbs-collapsed.go:98-314 as a pure function over the flattened model
gb_model* collapse_model(const gb_model* src, int32_t var, uint64_t seed, int32_t* var_out, double* marg_out) {
    const gb::HostModel& h = src->h;
    if (src->device < 0) throw gb::Err("collapse needs a device-resident model (there is no CPU fallback)");
    require_device(src->device);
    if (var < 0) {  // lines 102-120: up to n tries for a tractable variable
        std::vector<int32_t> elig;
        for (int v = 0; v < h.n_vars; v++)
            if (h.fixed[v] < 0 && !h.collapsed[v]) elig.push_back(v);
        if (elig.empty()) throw gb::Err("Failure selecting random variable to collapse: No Variables to select");
        for (int t = 0; t < h.n_vars; t++) {
            int32_t pick = elig[0];
            if (elig.size() > 1) {
                gb::Philox4 r = gb::philox4x32_10((uint32_t)t, 0u, 0u, gb::kTagCollapse, (uint32_t)seed, (uint32_t)(seed >> 32));
                pick = elig[(size_t)(((uint64_t)r.x * elig.size()) >> 32)];
            }
            if ((int)h.nbrs[pick].size() <= gb::kNeighborVarMax) {
                var = pick;
                break;
            }
        }
        if (var < 0) throw gb::Err("Failed to randomly select a variable to collapse");
    }
    if (var >= h.n_vars) throw gb::Err("Invalid variable index: max is " + std::to_string(h.n_vars - 1));
    if (h.fixed[var] >= 0) throw gb::Err("Can not collapse Fixed Val variable " + std::to_string(var));
    if (h.collapsed[var]) throw gb::Err("Already collapsed variable " + std::to_string(var));

    std::vector<int32_t> blanket;  // without var, ascending (the reference's map order is random)
    bool self = false;
    for (int32_t u : h.nbrs[var]) {
        if (u == var) self = true;
        else blanket.push_back(u);
    }
    if (!self) throw gb::Err("Collapsing variable not in its own blanket");
    if (blanket.empty()) throw gb::Err("New function would have 0 variables");
    int64_t new_size = 1;
    for (int32_t u : blanket) {
        new_size *= h.card[u];
        if (new_size > gb::kMaxTabSize)
            throw gb::Err("Function over " + std::to_string(blanket.size()) + " vars has size > " + std::to_string(gb::kMaxTabSize));
    }
    if ((int)blanket.size() > gb::kNeighborVarMaxDev) throw gb::Err("blanket exceeds the device limit");  // unreachable: 2^23 entries cap it at 23

    const auto& vf = h.var_funcs[var];
    gb::CollapsePlan pl{};
    pl.n_b = (int32_t)blanket.size();
    pl.n_f = (int32_t)vf.size();
    pl.card_v = h.card[var];
    pl.new_size = new_size;
    for (int b = 0; b < pl.n_b; b++) {
        pl.bcard[b] = h.card[blanket[b]];
        pl.bfixed[b] = h.fixed[blanket[b]];
    }
    std::vector<int32_t> f_tab_off, f_stride_v, f_stride_b((size_t)pl.n_f * pl.n_b, 0);
    for (int fi = 0; fi < pl.n_f; fi++) {
        const gb::Factor& f = h.funcs[vf[fi]];
        f_tab_off.push_back((int32_t)f.off);
        int32_t sv = 0;
        for (size_t i = 0; i < f.vars.size(); i++) {
            if (f.vars[i] == var) {
                sv = (int32_t)f.strides[i];  // Eval reads the value of every scope slot from the state
                continue;
            }
            int b = (int)(std::lower_bound(blanket.begin(), blanket.end(), f.vars[i]) - blanket.begin());
            f_stride_b[(size_t)fi * pl.n_b + b] += (int32_t)f.strides[i];
        }
        f_stride_v.push_back(sv);
    }
    struct DevBuf {  // scratch of this call, released on every exit path
        void* p = nullptr;
        ~DevBuf() { cudaFree(p); }
    } b_off, b_sv, b_sb, b_new, b_marg;
    b_off.p = dev_upload(f_tab_off);
    b_sv.p = dev_upload(f_stride_v);
    b_sb.p = dev_upload(f_stride_b);
    pl.f_tab_off = static_cast<int32_t*>(b_off.p);
    pl.f_stride_v = static_cast<int32_t*>(b_sv.p);
    pl.f_stride_b = static_cast<int32_t*>(b_sb.p);
    CUDA_CHECK(cudaMalloc(&b_new.p, (size_t)new_size * sizeof(double)));
    std::vector<double> marg(pl.card_v, 1e-12);  // line 138-140
    b_marg.p = dev_upload(marg);
    double *d_new = static_cast<double*>(b_new.p), *d_marg = static_cast<double*>(b_marg.p);
    gb::k_collapse<<<(int)((new_size + 255) / 256), 256>>>(pl, src->dev.tab64, d_new, d_marg);
    CUDA_CHECK(cudaGetLastError());
    std::vector<double> new_tab((size_t)new_size);
    CUDA_CHECK(cudaMemcpy(new_tab.data(), d_new, (size_t)new_size * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(marg.data(), d_marg, marg.size() * sizeof(double), cudaMemcpyDeviceToHost));
    gb::norm_marginal(marg);  // line 263

    // lines 275-290: append the new factor, drop the variable's old factors, keep order
    auto out = std::make_unique<gb_model>();
    gb::HostModel& n = out->h;
    n.n_vars = h.n_vars;
    n.card = h.card;
    n.fixed = h.fixed;
    n.collapsed = h.collapsed;
    n.coll_marg = h.coll_marg;
    std::vector<uint8_t> drop(h.funcs.size(), 0);
    for (int32_t fi : vf) drop[fi] = 1;
    for (size_t fi = 0; fi < h.funcs.size(); fi++) {
        if (drop[fi]) continue;
        const gb::Factor& f = h.funcs[fi];
        n.add_factor(f.vars, h.log_tab.data() + f.off, f.size, true);
    }
    n.add_factor(blanket, new_tab.data(), new_size, true);
    n.collapsed[var] = 1;
    n.coll_marg[var] = marg;  // lines 310-313
    n.build_derived();
    out->upload(src->device);
    if (var_out) *var_out = var;
    if (marg_out) std::copy(marg.begin(), marg.end(), marg_out);
    return out.release();
}

}  // namespace

extern "C" {

const char* gb_last_error(void) { return g_err.c_str(); }
int gb_version(void) { return 100; }
int gb_device_count(int* n_out) {
    GB_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    *n_out = n;
    GB_END
}

// ------------------------------------------------------------------ model
int gb_model_create(int32_t n_vars, const int32_t* card, const int32_t* fixed, int32_t n_funcs,
                    const int32_t* scope_off, const int32_t* scope_vars, const int64_t* tab_off,
                    const double* tables_raw, int device, gb_model** out) {
    GB_TRY
    auto m = std::make_unique<gb_model>();
    m->h = gb::make_model(n_vars, card, fixed, n_funcs, scope_off, scope_vars, tab_off, tables_raw);
    m->upload(device);
    *out = m.release();
    GB_END
}
int gb_model_load_uai(const char* uai_path, const char* evid_path, int device, gb_model** out) {
    GB_TRY
    auto m = std::make_unique<gb_model>();
    m->h = gb::load_uai(uai_path, evid_path);
    m->upload(device);
    *out = m.release();
    GB_END
}
void gb_model_destroy(gb_model* m) { delete m; }

int gb_model_n_vars(const gb_model* m, int32_t* out) { *out = m->h.n_vars; return 0; }
int gb_model_n_funcs(const gb_model* m, int32_t* out) { *out = (int32_t)m->h.funcs.size(); return 0; }
int gb_model_total_card(const gb_model* m, int32_t* out) { *out = m->h.total_card; return 0; }
int gb_model_cards(const gb_model* m, int32_t* out) { std::copy(m->h.card.begin(), m->h.card.end(), out); return 0; }
int gb_model_fixed(const gb_model* m, int32_t* out) { std::copy(m->h.fixed.begin(), m->h.fixed.end(), out); return 0; }
int gb_model_collapsed(const gb_model* m, int32_t* out) {
    for (int v = 0; v < m->h.n_vars; v++) out[v] = m->h.collapsed[v];
    return 0;
}
#define GB_FUNC_CHECK(f) \
    if ((f) < 0 || (f) >= (int32_t)m->h.funcs.size()) throw gb::Err("function index out of range")
int gb_model_func_arity(const gb_model* m, int32_t f, int32_t* out) {
    GB_TRY GB_FUNC_CHECK(f);
    *out = (int32_t)m->h.funcs[f].vars.size();
    GB_END
}
int gb_model_func_scope(const gb_model* m, int32_t f, int32_t* out) {
    GB_TRY GB_FUNC_CHECK(f);
    std::copy(m->h.funcs[f].vars.begin(), m->h.funcs[f].vars.end(), out);
    GB_END
}
int gb_model_func_table_size(const gb_model* m, int32_t f, int64_t* out) {
    GB_TRY GB_FUNC_CHECK(f);
    *out = m->h.funcs[f].size;
    GB_END
}
int gb_model_func_log_table(const gb_model* m, int32_t f, double* out) {
    GB_TRY GB_FUNC_CHECK(f);
    const gb::Factor& fn = m->h.funcs[f];
    std::copy(m->h.log_tab.begin() + fn.off, m->h.log_tab.begin() + fn.off + fn.size, out);
    GB_END
}
int gb_model_blanket_size(const gb_model* m, int32_t var, int32_t* out) {
    GB_TRY
    if (var < 0 || var >= m->h.n_vars) throw gb::Err("Invalid variable index");
    *out = (int32_t)m->h.nbrs[var].size();
    GB_END
}
int gb_model_function_count(const gb_model* m, int32_t var, int32_t* out) {
    GB_TRY
    if (var < 0 || var >= m->h.n_vars) throw gb::Err("Invalid variable index");
    *out = (int32_t)m->h.var_funcs[var].size();
    GB_END
}
int gb_model_schedule(const gb_model* m, int32_t* n_order, int32_t* n_colours, int32_t* order, int32_t* colour_off) {
    GB_TRY
    if (n_order) *n_order = (int32_t)m->h.order.size();
    if (n_colours) *n_colours = (int32_t)m->h.colour_off.size() - 1;
    if (order) std::copy(m->h.order.begin(), m->h.order.end(), order);
    if (colour_off) std::copy(m->h.colour_off.begin(), m->h.colour_off.end(), colour_off);
    GB_END
}
int gb_model_table_mode(gb_model* m, int32_t* ok_out, int64_t* n_thresholds_out) {
    GB_TRY
    if (ok_out) *ok_out = m->h.tab_ok ? 1 : 0;
    if (n_thresholds_out) *n_thresholds_out = m->h.n_thresholds;
    GB_END
}
int gb_model_bits_mode(const gb_model* m, int32_t* ok_out) {
    *ok_out = m->h.bits_ok ? 1 : 0;
    return 0;
}
int gb_model_hybrid_mask(const gb_model* m, int32_t* mask_out) {
    GB_TRY
    const bool on = m->hybrid_tables();
    for (int v = 0; v < m->h.n_vars; v++) mask_out[v] = (on && m->h.tp_off[v] >= 0) ? 1 : 0;
    GB_END
}
int gb_model_thresholds(gb_model* m, int32_t var, int32_t* n_out, uint32_t* out) {
    GB_TRY
    if (var < 0 || var >= m->h.n_vars) throw gb::Err("Invalid variable index");
    m->ensure_thresholds();
    if (m->h.tp_off[var] < 0) throw gb::Err("variable has no threshold table (not sampled, cardinality above 4, or too many neighbour configurations)");
    const int32_t* tp = m->h.tprog.data() + m->h.tp_off[var];
    int n = m->h.card[var] - 1;  // card - 1 cumulative thresholds per configuration
    for (int i = 0; i < tp[0]; i++) n *= m->h.card[tp[2 + 2 * i]];
    if (n_out) *n_out = n;
    if (out) CUDA_CHECK(cudaMemcpy(out, m->tab.thr + tp[1], (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    GB_END
}
int gb_model_collapse(const gb_model* src, int32_t var, uint64_t seed, int32_t* collapsed_var_out,
                      double* marginal_out, gb_model** out) {
    if (collapsed_var_out) *collapsed_var_out = -1;
    GB_TRY *out = collapse_model(src, var, seed, collapsed_var_out, marginal_out);
    GB_END
}

// K5 on validated inputs: floored un-normalised weights e[k] (gibbs-simple.go:171-258) of (state, variable) pairs
static void conditional_impl(const gb_model* m, int precision, int32_t n_states, const int32_t* states, const int32_t* vars,
                             double* out) {
    if (m->device < 0) throw gb::Err("gb_conditional needs a device-resident model (there is no CPU fallback)");
    require_device(m->device);
    const gb::HostModel& h = m->h;
    for (int s = 0; s < n_states; s++) {
        int v = vars[s];
        if (v < 0 || v >= h.n_vars) throw gb::Err("Invalid variable index");
        if (h.fixed[v] >= 0) throw gb::Err("Selected sample variable " + std::to_string(v) + " which has FixedVal=" + std::to_string(h.fixed[v]));
        if (h.prog_off[v] < 0) throw gb::Err("variable " + std::to_string(v) + " is collapsed: it has no factors to sample from");
        for (int u = 0; u < h.n_vars; u++) {
            int x = states[(size_t)s * h.n_vars + u];
            if (x < 0 || x >= h.card[u]) throw gb::Err("Value " + std::to_string(x) + " invalid for cardinality " + std::to_string(h.card[u]));
        }
    }
    struct DevBuf {
        void* p = nullptr;
        ~DevBuf() { cudaFree(p); }
    } b_states, b_vars, b_out;
    const size_t ns = (size_t)n_states;
    CUDA_CHECK(cudaMalloc(&b_states.p, ns * h.n_vars * sizeof(int32_t)));
    CUDA_CHECK(cudaMalloc(&b_vars.p, ns * sizeof(int32_t)));
    CUDA_CHECK(cudaMalloc(&b_out.p, ns * gb::kProbeStride * sizeof(double)));
    int32_t *d_states = static_cast<int32_t*>(b_states.p), *d_vars = static_cast<int32_t*>(b_vars.p);
    double* d_out = static_cast<double*>(b_out.p);
    CUDA_CHECK(cudaMemcpy(d_states, states, ns * h.n_vars * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_vars, vars, ns * sizeof(int32_t), cudaMemcpyHostToDevice));
    const int blocks = (n_states + 127) / 128;
    if (precision == GB_F32) gb::k_conditional<float><<<blocks, 128>>>(m->dev, n_states, d_states, d_vars, d_out);
    else gb::k_conditional<double><<<blocks, 128>>>(m->dev, n_states, d_states, d_vars, d_out);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(out, d_out, ns * gb::kProbeStride * sizeof(double), cudaMemcpyDeviceToHost));
}

int gb_conditional(const gb_model* m, int precision, int32_t n_states, const int32_t* states,
                   const int32_t* vars, double* out) {
    GB_TRY
    conditional_impl(m, precision, n_states, states, vars, out);
    GB_END
}

// (*GibbsSimple).Sample / SampleVar, (*GibbsCollapsed).Sample for ONE caller-held state (gibbs-simple.go:148-271,
// gibbs-collapsed.go:317-334): variable choice on the host, conditional on the device (K5), inverse-CDF draw on the host
int gb_model_sample(const gb_model* m, int precision, int32_t var, int exclude_collapsed, uint64_t seed, uint64_t step,
                    int32_t* state_inout, int32_t* var_out) {
    if (var_out) *var_out = -1;  // the reference returns index -1 on failure
    GB_TRY
    const gb::HostModel& h = m->h;
    const gb::Philox4 r = gb::philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 0u, gb::kTagScan, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (var < 0) {  // UniformSampler.VarSample (sampler.go:135-174): uniform over FixedVal < 0 (and not Collapsed)
        std::vector<int32_t> elig;
        for (int v = 0; v < h.n_vars; v++)
            if (h.fixed[v] < 0 && !(exclude_collapsed && h.collapsed[v])) elig.push_back(v);
        if (elig.empty()) throw gb::Err("No Variables to select");
        var = elig[(size_t)(((uint64_t)r.x * elig.size()) >> 32)];
    }
    if (var >= h.n_vars) throw gb::Err("Invalid variable index");
    double e[gb::kProbeStride];
    conditional_impl(m, precision, 1, state_inout, &var, e);
    // UniformSampler.WeightedSample (sampler.go:107-123): re-sum, r = U * tot, first k with r <= w[k]
    const int card = h.card[var];
    double tot = 0.0;
    for (int k = 0; k < card; k++) tot += e[k];
    double u = gb::u53(r.z, r.w) * tot;
    int sel = card - 1;
    for (int k = 0; k < card; k++) {
        if (u <= e[k]) {
            sel = k;
            break;
        }
        u -= e[k];
    }
    state_inout[var] = sel;
    if (var_out) *var_out = var;
    GB_END
}

// ------------------------------------------------------------------ chains
int gb_chains_create(int32_t n_groups, gb_model* const* models, const int32_t* chains_per_model,
                     uint64_t seed, uint64_t first_chain_id, int precision, uint32_t flags, int device,
                     gb_chains** out) {
    GB_TRY
    if (n_groups < 1) throw gb::Err("at least one chain group is required");
    if (precision != GB_F64 && precision != GB_F32 && precision != GB_TABLE && precision != GB_HYBRID && precision != GB_TABLE_BITS)
        throw gb::Err("unknown precision");
    if ((flags & GB_CHAINS_RAO_BLACKWELL) && precision == GB_TABLE_BITS)
        throw gb::Err("GB_CHAINS_RAO_BLACKWELL is not available under GB_TABLE_BITS (its thresholds are bit-sliced): use GB_TABLE or GB_HYBRID");
    require_device(device);
    auto c = std::make_unique<gb_chains>();
    c->device = device;
    c->seed = seed;
    c->precision = precision;
    c->flags = flags;
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    uint64_t first = first_chain_id;
    for (int g = 0; g < n_groups; g++) {
        add_group(c.get(), models[g], chains_per_model[g], first, false);
        first += (uint64_t)((chains_per_model[g] + 7) / 8 * 8);
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = c.release();
    GB_END
}
int gb_chains_add_group(gb_chains* c, gb_model* model, int32_t n_chains, uint64_t first_chain_id) {
    GB_TRY
    GB_LOCK(c);
    CUDA_CHECK(cudaSetDevice(c->device));
    add_group(c, model, n_chains, first_chain_id, false);
    GB_END
}
void gb_chains_destroy(gb_chains* c) { delete c; }
int gb_chains_n_groups(const gb_chains* cc, int32_t* out) {
    gb_chains* c = const_cast<gb_chains*>(cc);
    GB_LOCK(c);
    *out = (int32_t)c->groups.size();
    return 0;
}
int gb_chains_n_chains(const gb_chains* cc, int64_t* out) {
    gb_chains* c = const_cast<gb_chains*>(cc);
    GB_LOCK(c);
    *out = local_chains(c);
    return 0;
}

int gb_chains_sweep(gb_chains* c, int64_t n_sweeps, int record) {
    GB_TRY
    GB_LOCK(c);
    sweeps(c, n_sweeps, record);
    GB_END
}
int gb_chains_sweep_timed(gb_chains* c, int64_t n_sweeps, int record, float* ms_out) {
    GB_TRY
    GB_LOCK(c);
    CUDA_CHECK(cudaSetDevice(c->device));
    if (!c->ev0) {
        CUDA_CHECK(cudaEventCreate(&c->ev0));
        CUDA_CHECK(cudaEventCreate(&c->ev1));
    }
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    sweeps(c, n_sweeps, record);
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    CUDA_CHECK(cudaEventSynchronize(c->ev1));
    CUDA_CHECK(cudaEventElapsedTime(ms_out, c->ev0, c->ev1));
    GB_END
}
int gb_chains_launch_count(const gb_chains* cc, int64_t* out) {
    gb_chains* c = const_cast<gb_chains*>(cc);
    GB_LOCK(c);
    *out = c->launches;
    return 0;
}
int gb_chains_scan(gb_chains* c, int64_t n_steps, int record) {
    GB_TRY
    GB_LOCK(c);
    if (n_steps < 0) throw gb::Err("Invalid step count");
    if (c->flags & GB_CHAINS_RAO_BLACKWELL) throw gb::Err("the random-scan parity mode records plain counts (no GB_CHAINS_RAO_BLACKWELL)");
    if (c->precision == GB_TABLE_BITS) throw gb::Err("the random-scan parity mode needs byte state: not available under GB_TABLE_BITS");
    CUDA_CHECK(cudaSetDevice(c->device));
    for (auto& g : c->groups) {
        const gb::HostModel& h = g.model->h;
        const int blocks = (g.n_chains + 127) / 128;
        const int32_t n_order = (int32_t)h.order.size();
        const int mc = h.max_card;
        if (mc <= 2) gb::k_random_scan<2><<<blocks, 128, 0, c->stream>>>(g.model->dev, g.dev, g.model->d_order, n_order, g.scan_step, n_steps, record);
        else if (mc <= 4) gb::k_random_scan<4><<<blocks, 128, 0, c->stream>>>(g.model->dev, g.dev, g.model->d_order, n_order, g.scan_step, n_steps, record);
        else if (mc <= 16) gb::k_random_scan<16><<<blocks, 128, 0, c->stream>>>(g.model->dev, g.dev, g.model->d_order, n_order, g.scan_step, n_steps, record);
        else gb::k_random_scan<64><<<blocks, 128, 0, c->stream>>>(g.model->dev, g.dev, g.model->d_order, n_order, g.scan_step, n_steps, record);
        c->launches++;
        g.scan_step += (uint64_t)n_steps;
        if (record) {
            g.total_samples += n_steps * g.n_chains;
            c->total_samples += n_steps * g.n_chains;
        }
    }
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_burnin(gb_chains* c, int64_t n_sweeps) {
    GB_TRY
    GB_LOCK(c);
    sweeps(c, n_sweeps, 0);
    GB_END
}
int gb_chains_advance(gb_chains* c, int32_t cw) {
    GB_TRY
    GB_LOCK(c);
    if (cw < 0) throw gb::Err("Invalid convergence window");
    CUDA_CHECK(cudaSetDevice(c->device));
    c->last_cw = cw;
    for_each_group_concurrent(c, [&](Group& g) { advance_group(c, g, cw); });
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_total_samples(const gb_chains* cc, int64_t* out) {
    gb_chains* c = const_cast<gb_chains*>(cc);
    GB_LOCK(c);
    *out = c->total_samples;
    return 0;
}

// ---- per-group forms: one reference Chain maps to one group of replica chains
static Group& group_at(gb_chains* c, int32_t group) {
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    return c->groups[group];
}
int gb_chains_group_sweep(gb_chains* c, int32_t group, int64_t n_sweeps, int record) {
    GB_TRY
    GB_LOCK(c);
    CUDA_CHECK(cudaSetDevice(c->device));
    run_group(c, group_at(c, group), n_sweeps, record, 0, -1);
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_group_advance(gb_chains* c, int32_t group, int32_t cw) {
    GB_TRY
    GB_LOCK(c);
    if (cw < 0) throw gb::Err("Invalid convergence window");
    CUDA_CHECK(cudaSetDevice(c->device));
    c->last_cw = cw;
    advance_group(c, group_at(c, group), cw);
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_group_info(gb_chains* c, int32_t group, int32_t* n_chains_out, int64_t* total_samples_out,
                         gb_model** model_out) {
    GB_TRY
    GB_LOCK(c);
    Group& g = group_at(c, group);
    if (n_chains_out) *n_chains_out = g.n_chains;
    if (total_samples_out) *total_samples_out = g.total_samples;
    if (model_out) *model_out = g.model;
    GB_END
}
// The wait itself happens OUTSIDE the handle's lock: a goroutine parked in wg.Wait() must not keep the others from
// enqueuing work (everything enqueued before this call is covered; what arrives later is not waited for).
int gb_chains_synchronize(gb_chains* c) {
    GB_TRY
    cudaEvent_t ev = nullptr;
    {
        GB_LOCK(c);
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(ev, c->stream));
    }
    const cudaError_t e = cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
    CUDA_CHECK(e);
    GB_END
}

int gb_chains_merged_marginals(gb_chains* c, double* out, int32_t* collapsed_out) {
    GB_TRY
    GB_LOCK(c);
    merged_marginals(c, out, collapsed_out);
    GB_END
}
int gb_chains_merge_begin(gb_chains* c, double* out, int32_t* collapsed_out) {
    GB_TRY
    GB_LOCK(c);
    merge_snapshot(c);
    merge_reduce(c);
    merge_collect(c, out, collapsed_out);
    GB_END
}
int gb_chains_merge_end(gb_chains* c, int64_t* total_chains_out, int64_t* total_samples_out) {
    GB_TRY
    GB_LOCK(c);
    merge_wait(c);
    if (total_chains_out) *total_chains_out = c->global_chains;
    if (total_samples_out) *total_samples_out = c->global_samples;
    GB_END
}
int gb_chains_merge_timing(gb_chains* c, float* ms_out) {
    GB_TRY
    GB_LOCK(c);
    if (c->last_done_slot < 0) throw gb::Err("no completed merge to report on");
    CUDA_CHECK(cudaSetDevice(c->device));
    const auto& sl = c->slots[c->last_done_slot];
    for (int i = 0; i < 4; i++) CUDA_CHECK(cudaEventElapsedTime(ms_out + i, sl.ev_t[i], sl.ev_t[i + 1]));
    GB_END
}
int gb_chains_merge_partial_dev(gb_chains* c, double** dev_ptr_out, int64_t* n_out) {
    GB_TRY
    GB_LOCK(c);
    merge_partial(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the caller all-reduces the buffer on ITS stream
    *dev_ptr_out = c->d_merge;
    *n_out = c->base().total_card;
    GB_END
}
int gb_chains_merge_finalize(gb_chains* c, double* out, int32_t* collapsed_out) {
    GB_TRY
    GB_LOCK(c);
    merge_finalize(c, out, collapsed_out);
    GB_END
}

int gb_chains_convergence(gb_chains* c, int measure, const double* merged, double* out) {
    GB_TRY
    GB_LOCK(c);
    convergence_scores(c, measure, merged, out);
    GB_END
}
int gb_chains_convergence_partial_dev(gb_chains* c, int measure, const double* merged, double** dev_ptr_out,
                                      int64_t* n_out) {
    GB_TRY
    GB_LOCK(c);
    convergence_partial(c, measure, merged);
    *dev_ptr_out = c->d_wb;
    *n_out = (int64_t)2 * c->base().n_vars;
    GB_END
}
int gb_convergence_finalize(const gb_model* base, const double* wb, int32_t cw, int64_t total_chains,
                            const int32_t* collapsed, double* out) {
    GB_TRY
    const gb::HostModel& h = base->h;
    std::vector<uint8_t> col(h.n_vars, 0);
    for (int v = 0; v < h.n_vars; v++) col[v] = collapsed ? (collapsed[v] != 0) : 0;
    convergence_finalize(h, wb, cw, total_chains, col.data(), out);
    GB_END
}

// ---- (*ConvergenceSampler).Adapt (adaptive.go:57-157) in steps, so that a fleet can run them in lockstep over its devices
struct AdaptPlan {
    bool noop = false;          // adaptive.go:62-64 (MaxChains reached) or no candidates (:88-91)
    bool need_scores = false;   // more candidates than new chains: ChainConvergence decides (:100-119)
    std::vector<int32_t> cand;  // adaptive.go:81-87: not fixed, not collapsed in any chain, 1 < blanket <= 12 on the ORIGINAL graph
};
static AdaptPlan adapt_plan(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t max_groups) {
    AdaptPlan p;
    if ((int32_t)c->groups.size() >= max_groups) {
        p.noop = true;
        return p;
    }
    const gb::HostModel& b = base->h;
    std::vector<uint8_t> col = collapsed_any(c);
    for (int v = 0; v < b.n_vars; v++) {
        int sz = (int)b.nbrs[v].size();
        if (b.fixed[v] < 0 && !col[v] && sz > 1 && sz <= gb::kNeighborVarMax) p.cand.push_back(v);
    }
    if (p.cand.empty()) p.noop = true;
    else p.need_scores = (int32_t)p.cand.size() > new_chain_count;
    return p;
}
// adaptive.go:111-119: sort descending, take from the END (= lowest scores); ties by id
static std::vector<int32_t> adapt_pick(const AdaptPlan& p, const double* conv, int32_t new_chain_count) {
    if (!p.need_scores) return p.cand;
    std::vector<int32_t> cand = p.cand, targets;
    std::stable_sort(cand.begin(), cand.end(), [&](int32_t x, int32_t y) { return conv[x] > conv[y]; });
    for (int i = 0; i < new_chain_count; i++) targets.push_back(cand[cand.size() - 1 - i]);
    return targets;
}
// adaptive.go:130-154: one new group per chosen variable over a fresh clone of `base` with that variable collapsed
static int adapt_apply(gb_chains* c, const gb_model* base, const std::vector<int32_t>& targets, int32_t chains_local,
                       uint64_t first_chain_id, uint64_t id_stride, int32_t* chosen_out) {
    uint64_t first = first_chain_id;
    int n_done = 0;
    for (int32_t v : targets) {
        gb_model* nm = collapse_model(base, v, 0, nullptr, nullptr);
        try {
            add_group(c, nm, chains_local, first, true);
        } catch (...) {
            delete nm;
            throw;
        }
        first += id_stride ? id_stride : (uint64_t)((chains_local + 7) / 8 * 8);
        // adaptive.go:145: NewChain(..., burnIn=2) — two single-variable steps; one un-recorded
        // sweep (>= 2 updates) is the sweep-granular equivalent
        run_group(c, c->groups.back(), 1, 0, 0, -1);
        if (chosen_out) chosen_out[n_done] = v;
        n_done++;
    }
    CUDA_CHECK(cudaGetLastError());
    return n_done;
}
// this rank's share of `total` chains: contiguous, aligned to the blocks of 8 chains that share Philox calls
static void shard_chains(int64_t total, int world, int rank, uint64_t* first_out, int32_t* n_out) {
    const int64_t blocks = (total + 7) / 8, per = (blocks + world - 1) / world;
    const int64_t first = std::min<int64_t>((int64_t)rank * per, blocks) * 8, last = std::min<int64_t>((int64_t)(rank + 1) * per, blocks) * 8;
    *first_out = (uint64_t)first;
    *n_out = (int32_t)std::max<int64_t>(0, std::min<int64_t>(last, total) - first);
}

// scores == nullptr: ChainConvergence over this handle's chains — over the chains of every rank when a communicator
// is attached (collective; chains_per_new_model then counts a new variant's chains over ALL ranks and this rank takes
// its shard); otherwise the caller's per-variable scores with an explicit local share
static int adapt_impl(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                      int measure, int32_t cw, const double* scores, int64_t total_chains_hint, int32_t max_groups,
                      uint64_t first_chain_id, uint64_t id_stride, int32_t* chosen_out, int32_t* n_chosen_out) {
    if (n_chosen_out) *n_chosen_out = 0;
    GB_TRY
    GB_LOCK(c);
    CUDA_CHECK(cudaSetDevice(c->device));
    const bool collective = !scores && has_peers(c);
    if (!collective) {
        const int64_t total = total_chains_hint >= 0 ? total_chains_hint : local_chains(c);
        if (total < 2) throw gb::Err("At least 2 chains required for adaptation");
    }
    const AdaptPlan plan = adapt_plan(c, base, new_chain_count, max_groups);  // identical on every rank: same group list
    if (plan.noop) return 0;
    std::vector<double> conv;
    if (plan.need_scores) {
        if (scores) conv.assign(scores, scores + base->h.n_vars);
        else {
            conv.resize(base->h.n_vars);
            if (cw >= 2) c->last_cw = cw;
            convergence_scores(c, measure, nullptr, conv.data());
        }
    }
    const std::vector<int32_t> targets = adapt_pick(plan, conv.data(), new_chain_count);
    int32_t local = chains_per_new_model;
    uint64_t first = first_chain_id, stride = id_stride;
    if (collective) {
        uint64_t off = 0;
        shard_chains(chains_per_new_model, c->comm->world, c->comm->rank, &off, &local);
        if (local < 1) throw gb::Err("fewer than 8 chains per rank in a new variant: rank " + std::to_string(c->comm->rank) + " would hold none");
        first += off;
        stride = (uint64_t)((chains_per_new_model + 7) / 8 * 8);
    }
    const int n_done = adapt_apply(c, base, targets, local, first, stride, chosen_out);
    if (n_chosen_out) *n_chosen_out = n_done;
    GB_END
}

int gb_chains_adapt(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                    int measure, int32_t cw, int32_t max_groups, uint64_t first_chain_id, int32_t* chosen_out,
                    int32_t* n_chosen_out) {
    return adapt_impl(c, base, new_chain_count, chains_per_new_model, measure, cw, nullptr, -1, max_groups, first_chain_id, 0,
                      chosen_out, n_chosen_out);
}

int gb_chains_adapt_scores(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                           const double* scores, int64_t total_chains, int32_t max_groups, uint64_t first_chain_id,
                           uint64_t id_stride, int32_t* chosen_out, int32_t* n_chosen_out) {
    if (!scores) {
        g_err = "scores must not be NULL";
        return 1;
    }
    return adapt_impl(c, base, new_chain_count, chains_per_new_model, GB_HELLINGER, 0, scores, total_chains, max_groups,
                      first_chain_id, id_stride, chosen_out, n_chosen_out);
}

int gb_chains_get_state(gb_chains* c, int32_t group, int32_t* out) {
    GB_TRY
    GB_LOCK(c);
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    const int nv = g.model->h.n_vars;
    if (g.d_bits) {
        std::vector<uint32_t> wb((size_t)nv * g.n_words);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaMemcpy(wb.data(), g.d_bits, wb.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int ch = 0; ch < g.n_chains; ch++)
            for (int v = 0; v < nv; v++) out[(size_t)ch * nv + v] = (int32_t)((wb[(size_t)v * g.n_words + (ch >> 5)] >> (ch & 31)) & 1u);
        return 0;
    }
    std::vector<uint8_t> st((size_t)nv * g.n_pad);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(st.data(), g.d_state, st.size(), cudaMemcpyDeviceToHost));
    for (int ch = 0; ch < g.n_chains; ch++)
        for (int v = 0; v < nv; v++) out[(size_t)ch * nv + v] = st[(size_t)v * g.n_pad + ch];
    GB_END
}
int gb_chains_set_state(gb_chains* c, int32_t group, const int32_t* in) {
    GB_TRY
    GB_LOCK(c);
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    const gb::HostModel& h = g.model->h;
    if (g.d_bits) {
        std::vector<uint32_t> wb((size_t)h.n_vars * g.n_words, 0u);
        for (int ch = 0; ch < g.n_chains; ch++)
            for (int v = 0; v < h.n_vars; v++) {
                const int x = in[(size_t)ch * h.n_vars + v];
                if (x < 0 || x >= h.card[v]) throw gb::Err("Value " + std::to_string(x) + " invalid for cardinality " + std::to_string(h.card[v]));
                if (h.fixed[v] >= 0 && x != h.fixed[v]) throw gb::Err("state contradicts FixedVal of variable " + std::to_string(v));
                wb[(size_t)v * g.n_words + (ch >> 5)] |= (uint32_t)x << (ch & 31);
            }
        for (int v = 0; v < h.n_vars; v++)  // padding chains of the last word carry a fixed variable's value too
            if (h.fixed[v] > 0 && (g.n_chains & 31)) wb[(size_t)v * g.n_words + g.n_words - 1] |= ~0u << (g.n_chains & 31);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaMemcpy(g.d_bits, wb.data(), wb.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        return 0;
    }
    std::vector<uint8_t> st((size_t)h.n_vars * g.n_pad, 0);
    for (int ch = 0; ch < g.n_chains; ch++)
        for (int v = 0; v < h.n_vars; v++) {
            int x = in[(size_t)ch * h.n_vars + v];
            if (x < 0 || x >= h.card[v]) throw gb::Err("Value " + std::to_string(x) + " invalid for cardinality " + std::to_string(h.card[v]));
            if (h.fixed[v] >= 0 && x != h.fixed[v]) throw gb::Err("state contradicts FixedVal of variable " + std::to_string(v));
            st[(size_t)v * g.n_pad + ch] = (uint8_t)x;
        }
    for (int ch = g.n_chains; ch < g.n_pad; ch++)  // padding chains mirror the last real chain
        for (int v = 0; v < h.n_vars; v++) st[(size_t)v * g.n_pad + ch] = st[(size_t)v * g.n_pad + g.n_chains - 1];
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(g.d_state, st.data(), st.size(), cudaMemcpyHostToDevice));
    GB_END
}
int gb_chains_group_counts(gb_chains* c, int32_t group, uint64_t* out) {
    GB_TRY
    GB_LOCK(c);
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(out, g.d_counts, (size_t)g.model->h.total_card * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    GB_END
}
int gb_chains_group_history(gb_chains* c, int32_t group, uint16_t* out) {
    GB_TRY
    GB_LOCK(c);
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    if (!(c->flags & GB_CHAINS_HISTORY)) throw gb::Err("chains were created without GB_CHAINS_HISTORY");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    const int tc = g.model->h.total_card;
    std::vector<uint16_t> hh((size_t)2 * tc * g.n_pad);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(hh.data(), g.d_hist, hh.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < (size_t)2 * tc; r++)
        for (int ch = 0; ch < g.n_chains; ch++) out[r * g.n_chains + ch] = hh[r * g.n_pad + ch];
    GB_END
}

// ------------------------------------------------------------------ communicators and fleets
int gb_comm_unique_id(uint8_t* id_out) {
    GB_TRY
    gbn::UniqueId id;
    NCCL_CHECK(gbn::api().GetUniqueId(&id));
    std::memcpy(id_out, id.internal, gbn::kUniqueIdBytes);
    GB_END
}
int gb_comm_init_rank(const uint8_t* id, int32_t world, int32_t rank, int device, gb_comm** out) {
    GB_TRY
    if (world < 1 || rank < 0 || rank >= world) throw gb::Err("invalid communicator rank / size");
    require_device(device);
    auto cm = std::make_unique<gb_comm>();
    cm->world = world;
    cm->rank = rank;
    cm->device = device;
    if (world > 1) {
        if (!id) throw gb::Err("a communicator of more than one rank needs the unique id of gb_comm_unique_id");
        gbn::UniqueId uid;
        std::memcpy(uid.internal, id, gbn::kUniqueIdBytes);
        NCCL_CHECK(gbn::api().CommInitRank(&cm->nccl, world, uid, rank));
    }
    *out = cm.release();
    GB_END
}
int gb_comm_info(const gb_comm* cm, int32_t* world_out, int32_t* rank_out, int* device_out) {
    if (world_out) *world_out = cm->world;
    if (rank_out) *rank_out = cm->rank;
    if (device_out) *device_out = cm->device;
    return 0;
}
void gb_comm_destroy(gb_comm* cm) {
    if (!cm) return;
    if (cm->nccl) {
        cudaSetDevice(cm->device);
        try {
            gbn::api().CommDestroy(cm->nccl);
        } catch (...) {
        }
    }
    delete cm;
}
int gb_shard(int64_t total_chains, int32_t world, int32_t rank, uint64_t* first_out, int32_t* n_out) {
    GB_TRY
    if (total_chains < 0 || world < 1 || rank < 0 || rank >= world) throw gb::Err("invalid shard request");
    shard_chains(total_chains, world, rank, first_out, n_out);
    GB_END
}
int gb_chains_attach_comm(gb_chains* c, gb_comm* cm) {
    GB_TRY
    GB_LOCK(c);
    if (cm && cm->device != c->device) throw gb::Err("communicator rank is bound to a different device than the chains");
    if (c->slots_pending) throw gb::Err("a merge is pending on this handle");
    c->comm = cm;
    GB_END
}
int gb_chains_global_totals(const gb_chains* cc, int64_t* total_chains_out, int64_t* total_samples_out) {
    gb_chains* c = const_cast<gb_chains*>(cc);
    GB_TRY
    GB_LOCK(c);
    if (c->global_chains < 0) throw gb::Err("no merge has completed on this handle yet");
    if (total_chains_out) *total_chains_out = c->global_chains;
    if (total_samples_out) *total_samples_out = c->global_samples;
    GB_END
}

}  // extern "C"

// A fleet = the chain handles of ONE process's devices joined by a single-process communicator (ncclCommInitAll).
// Its collectives run each phase on every device before the next, with the NCCL calls of a phase inside one group, which
// is how a single thread drives several ranks.  Every device ends up with the same merged marginals / scores.
struct gb_fleet {
    std::mutex mu;
    std::vector<gb_comm*> comms;
    std::vector<gb_chains*> chains;  // borrowed, one per device (nullptr until attached)
    ~gb_fleet() {
        for (auto* cm : comms) gb_comm_destroy(cm);
    }
};
namespace {
struct FleetLock {  // the fleet's own lock, then every handle's, in device order
    std::lock_guard<std::mutex> f;
    std::vector<std::unique_lock<std::recursive_mutex>> h;
    explicit FleetLock(gb_fleet* fl) : f(fl->mu) {
        for (auto* c : fl->chains) {
            if (!c) throw gb::Err("fleet: a device has no chains attached");
            h.emplace_back(c->mu);
        }
    }
};
struct NcclGroup {
    bool on;
    explicit NcclGroup(bool enable) : on(enable) {
        if (on) NCCL_CHECK(gbn::api().GroupStart());
    }
    void end() {
        if (on) {
            on = false;
            NCCL_CHECK(gbn::api().GroupEnd());
        }
    }
    ~NcclGroup() {
        if (on) gbn::api().GroupEnd();
    }
};
void fleet_merge_begin(gb_fleet* f, double* out, int32_t* collapsed_out) {
    for (auto* c : f->chains) merge_snapshot(c);
    NcclGroup grp(f->chains.size() > 1);
    for (auto* c : f->chains) merge_reduce(c);
    grp.end();
    for (size_t i = 0; i < f->chains.size(); i++) merge_collect(f->chains[i], i == 0 ? out : nullptr, i == 0 ? collapsed_out : nullptr);
}
void fleet_merge_end(gb_fleet* f) {
    for (auto* c : f->chains) merge_wait(c);
}
void fleet_convergence(gb_fleet* f, int measure, const double* merged, double* out) {
    std::vector<double> tmp;
    if (!merged) {
        tmp.resize(f->chains[0]->base().total_card);
        fleet_merge_begin(f, tmp.data(), nullptr);
        fleet_merge_end(f);
        merged = tmp.data();
    }
    for (auto* c : f->chains) {
        if (c->last_cw < 2) throw gb::Err("Total seen < Convergence Window: run gb_fleet_advance first");
        for (const auto& g : c->groups)
            if (!g.window_filled) throw gb::Err("Total seen < Convergence Window: a chain group has not advanced since it was added");
        convergence_enqueue(c, measure, merged);
    }
    NcclGroup grp(f->chains.size() > 1);
    for (auto* c : f->chains) convergence_reduce(c);
    grp.end();
    std::vector<double> other(f->chains[0]->base().n_vars);
    for (size_t i = 0; i < f->chains.size(); i++) convergence_collect(f->chains[i], i == 0 ? out : other.data());
}
}  // namespace

extern "C" {

int gb_fleet_create(int32_t n_dev, const int* devices, gb_fleet** out) {
    GB_TRY
    if (n_dev < 1) throw gb::Err("a fleet needs at least one device");
    for (int i = 0; i < n_dev; i++) require_device(devices[i]);
    auto f = std::make_unique<gb_fleet>();
    std::vector<gbn::Comm> raw(n_dev, nullptr);
    if (n_dev > 1) NCCL_CHECK(gbn::api().CommInitAll(raw.data(), n_dev, devices));
    for (int i = 0; i < n_dev; i++) {
        auto* cm = new gb_comm();
        cm->nccl = raw[i];
        cm->world = n_dev;
        cm->rank = i;
        cm->device = devices[i];
        cm->in_fleet = true;
        f->comms.push_back(cm);
    }
    f->chains.assign(n_dev, nullptr);
    *out = f.release();
    GB_END
}
void gb_fleet_destroy(gb_fleet* f) { delete f; }
int gb_fleet_size(const gb_fleet* f, int32_t* n_out) {
    *n_out = (int32_t)f->comms.size();
    return 0;
}
int gb_fleet_attach(gb_fleet* f, int32_t i, gb_chains* c) {
    GB_TRY
    std::lock_guard<std::mutex> lk(f->mu);
    if (i < 0 || i >= (int32_t)f->comms.size()) throw gb::Err("fleet: device slot out of range");
    if (!c) throw gb::Err("fleet: no chains supplied");
    if (gb_chains_attach_comm(c, f->comms[i]) != 0) throw gb::Err(g_err);
    f->chains[i] = c;
    GB_END
}
// the per-device calls only enqueue work, so one thread keeps every device busy
int gb_fleet_sweep(gb_fleet* f, int64_t n_sweeps, int record) {
    GB_TRY
    FleetLock lk(f);
    for (auto* c : f->chains) sweeps(c, n_sweeps, record);
    GB_END
}
int gb_fleet_advance(gb_fleet* f, int32_t cw) {
    GB_TRY
    FleetLock lk(f);
    for (auto* c : f->chains)
        if (gb_chains_advance(c, cw) != 0) throw gb::Err(g_err);
    GB_END
}
int gb_fleet_synchronize(gb_fleet* f) {
    GB_TRY
    FleetLock lk(f);
    for (auto* c : f->chains) {
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    GB_END
}
int gb_fleet_merge_begin(gb_fleet* f, double* out, int32_t* collapsed_out) {
    GB_TRY
    FleetLock lk(f);
    fleet_merge_begin(f, out, collapsed_out);
    GB_END
}
int gb_fleet_merge_end(gb_fleet* f, int64_t* total_chains_out, int64_t* total_samples_out) {
    GB_TRY
    FleetLock lk(f);
    fleet_merge_end(f);
    if (total_chains_out) *total_chains_out = f->chains[0]->global_chains;
    if (total_samples_out) *total_samples_out = f->chains[0]->global_samples;
    GB_END
}
int gb_fleet_merged_marginals(gb_fleet* f, double* out, int32_t* collapsed_out) {
    GB_TRY
    FleetLock lk(f);
    fleet_merge_begin(f, out, collapsed_out);
    fleet_merge_end(f);
    GB_END
}
int gb_fleet_convergence(gb_fleet* f, int measure, const double* merged, double* out) {
    GB_TRY
    FleetLock lk(f);
    fleet_convergence(f, measure, merged, out);
    GB_END
}
int gb_fleet_adapt(gb_fleet* f, gb_model* const* bases, int32_t new_chain_count, int32_t chains_per_new_model, int measure,
                   int32_t cw, int32_t max_groups, uint64_t first_chain_id, int32_t* chosen_out, int32_t* n_chosen_out) {
    if (n_chosen_out) *n_chosen_out = 0;
    GB_TRY
    FleetLock lk(f);
    const int world = (int)f->chains.size();
    const AdaptPlan plan = adapt_plan(f->chains[0], bases[0], new_chain_count, max_groups);
    if (plan.noop) return 0;
    std::vector<double> conv(bases[0]->h.n_vars, 0.0);
    if (plan.need_scores) {
        if (cw >= 2)
            for (auto* c : f->chains) c->last_cw = cw;
        fleet_convergence(f, measure, nullptr, conv.data());
    }
    const std::vector<int32_t> targets = adapt_pick(plan, conv.data(), new_chain_count);
    const uint64_t stride = (uint64_t)((chains_per_new_model + 7) / 8 * 8);
    int n_done = 0;
    for (int i = 0; i < world; i++) {
        uint64_t off = 0;
        int32_t local = 0;
        shard_chains(chains_per_new_model, world, i, &off, &local);
        if (local < 1) throw gb::Err("fewer than 8 chains per device in a new variant: device slot " + std::to_string(i) + " would hold none");
        CUDA_CHECK(cudaSetDevice(f->chains[i]->device));
        n_done = adapt_apply(f->chains[i], bases[i], targets, local, first_chain_id + off, stride, i == 0 ? chosen_out : nullptr);
    }
    if (n_chosen_out) *n_chosen_out = n_done;
    GB_END
}

// ------------------------------------------------------------------ scoring (host)
int gb_error_suite(int32_t n_vars, const int32_t* card, const int32_t* fixed1, const double* marg1,
                   const int32_t* fixed2, const double* marg2, double* out8) {
    GB_TRY
    int cnt = 0;
    for (int v = 0; v < n_vars; v++)
        if ((!fixed1 || fixed1[v] < 0) && (!fixed2 || fixed2[v] < 0)) cnt++;
    if (cnt < 1) throw gb::Err("No un-fixed vars to score");
    double mean[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
    size_t off = 0;
    for (int v = 0; v < n_vars; v++) {
        const bool fx = (fixed1 && fixed1[v] >= 0) || (fixed2 && fixed2[v] >= 0);
        // error.go order: MeanAbs, MaxAbs, Hellinger, JS
        const int which[4] = {GB_MEAN_ABS, GB_MAX_ABS, GB_HELLINGER, GB_JS};
        for (int i = 0; i < 4; i++) {
            double d = fx ? 0.0 : gb::measure_host(which[i], marg1 + off, marg2 + off, card[v]);
            mean[i] += d;
            mx[i] = std::fmax(d, mx[i]);
        }
        off += card[v];
    }
    for (int i = 0; i < 4; i++) mean[i] /= (double)cnt;
    out8[0] = mean[0]; out8[1] = mx[0]; out8[2] = mean[1]; out8[3] = mx[1];
    out8[4] = mean[2]; out8[5] = mx[2]; out8[6] = mean[3]; out8[7] = mx[3];
    GB_END
}
int gb_mar_load(const char* path, int32_t* n_vars_out, int32_t* total_card_out, int32_t* card_out, double* marg_out) {
    GB_TRY
    std::vector<int32_t> card;
    std::vector<double> marg;
    gb::load_mar(path, card, marg);
    if (n_vars_out) *n_vars_out = (int32_t)card.size();
    if (total_card_out) *total_card_out = (int32_t)marg.size();
    if (card_out) std::copy(card.begin(), card.end(), card_out);
    if (marg_out) std::copy(marg.begin(), marg.end(), marg_out);
    GB_END
}

}  // extern "C"
#endif  // GB_MAIN
