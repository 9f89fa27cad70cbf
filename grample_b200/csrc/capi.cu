// C ABI of grample_b200 (include/grample_b200.h): host orchestration around the kernels in
// kernels.cuh.  There is no CPU fallback: every compute entry needs a CUDA device and fails
// loudly without one.  Host-only models (device = -1) exist for schedule / blanket inspection.
#include "../../include/grample_b200.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "host_model.hpp"
#include "kernels.cuh"

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

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1)
        throw gb::Err("grample_b200: no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= n) throw gb::Err("grample_b200: invalid device index " + std::to_string(device));
    CUDA_CHECK(cudaSetDevice(device));
}

int grid_for(int64_t items, int threads) {
    int64_t blocks = (items + threads - 1) / threads;
    const int64_t cap = 148 * 8;  // B200: 148 SMs x 8 resident 256-thread CTAs
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

constexpr int kMaxDevices = 64;  // per-device caches of launch configuration

}  // namespace

struct gb_model {
    gb::HostModel h;
    int device = -1;
    std::vector<void*> allocs;
    gb::DevModel dev{};
    int32_t* d_order = nullptr;
    int32_t* d_colour_off = nullptr;
    gb::DevTab tab{};
    bool tab_built = false;

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
    void ensure_tab() {
        if (!h.tab_ok) throw gb::Err("table mode does not apply to this model: " + h.tab_why);
        if (tab_built) return;
        ensure_thresholds();
        tab.trec = up(h.trec);
        tab_built = true;
    }
    // hybrid mode tabulates what qualifies (binary, <= 65536 configurations) in models whose cardinalities are <= 4
    bool hybrid_tables() const { return h.max_card <= 4 && h.n_tab_vars > 0; }
    // hybrid mode on a model where EVERY sampled variable got a table (plain binary models and their
    // single-collapsed variants): the resident table kernel runs it, wide variables through their tprog entry
    bool hybrid_all_tables() const { return hybrid_tables() && h.tab_all; }
    void ensure_hybrid() {
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
    uint8_t* d_state = nullptr;
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
            cudaFree(g.d_counts);
            cudaFree(g.d_hist);
            if (g.owns_model) delete g.model;
        }
        cudaFree(d_merge);
        cudaFree(d_wb);
        cudaFree(d_skip);
        cudaFree(d_merged_in);
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
    if (c->precision == GB_TABLE) model->ensure_tab();
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
    CUDA_CHECK(cudaMalloc(&g.d_state, (size_t)h.n_vars * g.n_pad));
    CUDA_CHECK(cudaMalloc(&g.d_counts, (size_t)h.total_card * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemsetAsync(g.d_counts, 0, (size_t)h.total_card * sizeof(unsigned long long), c->stream));
    if (c->flags & GB_CHAINS_HISTORY) {
        size_t hb = (size_t)2 * h.total_card * g.n_pad * sizeof(uint16_t);
        CUDA_CHECK(cudaMalloc(&g.d_hist, hb));
        CUDA_CHECK(cudaMemsetAsync(g.d_hist, 0, hb, c->stream));
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
    const int64_t items = (int64_t)h.n_vars * (g.n_pad / 4);
    gb::k_init_state<<<grid_for(items, 256), 256, 0, c->stream>>>(model->dev, g.dev);
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
    int& resident = resident_dev[c->device % kMaxDevices];
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
        } else if (c->precision == GB_F32) {
            gbh::lse_colour_f32(c, g, dv, n, record, hist_half);
        } else {
            gbh::lse_colour_f64(c, g, dv, n, record, hist_half);
        }
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
    size_t& configured = configured_dev[c->device % kMaxDevices];
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
    static int disabled = -1, no_ts = -1;
    if (disabled < 0) disabled = std::getenv("GB_NO_RESIDENT") ? 1 : 0;
    if (no_ts < 0) no_ts = std::getenv("GB_NO_SMEM_TABLES") ? 1 : 0;  // A/B knob
    ResidentPlan p;
    if (disabled || (c->flags & GB_CHAINS_PER_COLOUR)) return p;
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
    // 1) tables staged in shared memory, if every CTA of the launch is still co-resident (these models
    //    expose little parallelism per colour, so a second wave of CTAs costs more than LDS tables gain)
    if (!no_ts)
        for (int ch : {8, 16, 32, 64}) {
            if (g.n_pad % ch) continue;
            const size_t smem = base(ch) + tab_bytes;
            if (smem > kSmemPerCta) break;
            const int64_t per_sm = std::min<int64_t>(8, (int64_t)(kSmemPerSm / (smem + 1024)));
            if (ch < 64 && ctas(ch) > per_sm * sms) continue;
            if ((int64_t)(g.n_pad / ch) > per_sm * sms) break;
            p.ch = ch; p.smem = smem; p.ts = true;
            p.n_stage = (int32_t)(((size_t)h.log_tab.size() + 3) & ~(size_t)3);
            return p;
        }
    // 1b) one-thread-per-chain kernels (cardinality >= 8): stage the longest PREFIX of the tables that fits and ends on
    //     a factor boundary.  A collapsed variant keeps the model's small factors first and appends the large factor
    //     over the collapsed variable's blanket (up to 11^6 entries on ObjectDetection_11), so every small factor is
    //     served from shared memory and only the large one goes through L1/L2.
    if (!no_ts && h.max_card >= 8)
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
            if (ch < 64 && ctas(ch) > per_sm * sms) continue;
            if ((int64_t)(g.n_pad / ch) > per_sm * sms) break;
            p.ch = ch; p.smem = smem; p.n_stage = (int32_t)n;
            return p;
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
    static size_t configured_dev[2][kMaxDevices] = {};  // function attributes are per device
    const bool wide = h.tab_max_nbr > 8;
    size_t& configured = configured_dev[wide][c->device % kMaxDevices];
    auto kernel = wide ? gb::k_sweep_tab_resident<true> : gb::k_sweep_tab_resident<false>;
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

// (*Chain).AdvanceChain for one group: cw + 1 recorded sweeps, the last 2 * (cw / 2) into the window
void advance_group(gb_chains* c, Group& g, int32_t cw) {
    const int32_t half = cw / 2;
    if (c->flags & GB_CHAINS_HISTORY) {
        if (half > 65535) throw gb::Err("convergence window too large for 16-bit half-window histograms");
        CUDA_CHECK(cudaMemsetAsync(g.d_hist, 0, (size_t)2 * g.model->h.total_card * g.n_pad * sizeof(uint16_t), c->stream));
    }
    run_group(c, g, (int64_t)cw + 1, 1, cw + 1 - 2 * half, half);
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
    if (!c->d_wb) CUDA_CHECK(cudaMalloc(&c->d_wb, (size_t)2 * h.n_vars * sizeof(double)));
    if (!c->d_skip) CUDA_CHECK(cudaMalloc(&c->d_skip, (size_t)h.n_vars));
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
    CUDA_CHECK(cudaMemcpyAsync(c->d_skip, c->skip_bits.data(), c->skip_bits.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->col_any32.assign(c->col_any.begin(), c->col_any.end());
    c->skip_uploaded = true;
}

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
    // a caller buffer that is page-locked (cudaHostAlloc / cudaHostRegister / a pinned torch tensor) receives the DMA
    // directly; a pageable one goes through the handle's pinned staging buffer
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();  // (older drivers report an unregistered host pointer as an error)
    if (pinned) {
        CUDA_CHECK(cudaMemcpyAsync(out, c->d_merge, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
        if (!c->h_merge) CUDA_CHECK(cudaMallocHost(&c->h_merge, bytes));
        CUDA_CHECK(cudaMemcpyAsync(c->h_merge, c->d_merge, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::memcpy(out, c->h_merge, bytes);
    }
    refresh_collapsed_cache(c);
    if (collapsed_out) {
        upload_skip(c);
        std::memcpy(collapsed_out, c->col_any32.data(), (size_t)h.n_vars * sizeof(int32_t));
    }
    for (int v : c->col_vars) {
        const auto& m = c->groups[c->col_first_group[v]].model->h.coll_marg[v];  // chain.go:113-129: first chain found
        for (int k = 0; k < h.card[v]; k++) out[h.card_off[v] + k] = m[k];
    }
}

void convergence_partial(gb_chains* c, int measure, const double* merged) {
    CUDA_CHECK(cudaSetDevice(c->device));
    if (!(c->flags & GB_CHAINS_HISTORY)) throw gb::Err("chains were created without GB_CHAINS_HISTORY");
    if (measure < 0 || measure > 3) throw gb::Err("unknown measure");
    ensure_scratch(c);
    const gb::HostModel& h = c->base();
    std::vector<double> tmp;
    if (!merged) {
        merge_partial(c);
        tmp.resize(h.total_card);
        merge_finalize(c, tmp.data(), nullptr);
        merged = tmp.data();
    }
    upload_skip(c);
    CUDA_CHECK(cudaMemcpyAsync(c->d_merged_in, merged, (size_t)h.total_card * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->d_wb, 0, (size_t)2 * h.n_vars * sizeof(double), c->stream));
    for (auto& g : c->groups) {
        const int64_t items = (int64_t)h.n_vars * g.n_chains;
        gb::k_chain_dist<<<grid_for(items, 256), 256, 0, c->stream>>>(g.model->dev, g.dev, c->d_merged_in, c->d_skip,
                                                                     measure, c->d_wb);
        c->launches++;
    }
    CUDA_CHECK(cudaGetLastError());
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
    if ((int)blanket.size() > gb::kNeighborVarMaxDev) throw gb::Err("blanket exceeds the device limit");

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
    int32_t* d_off = dev_upload(f_tab_off);
    int32_t* d_sv = dev_upload(f_stride_v);
    int32_t* d_sb = dev_upload(f_stride_b);
    pl.f_tab_off = d_off;
    pl.f_stride_v = d_sv;
    pl.f_stride_b = d_sb;
    double *d_new = nullptr, *d_marg = nullptr;
    CUDA_CHECK(cudaMalloc(&d_new, (size_t)new_size * sizeof(double)));
    std::vector<double> marg(pl.card_v, 1e-12);  // line 138-140
    d_marg = dev_upload(marg);
    gb::k_collapse<<<(int)((new_size + 255) / 256), 256>>>(pl, src->dev.tab64, d_new, d_marg);
    CUDA_CHECK(cudaGetLastError());
    std::vector<double> new_tab((size_t)new_size);
    CUDA_CHECK(cudaMemcpy(new_tab.data(), d_new, (size_t)new_size * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(marg.data(), d_marg, marg.size() * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_off); cudaFree(d_sv); cudaFree(d_sb); cudaFree(d_new); cudaFree(d_marg);
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
    if (m->h.tp_off[var] < 0) throw gb::Err("variable has no threshold table (not sampled, not binary, or too many neighbour configurations)");
    const int32_t* tp = m->h.tprog.data() + m->h.tp_off[var];
    int n = 1;
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

int gb_conditional(const gb_model* m, int precision, int32_t n_states, const int32_t* states,
                   const int32_t* vars, double* out) {
    GB_TRY
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
    int32_t *d_states = nullptr, *d_vars = nullptr;
    double* d_out = nullptr;
    const size_t ns = (size_t)n_states;
    CUDA_CHECK(cudaMalloc(&d_states, ns * h.n_vars * sizeof(int32_t)));
    CUDA_CHECK(cudaMalloc(&d_vars, ns * sizeof(int32_t)));
    CUDA_CHECK(cudaMalloc(&d_out, ns * gb::kProbeStride * sizeof(double)));
    CUDA_CHECK(cudaMemcpy(d_states, states, ns * h.n_vars * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(d_vars, vars, ns * sizeof(int32_t), cudaMemcpyHostToDevice));
    const int blocks = (n_states + 127) / 128;
    if (precision == GB_F32) gb::k_conditional<float><<<blocks, 128>>>(m->dev, n_states, d_states, d_vars, d_out);
    else gb::k_conditional<double><<<blocks, 128>>>(m->dev, n_states, d_states, d_vars, d_out);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(out, d_out, ns * gb::kProbeStride * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_states); cudaFree(d_vars); cudaFree(d_out);
    GB_END
}

// ------------------------------------------------------------------ chains
int gb_chains_create(int32_t n_groups, gb_model* const* models, const int32_t* chains_per_model,
                     uint64_t seed, uint64_t first_chain_id, int precision, uint32_t flags, int device,
                     gb_chains** out) {
    GB_TRY
    if (n_groups < 1) throw gb::Err("at least one chain group is required");
    if (precision != GB_F64 && precision != GB_F32 && precision != GB_TABLE && precision != GB_HYBRID) throw gb::Err("unknown precision");
    if ((flags & GB_CHAINS_RAO_BLACKWELL) && precision != GB_F64 && precision != GB_F32)
        throw gb::Err("GB_CHAINS_RAO_BLACKWELL needs precision GB_F64 or GB_F32 (the estimator accumulates the log-sum-exp conditionals)");
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
    CUDA_CHECK(cudaSetDevice(c->device));
    add_group(c, model, n_chains, first_chain_id, false);
    GB_END
}
void gb_chains_destroy(gb_chains* c) { delete c; }
int gb_chains_n_groups(const gb_chains* c, int32_t* out) { *out = (int32_t)c->groups.size(); return 0; }
int gb_chains_n_chains(const gb_chains* c, int64_t* out) {
    int64_t n = 0;
    for (auto& g : c->groups) n += g.n_chains;
    *out = n;
    return 0;
}

int gb_chains_sweep(gb_chains* c, int64_t n_sweeps, int record) {
    GB_TRY sweeps(c, n_sweeps, record);
    GB_END
}
int gb_chains_sweep_timed(gb_chains* c, int64_t n_sweeps, int record, float* ms_out) {
    GB_TRY
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
int gb_chains_launch_count(const gb_chains* c, int64_t* out) { *out = c->launches; return 0; }
int gb_chains_scan(gb_chains* c, int64_t n_steps, int record) {
    GB_TRY
    if (n_steps < 0) throw gb::Err("Invalid step count");
    if (c->flags & GB_CHAINS_RAO_BLACKWELL) throw gb::Err("the random-scan parity mode records plain counts (no GB_CHAINS_RAO_BLACKWELL)");
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
    GB_TRY sweeps(c, n_sweeps, 0);
    GB_END
}
int gb_chains_advance(gb_chains* c, int32_t cw) {
    GB_TRY
    if (cw < 0) throw gb::Err("Invalid convergence window");
    CUDA_CHECK(cudaSetDevice(c->device));
    c->last_cw = cw;
    for_each_group_concurrent(c, [&](Group& g) { advance_group(c, g, cw); });
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_total_samples(const gb_chains* c, int64_t* out) { *out = c->total_samples; return 0; }

// ---- per-group forms: one reference Chain maps to one group of replica chains
static Group& group_at(gb_chains* c, int32_t group) {
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    return c->groups[group];
}
int gb_chains_group_sweep(gb_chains* c, int32_t group, int64_t n_sweeps, int record) {
    GB_TRY
    CUDA_CHECK(cudaSetDevice(c->device));
    run_group(c, group_at(c, group), n_sweeps, record, 0, -1);
    CUDA_CHECK(cudaGetLastError());
    GB_END
}
int gb_chains_group_advance(gb_chains* c, int32_t group, int32_t cw) {
    GB_TRY
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
    Group& g = group_at(c, group);
    if (n_chains_out) *n_chains_out = g.n_chains;
    if (total_samples_out) *total_samples_out = g.total_samples;
    if (model_out) *model_out = g.model;
    GB_END
}
int gb_chains_synchronize(gb_chains* c) {
    GB_TRY
    CUDA_CHECK(cudaSetDevice(c->device));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    GB_END
}

int gb_chains_merged_marginals(gb_chains* c, double* out, int32_t* collapsed_out) {
    GB_TRY
    merge_partial(c);
    merge_finalize(c, out, collapsed_out);
    GB_END
}
int gb_chains_merge_partial_dev(gb_chains* c, double** dev_ptr_out, int64_t* n_out) {
    GB_TRY
    merge_partial(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the caller all-reduces the buffer on ITS stream
    *dev_ptr_out = c->d_merge;
    *n_out = c->base().total_card;
    GB_END
}
int gb_chains_merge_finalize(gb_chains* c, double* out, int32_t* collapsed_out) {
    GB_TRY merge_finalize(c, out, collapsed_out);
    GB_END
}

int gb_chains_convergence(gb_chains* c, int measure, const double* merged, double* out) {
    GB_TRY
    int64_t total = 0;
    for (auto& g : c->groups) total += g.n_chains;
    if (total < 2) throw gb::Err("Convergence requires at least 2 chains");
    if (c->last_cw < 2) throw gb::Err("Total seen < Convergence Window: run gb_chains_advance first");
    convergence_partial(c, measure, merged);
    const gb::HostModel& h = c->base();
    std::vector<double> wb((size_t)2 * h.n_vars);
    CUDA_CHECK(cudaMemcpy(wb.data(), c->d_wb, wb.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<uint8_t> col = collapsed_any(c);
    convergence_finalize(h, wb.data(), c->last_cw, total, col.data(), out);
    GB_END
}
int gb_chains_convergence_partial_dev(gb_chains* c, int measure, const double* merged, double** dev_ptr_out,
                                      int64_t* n_out) {
    GB_TRY
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

// scores == nullptr: ChainConvergence over this handle's chains; otherwise the caller's per-variable
// scores (multi-GPU: computed from the all-reduced within/between sums, identical on every rank, so
// every rank chooses the same variables); total_chains_hint < 0: this handle's chain count
static int adapt_impl(gb_chains* c, const gb_model* base, int32_t new_chain_count, int32_t chains_per_new_model,
                      int measure, int32_t cw, const double* scores, int64_t total_chains_hint, int32_t max_groups,
                      uint64_t first_chain_id, uint64_t id_stride, int32_t* chosen_out, int32_t* n_chosen_out) {
    if (n_chosen_out) *n_chosen_out = 0;
    GB_TRY
    CUDA_CHECK(cudaSetDevice(c->device));
    int64_t total = 0;
    for (auto& g : c->groups) total += g.n_chains;
    if (total_chains_hint >= 0) total = total_chains_hint;
    if (total < 2) throw gb::Err("At least 2 chains required for adaptation");
    if ((int32_t)c->groups.size() >= max_groups) return 0;  // adaptive.go:62-64
    const gb::HostModel& b = base->h;
    std::vector<uint8_t> col = collapsed_any(c);
    std::vector<int32_t> cand;  // adaptive.go:81-87: blanket sizes on the ORIGINAL graph
    for (int v = 0; v < b.n_vars; v++) {
        int sz = (int)b.nbrs[v].size();
        if (b.fixed[v] < 0 && !col[v] && sz > 1 && sz <= gb::kNeighborVarMax) cand.push_back(v);
    }
    if (cand.empty()) return 0;
    std::vector<int32_t> targets;
    if ((int32_t)cand.size() <= new_chain_count) {
        targets = cand;
    } else {
        std::vector<double> conv(b.n_vars);
        if (scores) {
            conv.assign(scores, scores + b.n_vars);
        } else {
            convergence_partial(c, measure, nullptr);
            std::vector<double> wb((size_t)2 * b.n_vars);
            CUDA_CHECK(cudaMemcpy(wb.data(), c->d_wb, wb.size() * sizeof(double), cudaMemcpyDeviceToHost));
            convergence_finalize(b, wb.data(), cw, total, col.data(), conv.data());
        }
        // adaptive.go:111-119: sort descending, take from the END (= lowest scores); ties by id
        std::stable_sort(cand.begin(), cand.end(), [&](int32_t x, int32_t y) { return conv[x] > conv[y]; });
        for (int i = 0; i < new_chain_count; i++) targets.push_back(cand[cand.size() - 1 - i]);
    }
    uint64_t first = first_chain_id;
    int n_done = 0;
    for (int32_t v : targets) {
        gb_model* nm = collapse_model(base, v, 0, nullptr, nullptr);
        try {
            add_group(c, nm, chains_per_new_model, first, true);
        } catch (...) {
            delete nm;
            throw;
        }
        first += id_stride ? id_stride : (uint64_t)((chains_per_new_model + 7) / 8 * 8);
        // adaptive.go:145: NewChain(..., burnIn=2) — two single-variable steps; one un-recorded
        // sweep (>= 2 updates) is the sweep-granular equivalent
        run_group(c, c->groups.back(), 1, 0, 0, -1);
        if (chosen_out) chosen_out[n_done] = v;
        n_done++;
    }
    CUDA_CHECK(cudaGetLastError());
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
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    const int nv = g.model->h.n_vars;
    std::vector<uint8_t> st((size_t)nv * g.n_pad);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(st.data(), g.d_state, st.size(), cudaMemcpyDeviceToHost));
    for (int ch = 0; ch < g.n_chains; ch++)
        for (int v = 0; v < nv; v++) out[(size_t)ch * nv + v] = st[(size_t)v * g.n_pad + ch];
    GB_END
}
int gb_chains_set_state(gb_chains* c, int32_t group, const int32_t* in) {
    GB_TRY
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    const gb::HostModel& h = g.model->h;
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
    if (group < 0 || group >= (int32_t)c->groups.size()) throw gb::Err("group index out of range");
    CUDA_CHECK(cudaSetDevice(c->device));
    Group& g = c->groups[group];
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(out, g.d_counts, (size_t)g.model->h.total_card * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    GB_END
}
int gb_chains_group_history(gb_chains* c, int32_t group, uint16_t* out) {
    GB_TRY
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
