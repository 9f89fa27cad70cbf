// Host side of the flattened model: what the reference keeps in *model.Model plus the sampler
// bookkeeping of GibbsSimple / GibbsCollapsed, laid out as CSR arrays ready for upload.
//
// Reference behaviour implemented here (paths relative to the reference root):
//   model/function.go:126-142   log-space conversion with the `v < 1e-6 -> v += 1e-6` rule
//   model/function.go:180-202   table index: first scope variable most significant
//   sampler/gibbs-simple.go:73-99   var -> factor lists in m.Funcs order, validation
//   sampler/gibbs-collapsed.go:44-78  neighbour sets (blanket includes the variable itself)
//   model/uai.go:20-332, model/reader.go   UAI / evidence / MAR text formats
//   model/error.go:28-249       distance measures and the ErrorSuite
// New here (no reference counterpart): the greedy colouring that lets non-adjacent variables
// of one chain update concurrently, and the per-variable "update program" the kernels read.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace gb {

struct Err : std::runtime_error {
    using std::runtime_error::runtime_error;
};

constexpr int kNeighborVarMax = 12;
constexpr int64_t kMaxTabSize = 1 << 23;
constexpr int kMaxCard = 64;
constexpr int64_t kTabWideCfg = 65536;  // hybrid mode tabulates binary variables with up to this many neighbour configurations
constexpr int kTabTile = 64;  // sweep positions per CTA tile of k_sweep_tab (<= 4 neighbours); the locality order groups by it

struct Factor {
    std::vector<int32_t> vars;     // scope, most significant first
    std::vector<int64_t> strides;  // stride of each scope variable in the table
    int64_t off = 0;               // offset into HostModel::log_tab
    int64_t size = 0;
};

struct HostModel {
    int32_t n_vars = 0;
    std::vector<int32_t> card, fixed;
    std::vector<uint8_t> collapsed;
    std::vector<std::vector<double>> coll_marg;  // local marginal of collapsed variables
    std::vector<Factor> funcs;
    std::vector<double> log_tab;  // all factor tables, log space, concatenated

    // ---- derived (build_derived)
    std::vector<int32_t> card_off;                 // [n_vars+1]
    int32_t total_card = 0, max_card = 0;
    std::vector<std::vector<int32_t>> var_funcs;   // per var, in funcs order
    std::vector<std::vector<int32_t>> nbrs;        // per var, ascending, includes the var itself
    std::vector<int32_t> colour;                   // per var, -1 if never sampled
    std::vector<int32_t> order, colour_off;        // colour-sorted sweep schedule
    std::vector<int32_t> prog_off, prog;           // per-variable update program (see kernels.cuh)
    std::vector<int32_t> pw_off, pw_rec;           // pairwise fast path: per variable 4-word records, pw_off = -1 when a factor has arity > 2
    // ---- tabulated-conditional fast path (binary sampled variables, <= 256 neighbour configurations)
    bool tab_ok = false, tab_all = false, tab_all_binary = true;
    std::string tab_why;                           // why the fast path does not apply
    bool bits_ok = false;                          // GB_TABLE_BITS applies: tab_ok and every sampled variable has <= 4 free neighbours, all binary
    std::string bits_why;
    std::vector<int32_t> tp_off, tprog;            // per var: [n_nbr, thr_off, (nbr_var, stride) * n_nbr]
    std::vector<int32_t> trec;                     // per sweep position: kTabRec words (see kernels.cuh)
    int64_t n_thresholds = 0;
    int tab_max_nbr = 0;                           // largest n_nbr over the sampled variables
    int32_t n_tab_vars = 0;                        // sampled variables with a threshold table (binary, <= kTabWideCfg configurations)

    bool sampled(int v) const { return fixed[v] < 0 && !collapsed[v]; }

    void add_factor(const std::vector<int32_t>& scope, const double* raw, int64_t n_raw, bool already_log) {
        if (scope.empty()) throw Err("Empty variable list for function is invalid");
        Factor f;
        f.vars = scope;
        f.strides.assign(scope.size(), 0);
        int64_t sz = 1;
        for (int64_t i = (int64_t)scope.size() - 1; i >= 0; i--) {
            int32_t v = scope[i];
            if (v < 0 || v >= n_vars) throw Err("Invalid var idx " + std::to_string(v) + " in function scope");
            // The reference accepts a scope that names a variable twice and then reads BOTH slots from the state
            // (function.go:180-202), i.e. only the table's diagonal in that pair is reachable.  None of the UAI
            // fixtures does this and the fast paths here assume distinct scope variables, so it is rejected up front.
            for (int64_t j = i + 1; j < (int64_t)scope.size(); j++)
                if (scope[j] == v) throw Err("Variable " + std::to_string(v) + " appears twice in a function scope (not supported)");
            f.strides[i] = sz;
            sz *= card[v];
            if (sz > kMaxTabSize) throw Err("Function over " + std::to_string(scope.size()) + " vars has size > " + std::to_string(kMaxTabSize));
        }
        if (sz != n_raw) throw Err("Read table size " + std::to_string(n_raw) + " != Clique size " + std::to_string(sz));
        f.off = (int64_t)log_tab.size();
        f.size = sz;
        const double eps = 1e-6;
        for (int64_t i = 0; i < sz; i++) {
            double v = raw[i];
            if (!already_log) {
                if (v < eps) v += eps;
                v = std::log(v);
            }
            log_tab.push_back(v);
        }
        funcs.push_back(std::move(f));
    }

    void build_derived() {
        card_off.assign(n_vars + 1, 0);
        max_card = 0;
        for (int v = 0; v < n_vars; v++) {
            card_off[v + 1] = card_off[v] + card[v];
            max_card = std::max(max_card, card[v]);
        }
        total_card = card_off[n_vars];
        var_funcs.assign(n_vars, {});
        nbrs.assign(n_vars, {});
        for (size_t fi = 0; fi < funcs.size(); fi++)
            for (int32_t v : funcs[fi].vars) var_funcs[v].push_back((int32_t)fi);
        for (int v = 0; v < n_vars; v++) {
            auto& nb = nbrs[v];
            for (int32_t fi : var_funcs[v])
                for (int32_t u : funcs[fi].vars) nb.push_back(u);
            std::sort(nb.begin(), nb.end());
            nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        }
        for (int v = 0; v < n_vars; v++) {
            if (collapsed[v]) {
                if (!nbrs[v].empty()) throw Err("Var is collapsed but has a blanket");
            } else if (var_funcs[v].empty()) {
                throw Err("There are no functions for var (ID=" + std::to_string(v) + ")");
            }
        }
        build_colouring();
        build_programs();
        build_tab_programs();
    }

    // Tabulated conditionals: for a sampled variable of cardinality <= 4 whose distinct free neighbours span few joint
    // configurations, the whole conditional (gibbs-simple.go:171-258) depends only on that configuration, so it is
    // evaluated once per configuration (k_build_thresholds) and stored as card - 1 cumulative 32-bit inverse-CDF
    // thresholds; the sweep becomes: gather neighbour bytes -> configuration index -> thresholds -> compare.
    // Fixed neighbours are folded into the table (their value never changes).
    //   GB_TABLE  (tab_ok):  every sampled variable BINARY with <= 256 configurations (the integer bench kernels)
    //   GB_HYBRID (tp_off):  per variable, cardinality <= 4 and <= kTabWideCfg configurations; the rest by log-sum-exp
    void build_tab_programs() {
        tab_ok = true;
        tab_why.clear();
        tp_off.assign(n_vars, -1);
        tprog.clear();
        trec.clear();
        n_thresholds = 0;
        n_tab_vars = 0;
        tab_max_nbr = 0;
        tab_all_binary = true;
        for (int v : order) {
            std::string why;
            if (card[v] != 2) {
                why = "variable " + std::to_string(v) + " has cardinality " + std::to_string(card[v]) + " (table mode needs binary sampled variables)";
                tab_all_binary = false;
            }
            int64_t cfgs = 1;
            std::vector<int32_t> words;
            if (card[v] <= 4)
                for (int32_t u : nbrs[v]) {
                    if (u == v || fixed[u] >= 0 || card[u] == 1) continue;  // constant neighbours fold into the table
                    words.push_back(u);
                    words.push_back((int32_t)cfgs);
                    cfgs *= card[u];
                    if (cfgs > kTabWideCfg) break;
                }
            if (why.empty() && cfgs > 256) why = "variable " + std::to_string(v) + " has more than 256 neighbour configurations";
            if (!why.empty() && tab_ok) {
                tab_ok = false;  // GB_TABLE needs every sampled variable binary and narrow; GB_HYBRID takes what qualifies
                tab_why = why;
            }
            if (card[v] > 4 || cfgs > kTabWideCfg) continue;
            tp_off[v] = (int32_t)tprog.size();
            tprog.push_back((int32_t)words.size() / 2);
            tprog.push_back((int32_t)n_thresholds);
            tprog.insert(tprog.end(), words.begin(), words.end());
            n_thresholds += cfgs * (card[v] - 1);  // card - 1 cumulative thresholds per configuration
            n_tab_vars++;
        }
        // tab_all: every sampled variable has a table (any width) — the resident table kernel can run the whole
        // model (wide variables through their variable-length tprog entry); tab_ok additionally has them all binary and narrow
        tab_all = n_tab_vars == (int32_t)order.size() && n_tab_vars > 0;
        if (!tab_all) return;
        // fixed-size record per sweep position: {v, thr_off, n_nbr | card << 8, card_off, nbr[8], stride[8]}
        // (narrow = at most 8 neighbours and 256 configurations: the index fits the byte-packed arithmetic)
        trec.assign(order.size() * 20, 0);
        for (size_t j = 0; j < order.size(); j++) {
            const int v = order[j];
            const int32_t* tp = tprog.data() + tp_off[v];
            int32_t* r = trec.data() + j * 20;
            int64_t cfgs = 1;
            for (int i = 0; i < tp[0]; i++) cfgs *= card[tp[2 + 2 * i]];
            const bool wide = tp[0] > 8 || cfgs > 256;
            r[0] = v;
            r[1] = tp[1];
            r[2] = (wide ? std::max(tp[0], 9) : tp[0]) | (card[v] << 8);  // a wide record reports more than 8 neighbours
            tab_max_nbr = std::max(tab_max_nbr, wide ? std::max((int)tp[0], 9) : (int)tp[0]);
            r[3] = card_off[v];
            if (wide) {  // wide variable: the record only points at its tprog entry
                r[4] = tp_off[v];
                r[5] = tp[0];
                continue;
            }
            for (int i = 0; i < 8; i++) {
                r[4 + i] = i < tp[0] ? tp[2 + 2 * i] : v;
                r[12 + i] = i < tp[0] ? tp[3 + 2 * i] : 0;
            }
        }
        // bit-sliced variant: the configuration index must be the neighbours' bits themselves (strides 1, 2, 4, 8)
        bits_ok = tab_ok;
        bits_why = tab_why;
        for (size_t j = 0; j < order.size() && bits_ok; j++) {
            const int32_t* r = trec.data() + j * 20;
            const int nn = r[2] & 0xff;
            if (nn > 4) {
                bits_ok = false;
                bits_why = "variable " + std::to_string(r[0]) + " has more than 4 free neighbours";
            }
            for (int i = 0; i < nn && bits_ok; i++)
                if (card[r[4 + i]] != 2 || r[12 + i] != (1 << i)) {
                    bits_ok = false;
                    bits_why = "variable " + std::to_string(r[0]) + " has a non-binary neighbour";
                }
        }
    }

    // Colouring of the sampled variables (fixed / collapsed neighbours never change, so they do not constrain the
    // schedule): greedy in id order, or — when that needs strictly fewer colours — greedy in smallest-last
    // (degeneracy) order: repeatedly remove a variable of least remaining degree (ties: smallest id), colour in reverse
    // removal order.  A sweep costs one barrier-to-barrier step per colour on the resident kernels, so a colour saved is
    // a step saved (Pedigree_11 + evidence: 5 -> 4 colours); the id-order colouring stays wherever it is as good by
    // sweep_cost (the torus grids keep their checkerboard).  tests/golden/make_golden.py states the same rule independently.
    int greedy_colouring(const std::vector<int32_t>& seq, std::vector<int32_t>& col) const {
        col.assign(n_vars, -1);
        int n_col = 0;
        std::vector<int> used;
        for (int32_t v : seq) {
            used.assign(n_col + 1, 0);
            for (int32_t u : nbrs[v])
                if (u != v && col[u] >= 0) used[col[u]] = 1;
            int c = 0;
            while (used[c]) c++;
            col[v] = c;
            n_col = std::max(n_col, c + 1);
        }
        return n_col;
    }
    std::vector<int32_t> smallest_last_order() const {
        std::vector<int32_t> deg(n_vars, 0);
        std::vector<uint8_t> removed(n_vars, 1);
        std::vector<std::pair<int32_t, int32_t>> heap;  // min-heap on (remaining degree, id), lazily updated
        auto cmp = [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a > b; };
        for (int v = 0; v < n_vars; v++) {
            if (!sampled(v)) continue;
            removed[v] = 0;
            for (int32_t u : nbrs[v])
                if (u != v && sampled(u)) deg[v]++;
            heap.emplace_back(deg[v], v);
        }
        std::make_heap(heap.begin(), heap.end(), cmp);
        std::vector<int32_t> seq;
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            const auto top = heap.back();
            heap.pop_back();
            const int32_t v = top.second;
            if (removed[v] || top.first != deg[v]) continue;
            removed[v] = 1;
            seq.push_back(v);
            for (int32_t u : nbrs[v])
                if (u != v && !removed[u]) {
                    deg[u]--;
                    heap.emplace_back(deg[u], u);
                    std::push_heap(heap.begin(), heap.end(), cmp);
                }
        }
        std::reverse(seq.begin(), seq.end());
        return seq;
    }
    // What a sweep costs on the resident kernels, in units of one binary update: every colour is one barrier-to-barrier step
    // that lasts as long as its slowest update, and a non-binary variable's update (full-width draws, card - 1 thresholds
    // or a log-sum-exp) takes about 2.5 binary ones.  Pedigree_11 without evidence keeps its 5 id-order colours for this
    // reason: smallest-last needs 4 but spreads the 23 ternary variables over all of them (measured 15.0 vs 18.9 us/sweep).
    double sweep_cost(const std::vector<int32_t>& col, const int n_col) const {
        std::vector<double> worst(std::max(n_col, 1), 0.0);
        bool mixed = false;  // cardinalities differ among the sampled variables
        const int32_t c0 = card[order_probe(col)];
        for (int v = 0; v < n_vars; v++)
            if (col[v] >= 0 && card[v] != c0) mixed = true;
        for (int v = 0; v < n_vars; v++)
            if (col[v] >= 0) worst[col[v]] = std::max(worst[col[v]], (mixed && card[v] > 2) ? 2.5 : 1.0);
        double total = 0.0;
        for (double w : worst) total += w;
        return total;
    }
    int order_probe(const std::vector<int32_t>& col) const {  // any sampled variable (reference cardinality for `mixed`)
        for (int v = 0; v < n_vars; v++)
            if (col[v] >= 0) return v;
        return 0;
    }
    void build_colouring() {
        std::vector<int32_t> by_id;
        for (int v = 0; v < n_vars; v++)
            if (sampled(v)) by_id.push_back(v);
        int n_col = greedy_colouring(by_id, colour);
        if (!std::getenv("GB_COLOURING_ID_ORDER")) {  // (A/B knob: keep the id-order colouring)
            std::vector<int32_t> alt;
            const int n_alt = greedy_colouring(smallest_last_order(), alt);
            if (sweep_cost(alt, n_alt) < sweep_cost(colour, n_col)) {
                colour.swap(alt);
                n_col = n_alt;
            }
        }
        colour_off.assign(n_col + 1, 0);
        for (int v = 0; v < n_vars; v++)
            if (colour[v] >= 0) colour_off[colour[v] + 1]++;
        for (int c = 0; c < n_col; c++) colour_off[c + 1] += colour_off[c];
        order.assign(colour_off[n_col], 0);
        std::vector<int32_t> fill(colour_off.begin(), colour_off.end() - 1);
        for (int v = 0; v < n_vars; v++)
            if (colour[v] >= 0) order[fill[colour[v]]++] = v;
        localise_order(kTabTile);
    }

    // Reorder the variables INSIDE each colour (any order is the same sweep: they are mutually
    // non-adjacent) so that the `tile` consecutive sweep positions one CTA of k_sweep_tab processes
    // share as many neighbour rows as possible: a tile is grown greedily from the lowest unplaced
    // id, always adding the unplaced same-colour variable that shares the most neighbours with the
    // tile so far.  On a grid this yields compact 2-D patches instead of 1-D row segments, which
    // roughly halves the distinct neighbour rows a tile reads (L1 hits instead of L2/HBM reads).
    // GB_ORDER=id keeps the plain ascending-id order.
    void localise_order(int tile) {
        const char* env = std::getenv("GB_ORDER");
        if (env && std::string(env) == "id") return;
        const int n_col = (int)colour_off.size() - 1;
        std::vector<int32_t> score(n_vars, 0), sstamp(n_vars, -1), ustamp(n_vars, -1);
        std::vector<uint8_t> placed(n_vars, 0);
        std::vector<std::pair<int32_t, int32_t>> heap;  // (score, -id): max shared neighbours, then lowest id
        int32_t blob = 0;
        for (int c = 0; c < n_col; c++) {
            const std::vector<int32_t> seg(order.begin() + colour_off[c], order.begin() + colour_off[c + 1]);
            size_t seed = 0, out = (size_t)colour_off[c];
            const size_t end = (size_t)colour_off[c + 1];
            while (out < end) {
                while (placed[seg[seed]]) seed++;
                int32_t x = seg[seed];
                blob++;
                heap.clear();
                for (;;) {
                    placed[x] = 1;
                    order[out++] = x;
                    if ((out - (size_t)colour_off[c]) % (size_t)tile == 0 || out == end) break;
                    for (int32_t u : nbrs[x]) {
                        if (u == x || ustamp[u] == blob) continue;
                        ustamp[u] = blob;
                        for (int32_t y : nbrs[u]) {
                            if (colour[y] != c || placed[y]) continue;
                            if (sstamp[y] != blob) { sstamp[y] = blob; score[y] = 0; }
                            score[y]++;
                            heap.emplace_back(score[y], -y);
                            std::push_heap(heap.begin(), heap.end());
                        }
                    }
                    x = -1;
                    while (!heap.empty()) {
                        std::pop_heap(heap.begin(), heap.end());
                        const auto top = heap.back();
                        heap.pop_back();
                        const int32_t y = -top.second;
                        if (placed[y] || score[y] != top.first) continue;
                        x = y;
                        break;
                    }
                    if (x < 0) break;  // nothing left that shares a neighbour: next seed, same tile
                }
            }
        }
    }

    // Update program of variable v (int32 words):
    //   [n_factors] then per factor: [tab_off, stride_v, n_other, (other_var, other_stride) * n_other]
    // If v occurs twice in a scope the LAST position is the one varied (gibbs-simple.go:192-197).
    void build_programs() {
        prog_off.assign(n_vars, -1);
        prog.clear();
        for (int v = 0; v < n_vars; v++) {
            if (collapsed[v] || var_funcs[v].empty()) continue;  // programs for fixed vars exist (probe only)
            prog_off[v] = (int32_t)prog.size();
            prog.push_back((int32_t)var_funcs[v].size());
            for (int32_t fi : var_funcs[v]) {
                const Factor& f = funcs[fi];
                int pos = -1;
                for (size_t i = 0; i < f.vars.size(); i++)
                    if (f.vars[i] == v) pos = (int)i;
                if (f.off + f.size > INT32_MAX) throw Err("tables exceed 2^31 entries");
                prog.push_back((int32_t)f.off);
                prog.push_back((int32_t)f.strides[pos]);
                prog.push_back((int32_t)f.vars.size() - 1);
                for (size_t i = 0; i < f.vars.size(); i++) {
                    if ((int)i == pos) continue;
                    prog.push_back(f.vars[i]);
                    prog.push_back((int32_t)f.strides[i]);
                }
            }
        }
        // Pairwise fast path (grids, ObjectDetection-style MRFs): a variable whose factors all have arity <= 2
        // gets one 16-byte record per factor {tab_off, stride of v, other variable, stride of the other} in
        // m.Funcs order, so the device walks its blanket with one vector load per factor.  A unary factor names
        // the variable itself at stride 0.
        pw_off.assign(n_vars, -1);
        pw_rec.clear();
        for (int v = 0; v < n_vars; v++) {
            if (prog_off[v] < 0) continue;
            bool ok = true;
            for (int32_t fi : var_funcs[v]) ok = ok && funcs[fi].vars.size() <= 2;
            if (!ok) continue;
            pw_off[v] = (int32_t)pw_rec.size() / 4;
            for (int32_t fi : var_funcs[v]) {
                const Factor& f = funcs[fi];
                const int pos = f.vars[0] == v ? 0 : 1;
                pw_rec.push_back((int32_t)f.off);
                pw_rec.push_back((int32_t)f.strides[pos]);
                pw_rec.push_back(f.vars.size() == 2 ? f.vars[1 - pos] : v);
                pw_rec.push_back(f.vars.size() == 2 ? (int32_t)f.strides[1 - pos] : 0);
            }
        }
        if (pw_rec.empty()) pw_rec.assign(4, 0);
    }
};

inline HostModel make_model(int32_t n_vars, const int32_t* card, const int32_t* fixed, int32_t n_funcs,
                            const int32_t* scope_off, const int32_t* scope_vars, const int64_t* tab_off,
                            const double* tables_raw) {
    if (n_vars < 1) throw Err("Invalid variable count: " + std::to_string(n_vars));
    if (n_funcs < 1) throw Err("Invalid Clique count: " + std::to_string(n_funcs));
    HostModel m;
    m.n_vars = n_vars;
    m.card.assign(card, card + n_vars);
    m.fixed.assign(n_vars, -1);
    m.collapsed.assign(n_vars, 0);
    m.coll_marg.assign(n_vars, {});
    int n_fixed = 0;
    for (int v = 0; v < n_vars; v++) {
        if (card[v] < 1) throw Err("Invalid card " + std::to_string(card[v]) + " for var " + std::to_string(v));
        if (card[v] > kMaxCard) throw Err("Cardinality above " + std::to_string(kMaxCard) + " not supported on the device");
        if (fixed) {
            if (fixed[v] != -1 && (fixed[v] < 0 || fixed[v] >= card[v]))
                throw Err("Variable has fixed val " + std::to_string(fixed[v]) + " but must be -1 or match card");
            m.fixed[v] = fixed[v];
            if (fixed[v] >= 0) n_fixed++;
        }
    }
    if (n_fixed >= n_vars) throw Err("Fixed variable count is " + std::to_string(n_fixed) + " - all vars are fixed!");
    for (int f = 0; f < n_funcs; f++) {
        if (scope_off[f + 1] - scope_off[f] < 1) throw Err("Invalid variable count (<1) for Clique " + std::to_string(f));
        std::vector<int32_t> scope(scope_vars + scope_off[f], scope_vars + scope_off[f + 1]);
        m.add_factor(scope, tables_raw + tab_off[f], tab_off[f + 1] - tab_off[f], false);
    }
    m.build_derived();
    return m;
}

// ------------------------------------------------------------------ UAI text formats
struct Tokens {
    std::vector<std::string> tok;
    size_t pos = 0;
    int lines = 0;
    // model/uai.go:20-50: drop blank lines and lines starting with 'c'; optionally skip
    // everything before the first line starting with req_prefix
    Tokens(const std::string& data, const std::string& req_prefix) {
        bool started = req_prefix.empty();
        std::istringstream in(data);
        std::string ln;
        while (std::getline(in, ln)) {
            size_t a = ln.find_first_not_of(" \t\r\n\v\f");
            if (a == std::string::npos) continue;
            size_t b = ln.find_last_not_of(" \t\r\n\v\f");
            ln = ln.substr(a, b - a + 1);
            if (ln[0] == 'c') continue;
            if (!started) {
                if (ln.compare(0, req_prefix.size(), req_prefix) != 0) continue;
                started = true;
            }
            lines++;
            std::istringstream ls(ln);
            std::string t;
            while (ls >> t) tok.push_back(t);
        }
    }
    const std::string& next(const char* what) {
        if (pos >= tok.size()) throw Err(std::string("EOF while reading ") + what);
        return tok[pos++];
    }
    int64_t next_int(const char* what) {
        const std::string& s = next(what);
        char* e = nullptr;
        long long v = std::strtoll(s.c_str(), &e, 10);
        if (e == s.c_str() || *e) throw Err(std::string("Error reading ") + what + ": '" + s + "'");
        return v;
    }
    double next_float(const char* what) {
        const std::string& s = next(what);
        char* e = nullptr;
        double v = std::strtod(s.c_str(), &e);
        if (e == s.c_str() || *e) throw Err(std::string("Error reading ") + what + ": '" + s + "'");
        return v;
    }
};

inline std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Err("Could not READ " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// model/uai.go:183-249
inline void parse_evidence(const std::string& data, const std::vector<int32_t>& card, std::vector<int32_t>& fixed) {
    Tokens t(data, "");
    if (t.lines < 1) throw Err("Invalid data buffer: there is no data");
    if (t.lines > 2) throw Err("Found " + std::to_string(t.lines) + " lines: only understand evidence files with 1 or 2 lines");
    if (t.lines == 2) {
        int64_t sc = t.next_int("UAI evid file sample count");
        if (sc == 0) return;
        if (sc > 1) throw Err("Sample count is " + std::to_string(sc) + " - only single sample evidence currently supported");
    }
    int64_t n = t.next_int("UAI evid Variable Count");
    for (int64_t i = 0; i < n; i++) {
        int64_t idx = t.next_int("evid var");
        if (idx < 0 || idx >= (int64_t)card.size()) throw Err("Read incorrect variable index " + std::to_string(idx));
        if (fixed[idx] != -1) throw Err("variable[" + std::to_string(idx) + "] had previous fixedval");
        int64_t val = t.next_int("evid var value");
        if (val < 0 || val >= card[idx]) throw Err("Read invalid value " + std::to_string(val) + " for variable[" + std::to_string(idx) + "]");
        fixed[idx] = (int32_t)val;
    }
}

// model/uai.go:53-179 (+ evidence, model/model.go:52-112)
inline HostModel load_uai(const std::string& uai_path, const char* evid_path) {
    std::string data = read_file(uai_path);
    if (data.size() < 15) throw Err("Invalid data buffer: len<15");
    Tokens t(data, "");
    if (t.lines < 1) throw Err("No lines found in file");
    if (t.tok.size() < 6) throw Err("Invalid data: only " + std::to_string(t.tok.size()) + " fields found (<6)");
    std::string type = t.next("Type");
    if (type != "BAYES" && type != "MARKOV") throw Err("Unknown model type " + type);
    int64_t nv = t.next_int("Variable count");
    if (nv < 1) throw Err("Invalid variable count: " + std::to_string(nv));
    std::vector<int32_t> card(nv), fixed(nv, -1);
    for (int64_t i = 0; i < nv; i++) card[i] = (int32_t)t.next_int("Card");
    int64_t nf = t.next_int("Clique count");
    if (nf < 1) throw Err("Invalid Clique count: " + std::to_string(nf));
    std::vector<int32_t> scope_off(1, 0), scope_vars;
    for (int64_t f = 0; f < nf; f++) {
        int64_t k = t.next_int("Clique size");
        if (k < 1) throw Err("Invalid variable count (<1) for Clique " + std::to_string(f));
        for (int64_t j = 0; j < k; j++) {
            int64_t vi = t.next_int("Clique var idx");
            if (vi < 0 || vi >= nv) throw Err("Invalid var idx " + std::to_string(vi) + " for Clique " + std::to_string(f));
            scope_vars.push_back((int32_t)vi);
        }
        scope_off.push_back((int32_t)scope_vars.size());
    }
    std::vector<int64_t> tab_off(1, 0);
    std::vector<double> tables;
    for (int64_t f = 0; f < nf; f++) {
        int64_t ts = t.next_int("table size");
        int64_t exp = 1;
        for (int32_t i = scope_off[f]; i < scope_off[f + 1]; i++) {
            if (card[scope_vars[i]] < 1) throw Err("Invalid card for var " + std::to_string(scope_vars[i]));
            exp *= card[scope_vars[i]];
            if (exp > kMaxTabSize) throw Err("Function table size exceeds " + std::to_string(kMaxTabSize));
        }
        if (ts != exp) throw Err("Read table size " + std::to_string(ts) + " != previous Clique size " + std::to_string(exp) + " on function " + std::to_string(f));
        for (int64_t i = 0; i < ts; i++) tables.push_back(t.next_float("table entry"));
        tab_off.push_back((int64_t)tables.size());
    }
    if (evid_path && *evid_path) parse_evidence(read_file(evid_path), card, fixed);
    return make_model((int32_t)nv, card.data(), fixed.data(), (int32_t)nf, scope_off.data(), scope_vars.data(),
                      tab_off.data(), tables.data());
}

// model/variable.go:106-147 NormMarginal (early return when already normalised to 1e-8)
inline void norm_marginal(std::vector<double>& m) {
    if (m.empty()) return;
    if (m.size() == 1) m[0] = 1.0;
    double sum = 0.0;
    for (double p : m) sum += p;
    const double EPS = 1e-8;
    if (std::fabs(sum - 1.0) < EPS) return;
    if (std::fabs(sum) < EPS) {
        for (auto& p : m) p = 1.0 / (double)m.size();
        return;
    }
    for (auto& p : m) p /= sum;
}

// model/uai.go:252-332
inline void load_mar(const std::string& path, std::vector<int32_t>& card, std::vector<double>& marg) {
    std::string data = read_file(path);
    if (data.size() < 11) throw Err("Invalid data buffer: len<11");
    Tokens t(data, "MAR");
    if (t.lines < 1) throw Err("No lines in file");
    if (t.tok.size() < 4) throw Err("Invalid data: only " + std::to_string(t.tok.size()) + " fields found (<4)");
    std::string ty = t.next("solution type");
    if (ty != "MAR") throw Err("Unknown solution file type " + ty);
    int64_t nv = t.next_int("MAR Solution Variable Count");
    if (nv < 1) throw Err("Invalid variable count: " + std::to_string(nv));
    card.clear();
    marg.clear();
    for (int64_t i = 0; i < nv; i++) {
        int64_t c = t.next_int("Card");
        if (c < 1) throw Err("Invalid card " + std::to_string(c) + " for var " + std::to_string(i));
        std::vector<double> m(c);
        for (int64_t k = 0; k < c; k++) {
            m[k] = t.next_float("marg prob");
            if (m[k] < 0.0 || m[k] > 1.0) throw Err("Invalid marg prob on var " + std::to_string(i));
        }
        norm_marginal(m);
        card.push_back((int32_t)c);
        marg.insert(marg.end(), m.begin(), m.end());
    }
}

// ------------------------------------------------------------------ distance measures (host copies;
// the device copies live in kernels.cuh).  model/error.go:81-249
inline void totals(const double* a, const double* b, int card, double& t1, double& t2) {
    t1 = t2 = 0.0;
    for (int c = 0; c < card; c++) { t1 += a[c]; t2 += b[c]; }
    if (t1 < 1e-12) t1 = 1e-12;
    if (t2 < 1e-12) t2 = 1e-12;
}
inline double measure_host(int which, const double* a, const double* b, int card) {
    double t1, t2;
    totals(a, b, card, t1, t2);
    double acc = 0.0;
    switch (which) {
        case 0:  // MaxAbsDiff
            for (int c = 0; c < card; c++) {
                double e = std::fabs(a[c] / t1 - b[c] / t2);
                if (c == 0 || e > acc) acc = e;
            }
            return acc;
        case 1:  // MeanAbsDiff
            if (card < 1) return 0.0;
            for (int c = 0; c < card; c++) acc += std::fabs(a[c] / t1 - b[c] / t2);
            return acc / (double)card;
        case 2:  // HellingerDiff
            for (int c = 0; c < card; c++) {
                double d = std::sqrt(a[c] / t1) - std::sqrt(b[c] / t2);
                acc += d * d;
            }
            return std::sqrt(acc) / std::sqrt(2.0);
        case 3: {  // JSDivergence (base 2)
            double k1 = 0.0, k2 = 0.0;
            for (int c = 0; c < card; c++) {
                double p1 = a[c] / t1, p2 = b[c] / t2, mid = (p1 + p2) * 0.5;
                double q = mid < 1e-12 ? 1e-12 : mid;
                double x1 = p1 < 1e-12 ? 1e-12 : p1, x2 = p2 < 1e-12 ? 1e-12 : p2;
                k1 += x1 * std::log2(x1 / q);
                k2 += x2 * std::log2(x2 / q);
            }
            return 0.5 * (k1 + k2);
        }
    }
    throw Err("unknown measure");
}

}  // namespace gb
