// ORACLE — TEST INFRASTRUCTURE ONLY (see model.hpp header).
//
// CPU float64 restatement of the reference's hot path: Go packages `sampler` and
// `buffer`.  Citations are relative to /root/reference.
//
// Pinned by the reference's own tests (transcribed in tests/test_oracle_*.py):
//   gibbs-collapsed_test.go:14-48    collapse marginal 0.5/0.5 on deterministic.uai
//   gibbs-collapsed_test.go:51-111   collapse bookkeeping + mandatory third-collapse error
//   gibbs-simple_test.go:13-38       one.uai: Sample returns idx 0, both values drawn
//   chain_test.go:11-80              MergeChains known answers
//   sampler_test.go:69-220           UniSample / VarSample / WeightedSample conventions
//   circular_test.go:9-60            CircularInt halves
//   end-to-end: res/*.uai.MAR ground-truth marginals (statistical)
// PARITY UNPINNED by any reference test (the reference has none): the numerics of
// SampleVar (min-shift + sequential 1e-6 floor, gibbs-simple.go:227-258), the
// AdvanceChain thresholds (chain.go:180-218), ChainConvergence/ChainDist
// (chain.go:32-92, 253-290) and all of adaptive.go.  For those the only pin is that
// this restatement follows the cited lines (plus exact-inference checks on tiny models).
#pragma once
#include <algorithm>
#include <set>

#include "model.hpp"
#include "rng.hpp"

namespace oracle {

// ---------------------------------------------------------------- buffer.CircularInt
// buffer/circular.go:6-102
struct CircularInt {
    std::vector<int> buf;
    int pos = 0, buf_size = 0, count = 0;
    int64_t total_seen = 0;
    CircularInt() = default;
    explicit CircularInt(int total_size) {
        int half = total_size / 2;
        buf_size = half + half;
        buf.assign(buf_size, 0);
    }
    void add(int v) {
        total_seen++;
        buf[pos] = v;  // NOTE: like the reference this faults for buf_size == 0
        pos = (pos + 1) % buf_size;
        count++;
        if (count > buf_size) count = buf_size;
    }
    bool halves_valid() const { return count >= buf_size; }
    // oldest half, in insertion order (circular.go:53-64)
    std::vector<int> first_half() const {
        std::vector<int> out;
        int cur = pos;
        for (int r = buf_size / 2; r > 0; r--) { out.push_back(buf[cur]); cur = (cur + 1) % buf_size; }
        return out;
    }
    // newest half (circular.go:68-81)
    std::vector<int> second_half() const {
        std::vector<int> out;
        int half = buf_size / 2;
        int cur = (pos + half) % buf_size;
        for (int r = half; r > 0; r--) { out.push_back(buf[cur]); cur = (cur + 1) % buf_size; }
        return out;
    }
};

// ---------------------------------------------------------------- UniformSampler
// sampler/sampler.go:41-174
constexpr int kMaxCard = 1 << 30;  // sampler.go:69

struct UniformSampler {
    Generator* gen = nullptr;
    UniformSampler() = default;
    UniformSampler(Generator* g, int max_vars) : gen(g) {
        if (max_vars < 1) throw Error("Invalid max var count");
    }
    int uni_sample(int64_t card) const {  // sampler.go:72-86
        if (card < 1) throw Error("Can not sample if Cardinality < 1");
        if (card > kMaxCard) throw Error("Cardinality above 1<<30 not supported");
        if (card == 1) return 0;
        return (int)gen->int31n((int32_t)card);
    }
    int weighted_sample(int64_t card, const double* w, size_t len) const {  // sampler.go:90-130
        if (card < 1) throw Error("Can not sample if Cardinality < 1");
        if (card > kMaxCard) throw Error("Cardinality above 1<<30 not supported");
        if ((int64_t)len != card) throw Error("Weight array size must match cardinality");
        if (card == 1) return 0;
        double tot = 0.0;
        for (size_t i = 0; i < len; i++) {
            if (w[i] <= 0.0) throw Error("Weights must be > 0.0");
            tot += w[i];
        }
        double r = gen->float64() * tot;
        for (size_t i = 0; i < len; i++) {
            if (r <= w[i]) return (int)i;  // NOTE: <= (sampler.go:118)
            r -= w[i];
        }
        throw Error("Failed to sample");
    }
    int var_sample(const std::vector<Variable>& vs, bool exclude_collapsed) const {  // sampler.go:135-174
        if (vs.empty()) throw Error("Can not sample from an empty variable list");
        std::vector<int> idx;
        idx.reserve(vs.size());
        for (size_t i = 0; i < vs.size(); i++) {
            if (exclude_collapsed && vs[i].collapsed) continue;
            if (vs[i].fixed_val >= 0) continue;
            idx.push_back((int)i);
        }
        if (idx.empty()) throw Error("No Variables to select");
        if (idx.size() == 1) return idx[0];
        return idx[uni_sample((int64_t)idx.size())];
    }
};

// sampler/sampler.go:16-18
struct FullSampler {
    virtual ~FullSampler() = default;
    virtual int sample(std::vector<int>& s) = 0;
};

// ---------------------------------------------------------------- GibbsSimple
// sampler/gibbs-simple.go:13-271
struct GibbsSimple : FullSampler {
    Generator* gen;
    Model* pgm;
    UniformSampler uni;
    std::vector<std::vector<int>> var_funcs;  // per var: indices into pgm->funcs, in m.Funcs order
    std::vector<int> last;
    // "lean" bookkeeping: same draws, O(1) pick, no O(n) copy.
    bool lean = false;
    std::vector<int> eligible_all, eligible_uncollapsed;

    GibbsSimple(Generator* g, Model* m) : gen(g), pgm(m) {  // gibbs-simple.go:25-115
        if (!m) throw Error("No model supplied");
        uni = UniformSampler(g, (int)m->vars.size());
        var_funcs.assign(m->vars.size(), {});
        last.assign(m->vars.size(), 0);
        for (size_t fi = 0; fi < m->funcs.size(); fi++) {
            m->funcs[fi].use_log_space();  // errors if already log (double-call)
            for (int v : m->funcs[fi].vars) var_funcs[v].push_back((int)fi);
        }
        for (size_t i = 0; i < m->vars.size(); i++) {
            Variable& v = m->vars[i];
            v.state["Selections"] = 0.0;
            if ((int)i != v.id) throw Error("Invalid ID for var");
            if (var_funcs[i].empty()) throw Error("There are no functions for var");
            if (v.fixed_val >= 0) last[i] = v.fixed_val;
            else last[i] = uni.uni_sample(v.card);
        }
        rebuild_eligible();
    }

    void rebuild_eligible() {
        eligible_all.clear();
        eligible_uncollapsed.clear();
        for (size_t i = 0; i < pgm->vars.size(); i++) {
            if (pgm->vars[i].fixed_val >= 0) continue;
            eligible_all.push_back((int)i);
            if (!pgm->vars[i].collapsed) eligible_uncollapsed.push_back((int)i);
        }
    }

    void functions_changed() {  // gibbs-simple.go:119-145
        var_funcs.assign(pgm->vars.size(), {});
        for (size_t fi = 0; fi < pgm->funcs.size(); fi++) {
            if (!pgm->funcs[fi].is_log) throw Error("Function is not in log space on FunctionsChanged");
            for (int v : pgm->funcs[fi].vars) var_funcs[v].push_back((int)fi);
        }
        for (size_t i = 0; i < pgm->vars.size(); i++) {
            Variable& v = pgm->vars[i];
            if (v.fixed_val >= 0) last[i] = v.fixed_val;
            else last[i] = uni.weighted_sample(v.card, v.marginal.data(), v.marginal.size());
        }
        rebuild_eligible();
    }

    int pick_var(bool exclude_collapsed) {
        if (!lean) return uni.var_sample(pgm->vars, exclude_collapsed);
        const std::vector<int>& e = exclude_collapsed ? eligible_uncollapsed : eligible_all;
        if (e.empty()) throw Error("No Variables to select");
        if (e.size() == 1) return e[0];
        return e[uni.uni_sample((int64_t)e.size())];
    }

    int sample(std::vector<int>& s) override {  // gibbs-simple.go:148-160
        if (s.size() != pgm->vars.size()) throw Error("Sample size != Var size");
        int var_idx = pick_var(false);
        return sample_var(var_idx, s);
    }

    // gibbs-simple.go:171-258: the floored, un-normalised weights e[k] for `state`.
    // The conditional actually sampled is e[k] / sum(e) (WeightedSample re-sums).
    void conditional(int var_idx, const int* state, std::vector<double>& w) const {
        const Variable& sv = pgm->vars[var_idx];
        if (sv.fixed_val >= 0) throw Error("Selected sample variable which has FixedVal");
        w.assign(sv.card, 0.0);
        int call_vals[64];
        for (int fi : var_funcs[var_idx]) {
            const Function& f = pgm->funcs[fi];
            if (f.vars.size() > 64) throw Error("scope too large for oracle scratch");
            int call_idx = -1;
            for (size_t i = 0; i < f.vars.size(); i++) {
                call_vals[i] = state[f.vars[i]];
                if (f.vars[i] == sv.id) call_idx = (int)i;
            }
            if (call_idx < 0) throw Error("Var not in function var list?!");
            for (int k = 0; k < sv.card; k++) {
                call_vals[call_idx] = k;
                w[k] += f.eval(call_vals, f.vars.size());
            }
        }
        double mn = w[0];
        for (int k = 1; k < sv.card; k++)
            if (w[k] < mn) mn = w[k];
        if (mn < -8.0)
            for (int k = 0; k < sv.card; k++) w[k] = w[k] - (mn - 1.5);
        double tot = 0.0;
        for (int k = 0; k < sv.card; k++) {
            double v = std::exp(w[k]);
            tot += v;
            w[k] = v;
        }
        for (int k = 0; k < sv.card; k++) {
            if (w[k] / tot < 1e-6) {
                double delta = tot * 1e-6;
                if (delta <= 1e-12) throw Error("Logic error: delta <= 1e-12");
                tot += delta;  // sequential: later k see the adjusted total
                w[k] += delta;
            }
        }
    }

    int sample_var(int var_idx, std::vector<int>& s) {  // gibbs-simple.go:163-271
        Variable& sv = pgm->vars[var_idx];
        if (!lean) sv.state["Selections"] += 1.0;
        std::vector<double>& w = scratch_w;
        conditional(var_idx, last.data(), w);
        int next_val;
        try {
            next_val = uni.weighted_sample((int64_t)w.size(), w.data(), w.size());
        } catch (const Error&) {
            return -1;  // gibbs-simple.go:263-265 swallows the error
        }
        last[var_idx] = next_val;
        if (!lean) std::copy(last.begin(), last.end(), s.begin());  // O(n) copy, line 268
        else s[var_idx] = next_val;
        return var_idx;
    }
    std::vector<double> scratch_w;
};

// ---------------------------------------------------------------- GibbsCollapsed
// sampler/gibbs-collapsed.go:17-334
constexpr int kNeighborVarMax = 12;  // gibbs-collapsed.go:93

struct GibbsCollapsed : FullSampler {
    std::unique_ptr<GibbsSimple> base;
    std::vector<std::set<int>> var_neighbors;

    GibbsCollapsed(Generator* g, Model* m) {  // gibbs-collapsed.go:23-40
        base.reset(new GibbsSimple(g, m));
        functions_changed();
    }

    void functions_changed() {  // gibbs-collapsed.go:44-78
        Model* pgm = base->pgm;
        std::vector<std::set<int>> nb(pgm->vars.size());
        for (size_t i = 0; i < pgm->vars.size(); i++)
            if ((int)i != pgm->vars[i].id) throw Error("Invalid variable setup");
        for (size_t i = 0; i < base->var_funcs.size(); i++)
            for (int fi : base->var_funcs[i])
                for (int v : pgm->funcs[fi].vars) nb[i].insert(v);
        for (size_t i = 0; i < pgm->vars.size(); i++)
            if (pgm->vars[i].collapsed && !nb[i].empty()) throw Error("Var is collapsed but has a blanket");
        var_neighbors.swap(nb);
    }

    int blanket_size(int var_id) const { return (int)var_neighbors[var_id].size(); }      // :81-83
    int function_count(int var_id) const { return (int)base->var_funcs[var_id].size(); }  // :86-88

    // gibbs-collapsed.go:98-314.  Returns the index of the collapsed variable.
    // Blanket order: the reference iterates a Go map (random order, line 154); ascending
    // variable id is used here — the table LAYOUT of the new factor differs run to run in
    // the reference, its semantics do not.
    int collapse(int var_idx) {
        Model* pgm = base->pgm;
        if (var_idx < 0) {
            for (size_t i = 0; i < pgm->vars.size(); i++) {
                var_idx = base->pick_var(true);
                int n = (int)var_neighbors[var_idx].size();
                if (n <= kNeighborVarMax) break;
                var_idx = -1;
            }
        }
        if (var_idx < 0) throw Error("Failed to randomly select a variable to collapse");
        if (var_idx >= (int)pgm->vars.size()) throw Error("Invalid variable index");

        Variable coll = pgm->vars[var_idx];  // clone
        if (coll.fixed_val >= 0) throw Error("Can not collapse Fixed Val variable");
        if (coll.collapsed) throw Error("Already collapsed variable");
        for (int i = 0; i < coll.card; i++) coll.marginal[i] = 1e-12;

        std::vector<const Variable*> blanket;
        std::map<int, int> xref;
        int coll_idx = -1;
        std::vector<int> new_ids, new_cards;
        for (int vi : var_neighbors[var_idx]) {
            const Variable* v = &pgm->vars[vi];
            blanket.push_back(v);
            xref[v->id] = (int)blanket.size() - 1;
            if (coll.id == v->id) coll_idx = (int)blanket.size() - 1;
            else { new_ids.push_back(v->id); new_cards.push_back(v->card); }
        }
        if (coll_idx < 0) throw Error("Collapsing variable not in its own blanket");
        if (new_ids.size() != blanket.size() - 1) throw Error("New function size mismatch");
        if (new_ids.empty()) throw Error("New function would have 0 variables");

        const std::vector<int>& funcs = base->var_funcs[var_idx];
        std::set<std::string> del_names;
        for (int fi : funcs) {
            del_names.insert(pgm->funcs[fi].name);
            if (!pgm->funcs[fi].is_log) throw Error("Function is not set up for Log Space");
        }

        Function post = new_function((int)pgm->funcs.size(), new_ids, new_cards);  // cap 1<<23
        post.name = "COLLAPSE-" + coll.name;

        std::vector<int> call_vals(64), var_state(blanket.size());
        VariableIter it(blanket, true);
        for (;;) {
            it.val(var_state);
            int marginal_val = var_state[coll_idx];
            double fr = 0.0;
            for (int fi : funcs) {
                const Function& f = pgm->funcs[fi];
                for (size_t i = 0; i < f.vars.size(); i++) call_vals[i] = var_state[xref[f.vars[i]]];
                fr += f.eval(call_vals.data(), f.vars.size());
            }
            fr = std::exp(fr);
            coll.marginal[marginal_val] += fr;
            for (size_t i = 0; i < new_ids.size(); i++) call_vals[i] = var_state[xref[new_ids[i]]];
            post.add_value(call_vals.data(), new_ids.size(), fr);
            if (!it.next()) break;
        }
        coll.norm_marginal();
        post.use_log_space();

        pgm->funcs.push_back(post);
        std::vector<Function> kept;
        for (auto& f : pgm->funcs)
            if (!del_names.count(f.name)) kept.push_back(std::move(f));
        if (kept.empty()) throw Error("No functions left after collapse!");
        pgm->funcs.swap(kept);

        base->functions_changed();  // re-draws the whole state from current marginals
        functions_changed();
        pgm->check();

        Variable& dest = pgm->vars[var_idx];
        dest.collapsed = true;
        dest.marginal = coll.marginal;
        base->rebuild_eligible();
        return var_idx;
    }

    int sample(std::vector<int>& s) override {  // gibbs-collapsed.go:317-334
        if (s.size() != base->pgm->vars.size()) throw Error("Samples size is wrong");
        int var_idx = base->pick_var(true);
        return base->sample_var(var_idx, s);
    }
};

// ---------------------------------------------------------------- Chain
// sampler/chain.go:13-20, 151-246
struct Chain {
    Model* target;
    FullSampler* sampler;
    int cw;
    std::vector<CircularInt> history;
    int64_t total_sample_count = 0;
    std::vector<int> last_sample;

    Chain(Model* mod, FullSampler* samp, int cw_, int64_t burn_in) : target(mod), sampler(samp), cw(cw_) {
        history.assign(mod->vars.size(), CircularInt(cw_));
        last_sample.assign(mod->vars.size(), 0);
        for (int64_t i = 0; i < burn_in; i++) one_sample(false);
    }

    void one_sample(bool update_vars) {  // chain.go:221-246
        int var_idx = sampler->sample(last_sample);
        if (var_idx < 0 || target->vars[var_idx].fixed_val >= 0) throw Error("Invalid sample");
        if (update_vars) {
            int value = last_sample[var_idx];
            Variable& v = target->vars[var_idx];
            if (!v.collapsed) v.marginal[value] += 1.0;
            history[var_idx].add(value);
            total_sample_count++;
        }
    }

    // chain.go:180-218.  The reference runs this body in a goroutine joined by the caller's
    // WaitGroup; here it is synchronous (callers that want the concurrency use std::thread).
    void advance() {
        std::vector<int64_t> thresh(history.size());
        for (size_t i = 0; i < history.size(); i++) thresh[i] = history[i].total_seen + (int64_t)cw + 1;
        auto keep_running = [&]() {
            for (size_t i = 0; i < history.size(); i++) {
                const Variable& v = target->vars[i];
                if (!v.collapsed && v.fixed_val < 0 && history[i].total_seen < thresh[i]) return true;
            }
            return false;
        };
        size_t batch = target->vars.size() * 2;
        while (keep_running())
            for (size_t i = 0; i < batch; i++) one_sample(true);
    }

    // chain.go:253-290
    void chain_dist(int m, int var_idx, const Variable& merged, double& within, double& between) const {
        const CircularInt& h = history[var_idx];
        if (h.total_seen < (int64_t)cw) throw Error("Total seen < Convergence Window");
        const Variable& src = target->vars[var_idx];
        if (src.card != merged.card) throw Error("Variable mismatch");
        Variable v1 = src, v2 = src;
        for (int i = 0; i < src.card; i++) { v1.marginal[i] = 1e-8; v2.marginal[i] = 1e-8; }
        for (int val : h.first_half()) v1.marginal[val] += 1.0;
        for (int val : h.second_half()) v2.marginal[val] += 1.0;
        within = measure(m, v1, v2);
        for (int i = 0; i < src.card; i++) v1.marginal[i] += v2.marginal[i];
        between = measure(m, merged, v1);
    }
};

// chain.go:96-148
inline std::vector<Variable> merge_chains(const std::vector<Chain*>& chains) {
    if (chains.empty()) throw Error("Can not merge 0 chains");
    if (chains.size() == 1) return chains[0]->target->vars;
    size_t n = chains[0]->target->vars.size();
    std::vector<Variable> vars(n);
    std::vector<bool> is_coll(n, false);
    for (size_t vi = 0; vi < n; vi++) {
        const Variable* found = nullptr;
        for (auto* ch : chains)
            if (ch->target->vars[vi].collapsed) { found = &ch->target->vars[vi]; break; }
        if (found) { is_coll[vi] = true; vars[vi] = *found; }
        else vars[vi] = chains[0]->target->vars[vi];
    }
    for (size_t c = 1; c < chains.size(); c++) {
        if (chains[c]->target->vars.size() != n) throw Error("Cannot merge chains of different var counts");
        for (size_t vi = 0; vi < n; vi++) {
            if (is_coll[vi]) continue;
            const Variable& src = chains[c]->target->vars[vi];
            for (size_t k = 0; k < src.marginal.size(); k++) vars[vi].marginal[k] += src.marginal[k];
        }
    }
    return vars;
}

// chain.go:32-92
inline std::vector<double> chain_convergence(const std::vector<Chain*>& chains, int m, std::vector<Variable> merged) {
    if (chains.size() < 2) throw Error("Convergence requires at least 2 chains");
    if (merged.empty()) merged = merge_chains(chains);
    std::vector<double> vals(merged.size());
    double sample_count = (double)chains[0]->cw;
    double chain_count = (double)chains.size();
    double b_norm = sample_count / (chain_count - 1);
    double w_factor = (sample_count - 1) / sample_count;
    double b_factor = (chain_count + 1) / (chain_count * sample_count);
    for (size_t i = 0; i < merged.size(); i++) {
        const Variable& v = merged[i];
        if (v.collapsed || v.fixed_val >= 0) { vals[i] = 1.0; continue; }
        double W = 1e-8, B = 1e-8;
        for (auto* ch : chains) {
            double w1, b1;
            ch->chain_dist(m, (int)i, v, w1, b1);
            W += w1;
            B += b1;
        }
        W /= chain_count;
        B *= b_norm;
        double vhat = (w_factor * W) + (b_factor * B);
        vals[i] = std::sqrt((4.0 * vhat) / (2.0 * W));
    }
    return vals;
}

// ---------------------------------------------------------------- adaptive
// sampler/adaptive.go:28-157.  Owns the models/samplers/chains it creates (Go's GC does
// that in the reference).
struct OwnedChain {
    std::unique_ptr<Model> model;
    std::unique_ptr<FullSampler> sampler;
    std::unique_ptr<Chain> chain;
};

struct ConvergenceSampler {
    Model base_model;
    int dist = kHellinger;
    Generator* gen;
    int max_chains = 128;
    std::vector<std::unique_ptr<OwnedChain>> owned;
    std::vector<int> last_targets;  // for tests: variables collapsed by the last Adapt

    ConvergenceSampler(Generator* g, const Model& m, int measure_or_neg) : base_model(m), gen(g) {
        if (measure_or_neg >= 0) dist = measure_or_neg;
    }

    std::vector<Chain*> adapt(std::vector<Chain*> chains, int new_chain_count) {
        last_targets.clear();
        if (chains.size() < 2) throw Error("At least 2 chains required for adaptation");
        if ((int)chains.size() >= max_chains) return chains;

        auto oc = std::make_unique<OwnedChain>();
        oc->model.reset(new Model(base_model.clone()));
        auto* samp = new GibbsCollapsed(gen, oc->model.get());
        oc->sampler.reset(samp);

        std::vector<Variable> merged = merge_chains(chains);
        std::vector<const Variable*> vars;
        for (auto& v : merged) {
            int sz = samp->blanket_size(v.id);
            if (v.fixed_val < 0 && !v.collapsed && sz > 1 && sz <= kNeighborVarMax) vars.push_back(&v);
        }
        if (vars.empty()) return chains;

        std::vector<int> targets;
        if ((int)vars.size() <= new_chain_count) {
            for (auto* v : vars) targets.push_back(v->id);
        } else {
            std::vector<double> conv = chain_convergence(chains, dist, merged);
            // adaptive.go:111-119: sort DESCENDING then take from the END = lowest scores.
            // sort.Slice is unstable in Go; ties are broken by variable id here.
            std::stable_sort(vars.begin(), vars.end(),
                             [&](const Variable* a, const Variable* b) { return conv[a->id] > conv[b->id]; });
            int pos = (int)vars.size() - 1;
            for (int cc = 0; cc < new_chain_count; cc++) targets.push_back(vars[pos--]->id);
        }
        if (targets.empty()) return chains;

        Chain* last_chain = chains.back();
        for (int var_idx : targets) {
            if (!oc) {
                oc = std::make_unique<OwnedChain>();
                oc->model.reset(new Model(base_model.clone()));
                samp = new GibbsCollapsed(gen, oc->model.get());
                oc->sampler.reset(samp);
            }
            samp->collapse(var_idx);
            oc->chain.reset(new Chain(oc->model.get(), samp, last_chain->cw, 2));
            chains.push_back(oc->chain.get());
            owned.push_back(std::move(oc));
            last_targets.push_back(var_idx);
        }
        return chains;
    }
};

}  // namespace oracle
