// Host-side mirror of the reference's `model` and `sampler` package surface for the Gibbs hot
// path, written ABOVE the C ABI (include/grample_b200.h) and using nothing else.
//
// The reference is Go; this image has no Go toolchain, so the compiled host side is C++ with
// the same names, argument meaning and error behaviour (Go's `(value, error)` becomes a C++
// exception of type grample::Error).  The Go/cgo form of the same layer is go/sampler_cuda.go
// (written against this ABI, compiled elsewhere); INTEGRATION.md shows the binding.
//
// Mapping (reference file:line -> here):
//   model.NewModelFromFile          model/model.go:52-78      -> model::NewModelFromFile
//   model.NewSolutionFromFile/.Error model/solution.go:21-65   -> model::NewSolutionFromFile, Solution::Error
//   model.HellingerDiff ... (Measure) model/error.go:81-249    -> sampler::Measure enum values
//   rand.NewGenerator               rand/rand.go:47-49        -> rnd::NewGenerator (seed of the Philox stream)
//   sampler.NewGibbsSimple          gibbs-simple.go:25-115    -> sampler::NewGibbsSimple
//   sampler.NewGibbsCollapsed / .Collapse / .BlanketSize / .FunctionCount / NeighborVarMax
//                                   gibbs-collapsed.go:23-314 -> sampler::NewGibbsCollapsed ...
//   sampler.NewChain / (*Chain).AdvanceChain / TotalSampleCount / LastSample
//                                   chain.go:13-20,151-218    -> sampler::NewChain, Chain::AdvanceChain
//   sampler.MergeChains             chain.go:96-148           -> sampler::MergeChains
//   sampler.ChainConvergence        chain.go:32-92            -> sampler::ChainConvergence
//   sampler.NewConvergenceSampler / NewIdentitySampler / Adapt  adaptive.go:16-157 -> same names
//
// One reference *Chain maps to one GROUP of `Options::replicas` device chains over the chain's
// model (all chains of a process share one gb_chains handle and one CUDA stream).
// AdvanceChain only enqueues the group's sweeps — the analogue of the goroutine the reference
// spawns — and WaitGroup::Wait() is the stream synchronisation.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/grample_b200.h"

namespace grample {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc, const char* what) {
    if (rc != 0) throw Error(std::string(what) + ": " + gb_last_error());
}

struct Options {
    int device = 0;
    int replicas = 1;              // device chains per reference Chain
    int precision = GB_HYBRID;     // the one default of every host (CLI `auto`, Go shim): float64 reference arithmetic, tabulated where possible; GB_F64 = per-update log-sum-exp everywhere
    bool history = true;           // per-chain half-window histograms (needed by ChainConvergence)
};
inline Options& options() {
    static Options o;
    return o;
}

namespace rnd {
// rand.Generator: on the device the stream is counter-based Philox keyed by (seed, chain id);
// the generator object only carries the seed and hands out global chain ids.
struct Generator {
    uint64_t seed;
    uint64_t next_chain = 0;
};
inline std::shared_ptr<Generator> NewGenerator(int64_t seed) {
    auto g = std::make_shared<Generator>();
    g->seed = (uint64_t)seed;
    return g;
}
}  // namespace rnd

namespace model {

struct Variable {  // model/variable.go:10-18
    int ID = 0;
    int Card = 0;
    int FixedVal = -1;
    std::vector<double> Marginal;
    std::map<std::string, double> State;
    bool Collapsed = false;
};

struct ModelHandle {
    gb_model* h = nullptr;
    ~ModelHandle() { if (h) gb_model_destroy(h); }
};

// model.Model: variables live on the host (they are what callers read results from); the factor
// set lives behind the C ABI.
struct Model {
    std::string Type = "MARKOV", Name;
    std::vector<Variable> Vars;
    std::shared_ptr<ModelHandle> dev;

    // model.Model.Clone (model/model.go:32-49): deep copy of the variables; the flattened factor
    // graph is immutable behind the ABI (Collapse returns a new one), so it is shared.
    std::shared_ptr<Model> Clone() const { return std::make_shared<Model>(*this); }
    int FuncCount() const {
        int32_t n = 0;
        check(gb_model_n_funcs(dev->h, &n), "gb_model_n_funcs");
        return n;
    }
    void refresh_from_device() {
        int32_t n = 0;
        check(gb_model_n_vars(dev->h, &n), "gb_model_n_vars");
        std::vector<int32_t> card(n), fixed(n), col(n);
        gb_model_cards(dev->h, card.data());
        gb_model_fixed(dev->h, fixed.data());
        gb_model_collapsed(dev->h, col.data());
        if ((int)Vars.size() != n) Vars.assign(n, Variable());
        for (int i = 0; i < n; i++) {
            Variable& v = Vars[i];
            v.ID = i;
            v.Card = card[i];
            v.FixedVal = fixed[i];
            v.Collapsed = col[i] != 0;
            if ((int)v.Marginal.size() != card[i]) v.Marginal.assign(card[i], 1.0 / card[i]);  // variable.go:45
        }
    }
};

// model.NewModelFromFile(reader, filename, useEvidence)
inline std::shared_ptr<Model> NewModelFromFile(const std::string& filename, bool useEvidence) {
    auto m = std::make_shared<Model>();
    m->dev = std::make_shared<ModelHandle>();
    const std::string ev = filename + ".evid";
    check(gb_model_load_uai(filename.c_str(), useEvidence ? ev.c_str() : nullptr, options().device, &m->dev->h),
          "Could not PARSE model");
    size_t dot = filename.find_last_of('.');
    m->Name = dot == std::string::npos ? filename : filename.substr(0, dot);
    m->refresh_from_device();
    return m;
}

struct ErrorSuite {  // model/error.go:15-25
    double MeanMeanAbsError, MaxMeanAbsError, MeanMaxAbsError, MaxMaxAbsError, MeanHellinger, MaxHellinger,
        MeanJSDiverge, MaxJSDiverge;
};
inline ErrorSuite NewErrorSuite(const std::vector<Variable>& a, const std::vector<Variable>& b) {
    if (a.size() != b.size()) throw Error("Variable count mismatch");
    std::vector<int32_t> card, f1, f2;
    std::vector<double> m1, m2;
    for (size_t i = 0; i < a.size(); i++) {
        if (a[i].Card != b[i].Card) throw Error("Variable card mismatch");
        card.push_back(a[i].Card);
        f1.push_back(a[i].FixedVal);
        f2.push_back(b[i].FixedVal);
        m1.insert(m1.end(), a[i].Marginal.begin(), a[i].Marginal.end());
        m2.insert(m2.end(), b[i].Marginal.begin(), b[i].Marginal.end());
    }
    double o[8];
    check(gb_error_suite((int32_t)a.size(), card.data(), f1.data(), m1.data(), f2.data(), m2.data(), o), "NewErrorSuite");
    return ErrorSuite{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]};
}
struct Solution {  // model/solution.go:16-65
    std::vector<Variable> Vars;
    ErrorSuite Error(const std::vector<Variable>& vars) const { return NewErrorSuite(Vars, vars); }
};
inline Solution NewSolutionFromFile(const std::string& filename) {
    int32_t n = 0, tc = 0;
    check(gb_mar_load(filename.c_str(), &n, &tc, nullptr, nullptr), "Could not READ solution");
    std::vector<int32_t> card(n);
    std::vector<double> marg(tc);
    check(gb_mar_load(filename.c_str(), nullptr, nullptr, card.data(), marg.data()), "Could not PARSE solution");
    Solution s;
    size_t o = 0;
    for (int i = 0; i < n; i++) {
        Variable v;
        v.ID = i;
        v.Card = card[i];
        v.Marginal.assign(marg.begin() + o, marg.begin() + o + card[i]);
        o += card[i];
        s.Vars.push_back(v);
    }
    return s;
}
}  // namespace model

namespace sampler {

constexpr int NeighborVarMax = GB_NEIGHBOR_VAR_MAX;  // gibbs-collapsed.go:93
using Measure = int;                                  // chain.go:24; values: gb_measure
constexpr Measure MaxAbsDiff = GB_MAX_ABS, MeanAbsDiff = GB_MEAN_ABS, HellingerDiff = GB_HELLINGER, JSDivergence = GB_JS;

// the shared device population: every Chain of the process is a group of it
struct Pool {
    gb_chains* h = nullptr;
    uint64_t seed = 0;
    std::vector<std::shared_ptr<model::ModelHandle>> keep;  // models must outlive the groups
    ~Pool() { if (h) gb_chains_destroy(h); }
    void Wait() { if (h) check(gb_chains_synchronize(h), "gb_chains_synchronize"); }
};
inline std::shared_ptr<Pool>& pool() {
    static std::shared_ptr<Pool> p;
    return p;
}
inline void ResetPool() { pool().reset(); }

// sync.WaitGroup as used by cmd/root.go:475-479
struct WaitGroup {
    std::vector<std::shared_ptr<Pool>> pending;
    void Wait() {
        for (auto& p : pending) p->Wait();
        pending.clear();
    }
};

struct FullSampler {  // sampler.go:16-18
    virtual ~FullSampler() = default;
    std::shared_ptr<rnd::Generator> gen;
    std::shared_ptr<model::Model> pgm;
    uint64_t steps = 0;  // counter of the single-step Philox stream (one per Sample / SampleVar call)

    // FullSampler.Sample: one single-variable update of the caller-held sample `s`, variable drawn uniformly among the
    // free (and, for the collapsed sampler, un-collapsed) ones; returns its index (gibbs-simple.go:148-160,
    // gibbs-collapsed.go:317-334).  One small launch per call (gb_model_sample): the API-compatibility path that the
    // reference's tests and benchmarks drive; chains advance through Chain::AdvanceChain.
    virtual int Sample(std::vector<int>& s) { return sample_one(-1, s); }

  protected:
    virtual bool exclude_collapsed() const { return false; }
    int sample_one(int varIdx, std::vector<int>& s) {
        if (s.size() != pgm->Vars.size())  // gibbs-simple.go:149-151
            throw Error("Sample size " + std::to_string(s.size()) + " != Var size " + std::to_string(pgm->Vars.size()));
        std::vector<int32_t> st(s.begin(), s.end());
        int32_t v = -1;
        const int prec = options().precision == GB_F32 ? GB_F32 : GB_F64;
        check(gb_model_sample(pgm->dev->h, prec, varIdx, exclude_collapsed() ? 1 : 0, gen ? gen->seed : 0, ++steps, st.data(), &v),
              "Could not sample from var in model");
        pgm->Vars[v].State["Selections"] += 1.0;  // gibbs-simple.go:165
        s[v] = st[v];
        return v;
    }
};

struct GibbsSimple : FullSampler {
    // (*GibbsSimple).SampleVar (gibbs-simple.go:163-271)
    int SampleVar(int varIdx, std::vector<int>& s) {
        if (varIdx < 0 || varIdx >= (int)pgm->Vars.size()) throw Error("Invalid variable index " + std::to_string(varIdx));
        return sample_one(varIdx, s);
    }
};

// sampler.NewGibbsSimple(gen, m): validation happened when the model was flattened
// (gb_model_create / gb_model_load_uai); like the reference a nil model is an error.
inline std::shared_ptr<GibbsSimple> NewGibbsSimple(std::shared_ptr<rnd::Generator> gen, std::shared_ptr<model::Model> m) {
    if (!m) throw Error("No model supplied");
    auto s = std::make_shared<GibbsSimple>();
    s->gen = gen;
    s->pgm = m;
    for (auto& v : m->Vars) v.State["Selections"] = 0.0;
    return s;
}

struct GibbsCollapsed : FullSampler {
    bool exclude_collapsed() const override { return true; }  // gibbs-collapsed.go:326
    int BlanketSize(const model::Variable& v) const {  // gibbs-collapsed.go:81-83
        int32_t n = 0;
        check(gb_model_blanket_size(pgm->dev->h, v.ID, &n), "BlanketSize");
        return n;
    }
    int FunctionCount(const model::Variable& v) const {  // gibbs-collapsed.go:86-88
        int32_t n = 0;
        check(gb_model_function_count(pgm->dev->h, v.ID, &n), "FunctionCount");
        return n;
    }
    // (*GibbsCollapsed).Collapse(varIdx): mutates the sampler's model like the reference
    // (pgm.Funcs rewritten, Collapsed flag + exact local marginal stored) and returns the variable.
    model::Variable* Collapse(int varIdx) {
        int32_t v = -1;
        double marg[GB_MAX_CARD];
        gb_model* nm = nullptr;
        const uint64_t seed = gen ? gen->seed + 0x9E3779B97F4A7C15ull * (++collapse_calls) : collapse_calls;
        check(gb_model_collapse(pgm->dev->h, varIdx, seed, &v, marg, &nm), "Collapse");
        auto nh = std::make_shared<model::ModelHandle>();
        nh->h = nm;
        pgm->dev = nh;
        model::Variable& dest = pgm->Vars[v];
        dest.Collapsed = true;
        dest.Marginal.assign(marg, marg + dest.Card);
        return &dest;
    }
    uint64_t collapse_calls = 0;
};
inline std::shared_ptr<GibbsCollapsed> NewGibbsCollapsed(std::shared_ptr<rnd::Generator> gen, std::shared_ptr<model::Model> m) {
    if (!m) throw Error("Base simple Gibbs sampler could not be created: No model supplied");
    auto s = std::make_shared<GibbsCollapsed>();
    s->gen = gen;
    s->pgm = m;
    for (auto& v : m->Vars) v.State["Selections"] = 0.0;
    return s;
}

struct Chain {  // chain.go:13-20
    std::shared_ptr<model::Model> Target;
    std::shared_ptr<FullSampler> Sampler;
    int ConvergenceWindow = 0;
    int64_t TotalSampleCount = 0;
    std::vector<int> LastSample;
    std::shared_ptr<Pool> pool_;
    int group = -1;
    int replicas = 0;

    // (*Chain).AdvanceChain(wg): enqueue one round (cw+1 recorded sweeps) for this chain's group;
    // results become visible in Target.Vars / TotalSampleCount after wg.Wait() + Refresh().
    void AdvanceChain(WaitGroup& wg) {
        check(gb_chains_group_advance(pool_->h, group, ConvergenceWindow), "AdvanceChain");
        if (std::find(wg.pending.begin(), wg.pending.end(), pool_) == wg.pending.end()) wg.pending.push_back(pool_);
    }
    // copies counts / sample count / replica 0's state back into the host-visible fields
    void Refresh() {
        pool_->Wait();
        int32_t n = 0;
        check(gb_chains_group_info(pool_->h, group, &n, &TotalSampleCount, nullptr), "gb_chains_group_info");
        int32_t tc = 0;
        check(gb_model_total_card(Target->dev->h, &tc), "gb_model_total_card");
        std::vector<uint64_t> counts(tc);
        check(gb_chains_group_counts(pool_->h, group, counts.data()), "gb_chains_group_counts");
        size_t o = 0;
        for (auto& v : Target->Vars) {
            if (!v.Collapsed)  // chain.go:231-236: collapsed variables keep their exact marginal
                for (int k = 0; k < v.Card; k++) v.Marginal[k] = (double)replicas / v.Card + (double)counts[o + k];
            o += v.Card;
        }
        std::vector<int32_t> st((size_t)n * Target->Vars.size());
        check(gb_chains_get_state(pool_->h, group, st.data()), "gb_chains_get_state");
        LastSample.assign(st.begin(), st.begin() + Target->Vars.size());
    }
};

// sampler.NewChain(mod, samp, cw, burnIn): creates the group and performs the burn-in.
// burnIn counts single-variable steps (chain.go:167-172); a sweep performs n_free of them.
inline std::shared_ptr<Chain> NewChain(std::shared_ptr<model::Model> mod, std::shared_ptr<FullSampler> samp, int cw, int64_t burnIn) {
    if (!mod || !samp) throw Error("NewChain needs a model and a sampler");
    auto ch = std::make_shared<Chain>();
    ch->Target = mod;
    ch->Sampler = samp;
    ch->ConvergenceWindow = cw;
    ch->replicas = options().replicas;
    auto& gen = *samp->gen;
    const uint64_t first = gen.next_chain;
    gen.next_chain += (uint64_t)((ch->replicas + 7) / 8 * 8);
    auto& p = pool();
    if (!p) {
        p = std::make_shared<Pool>();
        p->seed = gen.seed;
        gb_model* ms[1] = {mod->dev->h};
        int32_t n[1] = {ch->replicas};
        check(gb_chains_create(1, ms, n, gen.seed, first, options().precision, options().history ? GB_CHAINS_HISTORY : 0u,
                               options().device, &p->h), "NewChain");
    } else {
        check(gb_chains_add_group(p->h, mod->dev->h, ch->replicas, first), "NewChain");
    }
    p->keep.push_back(mod->dev);
    int32_t ng = 0;
    gb_chains_n_groups(p->h, &ng);
    ch->pool_ = p;
    ch->group = ng - 1;
    int32_t n_order = 0;
    check(gb_model_schedule(mod->dev->h, &n_order, nullptr, nullptr, nullptr), "gb_model_schedule");
    const int64_t sweeps = burnIn <= 0 ? 0 : (burnIn + n_order - 1) / n_order;
    check(gb_chains_group_sweep(p->h, ch->group, sweeps, 0), "Failure during chain burn in");
    ch->LastSample.assign(mod->Vars.size(), 0);
    return ch;
}

// sampler.MergeChains(chains) (chain.go:96-148): 0 chains is an error; 1 chain returns its own
// variables; otherwise collapsed-in-any-chain wins (first chain in list order), else sums.
inline std::vector<model::Variable> MergeChains(const std::vector<std::shared_ptr<Chain>>& chains) {
    if (chains.empty()) throw Error("Can not merge 0 chains");
    for (auto& c : chains) c->Refresh();
    if (chains.size() == 1) return chains[0]->Target->Vars;
    const size_t n = chains[0]->Target->Vars.size();
    std::vector<model::Variable> vars(n);
    std::vector<bool> col(n, false);
    for (size_t i = 0; i < n; i++) {
        const model::Variable* found = nullptr;
        for (auto& c : chains)
            if (c->Target->Vars[i].Collapsed) { found = &c->Target->Vars[i]; break; }
        col[i] = found != nullptr;
        vars[i] = found ? *found : chains[0]->Target->Vars[i];
    }
    for (size_t c = 1; c < chains.size(); c++) {
        if (chains[c]->Target->Vars.size() != n) throw Error("Cannot merge chain with different variable count");
        for (size_t i = 0; i < n; i++) {
            if (col[i]) continue;
            const auto& src = chains[c]->Target->Vars[i].Marginal;
            for (size_t k = 0; k < src.size(); k++) vars[i].Marginal[k] += src[k];
        }
    }
    return vars;
}

// sampler.ChainConvergence(chains, distFunc, mergedVars) — K4 on the device.  The chain list
// must be every chain of the pool (that is how cmd/root.go and adaptive.go call it).
inline std::vector<double> ChainConvergence(const std::vector<std::shared_ptr<Chain>>& chains, Measure distFunc,
                                            std::vector<model::Variable> mergedVars) {
    if (chains.size() < 2 && (chains.empty() || chains[0]->replicas < 2))
        throw Error("Convergence requires at least 2 chains");
    auto& p = chains[0]->pool_;
    int32_t ng = 0;
    gb_chains_n_groups(p->h, &ng);
    if ((size_t)ng != chains.size()) throw Error("ChainConvergence expects every chain of the pool");
    if (mergedVars.empty()) mergedVars = MergeChains(chains);
    std::vector<double> merged;
    for (auto& v : mergedVars) merged.insert(merged.end(), v.Marginal.begin(), v.Marginal.end());
    std::vector<double> out(mergedVars.size());
    check(gb_chains_convergence(p->h, distFunc, merged.data(), out.data()), "ChainConvergence");
    return out;
}

struct AdaptiveSampler {  // sampler.go:23-25
    virtual ~AdaptiveSampler() = default;
    virtual std::vector<std::shared_ptr<Chain>> Adapt(std::vector<std::shared_ptr<Chain>> chains, int newChainCount) = 0;
};
struct IdentitySampler : AdaptiveSampler {  // adaptive.go:13-24
    std::vector<std::shared_ptr<Chain>> Adapt(std::vector<std::shared_ptr<Chain>> chains, int) override { return chains; }
};
inline std::shared_ptr<IdentitySampler> NewIdentitySampler() { return std::make_shared<IdentitySampler>(); }

struct ConvergenceSampler : AdaptiveSampler {  // adaptive.go:28-157
    std::shared_ptr<model::Model> BaseModel;
    Measure DistFunc = HellingerDiff;
    std::shared_ptr<rnd::Generator> Gen;
    int MaxChains = 128;

    std::vector<std::shared_ptr<Chain>> Adapt(std::vector<std::shared_ptr<Chain>> chains, int newChainCount) override {
        if (chains.size() < 2) throw Error("At least 2 chains required for adaptation");
        if ((int)chains.size() >= MaxChains) return chains;
        auto& p = chains[0]->pool_;
        const int replicas = chains.back()->replicas, cw = chains.back()->ConvergenceWindow;
        std::vector<int32_t> chosen(std::max(newChainCount, 1));
        int32_t n_chosen = 0;
        const uint64_t first = Gen->next_chain;
        check(gb_chains_adapt(p->h, BaseModel->dev->h, newChainCount, replicas, DistFunc, cw, MaxChains, first,
                              chosen.data(), &n_chosen), "Adapt");
        Gen->next_chain += (uint64_t)n_chosen * (uint64_t)((replicas + 7) / 8 * 8);
        int32_t ng = 0;
        gb_chains_n_groups(p->h, &ng);
        for (int i = 0; i < n_chosen; i++) {  // wrap the groups the device created as Chain objects
            auto ch = std::make_shared<Chain>();
            ch->Target = BaseModel->Clone();
            gb_model* gm = nullptr;
            ch->group = ng - n_chosen + i;
            check(gb_chains_group_info(p->h, ch->group, nullptr, nullptr, &gm), "gb_chains_group_info");
            // borrowed handle: the model of an adapted group is owned by the pool
            struct Borrow { static void none(model::ModelHandle* m) { m->h = nullptr; delete m; } };
            ch->Target->dev = std::shared_ptr<model::ModelHandle>(new model::ModelHandle(), Borrow::none);
            ch->Target->dev->h = gm;
            ch->Target->refresh_from_device();
            auto s = std::make_shared<GibbsCollapsed>();
            s->gen = Gen;
            s->pgm = ch->Target;
            ch->Sampler = s;
            ch->ConvergenceWindow = cw;
            ch->replicas = replicas;
            ch->pool_ = p;
            // the collapsed variable's exact local marginal comes from the merged view
            std::vector<double> merged;
            int32_t tc = 0;
            gb_model_total_card(gm, &tc);
            merged.resize(tc);
            std::vector<int32_t> col(ch->Target->Vars.size());
            check(gb_chains_merged_marginals(p->h, merged.data(), col.data()), "MergeChains");
            size_t o = 0;
            for (auto& v : ch->Target->Vars) {
                if (v.Collapsed) v.Marginal.assign(merged.begin() + o, merged.begin() + o + v.Card);
                o += v.Card;
            }
            ch->LastSample.assign(ch->Target->Vars.size(), 0);
            chains.push_back(ch);
        }
        return chains;
    }
};
inline std::shared_ptr<ConvergenceSampler> NewConvergenceSampler(std::shared_ptr<rnd::Generator> gen, std::shared_ptr<model::Model> m, Measure d = -1) {
    if (!m) throw Error("A full model is required for Adaptation");
    auto s = std::make_shared<ConvergenceSampler>();
    s->BaseModel = m;
    s->DistFunc = d < 0 ? HellingerDiff : d;  // adaptive.go:41-43
    s->Gen = gen;
    return s;
}

}  // namespace sampler
}  // namespace grample
