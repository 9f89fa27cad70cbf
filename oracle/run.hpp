// ORACLE — TEST INFRASTRUCTURE ONLY (see model.hpp header).
//
// Restatement of the caller's main loop, cmd/root.go:344-363 (defaults), 381-447 (chain
// construction per sampler kind), 475-561 (advance all chains -> barrier -> count ->
// adapt) and 565-571 (final merge + normalise).  Wall-clock driven behaviour (5 s status,
// maxSecs, adaptation stopping at maxSecs/2) is replaced by an explicit round budget so
// runs are reproducible; everything else follows the cited lines.
//
// Concurrency: the reference runs one goroutine per chain per round (chain.go:197-215)
// joined by a WaitGroup (root.go:476-479), all drawing from ONE generator goroutine
// through a channel (rand/rand.go:33-37, root.go:373).  Here: one std::thread per chain per
// round; `faithful` shares one mutex-guarded MT19937-64 (stand-in for the channel),
// `lean` gives every chain its own generator (seed + chain index) and uses the O(1)
// bookkeeping described in SURVEY §8d.
#pragma once
#include <chrono>
#include <mutex>
#include <thread>

#include "sampler.hpp"

namespace oracle {

struct LockedGenerator : Generator {
    std::mutex mu;
    explicit LockedGenerator(int64_t seed) : Generator(seed) {}
    int64_t int63() override {
        std::lock_guard<std::mutex> g(mu);
        return Generator::int63();
    }
};

enum SamplerKind { kSimple = 0, kCollapsed = 1, kAdaptive = 2 };

struct RunParams {
    int kind = kSimple;
    int n_chains = 4;       // root.go:358-363 (>= 2)
    int64_t burn_in = -1;   // <0 -> 2000*n (root.go:348-350)
    int64_t cw = 0;         // <=0 -> burn_in (root.go:351-353)
    int64_t max_iters = -1; // <0 -> 20000*n (root.go:354-356)
    int64_t seed = 1;
    bool lean = false;
    int n_threads = 0;      // 0 -> one thread per chain (reference behaviour)
    int chain_adds = 1;     // root.go flag -a
    int adapt_rounds = 1 << 30;  // stand-in for "stop adapting at maxSecs/2"
    int max_rounds = 1 << 30;    // stand-in for maxSecs
};

struct RunResult {
    std::vector<Variable> merged;  // un-normalised, as MergeChains returns
    int64_t samples = 0;
    int rounds = 0;
    int n_chains_final = 0;
    double burnin_seconds = 0, advance_seconds = 0;
    std::vector<int64_t> curve_samples;   // after every round
    std::vector<double> curve_mean_hel, curve_max_hel, curve_mean_abs;
};

inline RunResult run_marginals(const Model& mod, const Solution* sol, RunParams p) {
    using clk = std::chrono::steady_clock;
    const int64_t n = (int64_t)mod.vars.size();
    if (p.burn_in < 0) p.burn_in = 2000 * n;
    if (p.cw <= 0) p.cw = p.burn_in;
    if (p.max_iters < 0) p.max_iters = 20000 * n;
    if (p.n_chains < 2) p.n_chains = 2;
    if (p.kind != kAdaptive && p.chain_adds != 1) throw Error("Sampler is not adaptive: ChainAdds makes no sense");

    std::unique_ptr<LockedGenerator> shared;
    std::vector<std::unique_ptr<Generator>> gens;
    if (!p.lean) shared.reset(new LockedGenerator(p.seed));
    auto gen_for = [&](int idx) -> Generator* {
        if (!p.lean) return shared.get();
        gens.emplace_back(new Generator(p.seed + idx));
        return gens.back().get();
    };

    std::vector<std::unique_ptr<OwnedChain>> owned;
    std::vector<Chain*> chains;
    auto t0 = clk::now();
    for (int idx = 0; idx < p.n_chains; idx++) {  // root.go:383-430
        auto oc = std::make_unique<OwnedChain>();
        oc->model.reset(new Model(mod.clone()));
        Generator* g = gen_for(idx);
        if (p.kind == kSimple) {
            auto* s = new GibbsSimple(g, oc->model.get());
            s->lean = p.lean;
            oc->sampler.reset(s);
        } else {
            auto* s = new GibbsCollapsed(g, oc->model.get());
            s->base->lean = p.lean;
            oc->sampler.reset(s);
            if (p.kind == kCollapsed) s->collapse(-1);
        }
        oc->chain.reset(new Chain(oc->model.get(), oc->sampler.get(), (int)p.cw, p.burn_in));
        chains.push_back(oc->chain.get());
        owned.push_back(std::move(oc));
    }
    std::unique_ptr<ConvergenceSampler> adapt;
    if (p.kind == kAdaptive) adapt.reset(new ConvergenceSampler(gen_for(p.n_chains), mod.clone(), -1));
    auto t1 = clk::now();

    RunResult res;
    res.burnin_seconds = std::chrono::duration<double>(t1 - t0).count();
    bool keep_working = true;
    while (keep_working) {  // root.go:475-561
        auto ta = clk::now();
        if (p.n_threads == 1) {
            for (auto* ch : chains) ch->advance();
        } else {
            std::vector<std::thread> th;
            std::vector<std::string> errs(chains.size());
            for (size_t i = 0; i < chains.size(); i++)
                th.emplace_back([&, i] {
                    try { chains[i]->advance(); } catch (const std::exception& e) { errs[i] = e.what(); }
                });
            for (auto& t : th) t.join();
            for (auto& e : errs)
                if (!e.empty()) throw Error("Async sample generation failed: " + e);  // chain.go:209-212 panics
        }
        res.advance_seconds += std::chrono::duration<double>(clk::now() - ta).count();
        res.rounds++;
        int64_t cnt = 0;
        for (auto* ch : chains) cnt += ch->total_sample_count;
        res.samples = cnt;
        if (p.max_iters > 0 && cnt > p.max_iters) keep_working = false;
        if (res.rounds >= p.max_rounds) keep_working = false;
        if (sol) {
            ErrorSuite es = sol->error(merge_chains(chains));
            res.curve_samples.push_back(cnt);
            res.curve_mean_hel.push_back(es.mean_hellinger);
            res.curve_max_hel.push_back(es.max_hellinger);
            res.curve_mean_abs.push_back(es.mean_mean_abs);
        }
        if (keep_working && adapt && res.rounds <= p.adapt_rounds) chains = adapt->adapt(chains, p.chain_adds);
    }
    res.merged = merge_chains(chains);  // root.go:565-571 (normalisation left to the caller)
    res.n_chains_final = (int)chains.size();
    return res;
}

}  // namespace oracle
