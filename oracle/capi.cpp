// ORACLE — TEST INFRASTRUCTURE ONLY (see model.hpp header).
// extern "C" surface over the restatement so tests/ and bench.py's cpu_baseline leg can
// drive it through ctypes.  Every function returns 0 on success, non-zero on a reference
// error (message via orc_last_error()).
#include <cstring>

#include "model.hpp"
#include "rng.hpp"
#include "run.hpp"
#include "sampler.hpp"
#include "sweep.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_END                     \
    }                               \
    catch (const std::exception& e) { \
        g_err = e.what();           \
        return 1;                   \
    }                               \
    return 0;

struct SamplerH {
    std::unique_ptr<FullSampler> s;
    GibbsSimple* simple = nullptr;        // base sampler (always set)
    GibbsCollapsed* collapsed = nullptr;  // set for collapsed samplers
    Model* model = nullptr;
};
struct ChainH {
    std::unique_ptr<Model> owned_model;  // only for orc_chain_from_marginals
    std::unique_ptr<Chain> c;
    Chain* ref = nullptr;  // non-owning (chains produced by Adapt)
    Chain* get() { return c ? c.get() : ref; }
};

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ------------------------------------------------------------------ model
int orc_model_load(const char* path, int use_evidence, void** out) {
    ORC_TRY *out = new Model(model_from_file(path, use_evidence != 0));
    ORC_END
}
int orc_model_from_buffer(const char* data, long len, void** out) {
    ORC_TRY *out = new Model(model_from_buffer(std::string(data, (size_t)len)));
    ORC_END
}
int orc_model_apply_evidence(void* m, const char* data, long len) {
    ORC_TRY apply_evidence(std::string(data, (size_t)len), *(Model*)m);
    ORC_END
}
int orc_model_create(int n_vars, const int* card, const int* fixed, int n_funcs, const int* scope_off,
                     const int* scope_vars, const long long* tab_off, const double* tables, void** out) {
    ORC_TRY
    auto m = std::make_unique<Model>();
    m->type = "MARKOV";
    m->vars.reserve(n_vars);
    for (int i = 0; i < n_vars; i++) {
        m->vars.push_back(new_variable(i, card[i]));
        m->vars.back().fixed_val = fixed ? fixed[i] : -1;
    }
    m->funcs.reserve(n_funcs);
    for (int f = 0; f < n_funcs; f++) {
        std::vector<int> ids(scope_vars + scope_off[f], scope_vars + scope_off[f + 1]), cards;
        for (int v : ids) {
            if (v < 0 || v >= n_vars) throw Error("Invalid var idx for Clique");
            cards.push_back(card[v]);
        }
        Function fn = new_function(f, ids, cards);
        if ((long long)fn.table.size() != tab_off[f + 1] - tab_off[f]) throw Error("table size mismatch");
        std::copy(tables + tab_off[f], tables + tab_off[f + 1], fn.table.begin());
        m->funcs.push_back(std::move(fn));
    }
    m->check();
    *out = m.release();
    ORC_END
}
void orc_model_free(void* m) { delete (Model*)m; }
void* orc_model_clone(void* m) { return new Model(((Model*)m)->clone()); }
int orc_model_check(void* m) {
    ORC_TRY((Model*)m)->check();
    ORC_END
}
int orc_model_n_vars(void* m) { return (int)((Model*)m)->vars.size(); }
int orc_model_n_funcs(void* m) { return (int)((Model*)m)->funcs.size(); }
const char* orc_model_type(void* m) { return ((Model*)m)->type.c_str(); }
void orc_model_cards(void* m, int* out) {
    for (auto& v : ((Model*)m)->vars) *out++ = v.card;
}
void orc_model_fixed(void* m, int* out) {
    for (auto& v : ((Model*)m)->vars) *out++ = v.fixed_val;
}
void orc_model_collapsed(void* m, int* out) {
    for (auto& v : ((Model*)m)->vars) *out++ = v.collapsed ? 1 : 0;
}
void orc_model_set_fixed(void* m, int var, int val) { ((Model*)m)->vars[var].fixed_val = val; }
int orc_model_marginal_size(void* m) {
    int s = 0;
    for (auto& v : ((Model*)m)->vars) s += v.card;
    return s;
}
void orc_model_marginals(void* m, double* out) {
    for (auto& v : ((Model*)m)->vars)
        for (double p : v.marginal) *out++ = p;
}
void orc_model_set_marginals(void* m, const double* in) {
    for (auto& v : ((Model*)m)->vars)
        for (double& p : v.marginal) p = *in++;
}
int orc_model_func_arity(void* m, int f) { return (int)((Model*)m)->funcs[f].vars.size(); }
void orc_model_func_scope(void* m, int f, int* out) {
    for (int v : ((Model*)m)->funcs[f].vars) *out++ = v;
}
long long orc_model_func_tabsize(void* m, int f) { return (long long)((Model*)m)->funcs[f].table.size(); }
void orc_model_func_table(void* m, int f, double* out) {
    for (double t : ((Model*)m)->funcs[f].table) *out++ = t;
}
int orc_model_func_is_log(void* m, int f) { return ((Model*)m)->funcs[f].is_log ? 1 : 0; }
const char* orc_model_func_name(void* m, int f) { return ((Model*)m)->funcs[f].name.c_str(); }

// ------------------------------------------------------------------ Function ops (function_test.go)
int orc_func_eval(void* m, int f, const int* values, int n, double* out) {
    *out = std::nan("");
    ORC_TRY *out = ((Model*)m)->funcs[f].eval(values, (size_t)n);
    ORC_END
}
int orc_func_use_log_space(void* m, int f) {
    ORC_TRY((Model*)m)->funcs[f].use_log_space();
    ORC_END
}
int orc_func_add_value(void* m, int f, const int* values, int n, double inc) {
    ORC_TRY((Model*)m)->funcs[f].add_value(values, (size_t)n, inc);
    ORC_END
}
// stand-alone function over fresh variables (NewFunction + table) for the unit tests
int orc_model_single_function(int n, const int* cards, const double* table, long long tab_len, void** out) {
    ORC_TRY
    auto m = std::make_unique<Model>();
    m->type = "MARKOV";
    std::vector<int> ids, cs;
    for (int i = 0; i < n; i++) {
        m->vars.push_back(new_variable(i, cards[i]));
        ids.push_back(i);
        cs.push_back(cards[i]);
    }
    Function fn = new_function(0, ids, cs);
    if (table) {
        fn.table.assign(table, table + tab_len);  // tests build deliberately bad tables too
    }
    m->funcs.push_back(fn);
    *out = m.release();
    ORC_END
}
int orc_func_check(void* m, int f) {
    ORC_TRY((Model*)m)->funcs[f].check();
    ORC_END
}

// ------------------------------------------------------------------ VariableIter
int orc_variter_enumerate(int n, const int* cards, const int* fixed, int honor, int* out, int max_rows,
                          int* n_rows, int* final_state) {
    ORC_TRY
    std::vector<Variable> vs;
    std::vector<const Variable*> ps;
    for (int i = 0; i < n; i++) {
        vs.push_back(new_variable(i, cards[i]));
        vs.back().fixed_val = fixed[i];
    }
    for (auto& v : vs) ps.push_back(&v);
    VariableIter it(ps, honor != 0);
    std::vector<int> cur(n);
    int rows = 0;
    for (;;) {
        it.val(cur);
        if (rows < max_rows) std::copy(cur.begin(), cur.end(), out + (size_t)rows * n);
        rows++;
        if (!it.next()) break;
    }
    *n_rows = rows;
    for (int i = 0; i < n; i++) final_state[i] = it.last[i];
    ORC_END
}

// ------------------------------------------------------------------ measures / ErrorSuite
static std::vector<Variable> vars_from(int n, const int* cards, const int* fixed, const double* marg) {
    std::vector<Variable> vs;
    for (int i = 0; i < n; i++) {
        Variable v;
        v.id = i;
        v.card = cards[i];
        v.fixed_val = fixed ? fixed[i] : -1;
        v.marginal.assign(marg, marg + cards[i]);
        marg += cards[i];
        vs.push_back(v);
    }
    return vs;
}
int orc_error_suite(int n, const int* cards, const int* fixed1, const double* m1, const int* fixed2,
                    const double* m2, double* out8) {
    ORC_TRY
    ErrorSuite es = new_error_suite(vars_from(n, cards, fixed1, m1), vars_from(n, cards, fixed2, m2));
    out8[0] = es.mean_mean_abs; out8[1] = es.max_mean_abs; out8[2] = es.mean_max_abs; out8[3] = es.max_max_abs;
    out8[4] = es.mean_hellinger; out8[5] = es.max_hellinger; out8[6] = es.mean_js; out8[7] = es.max_js;
    ORC_END
}
int orc_measure(int which, int card, int fixed1, const double* m1, int fixed2, const double* m2, double* out) {
    ORC_TRY
    auto a = vars_from(1, &card, &fixed1, m1), b = vars_from(1, &card, &fixed2, m2);
    *out = measure(which, a[0], b[0]);
    ORC_END
}
int orc_norm_marginal(int card, double* m) {
    ORC_TRY
    Variable v;
    v.card = card;
    v.marginal.assign(m, m + card);
    v.norm_marginal();
    std::copy(v.marginal.begin(), v.marginal.end(), m);
    ORC_END
}

// ------------------------------------------------------------------ UAI helpers
int orc_uai_preprocess(const char* data, long len, const char* prefix, char* out, long cap, int* line_count) {
    ORC_TRY
    std::string s = uai_preprocess(std::string(data, (size_t)len), prefix, *line_count);
    if ((long)s.size() + 1 > cap) throw Error("buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
    ORC_END
}
// MAR solution -> a Model that only carries vars (marginals normalised as the reader does)
int orc_solution_load(const char* path, void** out) {
    ORC_TRY
    Solution s = read_marg_solution(slurp(path));
    auto m = std::make_unique<Model>();
    m->type = "MARKOV";
    m->vars = s.vars;
    *out = m.release();
    ORC_END
}
int orc_solution_from_buffer(const char* data, long len, void** out) {
    ORC_TRY
    Solution s = read_marg_solution(std::string(data, (size_t)len));
    auto m = std::make_unique<Model>();
    m->type = "MARKOV";
    m->vars = s.vars;
    *out = m.release();
    ORC_END
}
int orc_solution_check(void* sol, void* model) {
    ORC_TRY
    Solution s;
    s.vars = ((Model*)sol)->vars;
    s.check(*(Model*)model);
    ORC_END
}

// ------------------------------------------------------------------ rng
int orc_gen_new(const unsigned long long* seed, int n, void** out) {
    *out = nullptr;
    ORC_TRY
    std::vector<uint64_t> s(seed, seed + n);
    *out = new Generator(s);
    ORC_END
}
void orc_gen_free(void* g) { delete (Generator*)g; }
long long orc_gen_int63(void* g) { return ((Generator*)g)->int63(); }
int orc_gen_int31n(void* g, int n) { return ((Generator*)g)->int31n(n); }
double orc_gen_float64(void* g) { return ((Generator*)g)->float64(); }
void orc_philox(const unsigned* ctr, const unsigned* key, unsigned* out) { Philox4x32::gen(ctr, key, out); }
double orc_philox_uniform(unsigned long long seed, unsigned chain, unsigned sweep, unsigned var, int bits) {
    return philox_uniform(seed, chain, sweep, var, bits);
}
int orc_philox_init_value(unsigned long long seed, unsigned chain, unsigned var, int card) {
    return philox_init_value(seed, chain, var, card);
}

// ------------------------------------------------------------------ CircularInt
void* orc_circ_new(int size) { return new CircularInt(size); }
void orc_circ_free(void* c) { delete (CircularInt*)c; }
void orc_circ_add(void* c, int v) { ((CircularInt*)c)->add(v); }
int orc_circ_bufsize(void* c) { return ((CircularInt*)c)->buf_size; }
int orc_circ_count(void* c) { return ((CircularInt*)c)->count; }
// returns the number of values written, or -1 when the halves are not valid yet (nil iterator)
int orc_circ_half(void* c, int second, int* out) {
    auto* ci = (CircularInt*)c;
    if (!ci->halves_valid()) return -1;
    auto v = second ? ci->second_half() : ci->first_half();
    std::copy(v.begin(), v.end(), out);
    return (int)v.size();
}

// ------------------------------------------------------------------ UniformSampler conventions
int orc_uni_sample(void* gen, long long card, int* out) {
    *out = -1;
    ORC_TRY *out = UniformSampler((Generator*)gen, 32).uni_sample(card);
    ORC_END
}
int orc_weighted_sample(void* gen, long long card, const double* w, int len, int* out) {
    *out = -1;
    ORC_TRY *out = UniformSampler((Generator*)gen, 32).weighted_sample(card, w, (size_t)len);
    ORC_END
}
int orc_var_sample(void* gen, int n, const int* fixed, const int* collapsed, int exclude, int* out) {
    *out = -1;
    ORC_TRY
    std::vector<Variable> vs;
    for (int i = 0; i < n; i++) {
        vs.push_back(new_variable(i, 2));
        vs.back().fixed_val = fixed[i];
        vs.back().collapsed = collapsed[i] != 0;
    }
    *out = UniformSampler((Generator*)gen, 32).var_sample(vs, exclude != 0);
    ORC_END
}

// ------------------------------------------------------------------ samplers
int orc_gibbs_simple_new(void* gen, void* model, int lean, void** out) {
    ORC_TRY
    auto h = std::make_unique<SamplerH>();
    auto* s = new GibbsSimple((Generator*)gen, (Model*)model);
    s->lean = lean != 0;
    h->s.reset(s);
    h->simple = s;
    h->model = (Model*)model;
    *out = h.release();
    ORC_END
}
int orc_gibbs_collapsed_new(void* gen, void* model, int lean, void** out) {
    ORC_TRY
    auto h = std::make_unique<SamplerH>();
    auto* s = new GibbsCollapsed((Generator*)gen, (Model*)model);
    s->base->lean = lean != 0;
    h->s.reset(s);
    h->simple = s->base.get();
    h->collapsed = s;
    h->model = (Model*)model;
    *out = h.release();
    ORC_END
}
void orc_sampler_free(void* s) { delete (SamplerH*)s; }
int orc_sampler_sample(void* s, int* state, int n, int* var_idx) {
    *var_idx = -1;
    ORC_TRY
    std::vector<int> st(state, state + n);
    *var_idx = ((SamplerH*)s)->s->sample(st);
    std::copy(st.begin(), st.end(), state);
    ORC_END
}
int orc_sampler_sample_var(void* s, int var, int* state, int n, int* var_idx) {
    *var_idx = -1;
    ORC_TRY
    std::vector<int> st(state, state + n);
    *var_idx = ((SamplerH*)s)->simple->sample_var(var, st);
    std::copy(st.begin(), st.end(), state);
    ORC_END
}
// floored un-normalised weights e[k] (gibbs-simple.go:171-258) for a caller-supplied state
int orc_sampler_conditional(void* s, int var, const int* state, double* w_out) {
    ORC_TRY
    std::vector<double> w;
    ((SamplerH*)s)->simple->conditional(var, state, w);
    std::copy(w.begin(), w.end(), w_out);
    ORC_END
}
void orc_sampler_get_state(void* s, int* out) {
    auto& l = ((SamplerH*)s)->simple->last;
    std::copy(l.begin(), l.end(), out);
}
void orc_sampler_set_state(void* s, const int* in) {
    auto& l = ((SamplerH*)s)->simple->last;
    std::copy(in, in + l.size(), l.begin());
}
int orc_collapsed_collapse(void* s, int var_idx, int* collapsed_var, double* marginal_out) {
    *collapsed_var = -1;
    ORC_TRY
    auto* h = (SamplerH*)s;
    if (!h->collapsed) throw Error("not a collapsed sampler");
    int v = h->collapsed->collapse(var_idx);
    *collapsed_var = v;
    if (marginal_out) std::copy(h->model->vars[v].marginal.begin(), h->model->vars[v].marginal.end(), marginal_out);
    ORC_END
}
int orc_collapsed_blanket_size(void* s, int var) { return ((SamplerH*)s)->collapsed->blanket_size(var); }
int orc_collapsed_function_count(void* s, int var) { return ((SamplerH*)s)->collapsed->function_count(var); }
// neighbours (ascending) of var; returns count
int orc_collapsed_neighbors(void* s, int var, int* out) {
    auto& nb = ((SamplerH*)s)->collapsed->var_neighbors[var];
    int i = 0;
    for (int v : nb) out[i++] = v;
    return i;
}

// ------------------------------------------------------------------ chains
int orc_chain_new(void* model, void* sampler, int cw, long long burn_in, void** out) {
    ORC_TRY
    auto h = std::make_unique<ChainH>();
    FullSampler* fs = sampler ? ((SamplerH*)sampler)->s.get() : nullptr;
    h->c.reset(new Chain((Model*)model, fs, cw, burn_in));
    *out = h.release();
    ORC_END
}
// chain_test.go builds chains around hand-made variables with a nil sampler
int orc_chain_from_marginals(int n, const int* cards, const double* marg, const int* collapsed, int cw, void** out) {
    ORC_TRY
    auto h = std::make_unique<ChainH>();
    h->owned_model.reset(new Model());
    h->owned_model->type = "MARKOV";
    h->owned_model->vars = vars_from(n, cards, nullptr, marg);
    for (int i = 0; i < n; i++) h->owned_model->vars[i].collapsed = collapsed && collapsed[i];
    h->c.reset(new Chain(h->owned_model.get(), nullptr, cw, 0));
    *out = h.release();
    ORC_END
}
void orc_chain_free(void* c) { delete (ChainH*)c; }
int orc_chain_advance(void* c) {
    ORC_TRY((ChainH*)c)->get()->advance();
    ORC_END
}
int orc_chain_one_sample(void* c, int update) {
    ORC_TRY((ChainH*)c)->get()->one_sample(update != 0);
    ORC_END
}
long long orc_chain_total(void* c) { return ((ChainH*)c)->get()->total_sample_count; }
void orc_chain_marginals(void* c, double* out) { orc_model_marginals(((ChainH*)c)->get()->target, out); }
void orc_chain_collapsed(void* c, int* out) { orc_model_collapsed(((ChainH*)c)->get()->target, out); }
void orc_chain_last_sample(void* c, int* out) {
    auto& l = ((ChainH*)c)->get()->last_sample;
    std::copy(l.begin(), l.end(), out);
}
long long orc_chain_total_seen(void* c, int var) { return ((ChainH*)c)->get()->history[var].total_seen; }
int orc_chain_set_history(void* c, int var, const int* samples, int n) {
    ORC_TRY
    Chain* ch = ((ChainH*)c)->get();
    for (int i = 0; i < n; i++) ch->history[var].add(samples[i]);
    ORC_END
}
int orc_chain_dist(void* c, int measure_id, int var, int card, const double* merged, int merged_fixed,
                   int merged_collapsed, double* within, double* between) {
    ORC_TRY
    auto mv = vars_from(1, &card, &merged_fixed, merged);
    mv[0].collapsed = merged_collapsed != 0;
    ((ChainH*)c)->get()->chain_dist(measure_id, var, mv[0], *within, *between);
    ORC_END
}
static std::vector<Chain*> chain_vec(void** chains, int n) {
    std::vector<Chain*> v;
    for (int i = 0; i < n; i++) v.push_back(((ChainH*)chains[i])->get());
    return v;
}
int orc_merge_chains(void** chains, int n, double* marg_out, int* collapsed_out) {
    ORC_TRY
    auto merged = merge_chains(chain_vec(chains, n));
    for (auto& v : merged) {
        for (double p : v.marginal) *marg_out++ = p;
        if (collapsed_out) *collapsed_out++ = v.collapsed ? 1 : 0;
    }
    ORC_END
}
int orc_chain_convergence(void** chains, int n, int measure_id, double* out) {
    ORC_TRY
    auto vals = chain_convergence(chain_vec(chains, n), measure_id, {});
    std::copy(vals.begin(), vals.end(), out);
    ORC_END
}

// ------------------------------------------------------------------ adaptive
int orc_adapt_new(void* gen, void* model, int measure_or_neg, void** out) {
    ORC_TRY *out = new ConvergenceSampler((Generator*)gen, *(Model*)model, measure_or_neg);
    ORC_END
}
void orc_adapt_free(void* a) { delete (ConvergenceSampler*)a; }
// chains_out receives handles: the first n_in are the inputs, the rest are NEW non-owning
// handles (free with orc_chain_free; the chains themselves live as long as the adapter).
int orc_adapt_adapt(void* a, void** chains_in, int n_in, int new_count, void** chains_out, int cap, int* n_out,
                    int* targets_out, int* n_targets) {
    ORC_TRY
    auto* cs = (ConvergenceSampler*)a;
    auto res = cs->adapt(chain_vec(chains_in, n_in), new_count);
    if ((int)res.size() > cap) throw Error("chains_out too small");
    for (size_t i = 0; i < res.size(); i++) {
        if ((int)i < n_in) chains_out[i] = chains_in[i];
        else {
            auto* h = new ChainH();
            h->ref = res[i];
            chains_out[i] = h;
        }
    }
    *n_out = (int)res.size();
    *n_targets = (int)cs->last_targets.size();
    for (size_t i = 0; i < cs->last_targets.size(); i++) targets_out[i] = cs->last_targets[i];
    ORC_END
}

// ------------------------------------------------------------------ whole-run driver (root.go restated)
// out_info: [0]=samples [1]=rounds [2]=n_chains_final ; out_secs: [0]=burn-in [1]=advance
// curve_*: up to curve_cap points (samples, mean Hellinger, max Hellinger, mean abs error)
int orc_run(void* model, void* solution_or_null, int kind, int n_chains, long long burn_in, long long cw,
            long long max_iters, long long seed, int lean, int n_threads, int chain_adds, int adapt_rounds,
            int max_rounds, double* merged_out, int* collapsed_out, long long* out_info, double* out_secs,
            int curve_cap, long long* curve_samples, double* curve_mean_hel, double* curve_max_hel,
            double* curve_mean_abs, int* curve_n) {
    ORC_TRY
    RunParams p;
    p.kind = kind; p.n_chains = n_chains; p.burn_in = burn_in; p.cw = cw; p.max_iters = max_iters;
    p.seed = seed; p.lean = lean != 0; p.n_threads = n_threads; p.chain_adds = chain_adds;
    if (adapt_rounds >= 0) p.adapt_rounds = adapt_rounds;
    if (max_rounds > 0) p.max_rounds = max_rounds;
    Solution sol;
    if (solution_or_null) sol.vars = ((Model*)solution_or_null)->vars;
    RunResult r = run_marginals(*(Model*)model, solution_or_null ? &sol : nullptr, p);
    for (auto& v : r.merged) {
        for (double q : v.marginal) *merged_out++ = q;
        if (collapsed_out) *collapsed_out++ = v.collapsed ? 1 : 0;
    }
    out_info[0] = r.samples; out_info[1] = r.rounds; out_info[2] = r.n_chains_final;
    out_secs[0] = r.burnin_seconds; out_secs[1] = r.advance_seconds;
    int n = 0;
    if (curve_n) {
        for (size_t i = 0; i < r.curve_samples.size() && n < curve_cap; i++, n++) {
            curve_samples[n] = r.curve_samples[i];
            curve_mean_hel[n] = r.curve_mean_hel[i];
            curve_max_hel[n] = r.curve_max_hel[i];
            curve_mean_abs[n] = r.curve_mean_abs[i];
        }
        *curve_n = n;
    }
    ORC_END
}

// Throughput probe for the CPU baseline: `n_threads` independent chains (one std::thread
// each, own MT19937-64 seeded seed+i) over a model already in a sampler-ready clone; each
// runs `steps` recorded single-variable updates (Chain.oneSample(true), chain.go:221-246).
// Returns wall seconds of the stepping phase only (construction excluded).
int orc_throughput(void* model, int kind, int n_threads, long long steps, long long seed, int lean, int cw,
                   double* seconds_out, long long* updates_out) {
    ORC_TRY
    struct W {
        std::unique_ptr<Model> m;
        std::unique_ptr<Generator> g;
        std::unique_ptr<FullSampler> s;
        std::unique_ptr<Chain> c;
        std::string err;
    };
    std::vector<W> ws(n_threads);
    for (int i = 0; i < n_threads; i++) {
        ws[i].m.reset(new Model(((Model*)model)->clone()));
        ws[i].g.reset(new Generator(seed + i));
        if (kind == kSimple) {
            auto* s = new GibbsSimple(ws[i].g.get(), ws[i].m.get());
            s->lean = lean != 0;
            ws[i].s.reset(s);
        } else {
            auto* s = new GibbsCollapsed(ws[i].g.get(), ws[i].m.get());
            s->base->lean = lean != 0;
            ws[i].s.reset(s);
            if (kind == kCollapsed) s->collapse(-1);
        }
        ws[i].c.reset(new Chain(ws[i].m.get(), ws[i].s.get(), cw, 0));
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; i++)
        th.emplace_back([&, i] {
            try {
                for (long long k = 0; k < steps; k++) ws[i].c->one_sample(true);
            } catch (const std::exception& e) { ws[i].err = e.what(); }
        });
    for (auto& t : th) t.join();
    *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long tot = 0;
    for (auto& w : ws) {
        if (!w.err.empty()) throw Error(w.err);
        tot += w.c->total_sample_count;
    }
    *updates_out = tot;
    ORC_END
}

// ------------------------------------------------------------------ device-schedule sweeps
// states: [n_chains][n_vars] ints, in/out.  counts: [sum card] doubles, accumulated.
int orc_sweep_run(void* sampler, const int* order, int n_order, unsigned long long seed, unsigned chain0,
                  int n_chains, unsigned sweep0, unsigned n_sweeps, int bits, int record, int* states,
                  double* counts) {
    ORC_TRY
    auto* h = (SamplerH*)sampler;
    std::vector<int> ord(order, order + n_order), off;
    int acc = 0;
    for (auto& v : h->model->vars) { off.push_back(acc); acc += v.card; }
    size_t nv = h->model->vars.size();
    for (int c = 0; c < n_chains; c++)
        sweep_chain(*h->simple, ord, seed, chain0 + (unsigned)c, sweep0, n_sweeps, bits, record != 0,
                    states + (size_t)c * nv, off, counts, nullptr, record == 2);  // record 2 = Rao-Blackwell bins
    ORC_END
}

int orc_sweep_run_mixed(void* sampler, const int* order, int n_order, unsigned long long seed, unsigned chain0,
                        int n_chains, unsigned sweep0, unsigned n_sweeps, const int* var_bits, int record, int* states,
                        double* counts) {
    ORC_TRY
    auto* h = (SamplerH*)sampler;
    std::vector<int> ord(order, order + n_order), off;
    int acc = 0;
    for (auto& v : h->model->vars) { off.push_back(acc); acc += v.card; }
    size_t nv = h->model->vars.size();
    for (int c = 0; c < n_chains; c++)
        sweep_chain(*h->simple, ord, seed, chain0 + (unsigned)c, sweep0, n_sweeps, 53, record != 0,
                    states + (size_t)c * nv, off, counts, var_bits, record == 2);  // record 2 = Rao-Blackwell bins
    ORC_END
}

int orc_scan_run(void* sampler, const int* order, int n_order, unsigned long long seed, unsigned chain0, int n_chains,
                 unsigned long long step0, long long n_steps, int record, int* states, double* counts) {
    ORC_TRY
    auto* h = (SamplerH*)sampler;
    std::vector<int> ord(order, order + n_order), off;
    int acc = 0;
    for (auto& v : h->model->vars) { off.push_back(acc); acc += v.card; }
    size_t nv = h->model->vars.size();
    for (int c = 0; c < n_chains; c++)
        scan_chain(*h->simple, ord, seed, chain0 + (unsigned)c, step0, n_steps, record != 0, states + (size_t)c * nv, off, counts);
    ORC_END
}

}  // extern "C"
