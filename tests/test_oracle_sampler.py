"""Pins the oracle's sampler layer against the reference's own tests, plus exact-inference
checks for the parts the reference leaves untested (SampleVar numerics).

Each test names the reference test it transcribes (paths relative to /root/reference).
"""
import itertools
import math

import numpy as np
import pytest

import oracle
from oracle import OracleError


# ------------------------------------------------------------------ rand / buffer
def test_mt_canonical_seed():
    """rand/rand_test.go:17-39 TestMTCanonicalSeed (MT19937-64 init_by_array64 vector)"""
    gen = oracle.Generator([0x12345, 0x23456, 0x34567, 0x45678])
    seq = [7266447313870364031, 4946485549665804864, 16945909448695747420,
           16394063075524226720, 4873882236456199058]
    for v in seq:
        assert gen.int63() == (v & 0x7FFFFFFFFFFFFFFF)


def test_mt_bad_seed():
    """rand/rand_test.go:9-15 TestMTBadSeed"""
    with pytest.raises(OracleError):
        oracle.Generator(np.zeros(0, dtype=np.uint64))


def test_mt_single_seed_is_init_genrand64():
    """rand/rand.go:27-28: one-element seed -> Seed(); standard init_genrand64(5489) first output."""
    assert oracle.Generator(5489).int63() == (14514284786278117030 & 0x7FFFFFFFFFFFFFFF)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (the device stream restated in oracle/rng.hpp)."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_uniform_ranges():
    us = [oracle.philox_uniform(42, c, s, v, bits) for c in range(8) for s in range(4) for v in range(4) for bits in (32, 53)]
    assert all(0.0 <= u < 1.0 for u in us)
    assert len(set(us)) == len(us)
    # chains sharing one Philox call (quad / pair) still get distinct words
    assert oracle.philox_uniform(1, 0, 0, 0, 32) != oracle.philox_uniform(1, 1, 0, 0, 32)
    assert oracle.philox_uniform(1, 0, 0, 0, 53) != oracle.philox_uniform(1, 1, 0, 0, 53)


def test_circular_int():
    """buffer/circular_test.go:9-60 TestCircularInt"""
    ci = oracle.CircularInt(6)
    assert ci.buf_size == 6 and ci.count == 0
    for v in (1, 2, 3, 4, 5):
        ci.add(v)
    assert ci.buf_size == 6 and ci.count == 5
    assert ci.first_half() is None and ci.second_half() is None
    ci.add(6)
    assert ci.count == 6
    assert ci.first_half() + ci.second_half() == [1, 2, 3, 4, 5, 6]
    ci.add(8)
    ci.add(8)
    assert ci.first_half() + ci.second_half() == [3, 4, 5, 6, 8, 8]
    assert oracle.CircularInt(7).buf_size == 6  # circular.go:16-19 rounds down to even


# ------------------------------------------------------------------ UniformSampler conventions
def test_uniform_sampler():
    """sampler/sampler_test.go:69-130 TestUniformSampler"""
    gen = oracle.Generator(42)
    with pytest.raises(OracleError):
        gen.uni_sample(0)
    with pytest.raises(OracleError):
        gen.uni_sample((1 << 30) + 1)
    assert gen.uni_sample(1) == 0
    with pytest.raises(OracleError):
        gen.var_sample([], [], False)
    with pytest.raises(OracleError):
        gen.var_sample([-1], [1], True)
    assert gen.var_sample([-1], [1], False) == 0
    seen = set()
    for _ in range(2500):
        seen.add(gen.var_sample([-1, -1], [1, 0], False))
        if seen == {0, 1}:
            break
    assert seen == {0, 1}


def test_uniform_sampler_fixed():
    """sampler/sampler_test.go:132-165 TestUniformSamplerFixed"""
    gen = oracle.Generator(42)
    assert gen.var_sample([0, -1], [0, 0], False) == 1
    with pytest.raises(OracleError):
        gen.var_sample([0, 1], [0, 0], False)
    assert gen.var_sample([-1, -1], [1, 0], True) == 1
    with pytest.raises(OracleError):
        gen.var_sample([-1, -1], [1, 1], True)


def test_weighted_sampler():
    """sampler/sampler_test.go:167-220 TestWeightedSampler"""
    gen = oracle.Generator(42)
    with pytest.raises(OracleError):
        gen.weighted_sample(0, [])
    with pytest.raises(OracleError):
        gen.weighted_sample((1 << 30) + 1, [])
    with pytest.raises(OracleError):
        gen.weighted_sample(1, [])
    with pytest.raises(OracleError):
        gen.weighted_sample(1, [1.0, 1.0])
    with pytest.raises(OracleError):
        gen.weighted_sample(2, [1.0, -1.0])
    assert gen.weighted_sample(1, [1.0]) == 0
    heads = tails = flips = 0
    while heads < 100 or tails < 100:
        i = gen.weighted_sample(2, [100.1, 200.2])
        heads += i == 0
        tails += i == 1
        flips += 1
        assert flips <= 5000
    assert heads / tails == pytest.approx(0.5, rel=0.1)


# ------------------------------------------------------------------ Gibbs simple
def test_working_gibbs_simple(res):
    """sampler/gibbs-simple_test.go:13-38 TestWorkingGibbsSimple"""
    mod = oracle.Model.load(res("one.uai"))
    gen = oracle.Generator(42)
    samp = oracle.Sampler(gen, mod)
    counts = [0, 0]
    for _ in range(1024):
        idx, s = samp.sample([0])
        assert idx == 0
        counts[s[0]] += 1
    assert counts[0] > 0 and counts[1] > 0
    # stronger than the reference: the single factor is 0.25/0.75
    assert counts[1] / 1024 == pytest.approx(0.75, abs=0.06)


def test_gibbs_simple_double_log_space_errors(res):
    """gibbs-simple.go:73-77 + function.go:127-129: a second sampler on the same model fails."""
    mod = oracle.Model.load(res("one.uai"))
    gen = oracle.Generator(1)
    s1 = oracle.Sampler(gen, mod)
    with pytest.raises(OracleError):
        oracle.Sampler(gen, mod)
    del s1


def brute_force_conditional(model_raw, var, state):
    """exact p(x_var | rest) from the raw (non-log) tables with the reference's eps rule applied"""
    cards = model_raw.cards
    w = np.zeros(cards[var])
    for k in range(cards[var]):
        st = list(state)
        st[var] = k
        logp = 0.0
        for f in range(model_raw.n_funcs):
            sc = model_raw.func_scope(f)
            if var not in sc:
                continue
            t = model_raw.func_table(f)
            idx = 0
            for v in sc:
                idx = idx * cards[v] + st[v]
            val = t[idx]
            if val < 1e-6:
                val += 1e-6
            logp += math.log(val)
        w[k] = logp
    p = np.exp(w - w.max())
    return p / p.sum()


def apply_floor(p):
    """sequential 1e-6 floor of gibbs-simple.go:248-258 on an (any-scale) weight vector"""
    e = np.array(p, dtype=float)
    tot = e.sum()
    for k in range(len(e)):
        if e[k] / tot < 1e-6:
            d = tot * 1e-6
            tot += d
            e[k] += d
    return e / e.sum()


@pytest.mark.parametrize("name,evid", [("sample.uai", False), ("deterministic.uai", False), ("Grids_11.uai", False),
                                        ("Promedus_11.uai", True), ("Pedigree_11.uai", True),
                                        ("ObjectDetection_11.uai", False)])
def test_conditional_matches_independent_restatement(res, name, evid):
    """SampleVar numerics are untested in the reference; check the oracle's
    conditional against an independent numpy statement of the same lines (max-shift LSE +
    sequential floor; the floor is scale-invariant so min- vs max-shift must agree)."""
    raw = oracle.Model.load(res(name), use_evidence=evid)
    mod = raw.clone()
    samp = oracle.Sampler(oracle.Generator(7), mod)
    rng = np.random.default_rng(123)
    cards, fixed = raw.cards, raw.fixed
    free = [v for v in range(raw.n_vars) if fixed[v] < 0]
    for _ in range(25):
        state = np.array([fixed[v] if fixed[v] >= 0 else rng.integers(cards[v]) for v in range(raw.n_vars)], dtype=np.int32)
        var = int(rng.choice(free))
        e = samp.conditional(var, state)
        p = e / e.sum()
        ref = apply_floor(brute_force_conditional(raw, var, state))
        assert np.allclose(p, ref, rtol=1e-9, atol=0)
        assert p.min() >= 1e-6 * (1 - 1e-5)


def test_floor_is_sequential():
    """gibbs-simple.go:248-258: tot is updated inside the loop (later bins see the raised total)."""
    tables = np.array([1.0, 1e-9, 1e-9, 1e-9])  # one var card 4, one unary factor
    m = oracle.Model.create([4], [-1], [0, 1], [0], [0, 4], tables)
    s = oracle.Sampler(oracle.Generator(1), m)
    e = s.conditional(0, [0])
    # raw table -> eps rule: 1e-9 + 1e-6 ; log ; min = log(1.001e-6) < -8 -> shifted so min -> 1.5
    raw = np.array([1.0, 1.001e-6, 1.001e-6, 1.001e-6])
    w = np.log(raw)
    w = w - (w.min() - 1.5)
    ee = np.exp(w)
    tot = ee.sum()
    exp = ee.copy()
    for k in range(4):
        if exp[k] / tot < 1e-6:
            d = tot * 1e-6
            tot += d
            exp[k] += d
    assert np.allclose(e, exp, rtol=1e-13)


def exact_marginals(raw):
    """brute-force marginals of a tiny model (reference eps rule applied to the tables)"""
    cards = raw.cards
    n = raw.n_vars
    marg = [np.zeros(c) for c in cards]
    funcs = []
    for f in range(raw.n_funcs):
        t = raw.func_table(f).copy()
        t[t < 1e-6] += 1e-6
        funcs.append((raw.func_scope(f), t))
    fixed = raw.fixed
    for st in itertools.product(*[range(c) if fixed[i] < 0 else [fixed[i]] for i, c in enumerate(cards)]):
        p = 1.0
        for sc, t in funcs:
            idx = 0
            for v in sc:
                idx = idx * cards[v] + st[v]
            p *= t[idx]
        for v in range(n):
            marg[v][st[v]] += p
    return [m / m.sum() for m in marg]


def test_simple_chain_converges_to_exact_marginals(res):
    """End-to-end pin for gibbs-simple.go:163-271 + chain.go:221-246 on sample.uai."""
    raw = oracle.Model.load(res("sample.uai"))
    exact = exact_marginals(raw)
    out = oracle.run(raw, kind=oracle.SIMPLE, n_chains=4, burn_in=2000, cw=2000, max_iters=400000, seed=3, n_threads=1)
    o = 0
    for v, c in enumerate(raw.cards):
        m = out["merged"][o:o + c]
        o += c
        assert np.allclose(m / m.sum(), exact[v], atol=0.01)
    assert out["samples"] > 400000


# ------------------------------------------------------------------ Gibbs collapsed
def test_working_gibbs_collapsed(res):
    """sampler/gibbs-collapsed_test.go:14-48 TestWorkingGibbsCollapsed"""
    mod = oracle.Model.load(res("deterministic.uai"))
    gen = oracle.Generator(42)
    samp = oracle.Sampler(gen, mod.clone(), collapsed=True)
    assert samp.model.collapsed.tolist() == [0, 0, 0]
    for i in range(mod.n_vars):
        samp2 = oracle.Sampler(gen, mod.clone(), collapsed=True)
        v, marg = samp2.collapse(i)
        assert v == i
        assert samp2.model.collapsed.tolist() == [int(j == i) for j in range(3)]
        assert marg[0] == pytest.approx(0.5, rel=1e-5)
        assert marg[1] == pytest.approx(0.5, rel=1e-5)


def test_full_gibbs_collapsed(res):
    """sampler/gibbs-collapsed_test.go:51-111 TestFullGibbsCollapsed"""
    mod = oracle.Model.load(res("sample.uai"))
    gen = oracle.Generator(42)
    samp = oracle.Sampler(gen, mod.clone(), collapsed=True)
    assert samp.model.collapsed.tolist()[:2] == [0, 0]
    v, _ = samp.collapse(0)
    assert v == 0 and samp.model.collapsed.tolist() == [1, 0, 0]
    v, _ = samp.collapse(1)
    assert v == 1 and samp.model.collapsed.tolist() == [1, 1, 0]

    samp = oracle.Sampler(gen, mod, collapsed=True)
    assert int(samp.model.collapsed.sum()) == 0
    samp.collapse(-1)
    assert int(samp.model.collapsed.sum()) == 1
    samp.collapse(-1)
    assert int(samp.model.collapsed.sum()) == 2
    with pytest.raises(OracleError):  # at least one variable must remain uncollapsed
        samp.collapse(-1)
    assert int(samp.model.collapsed.sum()) == 2


def local_marginal(raw, var):
    """sum over the blanket of the product of ONLY the factors touching `var` (gibbs-collapsed.go:205-263)"""
    cards = raw.cards
    funcs = []
    for f in range(raw.n_funcs):
        sc = raw.func_scope(f)
        if var in sc:
            t = raw.func_table(f).copy()
            t[t < 1e-6] += 1e-6
            funcs.append((sc, t))
    blanket = sorted({int(v) for sc, _ in funcs for v in sc})
    fixed = raw.fixed
    m = np.full(cards[var], 1e-12)
    for cfg in itertools.product(*[range(cards[v]) if fixed[v] < 0 else [fixed[v]] for v in blanket]):
        st = dict(zip(blanket, cfg))
        p = 1.0
        for sc, t in funcs:
            idx = 0
            for v in sc:
                idx = idx * cards[v] + st[v]
            p *= t[idx]
        m[st[var]] += p
    return m / m.sum()


def test_collapse_marginal_is_local_marginal(res):
    """gibbs-collapsed.go:205-263: the stored marginal sums the product of the variable's OWN
    factors over its blanket (a local marginal: factors not touching the variable are ignored)."""
    raw = oracle.Model.load(res("sample.uai"))
    samp = oracle.Sampler(oracle.Generator(5), raw.clone(), collapsed=True)
    assert samp.neighbors(1) == [0, 1, 2]
    assert samp.blanket_size(1) == 3 and samp.function_count(1) == 2
    _, marg = samp.collapse(1)
    assert np.allclose(marg, local_marginal(raw, 1), rtol=1e-9)
    # new factor over (0, 2) replaced the two factors of var 1; var 0 keeps its unary
    assert samp.model.n_funcs == 2
    assert samp.model.func_name(1) == "COLLAPSE-B"
    assert samp.model.func_scope(1).tolist() == [0, 2]
    # when the variable's factors ARE the whole model the local marginal is the exact one
    det = oracle.Model.load(res("deterministic.uai"))
    s2 = oracle.Sampler(oracle.Generator(5), det.clone(), collapsed=True)
    _, m1 = s2.collapse(1)
    assert np.allclose(m1, exact_marginals(det)[1], rtol=1e-9)


def test_collapse_errors(res):
    """gibbs-collapsed.go:125-177 error cases"""
    raw = oracle.Model.load(res("Promedus_11.uai"), use_evidence=True)
    samp = oracle.Sampler(oracle.Generator(5), raw.clone(), collapsed=True)
    with pytest.raises(OracleError):
        samp.collapse(158)  # fixed by evidence
    with pytest.raises(OracleError):
        samp.collapse(100000)
    v, _ = samp.collapse(0)
    with pytest.raises(OracleError):
        samp.collapse(0)  # already collapsed


def test_collapsed_chain_converges(res):
    """collapsed sampler: un-collapsed variables keep their exact stationary marginals; a variable
    collapsed in any chain is REPORTED with that chain's local marginal (chain.go:113-129)."""
    raw = oracle.Model.load(res("sample.uai"))
    exact = exact_marginals(raw)
    out = oracle.run(raw, kind=oracle.COLLAPSED, n_chains=4, burn_in=2000, cw=2000, max_iters=300000, seed=11, n_threads=1)
    o = 0
    for v, c in enumerate(raw.cards):
        m = out["merged"][o:o + c]
        o += c
        if out["collapsed"][v]:
            assert np.allclose(m, local_marginal(raw, v), rtol=1e-9), v
        else:
            assert np.allclose(m / m.sum(), exact[v], atol=0.012), v
    assert out["collapsed"].sum() >= 1


# ------------------------------------------------------------------ MergeChains / convergence
def test_merge_chains():
    """sampler/chain_test.go:11-80 TestMergeChains"""
    with pytest.raises(OracleError):
        oracle.merge_chains([], 7, 3)
    cards = [2, 2, 3]
    base = [[0.5, 0.5], [5.1, 5.1], [1.1, 2.2, 3.3]]
    flat = np.concatenate(base)
    ch1 = oracle.Chain.from_marginals(cards, base)

    def one_var_test(chs):
        m, _ = oracle.merge_chains(chs, 7, 3)
        assert np.allclose(m, flat, atol=1e-8)

    m, _ = oracle.merge_chains([ch1, ch1], 7, 3)
    assert np.allclose(m, [1.0, 1.0, 10.2, 10.2, 2.2, 4.4, 6.6], atol=1e-8)
    one_var_test([ch1])
    ch2 = oracle.Chain.from_marginals(cards, base, collapsed=[1, 0, 0])
    for chs in ([ch1, ch2], [ch2, ch1]):
        m, col = oracle.merge_chains(chs, 7, 3)
        assert np.allclose(m, [0.5, 0.5, 10.2, 10.2, 2.2, 4.4, 6.6], atol=1e-8)
        assert col.tolist() == [1, 0, 0]
    one_var_test([ch1])
    one_var_test([ch2])
    ch3 = oracle.Chain.from_marginals(cards, base, collapsed=[1, 1, 1])
    one_var_test([ch1, ch2, ch3])
    one_var_test([ch1])


def test_chain_convergence_formula():
    """chain.go:32-92, 253-290 restated by hand for a 2-chain, 1-variable case."""
    cw = 8
    h = [[0, 0, 0, 1, 1, 1, 0, 1], [1, 1, 1, 1, 0, 1, 1, 1]]
    chains = []
    for seq in h:
        ch = oracle.Chain.from_marginals([2], [[5.5, 3.5]], cw=cw)
        ch.set_history(0, seq)
        chains.append(ch)
    merged = np.array([11.0, 7.0])

    def hell(a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        return math.sqrt(((np.sqrt(a / a.sum()) - np.sqrt(b / b.sum())) ** 2).sum()) / math.sqrt(2)

    W = B = 1e-8
    for seq in h:
        h1 = np.bincount(seq[:4], minlength=2) + 1e-8
        h2 = np.bincount(seq[4:], minlength=2) + 1e-8
        W += hell(h1, h2)
        B += hell(merged, h1 + h2)
    m, n = 2.0, float(cw)
    W /= m
    B *= n / (m - 1)
    vhat = (n - 1) / n * W + (m + 1) / (m * n) * B
    exp = math.sqrt(4 * vhat / (2 * W))
    got = oracle.chain_convergence(chains, oracle.HELLINGER, 1)
    assert got[0] == pytest.approx(exp, rel=1e-12)
    w0, b0 = chains[0].chain_dist(oracle.HELLINGER, 0, merged)
    assert w0 == pytest.approx(hell(np.array([3, 1]) + 1e-8, np.array([1, 3]) + 1e-8), rel=1e-12)
    with pytest.raises(OracleError):
        oracle.chain_convergence(chains[:1], oracle.HELLINGER, 1)
    short = oracle.Chain.from_marginals([2], [[1, 1]], cw=cw)
    short.set_history(0, [0, 1])
    with pytest.raises(OracleError):  # Total seen < Convergence Window
        oracle.chain_convergence([chains[0], short], oracle.HELLINGER, 1)


def test_advance_chain_thresholds(res):
    """chain.go:180-218: every free variable gains at least cw+1 recorded samples per round,
    in batches of 2n steps."""
    mod = oracle.Model.load(res("sample.uai"))
    samp = oracle.Sampler(oracle.Generator(9), mod)
    ch = oracle.Chain(mod, samp, cw=50, burn_in=10)
    assert ch.total_sample_count == 0  # burn-in is not recorded
    ch.advance()
    assert all(ch.total_seen(v) >= 51 for v in range(3))
    assert ch.total_sample_count % 6 == 0
    before = [ch.total_seen(v) for v in range(3)]
    ch.advance()
    assert all(ch.total_seen(v) >= before[v] + 51 for v in range(3))


# ------------------------------------------------------------------ adaptive
def test_adapt(res):
    """sampler/adaptive.go:57-157 (no reference test exists): new chains each collapse one
    eligible variable; lowest convergence scores are taken; no-op at MaxChains."""
    raw = oracle.Model.load(res("Pedigree_11.uai"), use_evidence=True)
    gen = oracle.Generator(17)
    chains, keep = [], []
    for _ in range(2):
        m = raw.clone()
        s = oracle.Sampler(gen, m, collapsed=True)
        c = oracle.Chain(m, s, cw=40, burn_in=500)
        c.advance()
        chains.append(c)
        keep.append((m, s))
    ad = oracle.ConvergenceSampler(gen, raw.clone())
    probe = oracle.Sampler(gen, raw.clone(), collapsed=True)
    fixed = raw.fixed
    cand = [v for v in range(raw.n_vars) if fixed[v] < 0 and 1 < probe.blanket_size(v) <= 12]
    conv = oracle.chain_convergence(chains, oracle.HELLINGER, raw.n_vars)
    new_chains, targets = ad.adapt(chains, 4)
    assert len(new_chains) == 6 and len(targets) == 4
    assert set(targets) <= set(cand)
    worst_kept = max(conv[t] for t in targets)
    others = [conv[v] for v in cand if v not in targets]
    assert worst_kept <= min(others) + 1e-12  # the LOWEST scores were selected (adaptive.go:111-119)
    tot = int(raw.cards.sum())
    merged, col = oracle.merge_chains(new_chains, tot, raw.n_vars)
    assert sorted(np.nonzero(col)[0].tolist()) == sorted(targets)
    with pytest.raises(OracleError):
        ad.adapt(chains[:1], 4)
