"""The committed golden vectors (tests/golden/make_golden.py) against the oracle and the host logic —
pins the parts of the path the reference's own tests leave open (SampleVar numerics, sweep schedule)
against accidental change.  CPU only; the CUDA path is compared with the same files in
test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_golden_files_present():
    assert len(_load("conditionals.json")) >= 100 and len(_load("trajectories.json")) == 4


def test_oracle_conditionals_match_golden(res):
    samplers = {}
    for rec in _load("conditionals.json"):
        key = (rec["model"], rec["evidence"])
        if key not in samplers:
            samplers[key] = oracle.Sampler(oracle.Generator(1), oracle.Model.load(res(rec["model"]), use_evidence=rec["evidence"]))
        e = samplers[key].conditional(rec["var"], np.asarray(rec["state"], dtype=np.int32))
        assert [repr(float(x)) for x in e] == rec["e"], (rec["model"], rec["var"])  # float64, bit for bit


def test_golden_conditionals_are_the_reference_formula(res):
    """independent numpy statement of gibbs-simple.go:186-258 on a few golden records"""
    for rec in _load("conditionals.json")[::7]:
        om = oracle.Model.load(res(rec["model"]), use_evidence=rec["evidence"])
        st, v = np.asarray(rec["state"]), rec["var"]
        card = int(om.cards[v])
        w = np.zeros(card)
        for f in range(om.n_funcs):
            scope = [int(x) for x in om.func_scope(f)]
            if v not in scope:
                continue
            tab = np.asarray(om.func_table(f), dtype=np.float64)
            tab = np.log(np.where(tab < 1e-6, tab + 1e-6, tab))  # function.go:126-142
            cards = [int(om.cards[u]) for u in scope]
            for k in range(card):
                idx = 0
                for u, c in zip(scope, cards):
                    idx = idx * c + (k if u == v else int(st[u]))
                w[k] += tab[idx]
        if w.min() < -8:
            w = w - (w.min() - 1.5)
        e = np.exp(w)
        tot = e.sum()
        for k in range(card):
            if e[k] / tot < 1e-6:
                d = tot * 1e-6
                tot += d
                e[k] += d
        np.testing.assert_allclose(e, [float(x) for x in rec["e"]], rtol=1e-12)


def test_oracle_trajectories_match_golden(res):
    for t in _load("trajectories.json"):
        order = [v for c in t["colours"] for v in c]
        for bits in (53, 32):
            g = t.get("bits%d" % bits)
            if g is None:
                continue
            s = oracle.Sampler(oracle.Generator(1), oracle.Model.load(res(t["model"]), use_evidence=t["evidence"]))
            st, counts = s.sweep_run(order, t["seed"], t["first_chain"], np.asarray(t["initial"], dtype=np.int32), 0,
                                     t["n_sweeps"], bits=bits, record=True)
            assert st.tolist() == g["final"], (t["model"], bits)
            assert [int(c) for c in counts] == g["counts"], (t["model"], bits)
            if "rb_bins" in g:
                _, bins = s.sweep_run(order, t["seed"], t["first_chain"], np.asarray(t["initial"], dtype=np.int32), 0,
                                      t["n_sweeps"], bits=bits, record=2)
                assert [int(b) for b in bins] == g["rb_bins"], t["model"]


def test_host_schedule_matches_golden_colouring(res):
    """HostModel::build_colouring (host-only model, no device): same colour classes as the golden schedule;
    the order INSIDE a colour is free (localise_order) and does not change a sweep"""
    gb = pytest.importorskip("grample_b200")
    for t in _load("trajectories.json"):
        m = gb.Model.from_uai(res(t["model"]), use_evidence=t["evidence"], device=-1)
        order, coff = m.schedule()
        got = [sorted(int(v) for v in order[coff[c]:coff[c + 1]]) for c in range(len(coff) - 1)]
        assert got == [sorted(c) for c in t["colours"]], t["model"]


def test_oracle_rao_blackwell_bins(res):
    """oracle replay of GB_CHAINS_RAO_BLACKWELL (sweep_run(record=2)): same trajectory as the counting run; every
    recorded update spreads 2^24 units over the variable's bins in proportion to the floored weights of
    gibbs-simple.go:239-258 — re-derived here from the golden initial state for the first update of a trajectory"""
    for t in _load("trajectories.json"):
        order = [v for c in t["colours"] for v in c]
        s = oracle.Sampler(oracle.Generator(1), oracle.Model.load(res(t["model"]), use_evidence=t["evidence"]))
        init = np.asarray(t["initial"], dtype=np.int32)
        st_c, counts = s.sweep_run(order, t["seed"], t["first_chain"], init, 0, t["n_sweeps"], bits=53, record=True)
        st_r, bins = s.sweep_run(order, t["seed"], t["first_chain"], init, 0, t["n_sweeps"], bits=53, record=2)
        assert np.array_equal(st_c, st_r)
        cards = s.model.cards
        off = np.concatenate([[0], np.cumsum(cards)])
        n_chains = init.shape[0]
        for v in order:
            assert counts[off[v]:off[v + 1]].sum() == t["n_sweeps"] * n_chains
            assert abs(bins[off[v]:off[v + 1]].sum() * 2.0 ** -24 - t["n_sweeps"] * n_chains) < 1e-4
        # first update of chain 0 in isolation: one variable, one sweep, state = the initial state
        v0 = order[0]
        _, b1 = s.sweep_run([v0], t["seed"], t["first_chain"], init[:1], 0, 1, bits=53, record=2)
        e = np.asarray(s.conditional(v0, init[0]), dtype=np.float64)
        expect = np.floor(e * (16777216.0 / e.sum()) + 0.5)
        assert np.array_equal(b1[off[v0]:off[v0 + 1]], expect), t["model"]
