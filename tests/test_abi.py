"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/grample_b200.h declares, fails loudly without a GPU, and the host-side model logic
(UAI parsing, flattening, colouring, scoring) agrees with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import grample_b200 as gb
import oracle
from grample_b200 import _lib

HAS_GPU = gb.device_count() > 0


def header_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/grample_b200.h but not exported"
    assert declared == gb.exported_symbols()  # the Python binding covers the whole header
    assert lib.gb_version() == 100


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(res):
    with pytest.raises(gb.GrampleError, match="no CUDA device"):
        gb.Model.from_uai(res("one.uai"), device=0)
    m = gb.Model.from_uai(res("Grids_11.uai"), device=-1)  # host-only metadata is allowed
    with pytest.raises(gb.GrampleError, match="no CPU fallback"):
        m.conditional(np.zeros((1, 100), np.int32), [0])
    with pytest.raises(gb.GrampleError, match="no CPU fallback"):
        m.collapse(0)
    with pytest.raises(gb.GrampleError):
        gb.Chains(m, 4, device=0)


@pytest.mark.parametrize("name,evid", [("one.uai", False), ("sample.uai", False), ("Grids_11.uai", False),
                                        ("Promedus_11.uai", True), ("Pedigree_11.uai", True),
                                        ("ObjectDetection_11.uai", False), ("dv-rel_1.uai", True)])
def test_uai_reader_and_log_tables_match_oracle(res, name, evid):
    dm = gb.Model.from_uai(res(name), use_evidence=evid, device=-1)
    om = oracle.Model.load(res(name), use_evidence=evid)
    assert dm.n_vars == om.n_vars and dm.n_funcs == om.n_funcs
    assert np.array_equal(dm.cards, om.cards) and np.array_equal(dm.fixed, om.fixed)
    osamp = oracle.Sampler(oracle.Generator(1), om, collapsed=True)  # converts the tables to log space
    for f in range(dm.n_funcs):
        assert np.array_equal(dm.func_scope(f), om.func_scope(f))
        assert np.array_equal(dm.func_log_table(f), om.func_table(f))  # same eps rule, same libm log
    for v in range(dm.n_vars):
        assert dm.blanket_size(v) == osamp.blanket_size(v)
        assert dm.function_count(v) == osamp.function_count(v)
    # from_arrays on the oracle's flattening gives the same model
    dm2 = gb.Model.from_arrays(*oracle.Model.load(res(name), use_evidence=evid).flatten(), device=-1)
    assert np.array_equal(dm2.schedule()[0], dm.schedule()[0])


def test_model_errors(res, tmp_path):
    """uai.go / model.go / function.go error cases"""
    bad = tmp_path / "bad.uai"
    for text in ("", "MARKOV\n1\n2\n1\n1 0\n3\n 0.1 0.2 0.7\n", "FOO\n1\n2\n1\n1 0\n2\n 0.5 0.5\n",
                 "MARKOV\n1\n2\n1\n1 5\n2\n 0.5 0.5\n", "MARKOV\n1\n2\n1\n1 0\n2\n 0.5 abc\n"):
        bad.write_text(text)
        with pytest.raises(gb.GrampleError):
            gb.Model.from_uai(str(bad), device=-1)
    with pytest.raises(gb.GrampleError):
        gb.Model.from_uai(str(tmp_path / "missing.uai"), device=-1)
    # all variables fixed (model.go:138-140); fixed value out of range (variable.go:84-88)
    with pytest.raises(gb.GrampleError):
        gb.Model.from_arrays([2], [1], [0, 1], [0], [0, 2], [0.5, 0.5], device=-1)
    with pytest.raises(gb.GrampleError):
        gb.Model.from_arrays([2, 2], [-1, 2], [0, 2], [0, 1], [0, 4], [1, 1, 1, 1], device=-1)
    # a variable in no factor (gibbs-simple.go:96-99)
    with pytest.raises(gb.GrampleError):
        gb.Model.from_arrays([2, 2], [-1, -1], [0, 1], [0], [0, 2], [0.5, 0.5], device=-1)
    # table larger than maxTabSize (function.go:77-79)
    with pytest.raises(gb.GrampleError):
        gb.Model.from_arrays([2] * 24, [-1] * 24, [0, 24], list(range(24)), [0, 1 << 24], np.ones(1 << 24), device=-1)
    # a scope that names a variable twice: the reference reads both slots from the state (only the diagonal is reachable);
    # the fast paths assume distinct scope variables, so it is refused with a clear message instead of sampled inconsistently
    with pytest.raises(gb.GrampleError, match="appears twice"):
        gb.Model.from_arrays([2, 2], [-1, -1], [0, 1, 3], [1, 0, 0], [0, 2, 6], [0.5, 0.5, 1, 2, 3, 4], device=-1)


def test_table_programs_cover_cardinality_up_to_four(res):
    """host side of table mode for cardinality 3 / 4 (function.go:180-202 index space): Pedigree_11 without evidence has 23
    ternary variables; every sampled variable gets a threshold table, so GB_HYBRID runs it on the integer kernels; GB_TABLE
    (binary only) and the bit-sliced mode still refuse it; Grids_11 qualifies for all three."""
    ped = gb.Model.from_uai(res("Pedigree_11.uai"), device=-1)
    order, _ = ped.schedule()
    assert (ped.cards[order] == 3).sum() == 23
    assert ped.hybrid_mask()[order].all()
    assert not ped.table_mode()[0] and not ped.bits_mode()
    grid = gb.Model.from_uai(res("Grids_11.uai"), device=-1)
    assert grid.table_mode() == (True, 1600) and grid.bits_mode()
    od = gb.Model.from_uai(res("ObjectDetection_11.uai"), device=-1)
    assert not od.hybrid_mask().any() and not od.bits_mode()


def test_schedule_is_proper_colouring(res):
    for name, evid, ncol in (("Grids_11.uai", False, 2), ("Promedus_11.uai", True, 4), ("Pedigree_11.uai", True, 4),
                             ("ObjectDetection_11.uai", False, 7)):
        dm = gb.Model.from_uai(res(name), use_evidence=evid, device=-1)
        order, coff = dm.schedule()
        assert len(coff) - 1 == ncol  # SURVEY §8a (Pedigree_11: 5 with the id-order greedy colouring, 4 with smallest-last)
        fixed = dm.fixed
        assert sorted(order.tolist()) == [v for v in range(dm.n_vars) if fixed[v] < 0]
        colour = {}
        for c in range(ncol):
            for v in order[coff[c]:coff[c + 1]]:
                colour[int(v)] = c
        for f in range(dm.n_funcs):
            sc = [int(v) for v in dm.func_scope(f) if int(v) in colour]
            assert len({colour[v] for v in sc}) == len(set(sc))


def test_ising_generator_layout():
    card, fixed, scope_off, scope_vars, tab_off, tables = gb.ising_torus(10, 10, seed=1)
    assert len(card) == 100 and len(scope_off) == 301 and tab_off[-1] == 1000
    scopes = [scope_vars[scope_off[f]:scope_off[f + 1]].tolist() for f in range(300)]
    assert scopes[:3] == [[0], [1], [2]]
    assert scopes[100:110] == [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 8], [8, 9], [0, 9]]
    assert scopes[200:210] == [[0, 10], [10, 20], [20, 30], [30, 40], [40, 50], [50, 60], [60, 70], [70, 80], [80, 90], [0, 90]]
    t = tables[tab_off[100]:tab_off[101]]
    assert t[0] == t[3] and t[1] == t[2] and t[0] * t[1] == pytest.approx(1.0)
    dm = gb.Model.from_arrays(card, fixed, scope_off, scope_vars, tab_off, tables, device=-1)
    order, coff = dm.schedule()
    assert coff.tolist() == [0, 50, 100]
    assert all(((v // 10) + (v % 10)) % 2 == 0 for v in order[:50])
    om = oracle.Model.create(card, fixed, scope_off, scope_vars, tab_off, tables)
    assert om.n_funcs == 300


def test_error_suite_and_mar_reader_match_reference_vectors(res):
    """model/error_test.go:92-119 and uai_test.go:144-171 through the product's host API"""
    es = gb.error_suite([3, 3], [30.0, 40.0, 30.0] * 2, [90.0, 5.0, 5.0, 60.0, 30.0, 10.0])
    exp = dict(MeanMeanAbsError=.30000000, MaxMeanAbsError=.39999999, MeanMaxAbsError=.45000000,
               MaxMaxAbsError=.60000000, MeanHellinger=.35109087, MaxHellinger=.46528369,
               MeanJSDiverge=.18806933, MaxJSDiverge=.29645726)
    for k, v in exp.items():
        assert es[k] == pytest.approx(v, rel=1e-7), k
    es = gb.error_suite([2, 2], [250.0, 750.0, 25.1, 75.3], [42.0, 42.0, 3.1, 3.1])
    assert es["MeanHellinger"] == pytest.approx(0.18459191128251448, rel=1e-8)
    assert es["MeanJSDiverge"] == pytest.approx(0.0487949406953985, rel=1e-8)
    with pytest.raises(gb.GrampleError):
        gb.error_suite([2], [1, 1], [1, 1], fixed1=[1])
    cards, marg = gb.mar_load(res("one.uai.MAR"))
    assert cards.tolist() == [2] and marg.tolist() == [0.25, 0.75]
    cards, marg = gb.mar_load(res("Grids_11.uai.merlin.MAR"))  # PR section is skipped
    assert len(cards) == 100 and marg[0] == pytest.approx(0.997878)
    for name in ("Grids_11", "Promedus_11", "Pedigree_11", "ObjectDetection_11"):
        cards, marg = gb.mar_load(res(f"{name}.uai.MAR"))
        s = oracle.solution_load(res(f"{name}.uai.MAR"))
        assert np.array_equal(cards, s.cards) and np.array_equal(marg, s.marginals)
        rng = np.random.default_rng(0)
        other = rng.random(len(marg)) + 0.01
        a = gb.error_suite(cards, marg, other)
        b = oracle.error_suite(cards, s.marginal_list(), np.split(other, np.cumsum(cards)[:-1]))
        for k in a:
            assert a[k] == pytest.approx(b[k], rel=1e-12)


# ------------------------------------------------------------------ CLI host logic (no device needed)
def test_cli_letter26_and_variable_json():
    """model/variable.go:167-189 naming and the encoding/json field order of model.Variable"""
    import json

    from grample_b200 import cli
    assert [cli.letter26(i) for i in (0, 1, 25, 26, 27, 51, 52, 701, 702)] == ["A", "B", "Z", "AA", "AB", "AZ", "BA", "ZZ", "AAA"]
    line = cli.variable_json(27, 2, -1, [0.25, 0.75], {"JS-Error": 0.5, "Hell-Convergence": 1.5}, True)
    assert line == '{"ID":27,"Name":"AB","Card":2,"FixedVal":-1,"Marginal":[0.25,0.75],"State":{"Hell-Convergence":1.5,"JS-Error":0.5},"Collapsed":true}'
    assert json.loads(line)["State"]["JS-Error"] == 0.5


def test_cli_per_var_measures_match_oracle():
    """the per-variable *-Error states (cmd/root.go:651-655) against the oracle's model/error.go restatement"""
    import numpy as np

    import oracle
    from grample_b200 import cli
    rng = np.random.default_rng(3)
    cards = np.array([2, 3, 5, 2], dtype=np.int32)
    p, q = rng.random(cards.sum()), rng.random(cards.sum())
    fixed = np.array([-1, -1, -1, 1], dtype=np.int32)
    got = cli.per_var_measures(cards, p, q, fixed_p=fixed)
    offs = np.concatenate([[0], np.cumsum(cards)])
    for name, which in (("MaxAD", 0), ("AvgAD", 1), ("Hell", 2), ("JS", 3)):
        for v in range(4):
            ref = oracle.measure(which, p[offs[v]:offs[v + 1]], q[offs[v]:offs[v + 1]], fixed1=int(fixed[v]))
            assert abs(got[name][v] - ref) < 1e-12, (name, v)
    assert got["Hell"][3] == 0.0


def test_cli_monitor_gauges_http():
    """cmd/monitor.go:44-67: expvar names, every path redirects to /debug/vars"""
    import json
    import urllib.request

    from grample_b200 import cli
    mon = cli.Monitor()
    assert len(mon.snapshot()) == 14
    mon.set("Burn-In", 7)
    mon.add("Total-Chain-Count", 2)
    mon.start("127.0.0.1:18765")
    try:
        body = json.loads(urllib.request.urlopen("http://127.0.0.1:18765/", timeout=5).read())
    finally:
        mon.stop()
    assert body["Burn-In"] == 7 and body["Total-Chain-Count"] == 2 and "Last-Max-JSD" in body
