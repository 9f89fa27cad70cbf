"""Pins the oracle's model layer against the reference's own known-answer tests.

Each test names the reference test it transcribes (paths relative to /root/reference).
"""
import math

import numpy as np
import pytest

import oracle
from oracle import OracleError

PASCAL = """MARKOV
3
2 2 3
3
1 0
2 0 1
2 1 2

2
 0.436 0.564

4
 0.128 0.872
 0.920 0.080

6
 0.210 0.333 0.457
 0.811 0.000 0.189
"""

TABLE_2x3 = [0.01, 1.02, 2.03, 3.04, 4.05, 5.06]


def test_func_eval_known_answers():
    """model/function_test.go:81-157 TestFuncTestEval"""
    m = oracle.Model.single_function([2, 3], TABLE_2x3)
    cases = [([0, 0], 0.01), ([0, 1], 1.02), ([0, 2], 2.03), ([1, 0], 3.04), ([1, 1], 4.05), ([1, 2], 5.06)]
    prod = 1.0
    for vals, exp in cases:
        v, err = m.func_eval(0, vals)
        assert err is None
        assert v == pytest.approx(exp, rel=1e-14)
        prod *= v
    for bad in ([], [0], [0, 0, 0], [2, 0], [0, 3]):
        v, err = m.func_eval(0, bad)
        assert err is not None
        assert math.isnan(v)
    assert not m.func_is_log(0)
    m.func_use_log_space(0)
    assert m.func_is_log(0)
    with pytest.raises(OracleError):
        m.func_use_log_space(0)
    assert m.func_is_log(0)
    log_sum = sum(m.func_eval(0, vals)[0] for vals, _ in cases)
    assert math.exp(log_sum) == pytest.approx(prod, rel=1e-14)


def test_func_log_space_eps_rule():
    """model/function.go:131-137: v += 1e-6 only when v < 1e-6 (not max)."""
    m = oracle.Model.single_function([4], [0.0, 5e-7, 1e-6, 0.3])
    m.func_use_log_space(0)
    t = m.func_table(0)
    assert t[0] == math.log(1e-6)
    assert t[1] == math.log(5e-7 + 1e-6)
    assert t[2] == math.log(1e-6)
    assert t[3] == math.log(0.3)


def test_func_buildup():
    """model/function_test.go:192-251 TestFuncBuildup"""
    m = oracle.Model.single_function([2, 3])
    rows, _ = oracle.variter_enumerate([2, 3], [-1, -1], False)
    for r in rows:
        m.func_add_value(0, r, 1.0)
    t = m.func_table(0)
    assert len(t) == 6
    assert np.allclose(t, 1.0, rtol=1e-6)
    for r in rows:
        m.func_add_value(0, r, 2.42)
    assert np.allclose(m.func_table(0), 3.42, rtol=1e-6)
    m.func_use_log_space(0)
    with pytest.raises(OracleError):
        m.func_add_value(0, [0, 0], 123.45)
    assert m.func_table(0)[0] == pytest.approx(math.log(3.42), rel=1e-6)


def test_func_checks():
    """model/function_test.go:19-78 TestFuncBadCheck / TestFuncGoodCheck"""
    good = oracle.Model.single_function([2, 3], TABLE_2x3)
    good.func_check(0)
    bad = oracle.Model.single_function([2, 3], TABLE_2x3[:5])
    with pytest.raises(OracleError):
        bad.func_check(0)
    with pytest.raises(OracleError):
        oracle.Model.single_function([])  # empty variable list
    with pytest.raises(OracleError):
        oracle.Model.single_function([2] * 24)  # 2^24 > maxTabSize 2^23
    oracle.Model.single_function([2] * 23)  # exactly the cap is allowed


def test_var_iter_order():
    """model/variable_iter_test.go:40-65 TestVarIter — last variable fastest"""
    rows, final = oracle.variter_enumerate([2, 3, 2], [-1, -1, -1], False)
    exp = [[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [0, 2, 0], [0, 2, 1],
           [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1], [1, 2, 0], [1, 2, 1]]
    assert rows.tolist() == exp
    assert final.tolist() == [0, 0, 0]  # wrapped around


def test_var_iter_corners_and_fixed():
    """model/variable_iter_test.go:67-138 TestVarIterCorners / TestVarIterFixedVals"""
    with pytest.raises(OracleError):
        oracle.variter_enumerate([], [], False)
    rows, _ = oracle.variter_enumerate([2], [-1], False)
    assert rows.tolist() == [[0], [1]]
    rows, final = oracle.variter_enumerate([2, 2], [1, 0], True)  # all fixed
    assert rows.tolist() == [[1, 0]]
    assert final.tolist() == [1, 0]
    rows, final = oracle.variter_enumerate([2, 2, 2], [1, -1, -1], True)
    assert rows.tolist() == [[1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]]
    assert final.tolist() == [1, 0, 0]
    rows, _ = oracle.variter_enumerate([2, 2, 2], [-1, 1, -1], True)
    assert rows.tolist() == [[0, 1, 0], [0, 1, 1], [1, 1, 0], [1, 1, 1]]
    rows, _ = oracle.variter_enumerate([2, 2, 2], [-1, -1, 1], True)
    assert rows.tolist() == [[0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 1]]
    # honorFixed=False ignores evidence
    rows, _ = oracle.variter_enumerate([2, 2], [1, 0], False)
    assert len(rows) == 4


def test_error_suite_normed():
    """model/error_test.go:11-89 TestErrorSuiteNormed"""
    hell = math.sqrt((math.sqrt(0.75) - math.sqrt(0.5)) ** 2 + (math.sqrt(0.25) - math.sqrt(0.5)) ** 2) / math.sqrt(2)
    assert hell == pytest.approx(0.18459191128251448, rel=1e-12)
    js = 0.0487949406953985
    v1 = [[250.0, 750.0], [25.1, 75.3]]
    v2 = [[42.0, 42.0], [3.1, 3.1]]

    def check(a, b):
        es = oracle.error_suite([2, 2], a, b)
        for k in ("MeanMeanAbsError", "MaxMeanAbsError", "MeanMaxAbsError", "MaxMaxAbsError"):
            assert es[k] == pytest.approx(0.25, rel=1e-8)
        assert es["MeanHellinger"] == pytest.approx(hell, rel=1e-8)
        assert es["MaxHellinger"] == pytest.approx(hell, rel=1e-8)
        assert es["MeanJSDiverge"] == pytest.approx(js, rel=1e-8)
        assert es["MaxJSDiverge"] == pytest.approx(js, rel=1e-8)

    check(v1, v2)
    v1n = [oracle.norm_marginal(v1[0]), v1[1]]
    v2n = [v2[0], oracle.norm_marginal(v2[1])]
    check(v1n, v2n)
    check([oracle.norm_marginal(x) for x in v1], [oracle.norm_marginal(x) for x in v2])


def test_error_suite_max_mean():
    """model/error_test.go:92-119 TestErrorSuiteMaxMean"""
    es = oracle.error_suite([3, 3], [[30.0, 40.0, 30.0]] * 2, [[90.0, 5.0, 5.0], [60.0, 30.0, 10.0]])
    exp = dict(MeanMeanAbsError=.30000000, MaxMeanAbsError=.39999999, MeanMaxAbsError=.45000000,
               MaxMaxAbsError=.60000000, MeanHellinger=.35109087, MaxHellinger=.46528369,
               MeanJSDiverge=.18806933, MaxJSDiverge=.29645726)
    for k, v in exp.items():
        assert es[k] == pytest.approx(v, rel=1e-7), k


def test_error_suite_fixed_semantics():
    """model/error.go:33-42, 82-84: fixed on either side scores 0 and is left out of the mean's divisor."""
    es = oracle.error_suite([2, 2], [[1, 3], [1, 1]], [[1, 1], [9, 1]], fixed1=[-1, 0])
    assert es["MeanMeanAbsError"] == pytest.approx(0.25)
    with pytest.raises(OracleError):
        oracle.error_suite([2], [[1, 1]], [[1, 1]], fixed1=[1])


def test_uai_preprocess():
    """model/uai_test.go:31-72 TestUAIPreproc"""
    P = oracle.uai_preprocess
    for pre in ("", "abc"):
        assert P("", pre) == ("", 0)
        assert P("\n\n\n", pre) == ("", 0)
        assert P("c\nc\ncnope", pre) == ("", 0)
        assert P(" abc ", pre) == ("abc", 1)
        assert P("abc\nc comment\n", pre) == ("abc", 1)
        assert P("\n\n\n\nc comment\n\n\nabc", pre) == ("abc", 1)
    assert P("hello\nworld") == ("hello\nworld", 2)
    assert P("hello\nworld\n") == ("hello\nworld", 2)
    assert P("\nhello\n\nworld\n") == ("hello\nworld", 2)
    assert P("c comment\n\nhello\nc again\nworld\nc last\n\n") == ("hello\nworld", 2)
    assert P("hello\nworld\nabc", "wor") == ("world\nabc", 2)
    assert P("\nhello\n\nworld\nabc", "wor") == ("world\nabc", 2)
    assert P("c comment\n\nhello\nc again\nworld\nabc\nc last\n\n", "wor") == ("world\nabc", 2)


def test_uai_doc():
    """model/uai_test.go:74-122 TestUAIDoc"""
    m = oracle.Model.from_buffer(PASCAL)
    m.check()
    assert m.type == "MARKOV"
    assert m.cards.tolist() == [2, 2, 3]
    assert m.n_funcs == 3
    cases = [([2], [0.436, 0.564]), ([2, 2], [0.128, 0.872, 0.920, 0.080]),
             ([2, 3], [0.210, 0.333, 0.457, 0.811, 0.000, 0.189])]
    for i, (cards, table) in enumerate(cases):
        assert [int(m.cards[v]) for v in m.func_scope(i)] == cards
        assert m.func_table(i).tolist() == table
    v, err = m.func_eval(2, [1, 2])
    assert err is None and abs(v - 0.189) < 1e-12


def test_uai_large_file(res):
    """model/uai_test.go:125-141 TestUAILargeFile"""
    m = oracle.Model.load(res("dv-rel_1.uai"))
    m.check()
    assert m.type == "MARKOV"
    assert m.n_vars == 120 and m.n_funcs == 40
    v, err = m.func_eval(39, [1, 1, 1])
    assert err is None and v == pytest.approx(2.038, rel=1e-12)


def test_uai_mar_sol_file(res):
    """model/uai_test.go:144-171 TestUAIMarSolFile"""
    m = oracle.Model.load(res("one.uai"))
    m.check()
    assert m.fixed.tolist() == [-1]
    s = oracle.solution_load(res("one.uai.MAR"))
    oracle.solution_check(s, m)
    es = oracle.error_suite(m.cards, s.marginal_list(), m.marginal_list())
    for k in ("MeanMeanAbsError", "MeanMaxAbsError", "MaxMeanAbsError", "MaxMaxAbsError"):
        assert es[k] == pytest.approx(0.25, rel=1e-8)
    hell = math.sqrt((math.sqrt(0.75) - math.sqrt(0.5)) ** 2 + (math.sqrt(0.25) - math.sqrt(0.5)) ** 2) / math.sqrt(2)
    assert es["MeanHellinger"] == pytest.approx(hell, rel=1e-8)
    assert es["MaxHellinger"] == pytest.approx(hell, rel=1e-8)


def test_uai_evidence(res):
    """model/uai_test.go:174-208 TestUAIMariEvidFile"""
    m = oracle.Model.load(res("one.uai"), use_evidence=True)
    m.check()
    assert m.fixed.tolist() == [-1]  # the default evid file holds no evidence
    with pytest.raises(OracleError):
        m.apply_evidence("2\n1 0 0\n1 0 1")
    assert m.fixed.tolist() == [-1]
    for evid, exp in (("1 0 0", 0), ("1\n1 0 1", 1)):
        m = oracle.Model.load(res("one.uai"))
        m.apply_evidence(evid)
        assert m.fixed.tolist() == [exp]
        with pytest.raises(OracleError):
            m.apply_evidence(evid)


def test_bundled_evidence_and_merlin_mar(res):
    """SURVEY §8d config 2 (8 fixed vars) and quirk 16 (merlin files carry a PR section first)."""
    m = oracle.Model.load(res("Promedus_11.uai"), use_evidence=True)
    fixed = {int(i): int(v) for i, v in enumerate(m.fixed) if v >= 0}
    assert fixed == {158: 1, 58: 1, 90: 1, 26: 1, 129: 1, 51: 1, 4: 1, 183: 1}
    p = oracle.Model.load(res("Pedigree_11.uai"), use_evidence=True)
    assert int((p.fixed >= 0).sum()) == 37
    s = oracle.solution_load(res("Grids_11.uai.merlin.MAR"))
    assert s.n_vars == 100
    assert s.marginal_list()[0] == pytest.approx([0.997878, 0.002122], abs=1e-9)
