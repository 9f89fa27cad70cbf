"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding over ``oracle/liboracle.so``, the CPU float64 restatement of the reference's
Gibbs hot path (see the headers of ``oracle/*.hpp`` for the file:line map and for what is
and is not pinned by the reference's own tests).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product (``grample_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_ABS, MEAN_ABS, HELLINGER, JS = 0, 1, 2, 3
SIMPLE, COLLAPSED, ADAPTIVE = 0, 1, 2


class OracleError(RuntimeError):
    pass


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) or f == "Makefile"]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_model_clone.restype = C.c_void_p
        _lib.orc_model_type.restype = C.c_char_p
        _lib.orc_model_func_name.restype = C.c_char_p
        _lib.orc_model_func_tabsize.restype = C.c_longlong
        _lib.orc_gen_int63.restype = C.c_longlong
        _lib.orc_gen_float64.restype = C.c_double
        _lib.orc_philox_uniform.restype = C.c_double
        _lib.orc_philox_uniform.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_int]
        _lib.orc_philox_init_value.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_int]
        _lib.orc_circ_new.restype = C.c_void_p
        _lib.orc_chain_total.restype = C.c_longlong
        _lib.orc_chain_total_seen.restype = C.c_longlong
    return _lib


def _chk(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def _ia(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _da(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Model:
    def __init__(self, handle, owned=True):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            lib().orc_model_free(self.h)
            self.h = None

    # -- constructors
    @staticmethod
    def load(path, use_evidence=False):
        out = C.c_void_p()
        _chk(lib().orc_model_load(path.encode(), int(use_evidence), C.byref(out)))
        return Model(out)

    @staticmethod
    def from_buffer(data):
        if isinstance(data, str):
            data = data.encode()
        out = C.c_void_p()
        _chk(lib().orc_model_from_buffer(data, C.c_long(len(data)), C.byref(out)))
        return Model(out)

    @staticmethod
    def create(card, fixed, scope_off, scope_vars, tab_off, tables):
        card, fixed, scope_off, scope_vars = _ia(card), _ia(fixed), _ia(scope_off), _ia(scope_vars)
        tab_off = np.ascontiguousarray(tab_off, dtype=np.int64)
        tables = _da(tables)
        out = C.c_void_p()
        _chk(lib().orc_model_create(len(card), _p(card, C.c_int), _p(fixed, C.c_int), len(scope_off) - 1,
                                    _p(scope_off, C.c_int), _p(scope_vars, C.c_int),
                                    _p(tab_off, C.c_longlong), _p(tables, C.c_double), C.byref(out)))
        return Model(out)

    @staticmethod
    def single_function(cards, table=None):
        cards = _ia(cards)
        out = C.c_void_p()
        if table is None:
            _chk(lib().orc_model_single_function(len(cards), _p(cards, C.c_int), None, C.c_longlong(0), C.byref(out)))
        else:
            t = _da(table)
            _chk(lib().orc_model_single_function(len(cards), _p(cards, C.c_int), _p(t, C.c_double),
                                                 C.c_longlong(len(t)), C.byref(out)))
        return Model(out)

    def clone(self):
        return Model(lib().orc_model_clone(self.h))

    def apply_evidence(self, data):
        if isinstance(data, str):
            data = data.encode()
        _chk(lib().orc_model_apply_evidence(self.h, data, C.c_long(len(data))))

    def check(self):
        _chk(lib().orc_model_check(self.h))

    # -- accessors
    @property
    def n_vars(self):
        return lib().orc_model_n_vars(self.h)

    @property
    def n_funcs(self):
        return lib().orc_model_n_funcs(self.h)

    @property
    def type(self):
        return lib().orc_model_type(self.h).decode()

    def _ivec(self, fn):
        out = np.zeros(self.n_vars, dtype=np.int32)
        fn(self.h, _p(out, C.c_int))
        return out

    @property
    def cards(self):
        return self._ivec(lib().orc_model_cards)

    @property
    def fixed(self):
        return self._ivec(lib().orc_model_fixed)

    @property
    def collapsed(self):
        return self._ivec(lib().orc_model_collapsed)

    def set_fixed(self, var, val):
        lib().orc_model_set_fixed(self.h, int(var), int(val))

    @property
    def marginals(self):
        out = np.zeros(lib().orc_model_marginal_size(self.h), dtype=np.float64)
        lib().orc_model_marginals(self.h, _p(out, C.c_double))
        return out

    def set_marginals(self, m):
        m = _da(m)
        lib().orc_model_set_marginals(self.h, _p(m, C.c_double))

    def marginal_list(self):
        m, out, o = self.marginals, [], 0
        for c in self.cards:
            out.append(m[o:o + c].copy())
            o += c
        return out

    def func_scope(self, f):
        out = np.zeros(lib().orc_model_func_arity(self.h, f), dtype=np.int32)
        lib().orc_model_func_scope(self.h, f, _p(out, C.c_int))
        return out

    def func_table(self, f):
        out = np.zeros(lib().orc_model_func_tabsize(self.h, f), dtype=np.float64)
        lib().orc_model_func_table(self.h, f, _p(out, C.c_double))
        return out

    def func_is_log(self, f):
        return bool(lib().orc_model_func_is_log(self.h, f))

    def func_name(self, f):
        return lib().orc_model_func_name(self.h, f).decode()

    def func_eval(self, f, values):
        v = _ia(values)
        out = C.c_double()
        rc = lib().orc_func_eval(self.h, f, _p(v, C.c_int), len(v), C.byref(out))
        return out.value, (None if rc == 0 else lib().orc_last_error().decode())

    def func_use_log_space(self, f):
        _chk(lib().orc_func_use_log_space(self.h, f))

    def func_add_value(self, f, values, inc):
        v = _ia(values)
        _chk(lib().orc_func_add_value(self.h, f, _p(v, C.c_int), len(v), C.c_double(inc)))

    def func_check(self, f):
        _chk(lib().orc_func_check(self.h, f))

    def flatten(self):
        """(card, fixed, scope_off, scope_vars, tab_off, tables) — the C-ABI model arrays."""
        scope_off, scope_vars, tab_off, tabs = [0], [], [0], []
        for f in range(self.n_funcs):
            s = self.func_scope(f)
            scope_vars.extend(int(x) for x in s)
            scope_off.append(len(scope_vars))
            t = self.func_table(f)
            tabs.append(t)
            tab_off.append(tab_off[-1] + len(t))
        return (self.cards.copy(), self.fixed.copy(), np.asarray(scope_off, np.int32),
                np.asarray(scope_vars, np.int32), np.asarray(tab_off, np.int64),
                np.concatenate(tabs) if tabs else np.zeros(0))


def solution_load(path):
    out = C.c_void_p()
    _chk(lib().orc_solution_load(path.encode(), C.byref(out)))
    return Model(out)


def solution_from_buffer(data):
    if isinstance(data, str):
        data = data.encode()
    out = C.c_void_p()
    _chk(lib().orc_solution_from_buffer(data, C.c_long(len(data)), C.byref(out)))
    return Model(out)


def solution_check(sol, model):
    _chk(lib().orc_solution_check(sol.h, model.h))


def uai_preprocess(data, prefix=""):
    if isinstance(data, str):
        data = data.encode()
    buf = C.create_string_buffer(len(data) + 16)
    n = C.c_int()
    _chk(lib().orc_uai_preprocess(data, C.c_long(len(data)), prefix.encode(), buf, C.c_long(len(buf)), C.byref(n)))
    return buf.value.decode(), n.value


def variter_enumerate(cards, fixed, honor, max_rows=4096):
    cards, fixed = _ia(cards), _ia(fixed)
    n = len(cards)
    out = np.zeros((max_rows, n), dtype=np.int32)
    rows = C.c_int()
    final = np.zeros(n, dtype=np.int32)
    _chk(lib().orc_variter_enumerate(n, _p(cards, C.c_int), _p(fixed, C.c_int), int(honor), _p(out, C.c_int),
                                     max_rows, C.byref(rows), _p(final, C.c_int)))
    return out[:rows.value].copy(), final


def error_suite(cards, marg1, marg2, fixed1=None, fixed2=None):
    """Returns dict of the 8 ErrorSuite numbers (model/error.go:15-25)."""
    cards = _ia(cards)
    n = len(cards)
    f1 = _ia(np.full(n, -1) if fixed1 is None else fixed1)
    f2 = _ia(np.full(n, -1) if fixed2 is None else fixed2)
    m1, m2 = _da(np.concatenate([np.asarray(x, float) for x in marg1])), _da(np.concatenate([np.asarray(x, float) for x in marg2]))
    out = np.zeros(8)
    _chk(lib().orc_error_suite(n, _p(cards, C.c_int), _p(f1, C.c_int), _p(m1, C.c_double), _p(f2, C.c_int),
                               _p(m2, C.c_double), _p(out, C.c_double)))
    keys = ["MeanMeanAbsError", "MaxMeanAbsError", "MeanMaxAbsError", "MaxMaxAbsError",
            "MeanHellinger", "MaxHellinger", "MeanJSDiverge", "MaxJSDiverge"]
    return dict(zip(keys, out.tolist()))


def measure(which, m1, m2, fixed1=-1, fixed2=-1):
    m1, m2 = _da(m1), _da(m2)
    out = C.c_double()
    _chk(lib().orc_measure(which, len(m1), fixed1, _p(m1, C.c_double), fixed2, _p(m2, C.c_double), C.byref(out)))
    return out.value


def norm_marginal(m):
    m = _da(m).copy()
    _chk(lib().orc_norm_marginal(len(m), _p(m, C.c_double)))
    return m


class Generator:
    def __init__(self, seed):
        s = np.ascontiguousarray(np.atleast_1d(np.asarray(seed, dtype=np.uint64)))
        self.h = C.c_void_p()
        rc = lib().orc_gen_new(_p(s, C.c_ulonglong), len(s), C.byref(self.h))
        if rc != 0:
            self.h = None
            _chk(rc)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gen_free(self.h)
            self.h = None

    def int63(self):
        return lib().orc_gen_int63(self.h)

    def int31n(self, n):
        return lib().orc_gen_int31n(self.h, n)

    def float64(self):
        return lib().orc_gen_float64(self.h)

    def uni_sample(self, card):
        out = C.c_int()
        _chk(lib().orc_uni_sample(self.h, C.c_longlong(card), C.byref(out)))
        return out.value

    def weighted_sample(self, card, weights):
        w = _da(weights)
        out = C.c_int()
        _chk(lib().orc_weighted_sample(self.h, C.c_longlong(card), _p(w, C.c_double), len(w), C.byref(out)))
        return out.value

    def var_sample(self, fixed, collapsed, exclude_collapsed):
        f, c = _ia(fixed), _ia(collapsed)
        out = C.c_int()
        _chk(lib().orc_var_sample(self.h, len(f), _p(f, C.c_int), _p(c, C.c_int), int(exclude_collapsed), C.byref(out)))
        return out.value


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c, C.c_uint), _p(k, C.c_uint), _p(out, C.c_uint))
    return out


def philox_uniform(seed, chain, sweep, var, bits):
    return lib().orc_philox_uniform(seed, chain, sweep, var, bits)


def philox_init_value(seed, chain, var, card):
    return lib().orc_philox_init_value(seed, chain, var, card)


class CircularInt:
    def __init__(self, size):
        self.h = C.c_void_p(lib().orc_circ_new(size))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_circ_free(self.h)
            self.h = None

    def add(self, v):
        lib().orc_circ_add(self.h, v)

    @property
    def buf_size(self):
        return lib().orc_circ_bufsize(self.h)

    @property
    def count(self):
        return lib().orc_circ_count(self.h)

    def _half(self, second):
        out = np.zeros(max(self.buf_size, 1), dtype=np.int32)
        n = lib().orc_circ_half(self.h, second, _p(out, C.c_int))
        return None if n < 0 else out[:n].tolist()

    def first_half(self):
        return self._half(0)

    def second_half(self):
        return self._half(1)


class Sampler:
    """GibbsSimple / GibbsCollapsed over a model (which it mutates, like the reference)."""

    def __init__(self, gen, model, collapsed=False, lean=False):
        self.gen, self.model = gen, model  # keep alive
        self.h = C.c_void_p()
        fn = lib().orc_gibbs_collapsed_new if collapsed else lib().orc_gibbs_simple_new
        _chk(fn(gen.h, model.h, int(lean), C.byref(self.h)))
        self.is_collapsed = collapsed

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sampler_free(self.h)
            self.h = None

    def sample(self, state):
        s = _ia(state).copy()
        idx = C.c_int()
        _chk(lib().orc_sampler_sample(self.h, _p(s, C.c_int), len(s), C.byref(idx)))
        return idx.value, s

    def conditional(self, var, state):
        """floored un-normalised weights e[k] of gibbs-simple.go:171-258"""
        s = _ia(state)
        out = np.zeros(int(self.model.cards[var]))
        _chk(lib().orc_sampler_conditional(self.h, int(var), _p(s, C.c_int), _p(out, C.c_double)))
        return out

    @property
    def state(self):
        out = np.zeros(self.model.n_vars, dtype=np.int32)
        lib().orc_sampler_get_state(self.h, _p(out, C.c_int))
        return out

    @state.setter
    def state(self, s):
        s = _ia(s)
        lib().orc_sampler_set_state(self.h, _p(s, C.c_int))

    def collapse(self, var_idx):
        v = C.c_int()
        marg = np.zeros(64)
        _chk(lib().orc_collapsed_collapse(self.h, int(var_idx), C.byref(v), _p(marg, C.c_double)))
        return v.value, marg[:int(self.model.cards[v.value])].copy()

    def blanket_size(self, var):
        return lib().orc_collapsed_blanket_size(self.h, int(var))

    def function_count(self, var):
        return lib().orc_collapsed_function_count(self.h, int(var))

    def neighbors(self, var):
        out = np.zeros(self.model.n_vars, dtype=np.int32)
        n = lib().orc_collapsed_neighbors(self.h, int(var), _p(out, C.c_int))
        return out[:n].tolist()

    def sweep_run(self, order, seed, chain0, states, sweep0, n_sweeps, bits=53, record=True, counts=None, var_bits=None):
        """Device-schedule sweeps (oracle/sweep.hpp).  states: [n_chains, n_vars] int32 (updated copy returned).
        var_bits: optional per-variable draw widths (32 / 53) — the device's hybrid mode."""
        order = _ia(order)
        st = _ia(states).copy()
        n_chains = st.shape[0]
        if counts is None:
            counts = np.zeros(int(self.model.cards.sum()))
        counts = _da(counts).copy()
        if var_bits is not None:
            vb = _ia(var_bits)
            _chk(lib().orc_sweep_run_mixed(self.h, _p(order, C.c_int), len(order), C.c_ulonglong(seed), C.c_uint(chain0),
                                           n_chains, C.c_uint(sweep0), C.c_uint(n_sweeps), _p(vb, C.c_int), int(record),
                                           _p(st, C.c_int), _p(counts, C.c_double)))
            return st, counts
        _chk(lib().orc_sweep_run(self.h, _p(order, C.c_int), len(order), C.c_ulonglong(seed), C.c_uint(chain0),
                                 n_chains, C.c_uint(sweep0), C.c_uint(n_sweeps), bits, int(record),
                                 _p(st, C.c_int), _p(counts, C.c_double)))
        return st, counts


def _scan_run(self, order, seed, chain0, states, step0, n_steps, record=True, counts=None):
    """Device random-scan parity mode (oracle/sweep.hpp::scan_chain)."""
    order = _ia(order)
    st = _ia(states).copy()
    if counts is None:
        counts = np.zeros(int(self.model.cards.sum()))
    counts = _da(counts).copy()
    _chk(lib().orc_scan_run(self.h, _p(order, C.c_int), len(order), C.c_ulonglong(seed), C.c_uint(chain0), st.shape[0],
                            C.c_ulonglong(step0), C.c_longlong(n_steps), int(record), _p(st, C.c_int), _p(counts, C.c_double)))
    return st, counts


Sampler.scan_run = _scan_run


class Chain:
    def __init__(self, model=None, sampler=None, cw=0, burn_in=0, _handle=None):
        self.model, self.sampler = model, sampler
        if _handle is not None:
            self.h = _handle
            return
        self.h = C.c_void_p()
        _chk(lib().orc_chain_new(model.h, sampler.h if sampler else None, int(cw), C.c_longlong(burn_in), C.byref(self.h)))

    @staticmethod
    def from_marginals(cards, marginals, collapsed=None, cw=0):
        cards = _ia(cards)
        m = _da(np.concatenate([np.asarray(x, float) for x in marginals]))
        col = _ia(np.zeros(len(cards)) if collapsed is None else collapsed)
        h = C.c_void_p()
        _chk(lib().orc_chain_from_marginals(len(cards), _p(cards, C.c_int), _p(m, C.c_double), _p(col, C.c_int), cw, C.byref(h)))
        ch = Chain(_handle=h)
        ch._cards = cards
        return ch

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_chain_free(self.h)
            self.h = None

    def advance(self):
        _chk(lib().orc_chain_advance(self.h))

    def one_sample(self, update=True):
        _chk(lib().orc_chain_one_sample(self.h, int(update)))

    @property
    def total_sample_count(self):
        return lib().orc_chain_total(self.h)

    def total_seen(self, var):
        return lib().orc_chain_total_seen(self.h, int(var))

    def set_history(self, var, samples):
        s = _ia(samples)
        _chk(lib().orc_chain_set_history(self.h, int(var), _p(s, C.c_int), len(s)))

    def chain_dist(self, which, var, merged, merged_fixed=-1, merged_collapsed=False):
        m = _da(merged)
        w, b = C.c_double(), C.c_double()
        _chk(lib().orc_chain_dist(self.h, which, int(var), len(m), _p(m, C.c_double), merged_fixed,
                                  int(merged_collapsed), C.byref(w), C.byref(b)))
        return w.value, b.value


def _harr(chains):
    arr = (C.c_void_p * len(chains))()
    for i, c in enumerate(chains):
        arr[i] = c.h.value if isinstance(c.h, C.c_void_p) else c.h
    return arr


def merge_chains(chains, total_card, n_vars):
    marg = np.zeros(total_card)
    col = np.zeros(n_vars, dtype=np.int32)
    _chk(lib().orc_merge_chains(_harr(chains), len(chains), _p(marg, C.c_double), _p(col, C.c_int)))
    return marg, col


def chain_convergence(chains, which, n_vars):
    out = np.zeros(n_vars)
    _chk(lib().orc_chain_convergence(_harr(chains), len(chains), which, _p(out, C.c_double)))
    return out


class ConvergenceSampler:
    def __init__(self, gen, model, measure_id=-1):
        self.gen, self.model = gen, model
        self.h = C.c_void_p()
        _chk(lib().orc_adapt_new(gen.h, model.h, measure_id, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_adapt_free(self.h)
            self.h = None

    def adapt(self, chains, new_chain_count):
        cap = len(chains) + new_chain_count + 8
        out = (C.c_void_p * cap)()
        n_out, n_t = C.c_int(), C.c_int()
        targets = np.zeros(cap, dtype=np.int32)
        _chk(lib().orc_adapt_adapt(self.h, _harr(chains), len(chains), new_chain_count, out, cap, C.byref(n_out),
                                   _p(targets, C.c_int), C.byref(n_t)))
        res = list(chains)
        for i in range(len(chains), n_out.value):
            res.append(Chain(_handle=C.c_void_p(out[i])))
        return res, targets[:n_t.value].tolist()


def run(model, solution=None, kind=SIMPLE, n_chains=4, burn_in=-1, cw=0, max_iters=-1, seed=1, lean=False,
        n_threads=0, chain_adds=1, adapt_rounds=-1, max_rounds=0, curve_cap=4096):
    """cmd/root.go main loop restated (oracle/run.hpp).  Returns dict."""
    tot = int(model.cards.sum())
    merged = np.zeros(tot)
    col = np.zeros(model.n_vars, dtype=np.int32)
    info = np.zeros(3, dtype=np.int64)
    secs = np.zeros(2)
    cs = np.zeros(curve_cap, dtype=np.int64)
    c1, c2, c3 = np.zeros(curve_cap), np.zeros(curve_cap), np.zeros(curve_cap)
    cn = C.c_int()
    _chk(lib().orc_run(model.h, solution.h if solution else None, kind, n_chains, C.c_longlong(burn_in),
                       C.c_longlong(cw), C.c_longlong(max_iters), C.c_longlong(seed), int(lean), n_threads,
                       chain_adds, adapt_rounds, max_rounds, _p(merged, C.c_double), _p(col, C.c_int),
                       _p(info, C.c_longlong), _p(secs, C.c_double), curve_cap, _p(cs, C.c_longlong),
                       _p(c1, C.c_double), _p(c2, C.c_double), _p(c3, C.c_double), C.byref(cn)))
    n = cn.value
    return dict(merged=merged, collapsed=col, samples=int(info[0]), rounds=int(info[1]), n_chains=int(info[2]),
                burnin_seconds=float(secs[0]), advance_seconds=float(secs[1]),
                curve=dict(samples=cs[:n].copy(), mean_hellinger=c1[:n].copy(), max_hellinger=c2[:n].copy(),
                           mean_abs=c3[:n].copy()))


def throughput(model, kind=SIMPLE, n_threads=1, steps=100000, seed=1, lean=True, cw=2):
    secs = C.c_double()
    ups = C.c_longlong()
    _chk(lib().orc_throughput(model.h, kind, n_threads, C.c_longlong(steps), C.c_longlong(seed), int(lean), cw,
                              C.byref(secs), C.byref(ups)))
    return ups.value, secs.value
