"""Thin object layer over the C ABI: `Model` (a flattened factor graph, possibly with collapsed
variables) and `Chains` (all chains of one device).  Every method is one C-ABI call; the
reference function each call replaces is cited in include/grample_b200.h.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, HELLINGER, TABLE, GrampleError, check, lib

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def device_count():
    n = C.c_int()
    check(lib().gb_device_count(C.byref(n)))
    return n.value


class Model:
    """Flattened model on a device (device=-1: host-only metadata, no compute)."""

    def __init__(self, handle, owned=True):
        self.h = handle
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and getattr(self, "h", None):
            lib().gb_model_destroy(self.h)
            self.h = None

    @staticmethod
    def from_arrays(card, fixed, scope_off, scope_vars, tab_off, tables, device=0):
        card, fixed, scope_off, scope_vars = _i32(card), _i32(fixed), _i32(scope_off), _i32(scope_vars)
        tab_off = np.ascontiguousarray(tab_off, dtype=np.int64)
        tables = _f64(tables)
        out = C.c_void_p()
        check(lib().gb_model_create(len(card), _ptr(card, _i32p), _ptr(fixed, _i32p), len(scope_off) - 1,
                                    _ptr(scope_off, _i32p), _ptr(scope_vars, _i32p), _ptr(tab_off, _i64p),
                                    _ptr(tables, _f64p), device, C.byref(out)))
        return Model(out)

    @staticmethod
    def from_uai(path, use_evidence=False, device=0):
        """model.NewModelFromFile(reader, filename, useEvidence) — evidence file is `path + '.evid'`."""
        out = C.c_void_p()
        evid = (path + ".evid").encode() if use_evidence else None
        check(lib().gb_model_load_uai(path.encode(), evid, device, C.byref(out)))
        return Model(out)

    def _int(self, fn, *args):
        v = C.c_int32()
        check(fn(self.h, *args, C.byref(v)))
        return v.value

    @property
    def n_vars(self):
        return self._int(lib().gb_model_n_vars)

    @property
    def n_funcs(self):
        return self._int(lib().gb_model_n_funcs)

    @property
    def total_card(self):
        return self._int(lib().gb_model_total_card)

    def _ivec(self, fn):
        out = np.zeros(self.n_vars, dtype=np.int32)
        check(fn(self.h, _ptr(out, _i32p)))
        return out

    @property
    def cards(self):
        return self._ivec(lib().gb_model_cards)

    @property
    def fixed(self):
        return self._ivec(lib().gb_model_fixed)

    @property
    def collapsed(self):
        return self._ivec(lib().gb_model_collapsed)

    def func_scope(self, f):
        out = np.zeros(self._int(lib().gb_model_func_arity, f), dtype=np.int32)
        check(lib().gb_model_func_scope(self.h, f, _ptr(out, _i32p)))
        return out

    def func_log_table(self, f):
        n = C.c_int64()
        check(lib().gb_model_func_table_size(self.h, f, C.byref(n)))
        out = np.zeros(n.value)
        check(lib().gb_model_func_log_table(self.h, f, _ptr(out, _f64p)))
        return out

    def blanket_size(self, var):
        return self._int(lib().gb_model_blanket_size, int(var))

    def function_count(self, var):
        return self._int(lib().gb_model_function_count, int(var))

    def schedule(self):
        """(order, colour_off): colour-sorted sweep order of the free, un-collapsed variables."""
        n, nc = C.c_int32(), C.c_int32()
        check(lib().gb_model_schedule(self.h, C.byref(n), C.byref(nc), None, None))
        order = np.zeros(n.value, dtype=np.int32)
        coff = np.zeros(nc.value + 1, dtype=np.int32)
        check(lib().gb_model_schedule(self.h, None, None, _ptr(order, _i32p), _ptr(coff, _i32p)))
        return order, coff

    def table_mode(self):
        """(applies?, number of tabulated configurations) for precision=TABLE"""
        ok, n = C.c_int32(), C.c_int64()
        check(lib().gb_model_table_mode(self.h, C.byref(ok), C.byref(n)))
        return bool(ok.value), n.value

    def bits_mode(self):
        """whether precision=TABLE_BITS (bit-packed state, bit-sliced sweep) applies"""
        ok = C.c_int32()
        check(lib().gb_model_bits_mode(self.h, C.byref(ok)))
        return bool(ok.value)

    def hybrid_mask(self):
        """int32 [n_vars]: 1 where precision=HYBRID samples the variable from a threshold table"""
        out = np.zeros(self.n_vars, dtype=np.int32)
        check(lib().gb_model_hybrid_mask(self.h, _ptr(out, _i32p)))
        return out

    def thresholds(self, var):
        n = C.c_int32()
        check(lib().gb_model_thresholds(self.h, int(var), C.byref(n), None))
        out = np.zeros(n.value, dtype=np.uint32)
        check(lib().gb_model_thresholds(self.h, int(var), None, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def collapse(self, var=-1, seed=0):
        """(*GibbsCollapsed).Collapse as a pure function: (new Model, collapsed var, its local marginal)."""
        v = C.c_int32()
        marg = np.zeros(_lib.MAX_CARD)
        out = C.c_void_p()
        check(lib().gb_model_collapse(self.h, int(var), C.c_uint64(seed), C.byref(v), _ptr(marg, _f64p), C.byref(out)))
        m = Model(out)
        return m, v.value, marg[:int(m.cards[v.value])].copy()

    def conditional(self, states, variables, precision=F64):
        """K5: floored un-normalised weights e[k] (gibbs-simple.go:171-258) per (state, var)."""
        states = _i32(np.atleast_2d(states))
        variables = _i32(np.atleast_1d(variables))
        n = states.shape[0]
        out = np.zeros((n, 64))
        check(lib().gb_conditional(self.h, precision, n, _ptr(states, _i32p), _ptr(variables, _i32p), _ptr(out, _f64p)))
        cards = self.cards
        return [out[i, :cards[variables[i]]].copy() for i in range(n)]


def _model_sample(self, state, step, var=-1, seed=1, precision=F64, exclude_collapsed=False):
    """(*GibbsSimple).Sample / SampleVar on one caller-held state: updates `state` (int32 [n_vars]) in place and returns
    the sampled variable's index."""
    assert state.dtype == np.int32 and state.flags["C_CONTIGUOUS"]
    v = C.c_int32()
    check(lib().gb_model_sample(self.h, precision, int(var), int(exclude_collapsed), C.c_uint64(seed), C.c_uint64(step),
                                _ptr(state, _i32p), C.byref(v)))
    return v.value


Model.sample = _model_sample


class Chains:
    """All chains of one device, grouped by model (one group per collapsed variant)."""

    def __init__(self, models, chains_per_model, seed=1, first_chain_id=0, precision=F64, history=False, device=0,
                 per_colour=False, rao_blackwell=False):
        if isinstance(models, Model):
            models, chains_per_model = [models], [chains_per_model]
        self.models = list(models)
        arr = (C.c_void_p * len(models))(*[m.h.value if isinstance(m.h, C.c_void_p) else m.h for m in models])
        cpm = _i32(chains_per_model)
        self.h = C.c_void_p()
        flags = ((_lib.CHAINS_HISTORY if history else 0) | (_lib.CHAINS_PER_COLOUR if per_colour else 0)
                 | (_lib.CHAINS_RAO_BLACKWELL if rao_blackwell else 0))
        check(lib().gb_chains_create(len(models), arr, _ptr(cpm, _i32p), C.c_uint64(seed), C.c_uint64(first_chain_id),
                                     precision, flags, device, C.byref(self.h)))
        self.base = self.models[0]
        self.device = device

    def __del__(self):
        if getattr(self, "h", None):
            lib().gb_chains_destroy(self.h)
            self.h = None

    def add_group(self, model, n_chains, first_chain_id):
        check(lib().gb_chains_add_group(self.h, model.h, n_chains, C.c_uint64(first_chain_id)))
        self.models.append(model)

    @property
    def n_groups(self):
        v = C.c_int32()
        check(lib().gb_chains_n_groups(self.h, C.byref(v)))
        return v.value

    @property
    def n_chains(self):
        v = C.c_int64()
        check(lib().gb_chains_n_chains(self.h, C.byref(v)))
        return v.value

    def sweep(self, n_sweeps, record=True):
        check(lib().gb_chains_sweep(self.h, int(n_sweeps), int(record)))

    def sweep_timed(self, n_sweeps, record=True):
        """sweeps bracketed by CUDA events on the library's stream; returns device milliseconds"""
        ms = C.c_float()
        check(lib().gb_chains_sweep_timed(self.h, int(n_sweeps), int(record), C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        v = C.c_int64()
        check(lib().gb_chains_launch_count(self.h, C.byref(v)))
        return v.value

    def scan(self, n_steps, record=True):
        """reference schedule: n_steps random-scan single-variable updates per chain"""
        check(lib().gb_chains_scan(self.h, int(n_steps), int(record)))

    def burnin(self, n_sweeps):
        check(lib().gb_chains_burnin(self.h, int(n_sweeps)))

    def advance(self, cw):
        check(lib().gb_chains_advance(self.h, int(cw)))

    def group_sweep(self, group, n_sweeps, record=True):
        check(lib().gb_chains_group_sweep(self.h, group, int(n_sweeps), int(record)))

    def group_advance(self, group, cw):
        check(lib().gb_chains_group_advance(self.h, group, int(cw)))

    def group_info(self, group):
        """(chains in the group, the group's TotalSampleCount)"""
        n, t = C.c_int32(), C.c_int64()
        check(lib().gb_chains_group_info(self.h, group, C.byref(n), C.byref(t), None))
        return n.value, t.value

    def synchronize(self):
        check(lib().gb_chains_synchronize(self.h))

    @property
    def total_samples(self):
        v = C.c_int64()
        check(lib().gb_chains_total_samples(self.h, C.byref(v)))
        return v.value

    def _merge_buffers(self, out):
        if out is not None:  # caller-owned (float64 [sum(card)], int32 [n_vars]) buffers, reused across intervals
            return out
        return np.empty(self.base.total_card), np.empty(self.base.n_vars, dtype=np.int32)

    def merged_marginals(self, out=None):
        out, col = self._merge_buffers(out)
        check(lib().gb_chains_merged_marginals(self.h, _ptr(out, _f64p), _ptr(col, _i32p)))
        return out, col

    def merge_begin(self, out=None):
        """MergeChains off the sweep stream: snapshot now, reduce (NCCL when a communicator is attached) and copy to the
        host on a side stream while later sweeps run.  `out` must stay alive until the matching merge_end(); up to two
        merges may be in flight; out="none" asks for no host copy on this rank."""
        if isinstance(out, str):  # "none": take part in the reduction, no host copy of the marginals on this rank
            check(lib().gb_chains_merge_begin(self.h, None, None))
            bufs = (None, None)
        else:
            bufs = self._merge_buffers(out)
            check(lib().gb_chains_merge_begin(self.h, _ptr(bufs[0], _f64p), _ptr(bufs[1], _i32p)))
        if not hasattr(self, "_pending_q"):
            self._pending_q = []
        self._pending_q.append(bufs)  # (only now: a refused call must not drop the buffers a pending merge still writes to)

    def merge_end(self):
        """-> (merged, collapsed flags, chains over all ranks, TotalSampleCount over all ranks)"""
        n, t = C.c_int64(), C.c_int64()
        check(lib().gb_chains_merge_end(self.h, C.byref(n), C.byref(t)))
        out, col = self._pending_q.pop(0)  # merges complete in the order they began
        return out, col, n.value, t.value

    def merge_timing(self):
        """device milliseconds of the last merge: [count sums, NCCL sum, conversion to marginals, copy to the host]"""
        ms = (C.c_float * 4)()
        check(lib().gb_chains_merge_timing(self.h, ms))
        return [round(float(x), 4) for x in ms]

    def global_totals(self):
        n, t = C.c_int64(), C.c_int64()
        check(lib().gb_chains_global_totals(self.h, C.byref(n), C.byref(t)))
        return n.value, t.value

    def attach_comm(self, comm):
        """in-library NCCL: merged_marginals / merge_begin / convergence / adapt become collective over the ranks"""
        check(lib().gb_chains_attach_comm(self.h, comm.h if comm is not None else None))
        self.comm = comm

    def merge_partial_dev(self):
        """(device pointer, n doubles) of this device's MergeChains contribution, for an all-reduce."""
        p, n = C.c_void_p(), C.c_int64()
        check(lib().gb_chains_merge_partial_dev(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def merge_finalize(self, out=None):
        out, col = self._merge_buffers(out)
        check(lib().gb_chains_merge_finalize(self.h, _ptr(out, _f64p), _ptr(col, _i32p)))
        return out, col

    def convergence(self, measure=HELLINGER, merged=None):
        out = np.zeros(self.base.n_vars)
        mp = None
        if merged is not None:
            merged = _f64(merged)
            mp = _ptr(merged, _f64p)
        check(lib().gb_chains_convergence(self.h, measure, mp, _ptr(out, _f64p)))
        return out

    def convergence_partial_dev(self, measure, merged):
        merged = _f64(merged)
        p, n = C.c_void_p(), C.c_int64()
        check(lib().gb_chains_convergence_partial_dev(self.h, measure, _ptr(merged, _f64p), C.byref(p), C.byref(n)))
        return p.value, n.value

    def convergence_finalize(self, wb, cw, total_chains, collapsed):
        return convergence_finalize(self.base, wb, cw, total_chains, collapsed)

    def adapt(self, base_model, new_chain_count, chains_per_new_model, cw, first_chain_id, measure=HELLINGER,
              max_groups=128):
        chosen = np.zeros(max(new_chain_count, 1), dtype=np.int32)
        n = C.c_int32()
        check(lib().gb_chains_adapt(self.h, base_model.h, new_chain_count, chains_per_new_model, measure, int(cw),
                                    max_groups, C.c_uint64(first_chain_id), _ptr(chosen, _i32p), C.byref(n)))
        return chosen[:n.value].tolist()

    def adapt_scores(self, base_model, new_chain_count, chains_per_new_model, scores, total_chains, first_chain_id,
                     id_stride=0, max_groups=128):
        """Adapt with caller-supplied (globally reduced) convergence scores — the multi-GPU form"""
        chosen = np.zeros(max(new_chain_count, 1), dtype=np.int32)
        n = C.c_int32()
        sc = _f64(scores)
        check(lib().gb_chains_adapt_scores(self.h, base_model.h, new_chain_count, chains_per_new_model, _ptr(sc, _f64p),
                                           int(total_chains), max_groups, C.c_uint64(first_chain_id), C.c_uint64(id_stride),
                                           _ptr(chosen, _i32p), C.byref(n)))
        return chosen[:n.value].tolist()

    def _group_chains(self, group):
        # groups keep their chain counts on the C side; recover via state size
        raise NotImplementedError

    def get_state(self, group, n_chains):
        out = np.zeros((n_chains, self.base.n_vars), dtype=np.int32)
        check(lib().gb_chains_get_state(self.h, group, _ptr(out, _i32p)))
        return out

    def set_state(self, group, states):
        st = _i32(states)
        check(lib().gb_chains_set_state(self.h, group, _ptr(st, _i32p)))

    def group_counts(self, group):
        out = np.zeros(self.base.total_card, dtype=np.uint64)
        check(lib().gb_chains_group_counts(self.h, group, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def group_history(self, group, n_chains):
        out = np.zeros((2, self.base.total_card, n_chains), dtype=np.uint16)
        check(lib().gb_chains_group_history(self.h, group, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out


class Comm:
    """One rank of an NCCL communicator owned by the library (gb_comm_*)."""

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().gb_comm_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def init_rank(unique_id, world, rank, device):
        out = C.c_void_p()
        buf = (C.c_uint8 * 128)(*unique_id) if unique_id is not None else None
        check(lib().gb_comm_init_rank(buf, world, rank, device, C.byref(out)))
        return Comm(out)

    @property
    def info(self):
        w, r, d = C.c_int32(), C.c_int32(), C.c_int()
        check(lib().gb_comm_info(self.h, C.byref(w), C.byref(r), C.byref(d)))
        return w.value, r.value, d.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().gb_comm_destroy(self.h)
            self.h = None


class Fleet:
    """The chain handles of one process's devices joined by a single-process communicator (gb_fleet_*)."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p()
        check(lib().gb_fleet_create(len(devices), devs, C.byref(self.h)))
        self.devices = list(devices)
        self.chains = [None] * len(devices)

    def __del__(self):
        if getattr(self, "h", None):
            self.chains = None  # (handles are destroyed by their owners; the fleet only borrows them)
            lib().gb_fleet_destroy(self.h)
            self.h = None

    def attach(self, slot, chains):
        check(lib().gb_fleet_attach(self.h, slot, chains.h))
        self.chains[slot] = chains

    def sweep(self, n_sweeps, record=True):
        check(lib().gb_fleet_sweep(self.h, int(n_sweeps), int(record)))

    def advance(self, cw):
        check(lib().gb_fleet_advance(self.h, int(cw)))

    def synchronize(self):
        check(lib().gb_fleet_synchronize(self.h))

    def merged_marginals(self, out=None):
        out, col = self.chains[0]._merge_buffers(out)
        check(lib().gb_fleet_merged_marginals(self.h, _ptr(out, _f64p), _ptr(col, _i32p)))
        return out, col

    def merge_begin(self, out=None):
        bufs = self.chains[0]._merge_buffers(out)
        check(lib().gb_fleet_merge_begin(self.h, _ptr(bufs[0], _f64p), _ptr(bufs[1], _i32p)))
        self._pending = bufs

    def merge_end(self):
        n, t = C.c_int64(), C.c_int64()
        check(lib().gb_fleet_merge_end(self.h, C.byref(n), C.byref(t)))
        out, col = self._pending
        self._pending = None
        return out, col, n.value, t.value

    def convergence(self, measure=HELLINGER, merged=None):
        out = np.zeros(self.chains[0].base.n_vars)
        mp = None
        if merged is not None:
            merged = _f64(merged)
            mp = _ptr(merged, _f64p)
        check(lib().gb_fleet_convergence(self.h, measure, mp, _ptr(out, _f64p)))
        return out

    def adapt(self, base_models, new_chain_count, chains_per_new_model, cw, first_chain_id, measure=HELLINGER, max_groups=128):
        chosen = np.zeros(max(new_chain_count, 1), dtype=np.int32)
        n = C.c_int32()
        arr = (C.c_void_p * len(base_models))(*[m.h.value if isinstance(m.h, C.c_void_p) else m.h for m in base_models])
        check(lib().gb_fleet_adapt(self.h, arr, new_chain_count, chains_per_new_model, measure, int(cw), max_groups,
                                   C.c_uint64(first_chain_id), _ptr(chosen, _i32p), C.byref(n)))
        return chosen[:n.value].tolist()


def convergence_finalize(base_model, wb, cw, total_chains, collapsed):
    """chain.go:46-59, 69-88 applied to all-reduced per-variable sums of within/between distances"""
    wb, collapsed = _f64(wb), _i32(collapsed)
    out = np.zeros(base_model.n_vars)
    check(lib().gb_convergence_finalize(base_model.h, _ptr(wb, _f64p), int(cw), int(total_chains),
                                        _ptr(collapsed, _i32p), _ptr(out, _f64p)))
    return out


def error_suite(cards, marg1, marg2, fixed1=None, fixed2=None):
    """model.NewErrorSuite — dict with the 8 ErrorSuite fields."""
    cards = _i32(cards)
    m1, m2 = _f64(marg1), _f64(marg2)
    f1 = _i32(fixed1) if fixed1 is not None else None
    f2 = _i32(fixed2) if fixed2 is not None else None
    out = np.zeros(8)
    check(lib().gb_error_suite(len(cards), _ptr(cards, _i32p), _ptr(f1, _i32p) if f1 is not None else None,
                               _ptr(m1, _f64p), _ptr(f2, _i32p) if f2 is not None else None, _ptr(m2, _f64p),
                               _ptr(out, _f64p)))
    keys = ["MeanMeanAbsError", "MaxMeanAbsError", "MeanMaxAbsError", "MaxMaxAbsError",
            "MeanHellinger", "MaxHellinger", "MeanJSDiverge", "MaxJSDiverge"]
    return dict(zip(keys, out.tolist()))


def mar_load(path):
    """UAIReader.ReadMargSolution: (cards, flat marginals)."""
    nv, tc = C.c_int32(), C.c_int32()
    check(lib().gb_mar_load(path.encode(), C.byref(nv), C.byref(tc), None, None))
    cards = np.zeros(nv.value, dtype=np.int32)
    marg = np.zeros(tc.value)
    check(lib().gb_mar_load(path.encode(), None, None, _ptr(cards, _i32p), _ptr(marg, _f64p)))
    return cards, marg
