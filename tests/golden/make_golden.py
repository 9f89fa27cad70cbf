#!/usr/bin/env python
"""Generates the committed golden vectors for the parts of the hot path the reference's own tests do
not pin (SampleVar numerics, the sweep schedule, collapse, convergence):

  tests/golden/conditionals.json   (model, state, var) -> floored un-normalised weights e[k]
                                   (gibbs-simple.go:171-258), float64, exact round trip through repr
  tests/golden/trajectories.json   (model, colour schedule, seed, initial states) -> states and marginal
                                   counts after n sweeps of the device schedule, for 53-bit (f64 kernels)
                                   and 32-bit (table kernels) draws, plus the Rao-Blackwell bins of the 53-bit run

Source of truth: the CPU oracle (oracle/, pinned by the reference's known-answer tests).  Run from the
repo root:  python tests/golden/make_golden.py      (no GPU needed; deterministic)
The tests compare BOTH the oracle (test_oracle_golden.py, CPU) and the CUDA path (test_gpu_parity.py)
with these files, so an accidental change of either side's arithmetic shows up without the other.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

MODELS = [("sample.uai", False), ("deterministic.uai", False), ("Grids_11.uai", False), ("Promedus_11.uai", True),
          ("Pedigree_11.uai", True), ("ObjectDetection_11.uai", False)]


def random_states(rng, cards, fixed, n):
    st = np.zeros((n, len(cards)), dtype=np.int32)
    for v, c in enumerate(cards):
        st[:, v] = fixed[v] if fixed[v] >= 0 else rng.integers(c, size=n)
    return st


def greedy_schedule(model):
    """colour-sorted sweep schedule (the rule HostModel::build_colouring implements, stated independently): greedy colouring
    of the sampled variables in id order, or — when that makes a sweep cheaper (fewer colour steps, weighted by the slowest
    update of each) — greedy in smallest-last order
    (repeatedly remove a variable of least remaining degree, ties by smallest id; colour in reverse removal order).
    Variables inside a colour in ascending id."""
    import heapq
    samp = oracle.Sampler(oracle.Generator(1), model, collapsed=True)
    fixed = model.fixed
    n = model.n_vars
    free = [v for v in range(n) if fixed[v] < 0]
    adj = {v: [u for u in samp.neighbors(v) if u != v and fixed[u] < 0] for v in free}

    def greedy(seq):
        colour = [-1] * n
        for v in seq:
            used = {colour[u] for u in adj[v] if colour[u] >= 0}
            c = 0
            while c in used:
                c += 1
            colour[v] = c
        return colour

    deg = {v: len(adj[v]) for v in free}
    heap = [(deg[v], v) for v in free]
    heapq.heapify(heap)
    removed, seq = set(), []
    while heap:
        d, v = heapq.heappop(heap)
        if v in removed or d != deg[v]:
            continue
        removed.add(v)
        seq.append(v)
        for u in adj[v]:
            if u not in removed:
                deg[u] -= 1
                heapq.heappush(heap, (deg[u], u))
    by_id, by_sl = greedy(free), greedy(seq[::-1])
    cards = model.cards
    mixed = len({int(cards[v]) for v in free}) > 1

    def sweep_cost(colour):  # one step per colour, as long as its slowest update (a non-binary update ~ 2.5 binary ones)
        return sum(max((2.5 if mixed and cards[v] > 2 else 1.0) for v in free if colour[v] == c) for c in range(max(colour) + 1))

    colour = by_sl if sweep_cost(by_sl) < sweep_cost(by_id) else by_id
    n_col = max(colour) + 1
    return [[v for v in range(n) if colour[v] == c] for c in range(n_col)]


def main():
    res = os.path.join(HERE, "res")
    rng = np.random.default_rng(20260101)
    cond, traj = [], []
    for name, evid in MODELS:
        om = oracle.Model.load(os.path.join(res, name), use_evidence=evid)
        cards, fixed = om.cards, om.fixed
        samp = oracle.Sampler(oracle.Generator(1), om)
        free = [v for v in range(om.n_vars) if fixed[v] < 0]
        states = random_states(rng, cards, fixed, 4)
        for st in states:
            for v in rng.choice(free, size=min(5, len(free)), replace=False):
                e = samp.conditional(int(v), st)
                cond.append({"model": name, "evidence": evid, "var": int(v), "state": st.tolist(), "e": [repr(float(x)) for x in e]})
        if name in ("Grids_11.uai", "Pedigree_11.uai", "ObjectDetection_11.uai", "sample.uai"):
            colours = greedy_schedule(oracle.Model.load(os.path.join(res, name), use_evidence=evid))
            order = [v for c in colours for v in c]
            n_chains, seed, first, n_sweeps = 8, 9001, 24, 3
            st0 = random_states(rng, cards, fixed, n_chains)
            entry = {"model": name, "evidence": evid, "colours": colours, "seed": seed, "first_chain": first,
                     "n_sweeps": n_sweeps, "initial": st0.tolist()}
            for bits in (53, 32):
                if bits == 32 and int(cards.max()) > 2:
                    continue
                s = oracle.Sampler(oracle.Generator(1), oracle.Model.load(os.path.join(res, name), use_evidence=evid))
                st, counts = s.sweep_run(order, seed, first, st0, 0, n_sweeps, bits=bits, record=True)
                entry["bits%d" % bits] = {"final": st.tolist(), "counts": [int(c) for c in counts]}
                if bits == 53:  # the Rao-Blackwell bins of the same trajectory (GB_CHAINS_RAO_BLACKWELL, units of 2^-24)
                    s = oracle.Sampler(oracle.Generator(1), oracle.Model.load(os.path.join(res, name), use_evidence=evid))
                    _, bins = s.sweep_run(order, seed, first, st0, 0, n_sweeps, bits=53, record=2)
                    entry["bits53"]["rb_bins"] = [int(b) for b in bins]
            traj.append(entry)
    with open(os.path.join(HERE, "conditionals.json"), "w") as f:
        json.dump(cond, f, separators=(",", ":"))
    with open(os.path.join(HERE, "trajectories.json"), "w") as f:
        json.dump(traj, f, separators=(",", ":"))
    print(f"wrote {len(cond)} conditionals and {len(traj)} trajectories")


if __name__ == "__main__":
    main()
