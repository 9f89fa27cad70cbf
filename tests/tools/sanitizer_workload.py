import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grample_b200 as gb
R=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'golden', 'res') + '/'
for name, evid in (('Grids_11.uai', False), ('Pedigree_11.uai', True), ('ObjectDetection_11.uai', False)):
    m = gb.Model.from_uai(R+name, use_evidence=evid, device=0)
    for prec in (gb.F64, gb.F32) + ((gb.TABLE,) if m.table_mode()[0] else ()):
        for pc in (False, True):  # resident multi-sweep kernels (TMA-staged tables) and per-colour launches (cp.async ring)
            ch = gb.Chains(m, 13, seed=1, precision=prec, history=True, device=0, per_colour=pc)
            ch.burnin(2); ch.advance(6); ch.scan(50) if prec == gb.F64 else None
            ch.merged_marginals(); ch.convergence()
    nm, v, marg = m.collapse(-1, seed=3)
    ch = gb.Chains([m, nm], [9, 17], seed=2, precision=gb.F32, history=True, device=0)
    ch.advance(4); ch.adapt(m, 2, 8, 4, first_chain_id=64)
    ch.advance(4); ch.merged_marginals()
# hybrid mode incl. wide records (single-collapsed variant of an all-binary model), Rao-Blackwell bins, prefix-staged tables
m = gb.Model.from_uai(R+'Promedus_11.uai', use_evidence=True, device=0)
fx = m.fixed
v = max((m.blanket_size(u), u) for u in range(m.n_vars) if fx[u] < 0 and m.blanket_size(u) <= 12)[1]
nm = m.collapse(v)[0]
for pc in (False, True):
    ch = gb.Chains([m, nm], [13, 21], seed=4, precision=gb.HYBRID, history=True, device=0, per_colour=pc)
    ch.advance(6); ch.merged_marginals(); ch.convergence()
m = gb.Model.from_uai(R+'ObjectDetection_11.uai', device=0)
by = {}
for u in range(m.n_vars):
    by.setdefault(m.blanket_size(u), u)
for b in (5, 6):
    nm = m.collapse(by[b])[0]
    for prec in (gb.F32, gb.F64):
        for rb in (False, True):
            ch = gb.Chains(nm, 37, seed=5, precision=prec, history=True, device=0, rao_blackwell=rb)
            ch.advance(4); ch.merged_marginals()
for pc in (False, True):
    ch = gb.Chains(m, 70, seed=6, precision=gb.F32, device=0, per_colour=pc, rao_blackwell=True); ch.sweep(3); ch.merged_marginals()
a = gb.ising_torus(32, 32)
m = gb.Model.from_arrays(*a, device=0)
for pc in (False, True):
    ch = gb.Chains(m, 2100, seed=1, precision=gb.TABLE, device=0, per_colour=pc); ch.sweep(3); ch.merged_marginals()
print('sanitizer workload ok')
