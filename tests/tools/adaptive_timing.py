"""Ad-hoc: time BASELINE configs[2] (Pedigree_11 + evidence, adaptive: base + collapsed variants x replicas)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import grample_b200 as gb
res = os.path.join(ROOT, "tests", "golden", "res")
prec = {"f32": gb.F32, "f64": gb.F64, "hybrid": gb.HYBRID}[sys.argv[1] if len(sys.argv) > 1 else "f32"]
replicas, cw = 64, 200
m = gb.Model.from_uai(os.path.join(res, "Pedigree_11.uai"), use_evidence=True, device=0)
ch = gb.Chains([m, m], [replicas, replicas], seed=1, precision=prec, history=True, device=0)
ch.burnin(50)
next_id = 2 * replicas
t0 = time.time()
while ch.n_groups < 128:
    ch.advance(cw)
    chosen = ch.adapt(m, 16, replicas, cw, first_chain_id=next_id)
    next_id += len(chosen) * replicas
    if not len(chosen):
        break
ch.synchronize()
print(f"grew to {ch.n_groups} groups / {ch.n_chains} chains in {time.time()-t0:.2f} s")
n_free = len(m.schedule()[0])
for rep in range(2):
    t0 = time.time(); ch.advance(cw); ch.synchronize(); dt = time.time() - t0
    print(f"advance(cw={cw}) over {ch.n_groups} groups: {dt*1e3:.1f} ms  -> {(cw+1)*ch.n_chains*n_free/dt:.3e} updates/s (upper bound: collapsed variants sample one variable fewer)")
