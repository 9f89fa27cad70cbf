"""Ad-hoc timing breakdown of the per-interval call sequence (sweep + merged marginals)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import grample_b200 as gb

chains_n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
arrays = gb.ising_torus(1024, 1024, wmax=4.9)
m = gb.Model.from_arrays(*arrays, device=0)
ch = gb.Chains(m, chains_n, seed=1, precision=gb.TABLE, device=0)
for _ in range(3):
    ch.sweep(1, record=True)
ch.synchronize()
for rep in range(3):
    t0 = time.time(); ch.sweep(1, record=True); t1 = time.time(); ch.synchronize(); t2 = time.time()
    out, col = ch.merged_marginals(); t3 = time.time()
    _ = ch.total_samples; t4 = time.time()
    print(f"launch {1e3*(t1-t0):.2f} ms, sweep wait {1e3*(t2-t1):.2f}, merged {1e3*(t3-t2):.2f}, total_samples {1e3*(t4-t3):.2f}")
t0 = time.time()
for _ in range(5):
    ch.sweep(1, record=True); out, col = ch.merged_marginals()
print(f"loop: {1e3*(time.time()-t0)/5:.2f} ms/step")
print("timed:", ch.sweep_timed(5, record=True) / 5, "ms/sweep")
