"""Ad-hoc: LSE per-colour kernel on the Ising workload with and without marginal recording."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
chains_n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
arrays = gb.ising_torus(1024, 1024, wmax=4.9)
m = gb.Model.from_arrays(*arrays, device=0)
for prec, name in ((gb.F32, "f32"), (gb.F64, "f64")):
    ch = gb.Chains(m, chains_n, seed=1, precision=prec, device=0)
    ch.sweep(1, record=True)
    for rec in (True, False):
        ms = ch.sweep_timed(2, record=rec) / 2
        print(f"{name} record={rec}: {ms:.1f} ms/sweep  {1048576*chains_n/ms/1e9:.4f}e12 updates/s")
    del ch
