"""Ad-hoc: a few launches of the per-colour LSE kernel on a 1024x1024 Ising torus (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
prec = {"f32": gb.F32, "f64": gb.F64}[sys.argv[1]]
m = gb.Model.from_arrays(*gb.ising_torus(1024, 1024, wmax=4.9), device=0)
ch = gb.Chains(m, int(sys.argv[2]), seed=1, precision=prec, device=0)
ch.sweep(2, record=True)
ch.synchronize()
