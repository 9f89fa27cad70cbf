"""Ad-hoc: time the table sweep with an alternative build of the library (argv[1] = .so name)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from grample_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1]:
    _lib.LIB_PATH = os.path.join(ROOT, "grample_b200", sys.argv[1])
import grample_b200 as gb
chains_n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
arrays = gb.ising_torus(1024, 1024, wmax=4.9)
m = gb.Model.from_arrays(*arrays, device=0)
ch = gb.Chains(m, chains_n, seed=1, precision=gb.TABLE, device=0)
ch.sweep(3, record=True)
ch.synchronize()
ms = ch.sweep_timed(10, record=True) / 10
print(f"{sys.argv[1:]} {ms:.3f} ms/sweep  {1048576*chains_n/ms/1e9:.4f}e12 updates/s")
