"""Ad-hoc: one launch of the resident LSE kernel on a bundled problem (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
name, evid, chains, prec = sys.argv[1], sys.argv[2] == "1", int(sys.argv[3]), {"f32": gb.F32, "f64": gb.F64, "table": gb.TABLE}[sys.argv[4]]
m = gb.Model.from_uai(os.path.join(ROOT, "tests", "golden", "res", name), use_evidence=evid, device=0)
ch = gb.Chains(m, chains, seed=1, precision=prec, device=0)
ch.sweep(20)
print(ch.sweep_timed(200) / 200 * 1e3, "us/sweep")
