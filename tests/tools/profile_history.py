"""Ad-hoc: one round of the resident table kernel with half-window histograms on Pedigree_11 + evidence (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
hist = len(sys.argv) < 2 or sys.argv[1] != "0"
m = gb.Model.from_uai(os.path.join(ROOT, "tests", "golden", "res", "Pedigree_11.uai"), use_evidence=True, device=0)
ch = gb.Chains(m, 2048, seed=1, precision=gb.HYBRID, history=hist, device=0)
ch.advance(100)
ch.advance(100)
ch.synchronize()
print("done")
