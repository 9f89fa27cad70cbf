"""Ad-hoc: cost of keeping the per-chain half-window histograms (GB_CHAINS_HISTORY) in a round, per problem."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
res = os.path.join(ROOT, "tests", "golden", "res")
cw = 400
for name, evid, chains, precs in (("ObjectDetection_11.uai", False, 8192, ("f32", "f64")), ("Pedigree_11.uai", True, 8192, ("f32", "hybrid")),
                                  ("Promedus_11.uai", True, 4096, ("f32", "hybrid"))):
    m = gb.Model.from_uai(os.path.join(res, name), use_evidence=evid, device=0)
    n_free = len(m.schedule()[0])
    for label in precs:
        prec = {"f32": gb.F32, "f64": gb.F64, "hybrid": gb.HYBRID}[label]
        out = []
        for hist in (False, True):
            ch = gb.Chains(m, chains, seed=1, precision=prec, history=hist, device=0)
            ch.advance(cw); ch.synchronize()
            t0 = time.time(); ch.advance(cw); ch.synchronize(); dt = time.time() - t0
            out.append(dt / (cw + 1) * 1e6)
        print(f"{name} {label} {chains} chains: round of cw={cw}: {out[0]:.1f} us/sweep without histories, {out[1]:.1f} with", flush=True)
