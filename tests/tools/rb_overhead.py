"""Ad-hoc: cost of the Rao-Blackwell estimator (GB_CHAINS_RAO_BLACKWELL) on the bundled problems, us per recorded sweep."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grample_b200 as gb
res = os.path.join(ROOT, "tests", "golden", "res")
for name, evid, chains in (("ObjectDetection_11.uai", False, 8192), ("Promedus_11.uai", True, 4096), ("Pedigree_11.uai", True, 8192)):
    m = gb.Model.from_uai(os.path.join(res, name), use_evidence=evid, device=0)
    for label, prec in (("f32", gb.F32), ("f64", gb.F64)):
        out = []
        for rb in (False, True):
            ch = gb.Chains(m, chains, seed=1, precision=prec, device=0, rao_blackwell=rb)
            ch.sweep(20)
            out.append(ch.sweep_timed(200) / 200 * 1e3)
        print(f"{name} {label}: counts {out[0]:.2f} us/sweep, rao-blackwell {out[1]:.2f} us/sweep")
