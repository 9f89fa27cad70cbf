import sys, os
sys.path.insert(0, "/root/repo")
import grample_b200 as gb
RES = "/root/repo/tests/golden/res"
for evid in (False, True):
    m = gb.Model.from_uai(os.path.join(RES, "Pedigree_11.uai"), use_evidence=evid, device=0)
    order, coff = m.schedule()
    for prec, name in ((gb.HYBRID, "hybrid"), (gb.F32, "f32")):
        ch = gb.Chains(m, 8192, seed=1, precision=prec, device=0)
        ch.sweep(20)
        ms = ch.sweep_timed(300)
        print("evid", evid, "colours", [int(coff[i+1]-coff[i]) for i in range(len(coff)-1)], name, round(1e3*ms/300, 2), "us/sweep", flush=True)
