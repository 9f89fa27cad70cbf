import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grample_b200 as gb
res = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'golden', 'res')
m = gb.Model.from_uai(os.path.join(res, "Pedigree_11.uai"), use_evidence=True, device=0)
cw = 400
for chains in (2048, 4096, 8192, 16384):
    out = []
    for hist in (False, True):
        ch = gb.Chains(m, chains, seed=1, precision=gb.HYBRID, history=hist, device=0)
        ch.advance(cw); ch.synchronize()
        t0 = time.time(); ch.advance(cw); ch.synchronize(); dt = time.time() - t0
        out.append(dt / (cw + 1) * 1e6)
    print(f"Pedigree hybrid {chains} chains: {out[0]:.2f} us/sweep without histories, {out[1]:.2f} with", flush=True)
