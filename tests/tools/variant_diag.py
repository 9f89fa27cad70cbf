"""Ad-hoc: how much of a collapsed Pedigree variant does hybrid mode tabulate, and what does a sweep cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import grample_b200 as gb
m = gb.Model.from_uai(os.path.join(ROOT, "tests", "golden", "res", "Pedigree_11.uai"), use_evidence=True, device=0)
cands = [v for v in range(m.n_vars) if m.fixed[v] < 0 and 1 < m.blanket_size(v) <= 12]
print("candidates", len(cands))
for v in cands[:: max(1, len(cands) // 6)][:6]:
    nm, _, _ = m.collapse(v)
    order, coff = nm.schedule()
    mask = nm.hybrid_mask()
    untab = [int(u) for u in order if not mask[u]]
    line = f"collapse {v} (blanket {m.blanket_size(v)}): sampled {len(order)}, colours {len(coff)-1}, untabulated {len(untab)} blankets {[nm.blanket_size(u) for u in untab][:8]}"
    for prec, name in ((gb.HYBRID, "hybrid"), (gb.F64, "f64"), (gb.F32, "f32")):
        ch = gb.Chains(nm, 8192, seed=1, precision=prec, device=0)
        ch.sweep(20)
        line += f" | {name} {1e3 * ch.sweep_timed(200) / 200:.1f} us"
    print(line)
