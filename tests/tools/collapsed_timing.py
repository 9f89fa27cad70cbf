"""Ad-hoc: BASELINE configs[3] kernel cost — ObjectDetection_11 with one collapsed variable per variant
(the new factor over the blanket has up to 11^6 entries: tables no longer fit shared memory)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import grample_b200 as gb
res = os.path.join(ROOT, "tests", "golden", "res")
m = gb.Model.from_uai(os.path.join(res, "ObjectDetection_11.uai"), device=0)
sizes = sorted((m.blanket_size(v), v) for v in range(m.n_vars))
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for b, v in [s for s in sizes if s[0] <= 7][::6]:
    var, _, _ = m.collapse(v)
    n_free = len(var.schedule()[0])
    for label, prec in (("f32", gb.F32), ("f64", gb.F64)):
        ch = gb.Chains(var, chains, seed=1, precision=prec, device=0)
        ch.sweep(5)
        us = ch.sweep_timed(50) / 50 * 1e3
        print(f"collapsed var {v} (blanket {b}), {chains} chains {label}: {us:.1f} us/sweep, {n_free * chains / us * 1e6:.3e} updates/s", flush=True)
