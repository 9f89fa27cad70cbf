#!/usr/bin/env python
"""bench_small.py — variable-updates/sec on the bundled UAI problems of BASELINE configs[1..3]
(small models: state and tables are L2/L1 resident, so these are launch- and issue-bound, not
HBM-bound).  Prints one JSON line per (problem, mode)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import grample_b200 as gb

RES = os.path.join(ROOT, "tests", "golden", "res")
CASES = [("Promedus_11.uai", True, 4096, 4.67), ("Pedigree_11.uai", True, 8192, 6.13), ("ObjectDetection_11.uai", False, 8192, 6.5),
         ("Grids_11.uai", False, 4096, 5.0)]


def main():
    sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    for name, evid, chains, bytes_per_update in CASES:
        m = gb.Model.from_uai(os.path.join(RES, name), use_evidence=evid, device=0)
        order, coff = m.schedule()
        modes = [("f32", gb.F32), ("f64", gb.F64), ("hybrid", gb.HYBRID)] + ([("table", gb.TABLE)] if m.table_mode()[0] else [])
        for label, prec in modes:
            ch = gb.Chains(m, chains, seed=1, precision=prec, device=0)
            ch.sweep(20)
            ms = ch.sweep_timed(sweeps)
            ups = len(order) * chains * sweeps / (ms * 1e-3)
            print(json.dumps({"problem": name, "mode": label, "chains": chains, "free_vars": len(order), "colours": len(coff) - 1,
                              "sweeps": sweeps, "us_per_sweep": 1e3 * ms / sweeps, "updates_per_sec": ups,
                              "gather_GBps": ups * bytes_per_update / 1e9}), flush=True)


if __name__ == "__main__":
    main()
