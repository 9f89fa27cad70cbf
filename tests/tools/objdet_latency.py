"""ObjectDetection_11 on the one-thread-per-chain resident kernel: time per sweep against the number of chains.
32 chains are ONE CTA (16 warps on one SM): the latency of the 7 colour steps of a sweep, the floor of this kernel's
design; 4736 = one CTA of 32 chains on every SM; 8192 = BASELINE's population (two CTAs on 108 SMs, one on 40)."""
import os, sys
sys.path.insert(0, "/root/repo")
import grample_b200 as gb
RES = "/root/repo/tests/golden/res"
m = gb.Model.from_uai(os.path.join(RES, "ObjectDetection_11.uai"), use_evidence=False, device=0)
order, coff = m.schedule()
print("colours", [int(coff[i + 1] - coff[i]) for i in range(len(coff) - 1)], flush=True)
for prec, name in ((gb.F32, "f32"), (gb.F64, "f64")):
    for chains in (32, 1024, 4736, 8192):
        ch = gb.Chains(m, chains, seed=5, precision=prec, device=0)
        ch.sweep(20)
        ms = ch.sweep_timed(400)
        print(name, chains, "chains", round(1e3 * ms / 400, 2), "us/sweep", flush=True)
