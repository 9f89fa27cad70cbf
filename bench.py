#!/usr/bin/env python
"""bench.py — variable-updates/sec of the Gibbs sweep hot path (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[4] — synthetic 1024x1024 binary Ising torus
(1 M variables, 3 M factors), 2-colour sweep, 65536 chains per GPU (64 GiB of uint8 state per
GPU, so inputs are far larger than the 126 MB L2 and no flush is needed between timed steps).
Default arithmetic (`--precision bits`, dtype "u32"): the reference's float64 conditional is
evaluated once per (variable, neighbour configuration) and stored as a 32-bit inverse-CDF
threshold; the sweep itself is integer work with 32-bit Philox draws on chain state packed one
bit per chain (8 GiB per GPU), 32 chains per machine word, bit-sliced (csrc/bits.cuh).
`--precision table` is the same arithmetic on uint8 state (64 GiB per GPU; round 1's kernel,
reported under `secondary.table_u8` at N = 1); `--precision f32|f64` run the per-update
log-sum-exp kernels (reported under `secondary.lse` at N = 1).
One "step" = one systematic sweep of every chain (n_vars x chains recorded single-variable
updates).  Chains shard across GPUs with no data-path collective.  `value` is the weak-scaling
figure (fixed 65536 chains per GPU); at N > 1 the `strong` block adds SURVEY 8d's split of the
same 65536 chains over the N GPUs.  The only exchange is the NCCL sum of the marginal counts at
the monitor interval — inside the library (gb_comm_*), overlapped with the next sweep, and part
of the e2e figure.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--chains C] [--side S]

Under torchrun (N > 1) every rank drives its own GPU; times are device times (CUDA events on
the library's stream), max over ranks.  `--impl reference` times the reference algorithm's CPU
restatement (oracle, "lean" variant, one thread per host core) on a bounded sample of the same
workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE line, the JSON result.  Native libraries write there too (NCCL prints its version banner on
# the first communicator of a process), so file descriptor 1 points at stderr for the whole run and the JSON line goes
# to a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "variable_updates_per_sec"
UNIT = "updates/s"
GATHER_BYTES_PER_UPDATE = 5.0  # 4 distinct neighbours read + 1 state byte written (uint8 state)
COMPULSORY_BYTES_PER_UPDATE = 2.0  # each state byte read once and written once per sweep (perfect neighbour reuse)
STATE_BYTES = {"bits": 0.125, "table": 1.0, "f32": 1.0, "f64": 1.0}  # bytes of chain state per (variable, chain)
FALLBACK_HBM_GBS = 6650.0      # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU")
    ap.add_argument("--side", type=int, default=1024, help="torus side (variables = side^2)")
    ap.add_argument("--wmax", type=float, default=4.9)
    ap.add_argument("--precision", default="bits", choices=["bits", "table", "f32", "f64"],
                    help="table: float64 conditionals tabulated per neighbour configuration, integer sweep")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the small-model secondary numbers")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.3 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


def cpu_reference_throughput(arrays, seconds, threads):
    """The reference algorithm's CPU restatement (oracle, lean variant) on a bounded sample of the
    workload: `threads` chains over the same model, one std::thread each (chain.go:197-215)."""
    import oracle
    om = oracle.Model.create(*arrays)
    # calibrate on a short run, then size the sample for ~`seconds`
    ups, secs = oracle.throughput(om, n_threads=threads, steps=20000, seed=1, lean=True)
    rate = ups / max(secs, 1e-9)
    steps = max(20000, int(rate * seconds / threads))
    ups, secs = oracle.throughput(om, n_threads=threads, steps=steps, seed=2, lean=True)
    return ups / secs, ups, secs


def small_model_rates(gb, dev, sweeps=500):
    """variable-updates/sec on the bundled problems of BASELINE.json configs[1..3] (state and tables are
    on-chip resident: these are issue/latency-bound, not HBM-bound, and are not the headline)"""
    res = os.path.join(ROOT, "tests", "golden", "res")
    out = {}
    for name, evid, n_chains in (("Promedus_11.uai", True, 4096), ("Pedigree_11.uai", True, 8192), ("Pedigree_11.uai", False, 8192),
                                 ("ObjectDetection_11.uai", False, 8192)):
        try:
            m = gb.Model.from_uai(os.path.join(res, name), use_evidence=evid, device=dev)
            n_free = len(m.schedule()[0])
            # hybrid = the hosts' default precision: threshold tables where a variable qualifies (cardinality <= 4), float64
            # log-sum-exp elsewhere (ObjectDetection_11: cardinality 11, so hybrid == f64 there)
            modes = [("f32", gb.F32), ("f64", gb.F64), ("hybrid", gb.HYBRID)] + ([("table", gb.TABLE)] if m.table_mode()[0] else [])
            entry = {"chains": n_chains, "free_vars": n_free}
            for label, prec in modes:
                ch = gb.Chains(m, n_chains, seed=1, precision=prec, device=dev)
                ch.sweep(20)
                ms = ch.sweep_timed(sweeps)
                entry[label] = {"us_per_sweep": round(1e3 * ms / sweeps, 2), "updates_per_sec": n_free * n_chains * sweeps / (ms * 1e-3)}
            out[name.replace(".uai", "") + ("+evid" if evid else "")] = entry
        except Exception as e:  # never lose the headline line over a secondary number
            out[name] = {"error": str(e)[:200]}
    return out


def config3_collapsed(gb, gbd, torch, dist, dev, rank, world, total_chains=65536, cw=2000, rounds=3):
    """BASELINE.json configs[3]: ObjectDetection_11 (cardinality 11), collapsed Gibbs, chains sharded over the N GPUs.
    Two collapsed variants (one collapsed variable each, gibbs-collapsed.go:98-314) x total_chains / 2 replicas; one
    "round" is the loop body of cmd/root.go:475-539 + 640-668: AdvanceChain of every chain (cw + 1 recorded sweeps with
    half-window histograms), MergeChains and ChainConvergence over ALL ranks (in-library NCCL), read on rank 0.
    Wall time per round, max over ranks."""
    res = os.path.join(ROOT, "tests", "golden", "res")
    out = {"problem": "ObjectDetection_11", "sampler": "collapsed", "variants": 2, "chains_total": total_chains, "cw": cw, "rounds": rounds}
    try:
        m = gb.Model.from_uai(os.path.join(res, "ObjectDetection_11.uai"), device=dev)
        picks = [v for v in range(m.n_vars) if m.blanket_size(v) <= 5][:2]  # small blankets: the new factor stays in shared memory
        variants = [m.collapse(v)[0] for v in picks]
        out["collapsed_variables"] = [int(v) for v in picks]
        per_variant = total_chains // 2
        first, n_local = gbd.shard(per_variant, world, rank)
        stride = (per_variant + 7) // 8 * 8
        for label, prec in (("f32", gb.F32), ("hybrid_f64", gb.HYBRID)):
            ch = gb.Chains(variants[0], n_local, seed=3, first_chain_id=first, precision=prec, history=True, device=dev)
            ch.add_group(variants[1], n_local, stride + first)
            gbd.attach(ch, dist)
            ch.burnin(200)
            ch.advance(cw)
            ch.merged_marginals()
            ch.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.time()
            for _ in range(rounds):
                ch.advance(cw)
                merged, col = ch.merged_marginals()
                conv = ch.convergence(gb.HELLINGER, merged)
            ch.synchronize()
            secs = time.time() - t0
            if dist is not None:
                t = torch.tensor([secs], device=f"cuda:{dev}", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                secs = float(t.item())
            n_free = sum(len(v.schedule()[0]) for v in variants) / 2.0
            updates = rounds * (cw + 1) * n_free * total_chains
            out[label] = {"updates_per_sec": updates / secs, "updates_per_sec_per_gpu": updates / secs / world,
                          "seconds_per_round": secs / rounds, "us_per_sweep": 1e6 * secs / (rounds * (cw + 1)),
                          "chains_per_gpu": 2 * n_local, "finite_scores": bool(np.isfinite(conv).all())}
            del ch
    except Exception as e:  # never lose the headline line over a secondary number
        out["error"] = str(e)[:300]
    return out


def samples_to_hellinger(gb, dev, threshold=0.01):
    """The second half of BASELINE.json's metric on the one bundled problem where it is finite:
    recorded updates until the mean Hellinger distance of the merged marginals to ObjectDetection_11.uai.MAR first drops
    below 0.01, checked every 38400 updates; 16 chains, float32, the reference's merge rule (per-chain uniform
    pseudo-count included).  Counting estimator (the reference's) and the Rao-Blackwell estimator, same trajectory."""
    res = os.path.join(ROOT, "tests", "golden", "res")
    out = {"problem": "ObjectDetection_11", "chains": 16, "precision": "f32", "threshold": threshold, "check_every_updates": 16 * 40 * 60,
           "reference_algorithm_cpu": {"value": 1278000, "se": 45000, "chains": 16, "seeds": 8,
                                       "source": "profiles/r02_accuracy_equal_chains.json: oracle (reference algorithm, random scan) with the SAME 16 chains, "
                                                 "mean +- s.e. over 8 seeds; device float64 1.33e6 +- 1.1e5, float32 1.37e6 +- 1.1e5 (equal within 1 s.e.); "
                                                 "with 8 chains on both sides: oracle 1.007e6 +- 3.3e4, device 1.02e6 +- 7.2e4.  The chain count is part of the "
                                                 "estimator: every chain adds a uniform 1/card pseudo-count (model/variable.go:45)"}}
    try:
        m = gb.Model.from_uai(os.path.join(res, "ObjectDetection_11.uai"), device=dev)
        cards, mar = gb.mar_load(os.path.join(res, "ObjectDetection_11.uai.MAR"))
        for label, rb in (("counts", False), ("rao_blackwell", True)):
            ch = gb.Chains(m, 16, seed=2024, precision=gb.F32, device=dev, rao_blackwell=rb)
            ch.burnin(2000)
            reached = None
            for _ in range(80):
                ch.sweep(40)
                h = gb.error_suite(cards, mar, ch.merged_marginals()[0])["MeanHellinger"]
                if h < threshold:
                    reached = int(ch.total_samples)
                    break
            out[label] = {"samples": reached, "mean_hellinger_at_stop": h}
    except Exception as e:  # never lose the headline line over a secondary number
        out["error"] = str(e)[:200]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import grample_b200 as gb
    threads = os.cpu_count() or 1
    arrays = gb.ising_torus(args.side, args.side, wmax=args.wmax)
    import oracle
    om = oracle.Model.create(*arrays)
    per_step = max(20000, int(2.0e5 * 1.0))  # recorded updates per thread per step (bounded sample)
    for _ in range(args.warmup):
        oracle.throughput(om, n_threads=threads, steps=per_step // 4, seed=1, lean=True)
    tot_u, tot_s = 0, 0.0
    for i in range(args.steps):
        ups, secs = oracle.throughput(om, n_threads=threads, steps=per_step, seed=10 + i, lean=True)
        tot_u += ups
        tot_s += secs
    value = tot_u / tot_s
    sample = (f"{threads} chains x {per_step} recorded single-variable updates per step on the {args.side}x{args.side} "
              f"Ising torus (random scan, float64, lean bookkeeping), {args.steps} steps")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ising_torus_{args.side}x{args.side}", "chains": threads, "schedule": "random scan"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def measure(gb, gbd, torch, dist, chains, model, args, updates_per_step, world, dev):
    """(device ms of K sweeps, launches, clocks, e2e wall ms of K intervals, last merged marginals)"""
    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        chains.sweep(1, record=True)
    chains.synchronize()
    # ---------------- timed region: exactly K steps, device time, max over ranks
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    launches0 = chains.launch_count
    barrier()
    t0 = time.time()
    ms = chains.sweep_timed(args.steps, record=True)
    barrier()
    t1 = time.time()
    launches = chains.launch_count - launches0
    clocks = sampler.stop(t0, t1)
    if dist is not None:
        t = torch.tensor([ms], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---------------- e2e: the call sequence a user makes each monitor interval (cmd/root.go:475-539): advance ->
    # MergeChains read back to the host -> score.  The merge of interval i (count sums, NCCL all-reduce over the ranks
    # inside the library, float64 marginals, D2H into the caller's pinned buffer) runs on the handle's side stream
    # while the sweep of interval i + 1 runs: gb_chains_merge_begin / gb_chains_merge_end.  Every interval's result is
    # delivered inside the timed region (the last one after the last sweep).
    total_card, n_vars = model.total_card, model.n_vars
    rank0 = dist is None or dist.get_rank() == 0
    # Only rank 0 reads the merged marginals on the host, as only the reference's main goroutine does (cmd/root.go:498-539):
    # the other ranks take part in the reduction and receive the totals, but ask for no device-to-host copy.
    bufs = [(torch.empty(total_card, dtype=torch.float64).pin_memory().numpy(), np.empty(n_vars, dtype=np.int32)) if rank0 else "none"
            for _ in range(3)]
    for i in range(args.warmup):  # warm the interval path too (first call creates the side stream and scratch)
        chains.sweep(1, record=True)
        chains.merge_begin(bufs[i % 3])
        chains.merge_end()
    barrier()
    e0 = time.time()
    marks, stages = [], []
    # two merges may be in flight: interval i's sweep AND snapshot are enqueued before the host waits for interval i - 1
    chains.sweep(1, record=True)
    chains.merge_begin(bufs[0])
    for i in range(1, args.steps):
        chains.sweep(1, record=True)
        chains.merge_begin(bufs[i % 3])
        merged, _, n_all, samples_all = chains.merge_end()   # interval i - 1's merged marginals arrive
        marks.append(time.time())
        stages.append(chains.merge_timing())
    merged, _, n_all, samples_all = chains.merge_end()
    marks.append(time.time())
    stages.append(chains.merge_timing())
    measure.merge_stage_ms = stages  # per interval: [count sums, NCCL sum incl. waiting for the slowest rank, conversion, D2H]
    barrier()
    e_ms = (time.time() - e0) * 1e3
    measure.interval_ms = [round((b - a) * 1e3, 3) for a, b in zip([e0] + marks[:-1], marks)]  # arrival times of the merged marginals
    if dist is not None:
        t = torch.tensor([e_ms], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    assert n_all == world * chains.n_chains, (n_all, world, chains.n_chains)
    return ms, launches, clocks, e_ms, merged


def other_rate(gb, model, args, dev, n_vars, peak, prec, kernel, steps):
    """another kernel on the same workload (uint8 state, SURVEY 8d's 5 B per update yardstick)"""
    try:
        ch = gb.Chains(model, args.chains, seed=20260101, precision=prec, device=dev)
        ch.sweep(1, record=True)
        ms = ch.sweep_timed(steps, record=True)
        v = n_vars * args.chains * steps / (ms * 1e-3)
        del ch
        return {"kernel": kernel, "value": v, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                "roofline_frac_at_5B": GATHER_BYTES_PER_UPDATE * v / 1e9 / peak}
    except Exception as e:  # never lose the headline line over a secondary number
        return {"error": str(e)[:200]}


def lse_rates(gb, model, args, dev, n_vars, peak):
    """The general per-update log-sum-exp path (`k_sweep_colour<Real,2,4>`) on the SAME workload: everything table
    mode rejects lands on these kernels, so the bench line carries them next to the headline (N = 1 only)."""
    out = {}
    for label, prec, steps in (("f32", gb.F32, 3), ("f64", gb.F64, 2)):
        try:
            ch = gb.Chains(model, args.chains, seed=20260101, precision=prec, device=dev)
            ch.sweep(1, record=True)
            ms = ch.sweep_timed(steps, record=True)
            v = n_vars * args.chains * steps / (ms * 1e-3)
            out[label] = {"kernel": "k_sweep_colour<%s,2,4>" % ("float" if label == "f32" else "double"), "value": v, "unit": UNIT,
                          "ms_per_step": ms / steps, "steps": steps, "roofline_frac": GATHER_BYTES_PER_UPDATE * v / 1e9 / peak}
            del ch
        except Exception as e:  # never lose the headline line over a secondary number
            out[label] = {"error": str(e)[:200]}
    return out


def run_native(args):
    import torch
    import grample_b200 as gb
    from grample_b200 import distributed as gbd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    prec = {"bits": gb.TABLE_BITS, "table": gb.TABLE, "f32": gb.F32, "f64": gb.F64}[args.precision]
    kernel = {"bits": "k_sweep_bits<2>", "table": "k_sweep_tab<64,4,false,3>", "f32": "k_sweep_colour<float,2,4>",
              "f64": "k_sweep_colour<double,2,4>"}[args.precision]
    dtype = {"bits": "u32", "table": "u32", "f32": "f32", "f64": "f64"}[args.precision]
    t_setup = time.time()
    arrays = gb.ising_torus(args.side, args.side, wmax=args.wmax)
    model = gb.Model.from_arrays(*arrays, device=dev)
    n_vars = model.n_vars
    order, coff = model.schedule()
    n_colours = len(coff) - 1
    chains = gb.Chains(model, args.chains, seed=20260101, first_chain_id=rank * args.chains, precision=prec, device=dev)
    gbd.attach(chains, dist)  # N > 1: the library's own NCCL communicator (torch.distributed only carries its 128-byte id)
    chains.synchronize()
    setup_s = time.time() - t_setup
    updates_per_step = n_vars * args.chains  # per GPU
    model_bytes = int(sum(a.nbytes for a in arrays))
    total_card = model.total_card

    ms, launches, clocks, e_ms, merged = measure(gb, gbd, torch, dist, chains, model, args, updates_per_step, world, dev)
    value = world * updates_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * updates_per_step * args.steps / (e_ms * 1e-3)
    weak_intervals, weak_stages = measure.interval_ms, measure.merge_stage_ms
    score = gb.error_suite(model.cards, np.full(total_card, 0.5), merged) if rank == 0 else None  # host scoring of the read-back (untimed sanity use)

    # ---------------- strong scaling (the SAME 65536 chains split over the N GPUs), N > 1 only
    strong = None
    if world > 1:
        del chains
        per = args.chains // world // 32 * 32  # (a bit-packed state word holds 32 chains: shards start on multiples of 32)
        sch = gb.Chains(model, per, seed=20260101, first_chain_id=rank * per, precision=prec, device=dev)
        gbd.attach(sch, dist)
        sch.synchronize()
        s_ms, s_launches, s_clocks, s_e_ms, _ = measure(gb, gbd, torch, dist, sch, model, args, n_vars * per, world, dev)
        s_updates = world * n_vars * per * args.steps
        strong = {"chains_total": per * world, "chains_per_gpu": per, "value": s_updates / (s_ms * 1e-3), "unit": UNIT,
                  "ms_per_step": s_ms / args.steps, "e2e": {"value": s_updates / (s_e_ms * 1e-3), "unit": UNIT, "ms_per_step": s_e_ms / args.steps,
                                                             "d2h_bytes_per_step": int(total_card * 8 + 16), "interval_ms": measure.interval_ms, "merge_stage_ms": measure.merge_stage_ms},
                  "gpu_launches": int(s_launches), "clocks": s_clocks,
                  "note": "efficiency = strong.value (or strong.e2e.value) / the N = 1 run's value (e2e.value): same total work"}
        chains = sch

    # ---------------- roofline of the dominant kernel (one launch = one colour of one sweep)
    peak, peak_src = hbm_peak()
    launch_ms = ms / (args.steps * n_colours)
    updates_per_launch = updates_per_step / n_colours
    sb = STATE_BYTES[args.precision]
    gather, compulsory = GATHER_BYTES_PER_UPDATE * sb, COMPULSORY_BYTES_PER_UPDATE * sb
    rate = updates_per_launch / (launch_ms * 1e-3)  # updates per second of one launch on one GPU
    achieved = gather * rate / 1e9
    roofline = {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": kernel,
                "algorithmic_bytes_per_update": gather, "updates_per_launch": updates_per_launch, "launch_ms": launch_ms,
                "compulsory_bytes_per_update": compulsory, "frac_of_compulsory_ceiling": compulsory * rate / 1e9 / peak,
                "frac_at_uint8_5B_yardstick": GATHER_BYTES_PER_UPDATE * rate / 1e9 / peak,
                "bound_note": "achieved / peak / frac: SURVEY 8d's gather bytes (4 neighbour reads + 1 write of %g B of state each) over the measured "
                              "HBM copy bandwidth; frac_at_uint8_5B_yardstick is the same rate on round 1's uint8 yardstick (5 B per update).  "
                              "ncu (profiles/) shows the kernel bound by instruction issue, not DRAM: `issue` holds the fraction of the "
                              "issue-slot ceiling (1 warp instruction per cycle per SM sub-partition)" % sb}
    # DRAM traffic and instruction counts of the dominant kernel from the committed ncu --set full capture
    # (profiles/traffic.json: per update at the capture's chain count, scaled to this launch's update count)
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr)).get(args.precision)
            if t:
                roofline["traffic"] = t["dram_bytes_per_update"] * updates_per_launch
                roofline["traffic_source"] = t["source"]
                sm_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
                issue_peak = 148 * 4 * sm_hz  # warp instructions per second: 148 SMs x 4 sub-partitions x 1 per cycle
                issued = t["lane_instructions_per_update"] / 32.0 * rate
                roofline["issue"] = {"achieved": issued / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s", "frac": issued / issue_peak,
                                     "lane_instructions_per_update": t["lane_instructions_per_update"], "ncu": t.get("ncu")}
        except Exception:
            pass

    # ---------------- BASELINE configs[3] at this N (collective: every rank takes part)
    cfg3 = None
    if not args.no_secondary:
        del chains
        chains = None
        cfg3 = config3_collapsed(gb, gbd, torch, dist, dev, rank, world)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- secondary workloads (reported next to the headline, N = 1 only): the general log-sum-exp path on
    # the same workload, and the bundled UAI problems of BASELINE.json configs[1..3] at their chain counts
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        if args.precision == "bits":
            secondary["table_u8"] = other_rate(gb, model, args, dev, n_vars, peak, gb.TABLE, "k_sweep_tab<64,4,false,3>", 3)
            secondary["lse"] = lse_rates(gb, model, args, dev, n_vars, peak)
        secondary.update(small_model_rates(gb, dev))
        secondary["samples_to_mean_hellinger_below_0.01"] = samples_to_hellinger(gb, dev)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, ups, secs = cpu_reference_throughput(arrays, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{threads} chains x {ups // threads} recorded updates on the same {args.side}x{args.side} Ising torus, "
                         f"random scan float64 oracle (lean bookkeeping), {secs:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"mode": args.precision, "workload": f"ising_torus_{args.side}x{args.side}", "variables": n_vars, "factors": 3 * n_vars,
                       "chains_per_gpu": args.chains, "colours": n_colours, "wmax": args.wmax,
                       "state": "1 bit per chain, uint32 [var][chain / 32]" if args.precision == "bits" else "uint8 [var][chain]",
                       "l2_policy": "inputs (state %.1f GiB per GPU) larger than L2, no flush" % (n_vars * args.chains * STATE_BYTES[args.precision] / 2**30),
                       "parallelism": f"chains sharded over {world} GPU(s), no data-path collective",
                       "setup_seconds": round(setup_s, 2)},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(total_card * 8 + 16), "ms_per_step": e_ms / args.steps, "interval_ms": weak_intervals, "merge_stage_ms": weak_stages,
                    "path": "per step: gb_chains_sweep, gb_chains_merge_begin, gb_chains_merge_end (the previous step's merged marginals "
                            "in rank 0's pinned host buffer); in-library NCCL sum of the uint64 counts when N > 1",
                    "note": "the interval loop of cmd/root.go has no per-interval host input: chain state is device-resident "
                            "(as each Go chain's state is resident in its goroutine); the model (CSR + tables, "
                            f"{model_bytes} bytes) is uploaded once from host arrays during setup_seconds, and the collapsed "
                            "flags the merge needs are uploaded once per change of the chain set"},
            "roofline": roofline, "sanity": {"mean_hellinger_vs_uniform": score["MeanHellinger"]}}
    if strong is not None:
        line["strong"] = strong
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if cfg3 is not None:
        secondary = secondary or {}
        secondary["config3_objectdetection_collapsed"] = cfg3
    if secondary is not None:
        line["secondary"] = secondary
    emit(line)
    if dist is not None:
        del chains  # (and with it the library's communicator) before torch tears its own down
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
