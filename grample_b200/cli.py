"""`python -m grample_b200 sample ...` — the `grample sample` command of the reference
(cmd/root.go:163-250 flags, :309-719 main loop) driving the device path.

Same flags and defaults as the reference where they make sense on the device; new knobs are
`--replicas` (device chains behind each reference chain), `--precision` and `--device`.
Status / error-report lines keep the reference's format (cmd/root.go:256-306, 498-504) and
`--experiment` writes the CSV the reference's script/trace_file_process.py parses
(cmd/root.go:455-458, 520-533).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

from . import core
from ._lib import F32, F64, HELLINGER, HYBRID, JS, MAX_ABS, MEAN_ABS, TABLE


def letter26(n):
    """model/variable.go:167-189: Excel-column style variable names (0 -> A, 25 -> Z, 26 -> AA)"""
    if n == 0:
        return "A"
    n += 1
    digits = []
    while n > 0:
        n, r = divmod(n - 1, 26)
        digits.append("ABCDEFGHIJKLMNOPQRSTUVWXYZ"[r])
    return "".join(reversed(digits))


def variable_json(vid, card, fixed, marginal, state, collapsed):
    """One line of the trace file: encoding/json of model.Variable (model/variable.go:10-18; struct field
    order, map keys sorted) — what script/trace_file_process.py reads from the `// VARS (ESTIMATED)` section."""
    rec = {"ID": int(vid), "Name": letter26(int(vid)), "Card": int(card), "FixedVal": int(fixed),
           "Marginal": [float(x) for x in marginal], "State": {k: float(state[k]) for k in sorted(state)},
           "Collapsed": bool(collapsed)}
    return json.dumps(rec, separators=(",", ":"))


def per_var_measures(cards, p, q, fixed_p=None, fixed_q=None):
    """model/error.go:81-249 per variable: dict measure-name -> array [n_vars] (0 where either side is fixed).
    Names follow the `*-Error` state keys of cmd/root.go:651-655."""
    cards = np.asarray(cards)
    offs = np.concatenate([[0], np.cumsum(cards)])
    n = len(cards)
    out = {k: np.zeros(n) for k in ("Hell", "JS", "MaxAD", "AvgAD")}
    for v in range(n):
        if (fixed_p is not None and fixed_p[v] >= 0) or (fixed_q is not None and fixed_q[v] >= 0):
            continue
        a, b = np.asarray(p[offs[v]:offs[v + 1]], dtype=np.float64), np.asarray(q[offs[v]:offs[v + 1]], dtype=np.float64)
        a, b = a / max(a.sum(), 1e-12), b / max(b.sum(), 1e-12)
        d = np.abs(a - b)
        out["MaxAD"][v], out["AvgAD"][v] = d.max(), d.sum() / len(d)
        out["Hell"][v] = math.sqrt(float(((np.sqrt(a) - np.sqrt(b)) ** 2).sum())) / math.sqrt(2.0)
        mid = np.maximum((a + b) * 0.5, 1e-12)
        ac, bc = np.maximum(a, 1e-12), np.maximum(b, 1e-12)
        out["JS"][v] = 0.5 * (float((ac * np.log2(ac / mid)).sum()) + float((bc * np.log2(bc / mid)).sum()))
    return out


class Monitor:
    """The gauges of cmd/monitor.go:53-67 under the reference's expvar names.  `snapshot()` is the JSON
    object expvar serves at /debug/vars; with an address it is served over HTTP like the reference's
    monitor (every path redirects there), otherwise it is only kept (and written to the trace file)."""
    NAMES = ["Burn-In", "Convergence-Window", "Base-Chain-Count", "Chain-Adds-For-Adaptive-Step", "Total-Chain-Count",
             "Max-Iterations", "Max-Seconds", "Run-Time", "Total-Samples", "Iterations", "Last-Mean-Hellinger",
             "Last-Max-Hellinger", "Last-Mean-JSD", "Last-Max-JSD"]

    def __init__(self):
        self.v = {k: 0 for k in self.NAMES}
        self.server = None

    def set(self, name, value):
        assert name in self.v, name
        self.v[name] = value

    def add(self, name, delta):
        self.v[name] += delta

    def score(self, es):  # cmd/root.go:262-265
        self.set("Last-Mean-Hellinger", es["MeanHellinger"])
        self.set("Last-Max-Hellinger", es["MaxHellinger"])
        self.set("Last-Mean-JSD", es["MeanJSDiverge"])
        self.set("Last-Max-JSD", es["MaxJSDiverge"])

    def snapshot(self):
        return dict(self.v)

    def start(self, addr):
        from http.server import BaseHTTPRequestHandler, HTTPServer
        host, _, port = addr.rpartition(":")
        mon = self

        class H(BaseHTTPRequestHandler):
            def do_GET(self):
                if self.path != "/debug/vars":
                    self.send_response(307)
                    self.send_header("Location", "/debug/vars")
                    self.end_headers()
                    return
                body = json.dumps(mon.snapshot()).encode()
                self.send_response(200)
                self.send_header("Content-Type", "application/json; charset=utf-8")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def log_message(self, *a):
                pass

        self.server = HTTPServer((host or "127.0.0.1", int(port)), H)
        threading.Thread(target=self.server.serve_forever, daemon=True).start()
        sys.stderr.write("HTTP now available at %s (see debug/vars/)\n" % addr)

    def stop(self):
        if self.server is not None:
            self.server.shutdown()
            self.server.server_close()
            self.server = None


def error_report(prefix, es, short, out):
    """cmd/root.go:256-306"""
    rows = [("MAE" if short else "MeanAbsError", es["MeanMeanAbsError"], es["MaxMeanAbsError"]),
            ("XAE" if short else "MaxAbsError", es["MeanMaxAbsError"], es["MaxMaxAbsError"]),
            ("HEL" if short else "Hellinger", es["MeanHellinger"], es["MaxHellinger"]),
            ("JSD" if short else "JS Diverge", es["MeanJSDiverge"], es["MaxJSDiverge"])]
    nl = lambda x: -math.log2(x) if x > 0 else float("inf")
    if short:
        out.write("".join("%s=>%.6f(%7.3f),X%.6f(%7.3f) | " % (t, a, nl(a), b, nl(b)) for t, a, b in rows) + "\n")
    else:
        out.write("%s ... M:mean(neg log), X:max(neg log)\n" % prefix)
        for t, a, b in rows:
            out.write("%15s => M:%.6f(%7.3f) X:%.6f(%7.3f)\n" % (t, a, nl(a), b, nl(b)))


def build_parser():
    ap = argparse.ArgumentParser(prog="grample_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("sample", help="estimate marginals by Gibbs sampling on the GPU")
    s.add_argument("-v", "--verbose", action="store_true")
    s.add_argument("-e", "--seed", type=int, default=0)
    s.add_argument("-t", "--trace", default="")
    s.add_argument("-s", "--sampler", default="simple", choices=["simple", "collapsed", "adaptive"])
    s.add_argument("-m", "--model", required=True)
    s.add_argument("-d", "--evidence", action="store_true", help="apply MODEL.evid")
    s.add_argument("-o", "--solution", action="store_true", help="score against MODEL.MAR")
    s.add_argument("-b", "--burnin", type=int, default=-1, help="single-variable steps per chain (default 2000*n)")
    s.add_argument("-w", "--cwin", type=int, default=0, help="convergence window (default 2000 on the device)")
    s.add_argument("-c", "--chains", type=int, default=0, help="base chains (default 2)")
    s.add_argument("-a", "--chainadds", type=int, default=1)
    s.add_argument("-i", "--maxiters", type=int, default=-1, help="recorded updates (default 20000*n per replica)")
    s.add_argument("-x", "--maxsecs", type=int, default=300)
    s.add_argument("-p", "--experiment", action="store_true")
    s.add_argument("--replicas", type=int, default=1024, help="device chains behind each reference chain")
    s.add_argument("--precision", default="auto", choices=["auto", "f64", "f32", "table", "hybrid"],
                   help="auto: hybrid (exact float64 conditionals from threshold tables) when every sampled variable of the "
                        "model gets a table, else f32; the Rao-Blackwell estimator implies f32")
    s.add_argument("--device", type=int, default=0)
    s.add_argument("--rao-blackwell", action="store_true",
                   help="marginals accumulate the sampled conditionals instead of counts (f32 / f64; not the reference's estimator)")
    s.add_argument("--addr", default="", help="ip:port for the expvar-style monitor (cmd/monitor.go); empty = no HTTP server")
    return ap


def dump_params(args, burn, cw, base, max_iters, seed, dst):
    """startupParams.dump (cmd/root.go:97-112)"""
    dst.write("Verbose:                %s\n" % str(bool(args.verbose)).lower())
    dst.write("Model:                  %s\n" % args.model)
    dst.write("Apply Evidence:         %s\n" % str(bool(args.evidence)).lower())
    dst.write("Solution:               %s\n" % str(bool(args.solution)).lower())
    dst.write("Sampler:                %s\n" % args.sampler)
    dst.write("Burn In:                %12d\n" % burn)
    dst.write("Converge Win:           %12d\n" % cw)
    dst.write("Num Base Chain:         %12d\n" % base)
    dst.write("Chains Added per Adapt: %12d\n" % args.chainadds)
    dst.write("Max Iters:              %12d\n" % max_iters)
    dst.write("Max Secs:               %12d\n" % args.maxsecs)
    dst.write("Rnd Seed:               %12d\n" % seed)
    dst.write("Monitor Addr:           %s\n" % args.addr)
    dst.write("Experiment Mode:        %s\n" % str(bool(args.experiment)).lower())


class _Null:
    def write(self, *_):
        pass

    def close(self):
        pass


def _dist_env():
    """(dist module or None, rank, world, local device).  Under torchrun (WORLD_SIZE > 1) the replicas of every
    reference chain are sharded over one process per GPU (NCCL); rank 0 reports."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 0, 1, None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, dist.get_rank(), world, local


def sample(args, out=sys.stdout, monitor=None):
    start = time.time()
    dist, rank, world, local = _dist_env()
    if local is not None:
        args.device = local
    if rank != 0:
        out = _Null()
    mon = monitor if monitor is not None else Monitor()
    if getattr(args, "addr", "") and rank == 0:
        mon.start(args.addr)
    try:
        return _sample(args, out, mon, start, dist, rank, world)
    finally:
        mon.stop()
        if dist is not None and dist.is_initialized() and os.environ.get("GB_KEEP_PROCESS_GROUP") is None:
            dist.destroy_process_group()


def _agree(dist, device, *flags):
    """rank 0's view of time-dependent loop decisions, so every rank takes the same branch"""
    if dist is None:
        return flags
    import torch
    t = torch.tensor([int(f) for f in flags], dtype=torch.int32, device=f"cuda:{device}")
    dist.broadcast(t, src=0)
    return tuple(bool(x) for x in t.tolist())


def _sample(args, out, mon, start, dist=None, rank=0, world=1):
    from . import distributed as gbd
    out.write("Reading model from %s\n" % args.model)
    mod = core.Model.from_uai(args.model, use_evidence=args.evidence, device=args.device)
    if args.precision == "auto":
        # ONE default for every host of the boundary (this CLI, the C++ mirror grample.hpp, the Go shim): GB_HYBRID =
        # the reference's float64 arithmetic throughout — variables whose conditional can be tabulated are sampled from
        # float64-derived thresholds, the rest by the float64 log-sum-exp kernels.  float32 is an explicit opt-in
        # (--precision f32).  The Rao-Blackwell estimator works under it too (tabulated conditionals are read back from the thresholds).
        args.precision = "hybrid"
        out.write("Precision: %s\n" % args.precision)
    prec = {"f64": F64, "f32": F32, "table": TABLE, "hybrid": HYBRID}[args.precision]
    n, cards, fixed = mod.n_vars, mod.cards, mod.fixed
    offs = np.concatenate([[0], np.cumsum(cards)])
    out.write("Model has %d vars and %d functions\n" % (n, mod.n_funcs))
    sol = None
    if args.solution:
        sol_cards, sol = core.mar_load(args.model + ".MAR")
        uniform = np.concatenate([np.full(c, 1.0 / c) for c in cards])
        es = core.error_suite(cards, sol, uniform, fixed2=fixed)
        mon.score(es)
        error_report("START", es, False, out)
    if args.experiment and not args.trace:
        raise core.GrampleError("Experiment mode requires a trace file")
    # defaults derived from n (cmd/root.go:344-363)
    seed = args.seed if args.seed >= 1 else int(time.time_ns() % (1 << 31))
    if dist is not None:  # a time-derived seed must be the same on every rank
        import torch
        t = torch.tensor([seed], dtype=torch.int64, device=f"cuda:{args.device}")
        dist.broadcast(t, src=0)
        seed = int(t.item())
    n_free = len(mod.schedule()[0])
    burn = args.burnin if args.burnin >= 0 else 2000 * n
    cw = args.cwin if args.cwin > 0 else 2000
    base = max(2, args.chains if args.chains > 0 else 2)
    max_iters = args.maxiters if args.maxiters >= 0 else 20000 * n * args.replicas
    mon.set("Burn-In", burn)  # cmd/root.go:367-370
    mon.set("Convergence-Window", cw)
    mon.set("Max-Iterations", max_iters)
    mon.set("Max-Seconds", args.maxsecs)
    if args.sampler != "adaptive" and args.chainadds != 1:
        raise core.GrampleError("Sampler is not adaptive: ChainAdds=%d makes no sense" % args.chainadds)
    mon.set("Chain-Adds-For-Adaptive-Step", args.chainadds)
    per = (args.replicas + 7) // 8 * 8          # global chain ids per reference chain / variant
    shard_first, n_local = gbd.shard(args.replicas, world, rank)
    if n_local == 0:
        raise core.GrampleError("--replicas %d leaves rank %d without chains (need >= %d)" % (args.replicas, rank, 8 * world))

    out.write("Creating chains and performing burn-in (%d)\n" % burn)
    models = []
    for idx in range(base):
        out.write(" ... Chain %3d out of %3d\n" % (idx + 1, base))
        if args.sampler == "collapsed":
            m, v, marg = mod.collapse(-1, seed=seed + idx)
            out.write("        - Collaped variable %d\nMARGINAL: %s\n" % (v, marg.tolist()))
            models.append(m)
        else:
            models.append(mod)
        mon.add("Base-Chain-Count", 1)  # cmd/root.go:428-429
        mon.add("Total-Chain-Count", 1)
    chains = core.Chains(models[0], n_local, seed=seed, first_chain_id=shard_first, precision=prec, history=True, device=args.device,
                         rao_blackwell=args.rao_blackwell)
    for idx in range(1, base):
        chains.add_group(models[idx], n_local, idx * per + shard_first)
    gbd.attach(chains, dist)  # multi-GPU: MergeChains / ChainConvergence / Adapt reduce inside the library (its own NCCL communicator)
    chains.burnin((burn + n_free - 1) // max(n_free, 1))
    next_id = base * per
    trace = open(args.trace, "w") if (args.trace and rank == 0) else (_Null() if args.trace else None)
    if args.experiment:
        trace.write("// EXPERIMENT RESULTS\nRunSecs, MaxHell, NegLogMaxHell, MaxJS, NegLogMaxJS, CollapseCount\n")
    nl = lambda x: -math.log2(x) if x > 0 else float("inf")

    out.write("Main Sampling Start\n")
    stop, next_status = start + args.maxsecs, start + 2.5
    keep_adapting, no_adapt = True, start + args.maxsecs / 2
    working = True
    while working:  # cmd/root.go:475-561
        chains.advance(cw)
        chains.synchronize()
        now = time.time()
        if args.maxsecs > 0 and now > stop:
            working = False
        count = gbd.total_samples(chains, dist)
        mon.set("Iterations", count)  # cmd/root.go:492
        mon.set("Total-Samples", count)
        if max_iters > 0 and count > max_iters:
            working = False
        status = now > next_status
        working, status = _agree(dist, args.device, working, status)
        if status or not working or args.experiment:
            if status or not working:
                mon.set("Run-Time", now - start)  # cmd/root.go:502
                out.write("  Samps: %12d | RT %12.2fsec\n" % (count, now - start))
            if sol is not None:
                merged, col = gbd.merged_marginals(chains, dist)
                score = core.error_suite(cards, sol, merged, fixed2=fixed)
                if status or not working:
                    mon.score(score)
                    error_report("", score, True, out)
                if args.experiment:
                    trace.write("%.1f, %.8f, %.5f, %.8f, %.5f, %d\n" % (now - start, score["MaxHellinger"], nl(score["MaxHellinger"]),
                                                                       score["MaxJSDiverge"], nl(score["MaxJSDiverge"]), int(col.sum())))
            if status or not working:
                next_status = now + 5
        (stop_adapt,) = _agree(dist, args.device, keep_adapting and now > no_adapt)
        if stop_adapt:
            out.write("STOPPING ADAPTATION\n")
            keep_adapting = False
        if working and keep_adapting and args.sampler == "adaptive":
            pre = chains.n_groups
            chosen, stride = gbd.adapt(chains, mod, args.chainadds, args.replicas, cw, next_id, HELLINGER, dist)
            next_id += len(chosen) * stride
            if chains.n_groups != pre:
                mon.set("Total-Chain-Count", chains.n_groups)  # cmd/root.go:557
                out.write("ADAPT: %d Chains (was %d)\n" % (chains.n_groups, pre))

    run_time = time.time() - start
    merged, col = gbd.merged_marginals(chains, dist)  # cmd/root.go:565-571
    final = merged.copy()
    for v in range(n):
        final[offs[v]:offs[v + 1]] /= final[offs[v]:offs[v + 1]].sum()
    out.write("DONE\n")
    state = [dict() for _ in range(n)]  # per-variable Variable.State (cmd/root.go:600-657)
    merlin = None
    if sol is not None:
        score = core.error_suite(cards, sol, final, fixed2=fixed)
        mon.score(score)
        error_report("FINAL", score, False, out)
        if args.experiment:  # cmd/root.go:582-597
            trace.write("%.1f, %.8f, %.5f, %.8f, %.5f, %d\n" % (run_time, score["MaxHellinger"], nl(score["MaxHellinger"]),
                                                               score["MaxJSDiverge"], nl(score["MaxJSDiverge"]), int(col.sum())))
            trace.write("// FINAL STATUS\n")
            error_report("FINAL", score, False, trace)
        for v in range(n):
            for c in range(cards[v]):
                state[v]["SOL-MAR[%d]" % c] = sol[offs[v] + c]
        merlin_path = args.model + ".merlin.MAR"  # cmd/root.go:609-635
        if os.path.exists(merlin_path):
            _, merlin = core.mar_load(merlin_path)
            ms = core.error_suite(cards, sol, merlin)
            error_report("MERLIN SCORE", ms, False, out)
            if args.experiment:
                trace.write("// MERLIN SCORES\n")
                error_report("MERLIN SCORE", ms, False, trace)
            error_report("OUR SCORE USING MERLIN AS SOLUTION", core.error_suite(cards, merlin, final, fixed2=fixed), False, out)
    # final convergence under all four measures and per-variable errors (cmd/root.go:638-657)
    for key, measure in (("Hell", HELLINGER), ("JS", JS), ("MaxAD", MAX_ABS), ("AvgAD", MEAN_ABS)):
        conv = gbd.convergence(chains, measure, merged, col, cw, dist)
        for v in range(n):
            state[v][key + "-Convergence"] = conv[v]
    if sol is not None:
        errs = per_var_measures(cards, final, sol, fixed_p=fixed)
        for key, arr in errs.items():
            for v in range(n):
                state[v][key + "-Error"] = arr[v]

    def line(v):
        return variable_json(v, cards[v], fixed[v], final[offs[v]:offs[v + 1]], state[v], col[v])

    if trace or args.verbose:  # cmd/root.go:659-710
        if trace:
            trace.write("// EVIDENCE\n")
        for v in range(n):
            if fixed[v] >= 0:
                if trace:
                    trace.write(line(v) + "\n")
                if args.verbose:
                    out.write("Variable[%d] %s (Card:%d, %s) EVID=%d\n" % (v, letter26(v), cards[v], state[v], fixed[v]))
        if trace:
            trace.write("// VARS (ESTIMATED)\n")
        for v in range(n):
            if fixed[v] < 0:
                if trace:
                    trace.write(line(v) + "\n")
                if args.verbose:
                    out.write("Variable[%d] %s (Card:%d, %s) %s\n" % (v, letter26(v), cards[v], state[v], final[offs[v]:offs[v + 1]].tolist()))
        if merlin is not None:
            mh = per_var_measures(cards, final, merlin, fixed_p=fixed)["Hell"]
            report = [v for v in range(n) if fixed[v] < 0]
            for v in report:
                state[v]["MerlinHellError"] = mh[v]
            report.sort(key=lambda v: state[v]["MerlinHellError"])
            if trace:
                trace.write("// VARS SORTED BY DIST FROM HELLINGER\n")
                for v in report:
                    trace.write(line(v) + "\n")
        if trace:
            trace.write("// OPERATING PARAMS\n")
            dump_params(args, burn, cw, base, max_iters, seed, trace)
            trace.write("// MONITOR\n" + json.dumps(mon.snapshot()) + "\n")
            # cmd/root.go:715-717: json of model.Model (Funcs are `json:"-"`): Type, Name, Vars
            trace.write("// ENTIRE MODEL\n")
            with open(args.model) as f:  # model type = first field of the UAI file (model/uai.go:73-80)
                mtype = next((ln.split()[0] for ln in f if ln.strip() and not ln.startswith("c")), "MARKOV")
            trace.write(json.dumps({"Type": mtype, "Name": os.path.splitext(args.model)[0],  # model/model.go:65
                                    "Vars": [json.loads(line(v)) for v in range(n)]}, indent=2) + "\n")
    if trace:
        trace.close()
    return final, col


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.cmd == "sample":
        try:
            sample(args)
        except core.GrampleError as e:
            sys.stderr.write("error: %s\n" % e)
            return 1
    return 0
