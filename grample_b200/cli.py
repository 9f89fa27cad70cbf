"""`python -m grample_b200 sample ...` — the `grample sample` command of the reference
(cmd/root.go:163-250 flags, :309-719 main loop) driving the device path.

Same flags and defaults as the reference where they make sense on the device; new knobs are
`--replicas` (device chains behind each reference chain), `--precision` and `--device`.
Status / error-report lines keep the reference's format (cmd/root.go:256-306, 498-504) and
`--experiment` writes the CSV the reference's script/trace_file_process.py parses
(cmd/root.go:455-458, 520-533).
"""
import argparse
import math
import sys
import time

import numpy as np

from . import core
from ._lib import F32, F64, HELLINGER, TABLE


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
    s.add_argument("--precision", default="f32", choices=["f64", "f32", "table"])
    s.add_argument("--device", type=int, default=0)
    return ap


def sample(args, out=sys.stdout):
    start = time.time()
    prec = {"f64": F64, "f32": F32, "table": TABLE}[args.precision]
    out.write("Reading model from %s\n" % args.model)
    mod = core.Model.from_uai(args.model, use_evidence=args.evidence, device=args.device)
    n, cards, fixed = mod.n_vars, mod.cards, mod.fixed
    out.write("Model has %d vars and %d functions\n" % (n, mod.n_funcs))
    sol = None
    if args.solution:
        sol_cards, sol = core.mar_load(args.model + ".MAR")
        uniform = np.concatenate([np.full(c, 1.0 / c) for c in cards])
        error_report("START", core.error_suite(cards, sol, uniform, fixed2=fixed), False, out)
    if args.experiment and not args.trace:
        raise core.GrampleError("Experiment mode requires a trace file")
    # defaults derived from n (cmd/root.go:344-363)
    seed = args.seed if args.seed >= 1 else int(time.time_ns() % (1 << 31))
    n_free = len(mod.schedule()[0])
    burn = args.burnin if args.burnin >= 0 else 2000 * n
    cw = args.cwin if args.cwin > 0 else 2000
    base = max(2, args.chains if args.chains > 0 else 2)
    max_iters = args.maxiters if args.maxiters >= 0 else 20000 * n * args.replicas
    if args.sampler != "adaptive" and args.chainadds != 1:
        raise core.GrampleError("Sampler is not adaptive: ChainAdds=%d makes no sense" % args.chainadds)
    per = (args.replicas + 7) // 8 * 8

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
    chains = core.Chains(models, [args.replicas] * base, seed=seed, precision=prec, history=True, device=args.device)
    chains.burnin((burn + n_free - 1) // max(n_free, 1))
    next_id = base * per
    trace = open(args.trace, "w") if args.trace else None
    if args.experiment:
        trace.write("// EXPERIMENT RESULTS\nRunSecs, MaxHell, NegLogMaxHell, MaxJS, NegLogMaxJS, CollapseCount\n")

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
        count = chains.total_samples
        if max_iters > 0 and count > max_iters:
            working = False
        if now > next_status or not working or args.experiment:
            if now > next_status or not working:
                out.write("  Samps: %12d | RT %12.2fsec\n" % (count, now - start))
            if sol is not None:
                merged, col = chains.merged_marginals()
                score = core.error_suite(cards, sol, merged, fixed2=fixed)
                if now > next_status or not working:
                    error_report("", score, True, out)
                if args.experiment:
                    nl = lambda x: -math.log2(x) if x > 0 else float("inf")
                    trace.write("%.1f, %.8f, %.5f, %.8f, %.5f, %d\n" % (now - start, score["MaxHellinger"], nl(score["MaxHellinger"]),
                                                                       score["MaxJSDiverge"], nl(score["MaxJSDiverge"]), int(col.sum())))
            if now > next_status or not working:
                next_status = now + 5
        if keep_adapting and now > no_adapt:
            out.write("STOPPING ADAPTATION\n")
            keep_adapting = False
        if working and keep_adapting and args.sampler == "adaptive":
            pre = chains.n_groups
            chosen = chains.adapt(mod, args.chainadds, args.replicas, cw, first_chain_id=next_id, measure=HELLINGER)
            next_id += len(chosen) * per
            if chains.n_groups != pre:
                out.write("ADAPT: %d Chains (was %d)\n" % (chains.n_groups, pre))

    merged, col = chains.merged_marginals()  # cmd/root.go:565-571
    offs = np.concatenate([[0], np.cumsum(cards)])
    final = merged.copy()
    for v in range(n):
        final[offs[v]:offs[v + 1]] /= final[offs[v]:offs[v + 1]].sum()
    out.write("DONE\n")
    if sol is not None:
        error_report("FINAL", core.error_suite(cards, sol, final, fixed2=fixed), False, out)
    conv = chains.convergence(HELLINGER, merged)
    if args.verbose or trace:
        dst = trace if trace else out
        dst.write("// VARS (ESTIMATED)\n")
        for v in range(n):
            if fixed[v] < 0:
                dst.write('{"ID":%d,"Card":%d,"Marginal":%s,"Collapsed":%s,"Hell-Convergence":%.6f}\n' % (
                    v, cards[v], final[offs[v]:offs[v + 1]].tolist(), "true" if col[v] else "false", conv[v]))
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
