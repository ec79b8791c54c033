#!/usr/bin/env python
"""Kernel-only A/B harness: extension-kernel time (CUDA events, inputs packed and resident in HBM, longest first) on slices
of the four synthetic workloads of BASELINE.md section 2.3. One JSON line per workload; used while tuning kernels -- the
numbers that are reported come from bench.py.

    python tools/kperf.py [--configs C1,C2,C3,C4] [--scale 1.0] [--reps 3] [--check]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {          # name: (profile, seed, pairs, band)
    "C1": (1, 1, 8192, 751),
    "C2": (2, 2, 16384, 751),
    "C3": (3, 3, 1024, 4095),
    "C4": (4, 4, 32768, 751),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare a 48-pair sample with the oracle")
    ap.add_argument("--tag", default="")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per workload (overrides the slice sizes)")
    args = ap.parse_args()
    import torch
    import agatha_b200 as ag
    dev = torch.device("cuda:0")
    for name in args.configs.split(","):
        prof, seed, n, W = CONFIGS[name]
        n = args.pairs or max(64, int(n * args.scale))
        dd = ag.synth_pairs(prof, seed, n)
        stq, qoff, qlen = ag.stage_batch(dd["qbuf"], dd["qoff"], dd["qlen"])
        stt, toff, tlen = ag.stage_batch(dd["tbuf"], dd["toff"], dd["tlen"])
        tq = torch.from_numpy(stq).to(dev); tt = torch.from_numpy(stt).to(dev)
        qp, tp = ag.pack_device(tq, tt)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
        order = d(ag.bucket_order(qlen, tlen, W))
        p = ag.make_params(band_width=W)
        a = (qp, tp, d(qoff), d(toff), d(qlen), d(tlen), p)
        out = ag.extend_device(*a, order=order)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ag.extend_device(*a, order=order, out=out); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        _, cells = ag.count_cells(qlen, tlen, W, out["dstop"].cpu().numpy())
        best = min(ms)
        line = {"tag": args.tag, "config": name, "pairs": n, "band": W, "kernel_ms": round(best, 3), "all_ms": [round(x, 3) for x in ms],
                "alignments_per_s": round(n / best * 1e3), "gcups": round(cells / best / 1e6, 1), "cells": int(cells)}
        if args.check:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle import oracle_py as op
            orc = op.Oracle()
            idx = np.random.default_rng(1).choice(n, min(n, 48), replace=False)
            sub = [(dd["qbuf"][int(dd["qoff"][i]):int(dd["qoff"][i]) + int(qlen[i])], dd["tbuf"][int(dd["toff"][i]):int(dd["toff"][i]) + int(tlen[i])]) for i in idx]
            exp = orc.align_pairs(sub, op.make_params(band_width=W))
            got = {k: out[k].cpu().numpy()[idx] for k in ("score", "query_end", "target_end", "stop", "dstop")}
            line["oracle_sample_ok"] = bool(all((got[k] == exp[e]).all() for k, e in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop"))))
        print(json.dumps(line), flush=True)
        del tq, tt, qp, tp


if __name__ == "__main__":
    main()
