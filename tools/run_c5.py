#!/usr/bin/env python
"""BASELINE config 5: 1,000,000 ONT-like pairs (seed 5) aligned by ONE process through agatha_align_job, sharded over
1, 2, 4 and 8 GPUs of the box by the host scheduler (strong scaling of a fixed job; no collective). Checks that every
device count returns identical results, spot-checks the oracle, and times the oracle on a bounded sample on the host cores.

    python tools/run_c5.py [--pairs 1000000] [--out profiles/c5_r01.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import agatha_b200 as ag                     # noqa: E402
from oracle import oracle_py as op           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "c5_r01.json"))
    args = ap.parse_args()
    ndev = ag.device_count()
    t0 = time.time()
    d = ag.synth_pairs(2, 5, args.pairs)
    gen_s = time.time() - t0
    p = ag.make_params()
    out = {"pairs": args.pairs, "generate_seconds": gen_s, "host_cores": os.cpu_count(), "devices_visible": ndev, "runs": {}}
    ref = None
    for k in (1, 2, 4, 8):
        if k > ndev:
            break
        devs = list(range(k))
        ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=devs)      # warm-up: allocations
        best = None
        for _ in range(2):
            t0 = time.time()
            res, st = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=devs)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
        if ref is None:
            ref = res
            _, cells = ag.count_cells(d["qlen"], d["tlen"], p.band_width, res["dstop"])
            out["needed_cells"] = cells
        same = bool((res == ref).all())
        out["runs"][str(k)] = {"seconds": best, "alignments_per_s": args.pairs / best, "gcups": out["needed_cells"] / best / 1e9,
                               "identical_to_1gpu": same, "batches": int(st["n_batches"]), "h2d_bytes": int(st["h2d_bytes"])}
        print(k, out["runs"][str(k)], flush=True)
    base = out["runs"]["1"]["alignments_per_s"]
    for k, v in out["runs"].items():
        v["speedup_vs_1gpu"] = v["alignments_per_s"] / base
    # oracle: spot check + CPU baseline on a bounded sample
    op.build(ref=False)
    orc = op.Oracle()
    idx = np.random.default_rng(0).choice(args.pairs, 256, replace=False)
    pairs = [(d["qbuf"][int(d["qoff"][i]):int(d["qoff"][i]) + int(d["qlen"][i])], d["tbuf"][int(d["toff"][i]):int(d["toff"][i]) + int(d["tlen"][i])]) for i in idx]
    t0 = time.time()
    exp = orc.align_pairs(pairs, op.make_params())
    cpu_s = time.time() - t0
    ok = all((ref[a][idx] == exp[b]).all() for a, b in (("score", "score"), ("query_end", "query_end"), ("target_end", "target_end"), ("stop", "stop"), ("dstop", "d_stop")))
    out["oracle_spot_check"] = {"pairs": 256, "identical": bool(ok)}
    out["cpu_baseline"] = {"alignments_per_s": 256 / cpu_s, "cores": os.cpu_count(), "kind": "port", "sample": "256 random pairs, %.1f s" % cpu_s}
    print(json.dumps(out["oracle_spot_check"]), json.dumps(out["cpu_baseline"]))
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
