#!/usr/bin/env python
"""Measure every workload of BASELINE.md section 2.3 (C1..C4) on one GPU: this engine through the C ABI (host buffers,
e2e), GCUPS from the needed cells, and the unmodified reference GPU program (oracle/_ref/agatha_ref_manual) on a slice
inside its valid domain, with a result comparison. Writes profiles/configs_rNN.json.

    python tools/run_configs.py [--out profiles/configs_r01.json] [--pairs 100000]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import agatha_b200 as ag                     # noqa: E402
from oracle import oracle_py as op           # noqa: E402

CONFIGS = [
    # name, profile, seed, pairs, band, reference slice
    ("C1_bundled_standin", 1, 1, 8192, 751, 8192),
    ("C2_ont_like", 2, 2, None, 751, 16384),
    ("C3_hifi_like_w4095", 3, 3, None, 4095, 2048),
    ("C4_heavy_tail_zdrop", 4, 4, None, 751, 16384),
]


def subset(d, idx):
    qs = [d["qbuf"][int(d["qoff"][i]):int(d["qoff"][i]) + int(d["qlen"][i])] for i in idx]
    ts = [d["tbuf"][int(d["toff"][i]):int(d["toff"][i]) + int(d["tlen"][i])] for i in idx]
    ql = d["qlen"][idx]; tl = d["tlen"][idx]
    qo = np.concatenate([[0], np.cumsum(ql[:-1], dtype=np.uint64)]).astype(np.uint64)
    to = np.concatenate([[0], np.cumsum(tl[:-1], dtype=np.uint64)]).astype(np.uint64)
    return dict(qbuf=np.concatenate(qs), tbuf=np.concatenate(ts), qoff=qo, toff=to, qlen=ql, tlen=tl)


def run_ref(d, W, tmp):
    if not os.path.exists(op.REF_GPU_BIN):
        return None, {"unavailable": "oracle/_ref/agatha_ref_manual not built"}
    qf, tf = os.path.join(tmp, "q.fa"), os.path.join(tmp, "t.fa")
    ag.write_fasta(qf, d["qbuf"], d["qoff"], d["qlen"]); ag.write_fasta(tf, d["tbuf"], d["toff"], d["tlen"])
    raw, score = os.path.join(tmp, "raw.log"), os.path.join(tmp, "score.log")
    if os.path.exists(raw):
        os.remove(raw)
    t0 = time.time()
    with open(score, "w") as so:
        r = subprocess.run([op.REF_GPU_BIN, "-p", "-m", "1", "-x", "4", "-q", "6", "-r", "2", "-s", "3", "-z", "400", "-w", str(W), qf, tf, raw],
                           stdout=so, stderr=subprocess.PIPE, text=True, timeout=1800)
    if r.returncode != 0:
        return None, {"error": r.stderr[-300:]}
    ms = sum(float(x) for x in open(raw).read().split())
    rows = [ln.split("\t") for ln in open(score).read().splitlines()]
    res = np.array([[int(a), int(b.split("=")[1]), int(c.split("=")[1])] for a, b, c in rows])
    return res, {"pairs": len(rows), "kernel_ms": ms, "alignments_per_s": len(rows) / (ms * 1e-3), "wall_s": time.time() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "configs_r01.json"))
    ap.add_argument("--pairs", type=int, default=100000)
    ap.add_argument("--only", default="")
    ap.add_argument("--ref-pairs", type=int, default=0, help="override the size of the slice given to the reference GPU program")
    args = ap.parse_args()
    out = {}
    for name, prof, seed, n, W, nref in CONFIGS:
        if args.only and args.only not in name:
            continue
        n = n or args.pairs
        d = ag.synth_pairs(prof, seed, n)
        p = ag.make_params(band_width=W)
        ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=[0])
        t0 = time.time()
        res, st = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=[0])
        dt = time.time() - t0
        _, cells = ag.count_cells(d["qlen"], d["tlen"], W, res["dstop"])
        _, full = ag.count_cells(d["qlen"], d["tlen"], W)
        e = {"pairs": n, "band_width": W, "mean_len": float(d["tlen"].mean()), "max_len": int(max(d["tlen"].max(), d["qlen"].max())),
             "e2e_seconds": dt, "alignments_per_s": n / dt, "gcups_needed_cells": cells / dt / 1e9, "needed_cells": cells, "full_band_cells": full,
             "stops": {"end": int((res["stop"] == 0).sum()), "zdrop": int((res["stop"] == 1).sum()), "bandexit": int((res["stop"] == 2).sum())},
             "h2d_bytes": int(st["h2d_bytes"]), "batches": int(st["n_batches"])}
        # reference GPU program on a slice inside its int16 domain
        ok = np.nonzero((d["qlen"] < 32768) & (d["tlen"] < 32768))[0][:(args.ref_pairs or nref)]
        sub = subset(d, ok)
        with tempfile.TemporaryDirectory() as tmp:
            ref_res, info = run_ref(sub, W, tmp)
        if ref_res is not None:
            mine = np.stack([res["score"][ok], res["query_end"][ok], res["target_end"][ok]], 1)
            same = (mine == ref_res).all(axis=1)
            info["identical_results"] = int(same.sum())
            # pairs whose score would overflow the reference's int16 are outside its domain
            info["outside_int16_domain"] = int((res["score"][ok] > 32767).sum())
            t1 = time.time()
            r2, _ = ag.align_job(sub["qbuf"], sub["qoff"], sub["qlen"], sub["tbuf"], sub["toff"], sub["tlen"], p, devices=[0])
            t1 = time.time()
            r2, s2 = ag.align_job(sub["qbuf"], sub["qoff"], sub["qlen"], sub["tbuf"], sub["toff"], sub["tlen"], p, devices=[0])
            info["ours_same_slice_alignments_per_s"] = len(ok) / (time.time() - t1)
            info["speedup_vs_reference_kernel"] = info["ours_same_slice_alignments_per_s"] / info["alignments_per_s"]
        e["reference_gpu"] = info
        out[name] = e
        print(name, json.dumps(e), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
