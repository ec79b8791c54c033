#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native guided-alignment engine.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                          (the reference arm: CPU, host cores, rank 0 only)

Workloads (BASELINE.json `configs`, generators in tools/synth/synth.cpp, default AGAThA.sh scoring -m 1 -x 4 -q 6 -r 2 -s 3 -z 400):
  N = 1   C2 = configs[1]: 100,000 synthetic ONT-like read/reference pairs (~10 kb, ~10 % error), -w 751.
  N > 1   C5 = configs[4]: the fixed 1,000,000-pair ONT-like set, rank r aligns pairs [r*1M/N, (r+1)*1M/N) -- STRONG scaling,
          pairs are independent so there is no collective on the data path; results are gathered to rank 0 inside the
          end-to-end region. Rank 0 then runs a small job through the library's own multi-device scheduler
          (agatha_align_job over all N devices) and compares it with single-device results.
A "step" is one pass of the hot path (banded extension with Z-drop, packed kernel + redo pass) over the whole set.

  value     alignments/s, inputs packed and resident in HBM, CUDA-event time, max over ranks
  e2e       the same through the C ABI agatha_align_job() with HOST buffers: host-side packing into pinned staging, H2D,
            kernels, D2H (and the gather at N > 1) inside the timed region
  roofline  integer-ALU bound: achieved GCUPS against the ceiling of the committed steady-state loop,
            ceiling = ALU-pipe lane-op rate MEASURED IN THIS RUN (agatha_measure_int_peak) / ALU instructions per lane-cell
            counted in the SASS of the shipped library (tools/sass_hot_loop.py); frac <= 1 by construction
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "alignments/s"
SCORING = "-m 1 -x 4 -q 6 -r 2 -s 3 -z 400"
WORKLOADS = {   # name: (profile, seed, pairs, band, description)
    "C1": (1, 1, 8192, 751, "C1: stand-in for the bundled dataset, 8,192 pairs of 1-8 kb, 5-15 % error, 1/3 with a random tail, -w 751"),
    "C2": (2, 2, 100000, 751, "C2: 100k synthetic ONT-like pairs (~10 kb lognormal, 4/3/3 % sub/ins/del), " + SCORING + " -w 751"),
    "C3": (3, 3, 100000, 4095, "C3: HiFi-like pairs (~15 kb, 0.4/0.3/0.3 % sub/ins/del), -w 4095"),
    "C4": (4, 4, 100000, 751, "C4: heavy-tailed lengths (1-100 kb), 10 % error, half of the reads turn random (early Z-drop), -w 751"),
    "C5": (2, 5, 1000000, 751, "C5: 1M synthetic ONT-like pairs sharded over the GPUs (strong scaling), " + SCORING + " -w 751"),
}
EXTRA_LEGS = {"C1": 8192, "C3": 8192, "C4": 100000}    # kernel-only legs at N = 1 (pairs per leg: a few seconds in total)


def load_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def hot_loop_stats(C, NW, JWS):
    """ALU-pipe instructions per lane-cell of the steady-state loop of the kernel variant that ran: counted live in the SASS of
    the library this process loaded (cuobjdump is in the image), else the committed count under profiles/."""
    name = "extend16_c%d%s_hot_loop_r02" % (C, "" if NW == 1 else "x%d" % NW)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_hot_loop as sh
        from agatha_b200._lib import lib_path
        ins = sh.disassemble(lib_path(), sh.kernel_symbol(C, NW, JWS))
        res = sh.analyse(ins, C) if ins else None
        if res:
            info = res[0]
            info.pop("opcode_mix", None)
            info["source"] = "cuobjdump -sass of the loaded library, this run"
            return info
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", name + ".json")) as f:
            info = json.load(f)
        info.pop("opcode_mix", None)
        info["source"] = "profiles/%s.json (committed)" % name
        return info
    except Exception:
        return None


def load_ncu(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[1])); smax = max(smax, float(p[2]))
                for nme, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place where bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------------------------
def cpu_leg(oracle_mod, data, params, budget_s, use_ref_host, cores):
    """Time the CPU checker on a bounded prefix of the workload with `cores` threads (set explicitly: torchrun exports
    OMP_NUM_THREADS=1). Returns (info, pairs, seconds)."""
    n_total = len(data["qlen"])
    ref = oracle_mod.RefHost() if (use_ref_host and oracle_mod.RefHost.available()) else None
    orc = None if ref is not None else oracle_mod.Oracle()
    p = oracle_mod.make_params(**params)

    def run(n):
        qoff = data["qoff"][:n].astype(np.uint32); toff = data["toff"][:n].astype(np.uint32)
        t0 = time.time()
        (ref or orc).align_batch(data["qbuf"], qoff, data["qlen"][:n], data["tbuf"], toff, data["tlen"][:n], p, nthreads=cores)
        return time.time() - t0
    probe_n = min(n_total, max(2 * cores, 8))
    t_probe = run(probe_n)
    n = int(min(n_total, max(probe_n, probe_n * budget_s / max(t_probe, 1e-3))))
    n = max(cores, (n // cores) * cores)
    t = run(n) if n != probe_n else t_probe
    kind = "reference" if ref is not None else "port"
    what = ("reference agatha_kernel.h compiled as host code (oracle/_ref), OpenMP over pairs" if ref is not None
            else "scalar C oracle (oracle/agatha_oracle.c), OpenMP over pairs")
    return {"value": n / t, "unit": METRIC, "cores": cores, "kind": kind, "sample": "first %d pairs of the workload, %.1f s, %s" % (n, t, what)}, n, t


def reference_arm(args, rank):
    """--impl reference: the reference's own kernel code as host code (oracle/_ref) on all host cores, bounded sample per step.
    Loads the workload generator and oracle/ only -- not the product library."""
    if rank != 0:
        return 0
    from agatha_b200.host_api import synth_pairs
    from agatha_b200._lib import DEFAULT_PARAMS
    from oracle import oracle_py as op
    op.build(ref=False)
    params = dict(DEFAULT_PARAMS)
    name = "C2" if args.gpus == 1 else "C5"
    prof, seed, _, W, desc = WORKLOADS[name]
    params["band_width"] = W
    cores = os.cpu_count() or 1
    data = synth_pairs(prof, seed, min(args.pairs or 4096, 4096))
    times, counts, info = [], [], None
    for it in range(args.warmup + args.steps):
        info, n, t = cpu_leg(op, data, params, min(args.cpu_budget, 12.0) if it >= args.warmup else 1.0, True, cores)
        if it >= args.warmup:
            times.append(t); counts.append(n)
    value = sum(counts) / sum(times)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "note": "CPU arm: the reference has no CPU implementation; its kernel header is compiled as host code (oracle/_ref) and "
                                                 "run on all %d host threads on a bounded sample per step; the host does not get faster with more GPUs" % cores},
            "cpu_baseline": info, "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------------------
# reference GPU program (extra leg at N = 1; not one of the contract's arms)
# --------------------------------------------------------------------------------------------------------------------
def run_reference_gpu(ag, data, params, n_pairs, tmpdir):
    """The UNMODIFIED reference GPU program built for sm_100 (oracle/_ref/agatha_ref_manual) on a slice of the same pairs, same
    box. Time = sum of raw.log (its own -p timing: bucketing round trip + agatha_kernel per batch, gasal_align.cu:219-236).
    Launch shapes: the defaults and a sweep with one job per 8-lane subwarp (a = b*t/8); the best one counts."""
    from oracle import oracle_py as op
    if not os.path.exists(op.REF_GPU_BIN):
        return {"unavailable": "oracle/_ref/agatha_ref_manual not built"}
    n = min(n_pairs, len(data["qlen"]))
    qf, tf = os.path.join(tmpdir, "ref_q.fasta"), os.path.join(tmpdir, "ref_t.fasta")
    ag.write_fasta(qf, data["qbuf"], data["qoff"], data["qlen"][:n])
    ag.write_fasta(tf, data["tbuf"], data["toff"], data["tlen"][:n])
    out = {"runs": {}}
    flags = ["-m", str(params["match"]), "-x", str(params["mismatch"]), "-q", str(params["gap_open"]), "-r", str(params["gap_extend"]),
             "-s", str(params["slice_width"]), "-z", str(params["z_threshold"]), "-w", str(params["band_width"])]
    shapes = [("default_b256_t256_a8192", [])]
    for b, t in ((148, 256), (296, 256), (444, 256), (592, 256), (296, 128), (592, 128), (148, 512)):
        a = b * t // 8
        if a <= 32767:
            shapes.append(("b%d_t%d_a%d" % (b, t, a), ["-b", str(b), "-t", str(t), "-a", str(a)]))
    best = None
    for name, extra in shapes:
        raw = os.path.join(tmpdir, "raw_%s.log" % name)
        score = os.path.join(tmpdir, "score_%s.log" % name)
        t0 = time.time()
        try:
            with open(score, "w") as so:
                r = subprocess.run([op.REF_GPU_BIN, "-p"] + flags + extra + [qf, tf, raw], stdout=so, stderr=subprocess.PIPE, text=True, timeout=300)
            if r.returncode != 0:
                out["runs"][name] = {"error": (r.stderr or "")[-200:]}
                continue
            ms = sum(float(x) for x in open(raw).read().split())
            out["runs"][name] = {"kernel_ms": ms, "alignments_per_s": n / (ms * 1e-3), "wall_s": round(time.time() - t0, 2)}
            if best is None or ms < best[1]:
                best = (name, ms, score)
        except Exception as e:  # noqa: BLE001
            out["runs"][name] = {"error": repr(e)[:200]}
    out["pairs"] = n
    if best:
        out["best"] = {"shape": best[0], "kernel_ms": best[1], "alignments_per_s": n / (best[1] * 1e-3)}
        out["_score_log"] = best[2]
    return out


def parse_ref_scores(path):
    sc = np.loadtxt(path, dtype=str, delimiter="\t")
    return (sc[:, 0].astype(np.int64), np.array([int(x.split("=")[1]) for x in sc[:, 1]]), np.array([int(x.split("=")[1]) for x in sc[:, 2]]))


# --------------------------------------------------------------------------------------------------------------------
# device-resident leg
# --------------------------------------------------------------------------------------------------------------------
class DeviceLeg:
    """One workload, packed on the host and resident in HBM; step() = the extension kernels (packed kernel + redo pass) over
    the whole set. Offsets are 32-bit counts of bases (the reference's batch layout), so a set with more than ~3.5 G bases per
    side is kept as several chunks and a step launches them back to back."""

    def __init__(self, ag, torch, dev, data, params):
        self.ag, self.torch, self.dev = ag, torch, dev
        W = params["band_width"]
        qlen, tlen = data["qlen"], data["tlen"]
        self.n = len(qlen)
        from agatha_b200._lib import check, lib
        self.check, self.L = check, lib()
        self.stream = torch.cuda.current_stream(dev)
        self.p = ag.make_params(**params)
        d32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
        pad = np.zeros(64, np.uint32)
        big = np.maximum(qlen, tlen).astype(np.int64) + 8
        bounds, acc, start = [], 0, 0
        for i in range(self.n):
            if acc + big[i] > 3_500_000_000:
                bounds.append((start, i)); start, acc = i, 0
            acc += big[i]
        bounds.append((start, self.n))
        self.chunks = []
        self.packed_bytes = 0
        for lo, hi in bounds:
            ids = np.arange(lo, hi, dtype=np.uint64)
            qw, qoff, ql = ag.pack_batch(data["qbuf"], data["qoff"], qlen, False, ids=ids, n_threads=8)
            tw, toff, tl = ag.pack_batch(data["tbuf"], data["toff"], tlen, True, ids=ids, n_threads=8)
            self.packed_bytes += 4 * (len(qw) + len(tw))
            order = ag.bucket_order(ql, tl, W)
            m = hi - lo
            self.chunks.append({"n": m, "qp": d32(np.concatenate([qw, pad])), "tp": d32(np.concatenate([tw, pad])),
                                "meta": [d32(x) for x in (qoff, toff, ql, tl, order)],
                                "out": {k: torch.empty(m, dtype=torch.int32, device=dev) for k in ("score", "query_end", "target_end", "stop", "dstop")},
                                "ws": torch.zeros(256, dtype=torch.uint8, device=dev)})

    def step(self):
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        for c in self.chunks:
            m, o = c["meta"], c["out"]
            self.check(self.L.agatha_extend_device(vp(c["qp"]), vp(c["tp"]), vp(m[0]), vp(m[1]), vp(m[2]), vp(m[3]), vp(m[4]), ctypes.c_uint32(c["n"]),
                                                   ctypes.byref(self.p), vp(o["score"]), vp(o["query_end"]), vp(o["target_end"]),
                                                   vp(o["stop"]), vp(o["dstop"]), vp(c["ws"]), ctypes.c_void_p(self.stream.cuda_stream)))

    def timed(self, steps, warmup, barrier=None):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        (barrier or (lambda: torch.cuda.synchronize(self.dev)))()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(self.stream)
        for a, b in ev:
            a.record(self.stream); self.step(); b.record(self.stream)
        e1.record(self.stream)
        (barrier or (lambda: torch.cuda.synchronize(self.dev)))()
        t1 = time.time()
        return e0.elapsed_time(e1), [a.elapsed_time(b) for a, b in ev], t0, t1

    def results(self):
        return {k: np.concatenate([c["out"][k].cpu().numpy() for c in self.chunks]) for k in ("score", "query_end", "target_end", "stop", "dstop")}


def roofline_for(ag, leg, res, kernel_ms, W, int_peak, qlen, tlen, ncu_name=None):
    """GCUPS against the ALU-pipe ceiling of the shipped steady-state loop (frac <= 1) + the HBM side for completeness."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    _, cells = ag.count_cells(qlen, tlen, W, res["dstop"])
    gcups = cells / (kernel_ms * 1e-3) / 1e9
    C, NW = (24, 1) if W == 751 else (32, 4)
    hl = hot_loop_stats(C, NW, W % C)
    lanes = 32 * NW
    band_fill = (2 * W + 1) / (2.0 * C * lanes)                      # share of a group's lane-cells that lie inside the band
    hbm_peak, hbm_src = load_hbm_peak()
    alg_bytes = int(((qlen.astype(np.int64) + 7) // 8 * 4 + (tlen.astype(np.int64) + 7) // 8 * 4 + 16 + 12).sum())
    r = {"bound": "int_alu", "kernel": "agatha::extend16_kernel<%d,%d,%d>" % (C, NW, W % C), "unit": "GCUPS", "achieved": gcups,
         "cells_per_launch": int(cells), "kernel_ms": kernel_ms,
         "alu_peak_tera_lane_ops": int_peak["alu"], "fma_peak_tera_lane_ops": int_peak["fma"], "mixed_peak_tera_lane_ops": int_peak["mixed"],
         "peak_source": "agatha_measure_int_peak in this run (VIADDMNMX.U16x2 / IMAD streams on all SMs); MEASURED_PEAKS.json has no integer figure",
         "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                 "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak}}
    if hl:
        ceiling = int_peak["alu"] * 1e12 / hl["alu_per_lane_cell"] * band_fill / 1e9
        r.update({"peak": ceiling, "frac": gcups / ceiling, "alu_instructions_per_lane_cell": hl["alu_per_lane_cell"], "band_fill": band_fill,
                  "hot_loop": hl,
                  "note": "ceiling = ALU lane-op rate measured in this run / ALU-pipe instructions per lane-cell of the steady-state loop in the shipped "
                          "SASS x share of lane-cells inside the band; prologue, tail, events and job switches are what keeps frac below the ncu ALU-pipe utilisation"})
    else:
        r.update({"peak": None, "frac": None, "note": "hot-loop statistics unavailable (no cuobjdump and no committed profile)"})
    ncu = load_ncu(ncu_name) if ncu_name else None
    # traffic is per launch like `achieved`: only a capture of a launch over the same number of pairs qualifies
    if ncu and ncu.get("pairs") not in (None, len(qlen)):
        ncu = None
    r["traffic"] = ncu.get("dram_bytes_per_launch") if ncu else None
    if ncu:
        r["ncu"] = {k: ncu.get(k) for k in ("alu_pipe_pct", "fma_pipe_pct", "issue_active_pct", "warps_active_pct", "source", "pairs")}
    return r, gcups, cells


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="agatha_b200", choices=["agatha_b200", "reference"])
    ap.add_argument("--pairs", type=int, default=0, help="override the pair count of the workload (debugging; the default is the named config)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the extra reference-GPU-binary leg")
    ap.add_argument("--ref-gpu-pairs", type=int, default=16384)
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the C1/C3/C4 kernel legs at N = 1")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--staging-threads", type=int, default=0, help="host packing threads per rank in the end-to-end leg (default: the cores of this rank, at most 8)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    # One rank per GPU on one box: give every rank its own slice of the host cores (and, by first touch, host memory on the
    # socket those cores belong to). Without it the ranks' packing threads and pinned staging land on arbitrary sockets and the
    # end-to-end legs of 8 ranks share the cross-socket link.
    pinned_cores = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            per = max(1, len(cores) // lw)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            pinned_cores = [mine[0], mine[-1]]
        except OSError:
            pass

    import agatha_b200 as ag
    from agatha_b200._lib import DEFAULT_PARAMS
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: agatha_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # pairs are independent: there is no GPU collective on this path. gloo carries the barrier, the max-over-ranks of the
        # device-measured times and the gather of the results (and keeps NCCL's banner off stdout: exactly one JSON line).
        dist.init_process_group("gloo")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    name = "C2" if world == 1 else "C5"
    prof, seed, total, W, desc = WORKLOADS[name]
    if args.pairs:
        total = args.pairs
    params = dict(DEFAULT_PARAMS); params["band_width"] = W
    lo, hi = rank * total // world, (rank + 1) * total // world          # this rank's shard of the fixed set
    n = hi - lo
    data = ag.synth_pairs(prof, seed, n, first_pair=lo)
    qlen, tlen = data["qlen"], data["tlen"]

    int_peak = ag.measure_int_peak(local_rank)
    leg = DeviceLeg(ag, torch, dev, data, params)
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    for _ in range(args.warmup):
        leg.step()
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ag.launch_count()
    ms_total, kernel_ms, t_wall0, t_wall1 = leg.timed(args.steps, 0, barrier)
    launches = ag.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    tmax = torch.tensor([ms_total], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = total / (ms_step * 1e-3)
    res_dev = leg.results()
    kms = statistics.mean(kernel_ms)
    roofline, gcups_rank, cells = roofline_for(ag, leg, res_dev, kms, W, int_peak, qlen, tlen, "extend16_c24_c2_r02_ncu.json")
    cells_t = torch.tensor([float(cells)], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(cells_t, op=dist.ReduceOp.SUM)
    stops = res_dev["stop"]

    # ---- e2e leg: host buffers through the C ABI (host packing into pinned staging + H2D + kernels + D2H [+ gather]) ---------
    e2e_steps = max(1, min(args.steps, 3))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    # host packing threads of this rank: the cores this process may run on (its slice at N > 1), at most 8
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8) // max(1, local_world)
    stg = args.staging_threads or max(1, min(8, avail))
    p = ag.make_params(**params)
    job = lambda: ag.align_job(data["qbuf"], data["qoff"], qlen, data["tbuf"], data["toff"], tlen, p, devices=[local_rank], staging_threads=stg)
    job()                                                                        # warm-up (allocations)
    barrier()
    t0 = time.time()
    h2d = d2h = 0
    gathered = None
    for _ in range(e2e_steps):
        res, stats = job()
        h2d, d2h = stats["h2d_bytes"], stats["d2h_bytes"]
        if dist is not None:                                                     # results of every shard to rank 0
            mine = torch.from_numpy(np.stack([res["score"], res["query_end"], res["target_end"]]).astype(np.int32))
            sizes = [((r + 1) * total // world) - (r * total // world) for r in range(world)]
            bufs = [torch.empty((3, s), dtype=torch.int32) for s in sizes] if rank == 0 else None
            dist.gather(mine, bufs, dst=0)
            gathered = bufs
    barrier()
    t_e2e = (time.time() - t0) / e2e_steps
    te = torch.tensor([t_e2e], dtype=torch.float64)
    hb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    e2e_value = total / float(te.item())
    same = bool((res["score"] == res_dev["score"]).all() and (res["query_end"] == res_dev["query_end"]).all() and (res["target_end"] == res_dev["target_end"]).all())
    same_t = torch.tensor([1.0 if same else 0.0], dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(same_t, op=dist.ReduceOp.MIN)

    # ---- N > 1: the library's own multi-device scheduler (one process, all N devices), checked against one device ----------
    multi = None
    if world > 1:
        if rank == 0:
            try:
                m = min(n, 16384)
                sub = {k: data[k] for k in data}
                one, _ = ag.align_job(sub["qbuf"], sub["qoff"][:m], qlen[:m], sub["tbuf"], sub["toff"][:m], tlen[:m], p, devices=[0])
                t0 = time.time()
                alln, st = ag.align_job(sub["qbuf"], sub["qoff"][:m], qlen[:m], sub["tbuf"], sub["toff"][:m], tlen[:m], p, devices=list(range(world)))
                dt_first = time.time() - t0                                   # includes CUDA contexts + pinned staging on N-1 more devices
                t0 = time.time()
                alln2, st = ag.align_job(sub["qbuf"], sub["qoff"][:m], qlen[:m], sub["tbuf"], sub["toff"][:m], tlen[:m], p, devices=list(range(world)))
                dt = time.time() - t0
                multi = {"api": "agatha_align_job, one process, devices 0..%d, LPT sharding" % (world - 1), "pairs": int(m), "devices": int(st["n_devices"]),
                         "identical_to_single_device": bool((one == alln).all() and (one == alln2).all()),
                         "seconds_first_call_with_context_and_staging_setup": dt_first, "seconds": dt}
            except Exception as e:  # noqa: BLE001
                multi = {"error": repr(e)[:300]}
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rank 0, N = 1 only: CPU baseline, other workloads, the reference GPU program -- all outside every timed region ------
    cpu = refgpu = extra = None
    if world == 1:
        from oracle import oracle_py as op
        try:
            op.build(ref=False)
            cpu, _, _ = cpu_leg(op, data, params, args.cpu_budget, False, os.cpu_count() or 1)
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": METRIC, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        if not args.no_extra_configs:
            extra = {}
            del leg
            for cname, cn in EXTRA_LEGS.items():
                try:
                    cprof, cseed, _, cW, cdesc = WORKLOADS[cname]
                    cd = ag.synth_pairs(cprof, cseed, cn)
                    cp = dict(DEFAULT_PARAMS); cp["band_width"] = cW
                    cl = DeviceLeg(ag, torch, dev, cd, cp)
                    _, kms_c, _, _ = cl.timed(3, 3)
                    cres = cl.results()
                    rf, g, _ = roofline_for(ag, cl, cres, min(kms_c), cW, int_peak, cd["qlen"], cd["tlen"],
                                            "extend16_c32x4_c3_r02_ncu.json" if cname == "C3" else None)
                    rf.pop("hot_loop", None)
                    extra[cname] = {"workload": cdesc, "pairs": cn, "kernel_ms": min(kms_c), "alignments_per_s": cn / (min(kms_c) * 1e-3), "gcups": g, "roofline": rf,
                                    "stops": {"end": int((cres["stop"] == 0).sum()), "zdrop": int((cres["stop"] == 1).sum()), "bandexit": int((cres["stop"] == 2).sum())}}
                    del cl
                except Exception as e:  # noqa: BLE001
                    extra[cname] = {"error": repr(e)[:300]}
        if not args.no_ref_gpu:
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                refgpu = run_reference_gpu(ag, data, params, args.ref_gpu_pairs, td)
                try:    # parity against the reference GPU kernel on the same pairs (informational; the tests gate it)
                    rs, rq, rt = parse_ref_scores(refgpu.pop("_score_log"))
                    m = len(rs)
                    eq = (rs == res["score"][:m]) & (rq == res["query_end"][:m]) & (rt == res["target_end"][:m])
                    refgpu["parity_vs_ours"] = {"pairs": int(m), "identical": int(eq.sum())}
                    refgpu["ours_over_best_reference_shape"] = {"device_resident": value / refgpu["best"]["alignments_per_s"], "e2e": e2e_value / refgpu["best"]["alignments_per_s"]}
                except Exception as e:  # noqa: BLE001
                    refgpu["parity_vs_ours"] = {"error": repr(e)[:200]}

    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32 results, 16-bit packed DP state", "data": "synthetic",
            "config": {"workload": desc, "pairs_total": total, "pairs_this_rank": n,
                       "l2": "packed inputs (%.2f GB per GPU) exceed the 126 MB L2" % (leg_bytes(data) / 1e9),
                       "parallelism": "independent pairs, %d GPU(s), rank r aligns pairs [r*T/N, (r+1)*T/N) of the fixed set, no collective on the data path" % world,
                       "host_cores_rank0": pinned_cores, "host_packing_threads_per_rank": stg,
                       "scaling_note": "N = 1 is C2 (100k pairs); N > 1 shard the fixed 1M-pair C5 set of the same generator: value(N) / (N x value(1)) is the strong-scaling efficiency",
                       "stops_rank0": {"end": int((stops == 0).sum()), "zdrop": int((stops == 1).sum()), "bandexit": int((stops == 2).sum())}},
            "gcups": float(cells_t.item()) / (ms_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": int(hb[0].item()), "d2h_bytes_per_step": int(hb[1].item()), "steps": e2e_steps,
                    "api": "agatha_align_job (C ABI): pageable host buffers -> 4-bit packing into pinned staging (host) -> H2D -> extend kernels -> D2H"
                           + (" -> gather to rank 0 (gloo)" if world > 1 else ""),
                    "matches_device_leg": bool(same_t.item() == 1.0),
                    "gathered_on_rank0": (int(sum(b.shape[1] for b in gathered)) if gathered else None)},
            "gpu_launches": int(launches), "multi_device_scheduler": multi, "other_workloads": extra, "reference_gpu": refgpu}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def leg_bytes(data):
    return ((data["qlen"].astype(np.int64) + 7) // 8 * 4).sum() + ((data["tlen"].astype(np.int64) + 7) // 8 * 4).sum()


if __name__ == "__main__":
    sys.exit(main())
