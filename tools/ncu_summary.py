#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (tools/ncu_kernel.sh) into the handful of numbers DESIGN.md and bench.py quote:
pipe utilisation, issue slots, stall reasons per issued instruction, occupancy, DRAM traffic. Prints JSON."""
import csv
import json
import sys


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    vals = rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))

    def f(name, default=None):
        v = d.get(name)
        if v in (None, "", "n/a"):
            return default
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return v
    out = {
        "kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"),
        "duration_ms": (f("gpu__time_duration.sum") or 0) / (1e6 if u.get("gpu__time_duration.sum") in ("ns", "nsecond") else 1e3 if u.get("gpu__time_duration.sum") in ("us", "usecond") else 1),
        "regs_per_thread": f("launch__registers_per_thread"),
        "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "fmaheavy_pipe_pct": f("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
        "uniform_pipe_pct": f("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_active") or f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "ipc": f("sm__inst_executed.avg.per_cycle_active"),
        "warp_insts": f("smsp__inst_executed.sum"),
        "thread_insts": f("smsp__thread_inst_executed.sum"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "icache_hit_pct": f("sm__icc_requests_lookup_hit.sum") and f("sm__icc_requests.sum") and 100.0 * f("sm__icc_requests_lookup_hit.sum") / f("sm__icc_requests.sum"),
        "dram_bytes_per_launch": sum((f(k) or 0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.get(k), 1)
                                     for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")),
        "local_ld": f("smsp__inst_executed_op_local_ld.sum"), "local_st": f("smsp__inst_executed_op_local_st.sum"),
        "shared_ld": f("smsp__inst_executed_op_shared_ld.sum"), "shared_st": f("smsp__inst_executed_op_shared_st.sum"),
    }
    stalls = {}
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            v = f(k)
            if v is not None:
                stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
    out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    out = {k: v for k, v in out.items() if v is not None}
    for a in extra:
        k, _, v = a.partition("=")
        out[k] = int(v) if v.isdigit() else v
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
