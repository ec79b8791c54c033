#!/bin/bash
# One ncu --set full capture of the first launch of a kernel (regex) inside a tools/kperf.py run, plus the raw-page CSV.
#   tools/ncu_kernel.sh OUT_PREFIX KERNEL_REGEX kperf args...
set -e
OUT=$1; KRE=$2; shift 2
mkdir -p "$(dirname "$OUT")"
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -c 1 -f -o "$OUT" python tools/kperf.py --reps 1 "$@" > "$OUT.log" 2>&1 || true
ncu -i "$OUT.ncu-rep" --page raw --csv > "$OUT.raw.csv" 2>/dev/null || true
