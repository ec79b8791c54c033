#!/bin/bash
# AGAThA.sh-compatible runner for the B200 engine (reference: AGAThA.sh:1-53, misc/avg_time.py).
# Same outputs: $OUTPUT_DIR/raw.log (kernel ms per run), score.log (alignment results), time.json (average).
#   tools/agatha.sh [-i ITER] [-d DATASET_DIR] [-o OUTPUT_DIR] [-g GPUS]
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
PROG="$HERE/agatha_b200/bin/agatha_manual"
DATASET_DIR="$HERE/dataset/"; OUTPUT_DIR="$HERE/output/"; ITER=1; GPUS=0
DATASET_NAME="test"; PROCESS="AGAThA"
while getopts "i:d:o:g:" opt; do
  case "$opt" in i) ITER="$OPTARG";; d) DATASET_DIR="$OPTARG/";; o) OUTPUT_DIR="$OPTARG/";; g) GPUS="$OPTARG";; esac
done
RAW_FILE="${OUTPUT_DIR}raw.log"; FINAL_FILE="${OUTPUT_DIR}time.json"; SCORE_FILE="${OUTPUT_DIR}score.log"
mkdir -p "$OUTPUT_DIR"; rm -f "$RAW_FILE" "$SCORE_FILE" "$FINAL_FILE"
echo ">>> Running $PROCESS for $ITER iterations."
for ((iter = 0; iter < ITER; iter++)); do
  echo ">> Iteration $((iter + 1))"
  # the reference passes ref.fasta as the query batch and query.fasta as the target batch (AGAThA.sh:44)
  "$PROG" -p -m 1 -x 4 -q 6 -r 2 -s 3 -z 400 -w 751 -g "$GPUS" "${DATASET_DIR}ref.fasta" "${DATASET_DIR}query.fasta" "$RAW_FILE" > "$SCORE_FILE" || exit 1
done
python3 - "$PROCESS" "$DATASET_NAME" "$RAW_FILE" "$FINAL_FILE" "$ITER" <<'PY'
import json, os, sys
process, dataset, raw, out, it = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
vals = [float(x) for x in open(raw).read().split()] if os.path.exists(raw) else []
avg = sum(vals) / it if vals else "NaN"
res = json.load(open(out)) if os.path.exists(out) else {}
res.setdefault(process, {})[dataset] = avg
json.dump(res, open(out, "w"))
PY
echo "Complete."
