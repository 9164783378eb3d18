#!/bin/bash
# round-2 call 13: super-row tile layout (one base per window) vs the two-base layout; both with unrolled groups
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/r02c13_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c13_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c13_$tag.json 2> $O/r02c13_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--workload detect"
run det
run det_twobase HT_LIB=$V/libht_twobase.so
run det_b
run det_twobase_b HT_LIB=$V/libht_twobase.so
BARGS="--pipeline 0"
run full
BARGS="--pipeline 1"
run full_pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
for f in $O/r02c13_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c13_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
# source-level capture of the new default k_cascade
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_cascade -c 1 -o $O/r02c13_cascade -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload detect > $O/r02c13_ncu_casc.log 2>&1
ncu -i $O/r02c13_cascade.ncu-rep --page source --print-source cuda,sass --csv > $O/r02c13_casc_cs.csv 2>/dev/null
ncu -i $O/r02c13_cascade.ncu-rep --page details --csv > $O/r02c13_cascade_details.csv 2>/dev/null
ls -la $O | grep r02c13_casc
