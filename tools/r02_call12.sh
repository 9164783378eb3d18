#!/bin/bash
# round-2 call 12: k_cascade with REDUX reductions / group lambda (default) vs four unrolled groups (HT_CT_GROUPS=1);
# source-level ncu capture of k_track<2,256> on the bench mix
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/r02c12_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c12_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c12_$tag.json 2> $O/r02c12_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--workload detect"
run det
run det_ctg HT_LIB=$V/libht_ctg.so
run det_b
run det_ctg_b HT_LIB=$V/libht_ctg.so
BARGS="--pipeline 0"
run full
run full_ctg HT_LIB=$V/libht_ctg.so
for f in $O/r02c12_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c12_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
HT_TRACK_HEAVY=0 HT_TRACK_MID=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_track$ --launch-skip 3 -c 1 -o $O/r02c12_track2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pipeline 0 > $O/r02c12_ncu_track.log 2>&1
tail -3 $O/r02c12_ncu_track.log
ncu -i $O/r02c12_track2.ncu-rep --page source --print-source cuda,sass --csv > $O/r02c12_track2_cs.csv 2>/dev/null
ncu -i $O/r02c12_track2.ncu-rep --page details --csv > $O/r02c12_track2_details.csv 2>/dev/null
ls -la $O | grep r02c12_track2
