#!/bin/bash
# round-2 call 8: late-stage 48-byte records; full GPU suite; strict k_track timeline; tier A/B; fresh ncu summaries
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r02c8_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c8_pytest.log
timeout 300 python tools/track_timeline.py 1024 > $O/r02c8_timeline_strict.txt 2>&1; tail -45 $O/r02c8_timeline_strict.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c8_$tag.json 2> $O/r02c8_$tag.err; }
BARGS=""
run full
run full_b
run full_mid32 HT_TRACK_MID=32
run full_mid32_b HT_TRACK_MID=32
run full_h128_16 HT_TRACK_HEAVY=128,16
BARGS="--workload detect"
run det
for f in $O/r02c8_full*.json $O/r02c8_det.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c8_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
# ncu: cascade with source-level counters, one capture of the whole-step kernel list
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_cascade -c 1 -o $O/r02c8_cascade -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload detect > $O/r02c8_ncu_casc.log 2>&1
ncu -i $O/r02c8_cascade.ncu-rep --page details --csv > $O/r02c8_cascade_details.csv 2>/dev/null
ncu -i $O/r02c8_cascade.ncu-rep --page source --csv > $O/r02c8_cascade_source.csv 2>/dev/null
ls -la $O | grep r02c8 | head -30
