#!/bin/bash
# round-2 call 15: zero-weight marking of the bin plane (k_bins_mask) + row-segment skip in k_track
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02c15_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02c15_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c15_$tag.json 2> $O/r02c15_$tag.err; }
BARGS="--pipeline 0"
run full
run full_nomask HT_TRACK_MASK=0
run full_b
run full_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
run full_h128 HT_TRACK_HEAVY=128
run full_mid32 HT_TRACK_MID=32
run full_320 
BARGS="--pipeline 1"
run pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
run pipe
for f in $O/r02c15_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c15_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"], d["track_stats"]["passes"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
timeout 300 python tools/track_timeline.py 1024 > $O/r02c15_timeline.txt 2>&1; head -6 $O/r02c15_timeline.txt; tail -9 $O/r02c15_timeline.txt
