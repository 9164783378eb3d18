#!/bin/bash
# round-2 call 22: single-CTA tier for the cheap streams, now that the launch is throughput-bound (CTA-time) and the tiers are prioritised
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline 0 > $O/r02c22_$tag.json 2> $O/r02c22_$tag.err; }
run base
run l2 HT_TRACK_LIGHT=2
run l1p3 HT_TRACK_LIGHT=1.3
run l1p1 HT_TRACK_LIGHT=1.1
run l1 HT_TRACK_LIGHT=1
run l1p3_128 HT_TRACK_LIGHT=1.3,128
run l1p3_512 HT_TRACK_LIGHT=1.3,512
run l1p3_h64 HT_TRACK_LIGHT=1.3 HT_TRACK_HEAVY=64
for f in $O/r02c22_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c22_")[1], round(d["value"]), round(d["ms_per_step"],3), d["kernel_ms_per_step"]["track"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
HT_TRACK_LIGHT=1.3 timeout 300 python tools/track_timeline.py 1024 > $O/r02c22_timeline_l1p3.txt 2>&1; head -5 $O/r02c22_timeline_l1p3.txt; tail -9 $O/r02c22_timeline_l1p3.txt
