#!/bin/bash
# round-2 call 7: full GPU suite with f3 (head epilogue) and f4 (ingest); k_track timeline on the bench mix; 16-CTA clusters
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r02c7_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/r02c7_pytest.log
timeout 300 python tools/track_timeline.py 1024 > $O/r02c7_timeline_default.txt 2>&1; tail -45 $O/r02c7_timeline_default.txt
timeout 300 python tools/track_timeline.py 1024 HT_TRACK_HEAVY=64,16 > $O/r02c7_timeline_h16.txt 2>&1; head -8 $O/r02c7_timeline_h16.txt | tail -5
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c7_$tag.json 2> $O/r02c7_$tag.err; }
BARGS=""
run full
run full_h64_16 HT_TRACK_HEAVY=64,16
run full_h128_16 HT_TRACK_HEAVY=128,16
run full_h128_16_mid32_8 HT_TRACK_HEAVY=128,16 HT_TRACK_MID=32,8
run full_h64_16_mid16_8 HT_TRACK_HEAVY=64,16 HT_TRACK_MID=16,8
run full_mid32 HT_TRACK_MID=32
run full_mid12 HT_TRACK_MID=12
for f in $O/r02c7_full*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c7_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
