#!/bin/bash
# round-2 call 17: selective zero-weight masking (streams that swept > 4 frames' worth of pixels last time); tier sizes
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02c17_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c17_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c17_$tag.json 2> $O/r02c17_$tag.err; }
BARGS="--pipeline 0"
run full
run full_maskall HT_TRACK_MASK=4,0
run full_nomask HT_TRACK_MASK=0
run full_m2 HT_TRACK_MASK=4,2
run full_m8 HT_TRACK_MASK=4,8
run full_mid32 HT_TRACK_MID=32
run full_mid32_h128 HT_TRACK_MID=32 HT_TRACK_HEAVY=128
run full_mid24 HT_TRACK_MID=24
run full_mid48 HT_TRACK_MID=48
BARGS="--pipeline 1"
run pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
for f in $O/r02c17_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c17_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
