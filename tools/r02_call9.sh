#!/bin/bash
# round-2 call 9: k_track tiers on prioritised side streams (heavy first); pipelined steps (tracking of step s under the
# detection of step s+1); GPU suite incl. tests/test_gpu_pipeline.py
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/r02c9_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $O/r02c9_pytest.log
timeout 300 python tools/track_timeline.py 1024 > $O/r02c9_timeline_prio.txt 2>&1; head -8 $O/r02c9_timeline_prio.txt; tail -9 $O/r02c9_timeline_prio.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c9_$tag.json 2> $O/r02c9_$tag.err; }
BARGS="--pipeline 0"
run nopipe
run nopipe_noprio HT_TRACK_PRIO=0
run nopipe_mid8 HT_TRACK_MID=8
run nopipe_h32 HT_TRACK_HEAVY=32
BARGS="--pipeline 1"
run pipe
run pipe_b
run pipe_noprio HT_TRACK_PRIO=0
run pipe_mid8 HT_TRACK_MID=8
run pipe_nt512 HT_TRACK_NT=512
for f in $O/r02c9_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c9_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
