#!/bin/bash
# round-2 call 14: k_track partial moments by st.async + mbarrier (no cluster barrier, one CTA barrier per pass) vs the
# cluster.sync exchange; 128-thread CTAs for the default tier
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02c14_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02c14_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c14_$tag.json 2> $O/r02c14_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--pipeline 0"
run full
run full_nombar HT_LIB=$V/libht_nombar.so
run full_b
run full_nombar_b HT_LIB=$V/libht_nombar.so
run full_nt128 HT_TRACK_NT=128
run full_c4 HT_TRACK_HEAVY=0 HT_TRACK_MID=0 HT_TRACK_CLUSTER=4
run full_h32 HT_TRACK_HEAVY=32
run full_mid8 HT_TRACK_MID=8
BARGS="--pipeline 1"
run pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
for f in $O/r02c14_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c14_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
timeout 300 python tools/track_timeline.py 1024 > $O/r02c14_timeline.txt 2>&1; head -6 $O/r02c14_timeline.txt; tail -4 $O/r02c14_timeline.txt
