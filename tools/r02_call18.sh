#!/bin/bash
# round-2 call 18: leaner k_track pixel loop (HT_TRACK_LOOP2) vs the previous one, new tier defaults (n/128 @ 8, n/32 @ 4)
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $O/r02c18_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02c18_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c18_$tag.json 2> $O/r02c18_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--pipeline 0"
run full
run full_loop1 HT_LIB=$V/libht_loop1.so
run full_b
run full_loop1_b HT_LIB=$V/libht_loop1.so
run full_mid48 HT_TRACK_MID=48
run full_h256 HT_TRACK_HEAVY=256
run full_mid24 HT_TRACK_MID=24
BARGS="--pipeline 1"
run pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
run pipe
for f in $O/r02c18_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c18_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
HT_LIB=$V/libht_ptrace.so timeout 300 python tools/track_timeline.py 1024 > $O/r02c18_phases.txt 2>&1; head -5 $O/r02c18_phases.txt; tail -5 $O/r02c18_phases.txt
