#!/bin/bash
# round-2 call 11: k_track with the maximum shared-memory carve-out (can its CTAs now share SMs with the detection kernels?);
# background mode: tracking below the main stream's priority, k_cascade at 3 CTAs/SM + 1 k_track CTA (64-register build)
O=gpurun_out; mkdir -p $O
HT_PIPE_BG=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 600 > $O/r02c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r02c11_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c11_$tag.json 2> $O/r02c11_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--pipeline 0"
run nopipe
BARGS="--pipeline 1"
run pipe
run pipe_noprio HT_TRACK_PRIO=0
run pipe_c1 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
run pipe_bg_trk4 HT_PIPE_BG=1 HT_BENCH_STREAM_PRIO=-10 HT_LIB=$V/libht_trk4.so
run pipe_bg_trk4_c1 HT_PIPE_BG=1 HT_BENCH_STREAM_PRIO=-10 HT_LIB=$V/libht_trk4.so HT_TRACK_HEAVY=0 HT_TRACK_MID=0
run pipe_bg_trk4_nt128 HT_PIPE_BG=1 HT_BENCH_STREAM_PRIO=-10 HT_LIB=$V/libht_trk4.so HT_TRACK_NT=128
run pipe_bg HT_PIPE_BG=1 HT_BENCH_STREAM_PRIO=-10
run pipe_bg_trk4_loprio HT_PIPE_BG=1 HT_LIB=$V/libht_trk4.so
for f in $O/r02c11_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c11_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
