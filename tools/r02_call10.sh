#!/bin/bash
# round-2 call 10: are the side streams falsely serialised by the 8 default hardware queues?  CUDA_DEVICE_MAX_CONNECTIONS=32;
# k_track at 64 registers (4 CTAs/SM); nth_bit32 instead of __fns in k_cascade
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_detect.py tests/test_gpu_quads.py -m gpu -q --timeout 600 > $O/r02c10_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r02c10_pytest.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c10_$tag.json 2> $O/r02c10_$tag.err; }
V=$PWD/headtrackr_b200/variants
BARGS="--pipeline 0"
run nopipe
run nopipe_c32 CUDA_DEVICE_MAX_CONNECTIONS=32
run nopipe_c32_trk4 CUDA_DEVICE_MAX_CONNECTIONS=32 HT_LIB=$V/libht_trk4.so
BARGS="--pipeline 1"
run pipe_c32 CUDA_DEVICE_MAX_CONNECTIONS=32
run pipe_c32_noprio CUDA_DEVICE_MAX_CONNECTIONS=32 HT_TRACK_PRIO=0
run pipe_c32_trk4 CUDA_DEVICE_MAX_CONNECTIONS=32 HT_LIB=$V/libht_trk4.so
run pipe_c32_trk4_noprio CUDA_DEVICE_MAX_CONNECTIONS=32 HT_LIB=$V/libht_trk4.so HT_TRACK_PRIO=0
run pipe_c32_c1 CUDA_DEVICE_MAX_CONNECTIONS=32 HT_TRACK_HEAVY=0 HT_TRACK_MID=0
for f in $O/r02c10_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c10_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/track_timeline.py 1024 > $O/r02c10_timeline_c32.txt 2>&1; head -8 $O/r02c10_timeline_c32.txt; tail -9 $O/r02c10_timeline_c32.txt
