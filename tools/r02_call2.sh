#!/bin/bash
# round-2 call 2: first run of the frame-quad detector on hardware - correctness first, then timing variants
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/r02c2_gpu.txt
# 1. memcheck of a tiny end-to-end call (catches out-of-bounds before anything long runs)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/r02c2_sanitizer.log 2>&1
echo "sanitizer rc=$?"; tail -5 $O/r02c2_sanitizer.log
# 2. parity tests
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r02c2_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $O/r02c2_pytest.log
HT_TMA=1 timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py -q --timeout 600 > $O/r02c2_pytest_tma.log 2>&1
echo "pytest (HT_TMA=1) rc=$?"; tail -4 $O/r02c2_pytest_tma.log
# 3. timing variants (detect workload isolates the new kernels)
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c2_$tag.json 2> $O/r02c2_$tag.err; }
BARGS="--workload detect"
run det_w32
run det_w16 HT_WAVE=16
run det_w64 HT_WAVE=64
run det_w1024 HT_WAVE=1024
run det_w32_pipe HT_DETECT_PIPE=1
run det_w16_pipe HT_WAVE=16 HT_DETECT_PIPE=1
run det_q3 HT_LIB=variants/libht_q3.so
run det_tma HT_TMA=1
BARGS=""
run full_w32
run full_w32_pipe HT_DETECT_PIPE=1
run full_nt512c1 HT_TRACK_NT=512 HT_TRACK_CLUSTER=1
run full_noheavy HT_TRACK_HEAVY=0
run full_nohist HT_TRACK_HISTORY=0
run full_heavy32 HT_TRACK_HEAVY=32
BARGS="--workload streams --streams 1 --stream-frames 120"
run streams1
BARGS="--workload streams --streams 16 --stream-frames 120"
run streams16
BARGS="--workload detect720 --interval 3"
run d720_i3
BARGS="--width 320 --height 240"
run full_320
for f in $O/r02c2_det_*.json $O/r02c2_full_*.json $O/r02c2_streams*.json $O/r02c2_d720*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c2_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
# 4. launch list + one full capture of k_cascade + whole-step DRAM bytes (256 frames)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r02c2_launches_256.csv python tools/profile_run.py --frames 256 --iters 1 > $O/r02c2_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cascade -s 2 -c 1 -f \
  -o $O/r02c2_cascade python tools/profile_run.py --frames 256 --iters 1 > $O/r02c2_ncu_cascade.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gray -s 2 -c 1 -f \
  -o $O/r02c2_gray python tools/profile_run.py --frames 256 --iters 1 > $O/r02c2_ncu_gray.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 20 -c 1 -f \
  -o $O/r02c2_resample python tools/profile_run.py --frames 256 --iters 1 > $O/r02c2_ncu_resample.log 2>&1
ls -la $O | grep r02c2 | tail -30
