#!/bin/bash
# round-2 call 3: parallel ordered sums + 16-byte tile staging; wave budget sweep; TMA level-1; realistic DRAM
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py tests/test_gpu_stream.py tests/test_gpu_golden.py -q --timeout 600 > $O/r02c3_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02c3_pytest.log
HT_TMA=1 timeout 600 python -m pytest tests/test_gpu_detect.py -q --timeout 600 > $O/r02c3_pytest_tma.log 2>&1
echo "pytest (HT_TMA=1) rc=$?"; tail -3 $O/r02c3_pytest_tma.log
HT_LIB=variants/libht_v3_th8_t256.so timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_quads.py tests/test_gpu_stream.py -q --timeout 600 > $O/r02c3_pytest_v3.log 2>&1
echo "pytest (v3 th8) rc=$?"; tail -4 $O/r02c3_pytest_v3.log
HT_LIB=variants/libht_v3_th16_t512.so timeout 600 python -m pytest tests/test_gpu_detect.py -q --timeout 600 > $O/r02c3_pytest_v3b.log 2>&1
echo "pytest (v3 th16 t512) rc=$?"; tail -2 $O/r02c3_pytest_v3b.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BARGS > $O/r02c3_$tag.json 2> $O/r02c3_$tag.err; }
BARGS="--workload detect"
run det_mb64
run det_mb128 HT_WAVE_MB=128
run det_mb256 HT_WAVE_MB=256
run det_all HT_WAVE=1024
run det_mb64_pipe HT_DETECT_PIPE=1
run det_mb128_pipe HT_WAVE_MB=128 HT_DETECT_PIPE=1
run det_all_tma HT_WAVE=1024 HT_TMA=1
run det_all_q3 HT_WAVE=1024 HT_LIB=variants/libht_q3.so
for v in th8_t256 th12_t256 th16_t256 th16_t512; do
  run det_all_v3_$v HT_WAVE=1024 HT_LIB=variants/libht_v3_$v.so
  run det_mb64_v3_$v HT_LIB=variants/libht_v3_$v.so
done
run det_mb64_pipe_v3_th8 HT_DETECT_PIPE=1 HT_LIB=variants/libht_v3_th8_t256.so
BARGS=""
run full_mb64
run full_all HT_WAVE=1024
BARGS="--width 320 --height 240"
run full_320
for f in $O/r02c3_det_*.json $O/r02c3_full_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split("r02c3_")[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["kernel_ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
# whole-step DRAM bytes WITHOUT ncu's cache flush between kernels (256 frames, default waves and one wave)
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r02c3_dram_mb64.csv python tools/profile_run.py --frames 256 --iters 2 > $O/r02c3_dram1.log 2>&1
HT_WAVE=1024 timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r02c3_dram_all.csv python tools/profile_run.py --frames 256 --iters 2 > $O/r02c3_dram2.log 2>&1
HT_WAVE=1024 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cascade -s 1 -c 1 -f \
  -o $O/r02c3_cascade python tools/profile_run.py --frames 256 --iters 2 > $O/r02c3_ncu_cascade.log 2>&1
HT_LIB=variants/libht_v3_th8_t256.so HT_WAVE=1024 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cascade -s 1 -c 1 -f \
  -o $O/r02c3_cascade_v3 python tools/profile_run.py --frames 256 --iters 2 > $O/r02c3_ncu_cascade_v3.log 2>&1
HT_WAVE=1024 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resample -s 8 -c 1 -f \
  -o $O/r02c3_resample python tools/profile_run.py --frames 256 --iters 2 > $O/r02c3_ncu_resample.log 2>&1
HT_WAVE=1024 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_track -s 2 -c 1 -f \
  -o $O/r02c3_track python tools/profile_run.py --frames 1024 --iters 2 > $O/r02c3_ncu_track.log 2>&1
ls -la $O | grep r02c3 | wc -l
