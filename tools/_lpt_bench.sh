set -x
python -m pytest tests/test_gpu_track.py -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['kernel_ms_per_step'])"; }
run HT_TRACK_LPT=0
run HT_TRACK_LPT=1
run HT_TRACK_NT=128
run HT_TRACK_NT=128 HT_TRACK_CLUSTER=4
run HT_TRACK_HEAVY=16
run HT_TRACK_HEAVY=8
run HT_TRACK_HEAVY=16 HT_TRACK_NT=128
run HT_TRACK_HEAVY=8 HT_TRACK_CLUSTER=1
run HT_TRACK_HEAVY=8 HT_TRACK_CLUSTER=1 HT_TRACK_NT=128
