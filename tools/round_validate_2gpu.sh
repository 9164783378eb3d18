#!/bin/bash
# gpurun --gpus 2 --timeout 600 -- 'bash tools/round_validate_2gpu.sh r01'
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu_$R.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${R}_n1_check.json 2> $O/bench_${R}_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_${R}_n2.json 2>> $O/bench_${R}_n2.err
tail -n 3 $O/bench_${R}_n2.err
python - <<PY
import json
for f in ("bench_${R}_n1_check", "bench_${R}_n2"):
    try:
        d = json.loads(open("$O/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "memo", round(d["memo"]["value"]), d["n_gpus"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
