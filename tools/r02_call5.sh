#!/bin/bash
# round-2 call 5 (2 GPUs): the N>1 path of bench.py (overlapped NCCL gather, shard check, per-rank report), both arms
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02c5_n2.json 2> $O/r02c5_n2.err
echo "n2 rc=$?"; tail -c 1500 $O/r02c5_n2.json; tail -5 $O/r02c5_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload streams --streams 8 > $O/r02c5_n2_streams.json 2> $O/r02c5_n2_streams.err
echo "n2 streams rc=$?"; tail -c 600 $O/r02c5_n2_streams.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --workload detect720 > $O/r02c5_n2_d720.json 2> $O/r02c5_n2_d720.err
echo "n2 d720 rc=$?"; tail -c 400 $O/r02c5_n2_d720.json
python bench.py --steps 5 --warmup 3 > $O/r02c5_n1.json 2> $O/r02c5_n1.err
echo "n1 rc=$?"; tail -c 800 $O/r02c5_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02c5_ref.json 2> $O/r02c5_ref.err
echo "ref rc=$?"; tail -c 400 $O/r02c5_ref.json
python __graft_entry__.py smoke 2>&1 | tail -2
