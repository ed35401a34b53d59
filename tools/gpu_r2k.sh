#!/bin/bash
# round 2, call K (8 GPUs): the bench under torchrun exactly as the driver launches it, and the reference arm beside it
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference --cpu-seconds 6 > $O/bench_8gpu_ref.json 2> $O/bench_8gpu_ref.err; echo "ref8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --config 4 --steps 2 --warmup 1 --no-e2e > $O/bench_8gpu_cfg4.json 2> $O/bench_8gpu_cfg4.err; echo "cfg4 rc=$?"
head -c 400 $O/bench_8gpu.json; echo; tail -n 3 $O/bench_8gpu.err
ls -la $O
