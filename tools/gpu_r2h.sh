#!/bin/bash
# round 2, call H (2 GPUs): the library's own NCCL path -- driver --gpus 0,1, bench under torchrun, the reference arm at N=2
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_driver_gpu.py -m gpu -x -q -k "two_gpus" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference --cpu-seconds 5 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; echo "ref2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4 --pairs 8000000 --steps 2 --warmup 1 --no-e2e > $O/bench_2gpu_cfg4.json 2> $O/bench_2gpu_cfg4.err; echo "cfg4 rc=$?"
tail -3 $O/*.err
ls -la $O
