#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/r2g_gpus.txt
echo "== NCCL world-2 sharded pipeline test"
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -8 | tee $OUT/r2g_pytest_distributed.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > $OUT/r2g_bench_n2.json; cut -c1-600 $OUT/r2g_bench_n2.json
echo "== reference arm under torchrun N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-300
