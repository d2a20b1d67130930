#!/usr/bin/env bash
# Session 3 (8 GPUs): multi-GPU tests at 8 ranks, full bench line at 8 GPUs, fp32 line, NCCL baseline
NG=${1:-8}; O=gpurun_out/s3; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
b() { tag=$1; shift; echo "== $tag: $*"; timeout 400 env "$@" > ${O}_$tag.json 2> ${O}_$tag.err; echo "rc=$?"; grep -o '"ms_per_step": [0-9.e-]*' ${O}_$tag.json | head -4; tail -2 ${O}_$tag.err; }
timeout 900 python -m pytest tests/test_distributed.py -m gpu -q > ${O}_tests.log 2>&1; tail -8 ${O}_tests.log
b full8  NNDT_X=1 $TR bench.py --gpus $NG --steps 20 --warmup 5
b f32_8  NNDT_X=1 $TR bench.py --gpus $NG --steps 20 --warmup 5 --dtype fp32 --no-extras
b f32_8k NNDT_X=1 $TR bench.py --gpus $NG --steps 1000 --warmup 20 --dtype fp32 --no-extras
b f32_1  NNDT_X=1 python bench.py --steps 20 --warmup 5 --dtype fp32 --no-extras
b f64_1  NNDT_X=1 python bench.py --steps 20 --warmup 5 --no-extras
b nccl8  NNDT_X=1 $TR bench.py --impl nccl --gpus $NG --steps 50 --warmup 5
