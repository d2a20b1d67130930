#!/usr/bin/env bash
# Session 9 (2 GPUs): distributed tests on both GPUs + the full 2-GPU bench line
NG=2; O=gpurun_out/s9; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > ${O}_tests.log 2>&1; tail -8 ${O}_tests.log
timeout 400 $TR bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench2.json 2> ${O}_bench2.err; echo "rc=$?"; grep -o '"ms_per_step": [0-9.e-]*' ${O}_bench2.json | head -4; tail -3 ${O}_bench2.err
