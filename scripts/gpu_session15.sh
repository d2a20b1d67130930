#!/usr/bin/env bash
# Session 15 (4 GPUs): full bench line (headline + one node per GPU sweep) and the NCCL baseline at N = 4
NG=4; mkdir -p gpurun_out; O=gpurun_out/s15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571"
timeout 400 $TR bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench4.json 2> ${O}_bench4.err; echo "bench4 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_bench4.json | head -4 | tr '\n' ' ')"
timeout 150 $TR bench.py --impl nccl --gpus $NG --steps 20 --warmup 5 > ${O}_nccl4.json 2> ${O}_nccl4.err; echo "nccl4 rc=$? $(grep -o '"ms_per_step[a-z_]*": [0-9.e-]*' ${O}_nccl4.json | head -3 | tr '\n' ' ')"
