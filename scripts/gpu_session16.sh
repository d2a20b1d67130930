#!/usr/bin/env bash
# Session 16 (8 GPUs): the full bench line with the final kernels / protocol
NG=8; mkdir -p gpurun_out; O=gpurun_out/s16
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29581"
timeout 200 $TR bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench8.json 2> ${O}_bench8.err; echo "bench8 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_bench8.json | head -4 | tr '\n' ' ')"
