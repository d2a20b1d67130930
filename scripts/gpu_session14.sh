#!/usr/bin/env bash
# Session 14 (2 GPUs): early announcement: distributed tests (incl. delayed ranks), full 2-GPU bench line, timeline
NG=2; mkdir -p gpurun_out; O=gpurun_out/s14
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561"
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 400 $TR bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench2.json 2> ${O}_bench2.err; echo "bench2 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_bench2.json | head -4 | tr '\n' ' ')"
NNDT_ANNOUNCE=end timeout 200 $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-extras > ${O}_end.json 2> ${O}_end.err; echo "announce=end rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_end.json | head -2 | tr '\n' ' ')"
timeout 200 $TR scripts/timeline_rounds.py --dtype fp64 --out gpurun_out/s14_tl > ${O}_tl2.log 2>&1; echo "tl2 rc=$?"
