#!/usr/bin/env bash
# Session 11 (1 GPU): one-wave update kernels (U = 4 variants): tests, 1-GPU bench lines, timeline
mkdir -p gpurun_out; O=gpurun_out/s11
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_tests.log 2>&1; tail -5 ${O}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > ${O}_f64.json 2> ${O}_f64.err; echo "f64 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f64.json | head -2 | tr '\n' ' ')"
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --dtype fp32 > ${O}_f32.json 2> ${O}_f32.err; echo "f32 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f32.json | head -2 | tr '\n' ' ')"
NNDT_MNIST_TC=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --dtype fp32 > ${O}_f32old.json 2> ${O}_f32old.err; echo "f32old rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f32old.json | head -2 | tr '\n' ' ')"
timeout 120 python scripts/timeline_rounds.py --dtype fp64 --out gpurun_out/s11_tl > ${O}_tl1.log 2>&1; echo "tl1 rc=$?"
timeout 120 python scripts/timeline_rounds.py --dtype fp32 --out gpurun_out/s11_tl > ${O}_tl1f.log 2>&1; echo "tl1f rc=$?"
