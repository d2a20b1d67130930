#!/usr/bin/env bash
# Session 20 (1 GPU): the two commands the driver runs at round end, on the final tree
mkdir -p gpurun_out; O=gpurun_out/s20
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > ${O}_ref.json 2> ${O}_ref.err; echo "ref rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_ref.json | head -2 | tr '\n' ' ')"
timeout 80 python bench.py > ${O}_default.json 2> ${O}_default.err; echo "default rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_default.json | head -2 | tr '\n' ' ') $(grep -o '"steps": [0-9]*' ${O}_default.json | head -1)"
