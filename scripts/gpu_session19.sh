#!/usr/bin/env bash
# Session 19 (1 GPU): final validation of the tree: GPU tests, smoke, 1-GPU bench lines, cl64 phases
mkdir -p gpurun_out; O=gpurun_out/s19
timeout 600 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > ${O}_f64.json 2> ${O}_f64.err; echo "f64 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f64.json | head -2 | tr '\n' ' ') $(grep -o '"host_placement": "[^"]*"' ${O}_f64.json)"
timeout 200 python scripts/profile_tc_phases.py --rounds 60 > ${O}_phases.txt 2>&1; grep -A18 "mnist_cl64" ${O}_phases.txt | tail -19
