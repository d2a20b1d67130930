#!/usr/bin/env bash
# Session 13 (1 GPU): cl64 conv-grad compaction + early announcement: tests, phase tables, 1-GPU bench lines
mkdir -p gpurun_out; O=gpurun_out/s13
timeout 900 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; tail -5 ${O}_tests.log
timeout 200 python scripts/profile_tc_phases.py --rounds 60 > ${O}_phases.txt 2>&1; grep -A18 "mnist_cl64" ${O}_phases.txt | tail -20
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > ${O}_f64.json 2> ${O}_f64.err; echo "f64 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f64.json | head -2 | tr '\n' ' ')"
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --dtype fp32 > ${O}_f32.json 2> ${O}_f32.err; echo "f32 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' ${O}_f32.json | head -2 | tr '\n' ' ')"
