#!/usr/bin/env bash
# Session 5 (1 GPU): whole single-GPU test-suite (no -x), tc / cl64 phase profiles, 1-GPU bench lines
O=gpurun_out/s5; mkdir -p gpurun_out
b() { tag=$1; shift; echo "== $tag: $*"; timeout 240 env "$@" > ${O}_$tag.json 2> ${O}_$tag.err; echo "rc=$?"; grep -o '"ms_per_step": [0-9.e-]*' ${O}_$tag.json | head -3; tail -2 ${O}_$tag.err; }
NNDT_MNIST_TC=1 timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_distributed.py > ${O}_tests.log 2>&1; tail -25 ${O}_tests.log
NNDT_MNIST_TC=1 timeout 200 python scripts/profile_tc_phases.py --rounds 60 > ${O}_tc_phases.txt 2>&1; tail -48 ${O}_tc_phases.txt
b tc1   NNDT_MNIST_TC=1 python bench.py --steps 20 --warmup 5 --dtype fp32 --no-extras
b f64   NNDT_X=1 python bench.py --steps 20 --warmup 5 --no-extras
