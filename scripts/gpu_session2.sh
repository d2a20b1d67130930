#!/usr/bin/env bash
# Session 2: tests (tc kernel), fp32 A/B at 1 GPU, 2-GPU overhead diagnosis
NG=2; O=gpurun_out/s2; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
b() { tag=$1; shift; echo "== $tag: $*"; timeout 300 env "$@" > ${O}_$tag.json 2> ${O}_$tag.err; echo "rc=$?"; grep -o '"ms_per_step": [0-9.e-]*' ${O}_$tag.json | head -3; tail -2 ${O}_$tag.err; }
timeout 1200 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; tail -15 ${O}_tests.log
b tc1   NNDT_X=1 python bench.py --steps 20 --warmup 5 --dtype fp32 --no-extras
b old1  NNDT_MNIST_TC=0 python bench.py --steps 20 --warmup 5 --dtype fp32 --no-extras
b tc1k  NNDT_X=1 python bench.py --steps 1000 --warmup 20 --dtype fp32 --no-extras
b tc2   NNDT_X=1 $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b tc2k  NNDT_X=1 $TR bench.py --gpus 2 --steps 1000 --warmup 20 --dtype fp32 --no-extras
b old2  NNDT_MNIST_TC=0 $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b old2res NNDT_MNIST_TC=0 NNDT_BENCH_PIPELINE=resident $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b old2dis NNDT_MNIST_TC=0 NNDT_BENCH_GRAPH=disjoint $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b old2ipc NNDT_MNIST_TC=0 NNDT_SYMM=ipc $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b old2nsp NNDT_MNIST_TC=0 NNDT_SEPARATE_PUBLISH=0 $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp32 --no-extras
b f64dis NNDT_BENCH_GRAPH=disjoint $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp64 --no-extras
b f64ipc NNDT_SYMM=ipc $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp64 --no-extras
b f64res NNDT_BENCH_PIPELINE=resident $TR bench.py --gpus 2 --steps 20 --warmup 5 --dtype fp64 --no-extras
