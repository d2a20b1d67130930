#!/usr/bin/env bash
# Run under gpurun (1 GPU): ncu captures of the hot kernels + launch list + clocks. Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list of the headline round (eager, no graph): shares per kernel
timeout 200 $NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file gpurun_out/launches_mnist.csv \
    python scripts/profile_round.py 24 > gpurun_out/launches_mnist.log 2>&1
# full captures, one launch each
timeout 300 $NCU --set full --import-source on -k regex:mnist_kernel -s 8 -c 1 -o gpurun_out/mnist_train \
    python scripts/profile_round.py 8 > gpurun_out/p1.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:dinno_update -s 8 -c 1 -o gpurun_out/dinno_update \
    python scripts/profile_round.py 8 > gpurun_out/p2.log 2>&1
# validation kernel: 0 training rounds, so the only mnist_kernel launch is the forward-only instantiation
EVAL=1 timeout 300 $NCU --set full --import-source on -k regex:mnist_kernel -s 0 -c 1 -o gpurun_out/mnist_eval \
    python scripts/profile_round.py 0 > gpurun_out/p3.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:mlp_train_kernel -s 2 -c 1 -o gpurun_out/mlp_train \
    python scripts/profile_mlp.py 4 > gpurun_out/p4.log 2>&1
# in-kernel phase timing (globaltimer stamps; not under ncu)
timeout 200 python scripts/profile_round_phases.py --per-step > gpurun_out/phases_per_step.txt 2>&1
timeout 200 python scripts/profile_round_phases.py > gpurun_out/phases_round_kernel.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.csv
tail -2 gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log gpurun_out/p4.log
