#!/usr/bin/env bash
# One gpurun session: GPU tests, then the bench arms.  Usage: scripts/gpu_session.sh <tag> <ngpus> [steps...]
TAG=${1:-s}; NG=${2:-1}
mkdir -p gpurun_out
O=gpurun_out/$TAG
run() { echo "== $*" ; timeout "$T" "$@"; echo "rc=$?"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > ${O}_smi.csv &
SMI=$!
if [[ " $* " == *" tests "* ]]; then
  T=900 run python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1
  tail -5 ${O}_tests.log
fi
if [[ " $* " == *" bench1 "* ]]; then
  T=300 run python bench.py --steps 20 --warmup 5 > ${O}_bench1.json 2> ${O}_bench1.err
  tail -c 600 ${O}_bench1.json
fi
if [[ " $* " == *" benchN "* ]]; then
  T=400 run $TR bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench${NG}.json 2> ${O}_bench${NG}.err
  tail -c 300 ${O}_bench${NG}.json
fi
if [[ " $* " == *" pull "* ]]; then
  NNDT_FLAG_MODE=pull T=300 run $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-extras > ${O}_bench${NG}_pull.json 2> ${O}_bench${NG}_pull.err
  NNDT_FLAG_MODE=pull T=300 run $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-extras --dtype fp32 > ${O}_bench${NG}_pull32.json 2> ${O}_bench${NG}_pull32.err
  NNDT_FLAG_MODE=push T=300 run $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-extras --dtype fp32 > ${O}_bench${NG}_push32.json 2> ${O}_bench${NG}_push32.err
fi
if [[ " $* " == *" nccl "* ]]; then
  T=300 run $TR bench.py --impl nccl --gpus $NG --steps 50 --warmup 5 > ${O}_nccl${NG}.json 2> ${O}_nccl${NG}.err
  tail -c 400 ${O}_nccl${NG}.json
fi
if [[ " $* " == *" ref "* ]]; then
  T=300 run python bench.py --impl reference --steps 20 --warmup 5 > ${O}_ref1.json 2> ${O}_ref1.err
  tail -c 400 ${O}_ref1.json
fi
kill $SMI
for f in ${O}_*.err; do echo "--- $f"; tail -5 $f; done
