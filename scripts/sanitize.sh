#!/usr/bin/env bash
# Race / memory checks of the fused kernels on the single-GPU virtual-node backend.
# Run under gpurun (1 GPU).  compute-sanitizer slows kernels ~50x: the targets are small on purpose.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
      python -m pytest tests/test_gpu_mnist.py -q -x -k "fwdbwd_matches_autograd[24-False] or fused_training_matches_torch_ops[cycle-DSGD" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$tool.log | tail -2 | tr '\n' ' ')"
done
