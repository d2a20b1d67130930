#!/usr/bin/env bash
# memcheck / racecheck of the round-2 cluster kernels (mnist_tc: tcgen05 + TMA + DSMEM; mnist_cl64: fp64 + DSMEM) and of the
# generic conv-net / MLP kernels, on small batches.  Run under gpurun (1 GPU).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
K='tc_kernel_matches_batch_split_kernel_and_autograd[8-False] or fp64_cluster_kernel_matches_generic_kernel_and_autograd[8]'
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 \
      python -m pytest tests/test_gpu_mnist.py -q -x -k "$K" > gpurun_out/sanitize_clusters_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitize_clusters_$tool.log | tail -2 | tr '\n' ' ')"
done
