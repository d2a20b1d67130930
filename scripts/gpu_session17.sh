#!/usr/bin/env bash
# Session 17 (1 GPU): final 1-GPU bench line; ncu re-captures of the kernels changed late in the round
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
timeout 200 python bench.py --steps 20 --warmup 5 > $O/s17_bench1.json 2> $O/s17_bench1.err; echo "bench1 rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' $O/s17_bench1.json | head -4 | tr '\n' ' ')"
timeout 100 $NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file $O/launches_tc.csv python scripts/profile_round.py 24 > $O/l_tc.log 2>&1
DTYPE=fp64 timeout 100 $NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file $O/launches_f64.csv python scripts/profile_round.py 24 > $O/l_f64.log 2>&1
DTYPE=fp64 timeout 150 $NCU --set full --import-source on -k regex:mnist_cl64 -s 8 -c 1 -f -o $O/mnist_cl64 python scripts/profile_round.py 8 > $O/p_c64.log 2>&1
DTYPE=fp64 timeout 150 $NCU --set full --import-source on -k regex:dinno_update -s 8 -c 1 -f -o $O/dinno_update_f64 python scripts/profile_round.py 8 > $O/p_u64.log 2>&1
timeout 150 $NCU --set full --import-source on -k regex:dinno_update -s 8 -c 1 -f -o $O/dinno_update python scripts/profile_round.py 8 > $O/p_u32.log 2>&1
tail -1 $O/l_tc.log $O/l_f64.log $O/p_c64.log $O/p_u64.log $O/p_u32.log
