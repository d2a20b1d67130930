#!/usr/bin/env bash
# Session 6 (1 GPU): ncu captures of the new kernels, precision parity study, online-density PAPER run on the real floor plan
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
NNDT_MNIST_TC=1 timeout 150 $NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file $O/launches_tc.csv python scripts/profile_round.py 24 > $O/l_tc.log 2>&1
DTYPE=fp64 timeout 150 $NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file $O/launches_f64.csv python scripts/profile_round.py 24 > $O/l_f64.log 2>&1
NNDT_MNIST_TC=1 timeout 200 $NCU --set full --import-source on -k regex:mnist_tc_train -s 8 -c 1 -o $O/mnist_tc python scripts/profile_round.py 8 > $O/p_tc.log 2>&1
DTYPE=fp64 timeout 200 $NCU --set full --import-source on -k regex:mnist_cl64 -s 8 -c 1 -o $O/mnist_cl64 python scripts/profile_round.py 8 > $O/p_c64.log 2>&1
tail -2 $O/p_tc.log $O/p_c64.log
timeout 420 python scripts/precision_parity.py --rounds 2000 --every 100 --algs dinno,dsgt,dsgd --ref-algs dinno > $O/parity.log 2>&1; tail -30 $O/parity.log
cp profiles/precision_parity.md profiles/precision_parity.json $O/ 2>/dev/null
mkdir -p $O/odense
( cd experiments && sed 's#output_metadir: ../results/#output_metadir: ../gpurun_out/odense/#' dist_online_dense_PAPER.yaml > /tmp/od.yaml && timeout 600 python dist_online_dense_ex.py /tmp/od.yaml > ../$O/odense.log 2>&1 )
tail -8 $O/odense.log
for d in $O/odense/*_dist_online_dense_PAPER; do python scripts/summarize_density_run.py $d > $O/odense_summary.md 2>&1; rm -f $d/*_models.pt; done
cat $O/odense_summary.md
du -sh $O
