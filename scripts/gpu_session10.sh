#!/usr/bin/env bash
# Session 10 (2 GPUs): where do the +18 us/round of the fp64 arm at 2 GPUs come from?  A/B switches, --no-extras
NG=2; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541"
run() { tag=$1; shift; env "$@" timeout 200 $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-extras > gpurun_out/s10_$tag.json 2> gpurun_out/s10_$tag.err; echo "$tag rc=$? $(grep -o '"ms_per_step": [0-9.e-]*' gpurun_out/s10_$tag.json | head -2 | tr '\n' ' ')"; }
run default NNDT_X=1
run disjoint NNDT_BENCH_GRAPH=disjoint
run resident NNDT_BENCH_PIPELINE=resident
run pull NNDT_FLAG_MODE=pull
run seppub NNDT_SEPARATE_PUBLISH=1
run nopdl NNDT_NO_PDL=1
run split1 NNDT_TC_SPLIT=1
timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/s10_one.json 2> gpurun_out/s10_one.err; echo "1gpu $(grep -o '"ms_per_step": [0-9.e-]*' gpurun_out/s10_one.json | head -2 | tr '\n' ' ')"
NNDT_BENCH_PIPELINE=resident timeout 200 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/s10_one_res.json 2> gpurun_out/s10_one_res.err; echo "1gpu resident $(grep -o '"ms_per_step": [0-9.e-]*' gpurun_out/s10_one_res.json | head -2 | tr '\n' ' ')"
# per-round timelines of the update kernels
timeout 120 python scripts/timeline_rounds.py --dtype fp64 --out gpurun_out/s10_tl > gpurun_out/s10_tl1.log 2>&1; echo "tl1 rc=$?"
timeout 200 $TR scripts/timeline_rounds.py --dtype fp64 --out gpurun_out/s10_tl > gpurun_out/s10_tl2.log 2>&1; echo "tl2 rc=$?"
timeout 200 $TR scripts/timeline_rounds.py --dtype fp32 --out gpurun_out/s10_tl > gpurun_out/s10_tl2f.log 2>&1; echo "tl2f rc=$?"
NNDT_BENCH_PIPELINE=resident timeout 200 $TR scripts/timeline_rounds.py --dtype fp64 --out gpurun_out/s10_tlres > gpurun_out/s10_tl2res.log 2>&1; echo "tl2res rc=$?"
