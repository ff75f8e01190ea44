#!/bin/bash
# microbenchmark + ncu --set full of the fill kernel at config 3's full single-GPU size (the headline bench's launch configuration)
cd $GRAFT_REPO_ROOT
O=gpurun_out/i; mkdir -p $O
timeout 120 scripts/microbench/ffma2_rate | tee $O/ffma2_rate.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fill_tma_kernel -s 6 -c 1 -f -o $O/full_fill_tma_cfg3_1gpu python bench.py --workload cfg3 --extras none --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full_cfg3.log 2>&1; echo "ncu rc $?"
tail -2 $O/ncu_full_cfg3.log
