#!/bin/bash
# ncu --set full of the batch kernel (config 5 at 600 k events) on the final tree, after the same command ran clean
cd $GRAFT_REPO_ROOT
O=gpurun_out/b; mkdir -p $O
timeout 600 python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $O/bench_cfg5_600k.json 2> $O/bench_cfg5_600k.err && \
timeout 600 ncu --clock-control none --set full --import-source on -k regex:fill_batch2_kernel -c 2 -o $O/full_fill_batch2_cfg5_600k -f python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $O/ncu_full_cfg5.log 2>&1
echo "rc $?"; tail -2 $O/ncu_full_cfg5.log
