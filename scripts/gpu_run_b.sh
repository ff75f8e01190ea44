#!/bin/bash
# GPU round-trip B: ncu evidence for the shipped configurations (each capture only after the same command ran clean)
cd $GRAFT_REPO_ROOT
O=gpurun_out/b; mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of the quick config-2 bench (every kernel, durations only)
timeout 600 python bench.py --workload cfg2 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err && \
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --workload cfg2 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_cfg2.log 2>&1
# 2. --set full of the fill kernel, shipped config-2 configuration
timeout 900 $NCU --set full --import-source on -k regex:fill_tma_kernel -s 20 -c 3 -o $O/full_fill_tma_cfg2 -f python bench.py --workload cfg2 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_full_cfg2.log 2>&1
# 3. the same at the 8-GPU shard size of config 3 (2.5 M events, 60 responses) and launch list
timeout 900 python bench.py --workload cfg3 --events 2500608 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg3_shard.json 2> $O/bench_cfg3_shard.err && \
timeout 900 $NCU --set full --import-source on -k regex:fill_tma_kernel -s 20 -c 3 -o $O/full_fill_tma_cfg3_shard -f python bench.py --workload cfg3 --events 2500608 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_full_cfg3.log 2>&1
# 4. batched proposals (config 5 at 600 k events): both kernel generations timed, then --set full of the new one
timeout 900 python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $O/bench_cfg5_600k.json 2> $O/bench_cfg5_600k.err && \
timeout 900 $NCU --set full --import-source on -k regex:fill_batch2_kernel -c 2 -o $O/full_fill_batch2_cfg5_600k -f python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $O/ncu_full_cfg5.log 2>&1
# 5. binned-spline workload (config 4, full size): launch list with DRAM / L2 sector counts, then --set full of the fill
timeout 900 python bench.py --workload cfg4 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err && \
timeout 900 $NCU --set full --import-source on -k regex:binned_fill_kernel -s 10 -c 2 -o $O/full_binned_fill_cfg4 -f python bench.py --workload cfg4 --no-cpu-baseline > $O/ncu_full_cfg4.log 2>&1
ls -la $O; for f in $O/bench_*.json; do echo $f; cut -c1-400 $f; done
