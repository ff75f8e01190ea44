#!/bin/bash
# GPU round-trip D: tests + smoke + quick bench lines of configs 2, 4 and 5 (600 k events, both batch-kernel generations)
cd $GRAFT_REPO_ROOT
bash scripts/gpu_run_a.sh
O=gpurun_out/d; mkdir -p $O
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 600 python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $O/bench_cfg5_600k.json 2> $O/bench_cfg5_600k.err
for f in $O/bench_*.json; do echo $f; cut -c1-300 $f; done
