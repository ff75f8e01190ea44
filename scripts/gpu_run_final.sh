#!/bin/bash
# the round's validation pass on one B200: full GPU test suite, smoke(), adapter timing, the reference arm and the default
# bench (config 3 + sub-records), then the ncu evidence for the kernels that changed last (each capture only after the same
# command ran clean without ncu)
cd $GRAFT_REPO_ROOT
bash scripts/gpu_run_a.sh
O=gpurun_out/c; mkdir -p $O
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc $?" >> $O/bench_reference.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "rc $?" >> $O/bench_default.err
cut -c1-600 $O/bench_default.json; tail -3 $O/bench_default.err
B=gpurun_out/b; mkdir -p $B
NCU="ncu --clock-control none"
timeout 600 python bench.py --workload cfg2 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $B/bench_cfg2.json 2> $B/bench_cfg2.err && \
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $B/launches_cfg2.csv python bench.py --workload cfg2 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $B/ncu_cfg2.log 2>&1
timeout 600 python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $B/bench_cfg5_600k.json 2> $B/bench_cfg5_600k.err && \
timeout 600 $NCU --set full --import-source on -k regex:fill_batch2_kernel -c 2 -o $B/full_fill_batch2_cfg5_600k -f python bench.py --workload cfg5 --events 600000 --no-cpu-baseline > $B/ncu_full_cfg5.log 2>&1
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline > $B/bench_cfg4.json 2> $B/bench_cfg4.err && \
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:binned -s 20 -c 6 --csv --log-file $B/launches_cfg4.csv python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $B/ncu_cfg4.log 2>&1
ls $B
