#!/bin/bash
# one GPU round-trip: tests + smoke + adapter timing (A), the default bench with its reference arm (C), ncu evidence (B)
cd $GRAFT_REPO_ROOT
bash scripts/gpu_run_a.sh
mkdir -p gpurun_out/c
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c/bench_reference.json 2> gpurun_out/c/bench_reference.err; echo "rc $?" >> gpurun_out/c/bench_reference.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c/bench_default.json 2> gpurun_out/c/bench_default.err; echo "rc $?" >> gpurun_out/c/bench_default.err
cut -c1-600 gpurun_out/c/bench_default.json; tail -3 gpurun_out/c/bench_default.err
bash scripts/gpu_run_b.sh
