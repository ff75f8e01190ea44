#!/bin/bash
# GPU round-trip A: the full GPU test suite, smoke(), a quick config-2 bench line and the adapter's real step cost
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/a
nvidia-smi -L > gpurun_out/a/smi.txt 2>&1
nproc >> gpurun_out/a/smi.txt
free -g >> gpurun_out/a/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/a/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/a/smoke.log
timeout 600 python bench.py --workload cfg2 --extras none --steps 20 --warmup 5 > gpurun_out/a/bench_cfg2.json 2> gpurun_out/a/bench_cfg2.err; echo "rc $?" >> gpurun_out/a/bench_cfg2.err
timeout 300 oracle/_ref/adapter_test 1000000 poisson 1 time > gpurun_out/a/adapter_time_1m.log 2>&1
timeout 300 oracle/_ref/adapter_test 1000000 poisson 2 time > gpurun_out/a/adapter_time_1m_2members.log 2>&1
tail -15 gpurun_out/a/pytest.log; tail -4 gpurun_out/a/smoke.log; cut -c1-700 gpurun_out/a/bench_cfg2.json; tail -3 gpurun_out/a/bench_cfg2.err; tail -4 gpurun_out/a/adapter_time_1m.log
