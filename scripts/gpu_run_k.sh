#!/bin/bash
# indexed oscillator table path: full GPU suite, the binned-oscillator e2e step, adapter timing, and the headline kernels for regressions
cd $GRAFT_REPO_ROOT
O=gpurun_out/k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "
import argparse, json, bench
a = argparse.Namespace(tile=0)
print(json.dumps(bench.measure_binned_osc(a, 0, 5, 100)))" 2>&1 | tail -1 | cut -c1-260 | tee $O/binned_osc.json
timeout 300 oracle/_ref/adapter_test 1000000 poisson 1 time 2>&1 | grep "adapter step cost" | cut -c1-200
timeout 600 python bench.py --workload cfg2 --extras none --steps 50 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --workload cfg3 --extras none --steps 20 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
for c in cfg2 cfg3; do python -c "
import json; d=json.load(open('$O/bench_$c.json')); print('$c', 'step_ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"; done
