#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/k; mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_adapter_gpu.py -m gpu -q --timeout 120 -k "indexed or adapter or osc" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python -c "
import argparse, json, bench
a = argparse.Namespace(tile=0)
print(json.dumps(bench.measure_binned_osc(a, 0, 5, 100)))" 2>&1 | tail -2 | tee $O/binned_osc.json
timeout 300 oracle/_ref/adapter_test 1000000 poisson 1 time 2>&1 | tail -5
