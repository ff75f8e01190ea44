#!/bin/bash
# binned path: parity tests (incl. the layout-independence cases), config 4 timing with and without the eval->fill
# programmatic dependent launch (experiments library for the "without")
cd $GRAFT_REPO_ROOT
O=gpurun_out/g; mkdir -p $O
timeout 600 python -m pytest tests/test_binned_gpu.py tests/test_adapter_gpu.py tests/test_reference_path.py tests/test_selection.py tests/test_shifts.py tests/test_group_gpu.py -m gpu -q --timeout 120 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -3 $O/pytest.log
for v in pdl nopdl pdl2 nopdl2; do
  case $v in nopdl*) E="M3B_LIB=mach3_b200/libm3b200_exp.so M3B_NO_BINNED_PDL=1";; *) E="M3B_LIB=mach3_b200/libm3b200.so";; esac
  env $E timeout 300 python bench.py --workload cfg4 --no-cpu-baseline --steps 50 > $O/bench_cfg4_$v.json 2> $O/bench_cfg4_$v.err
  python -c "
import json; d=json.load(open('$O/bench_cfg4_$v.json')); print('$v', 'step_ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'queued', round(d['extra']['queued']['ms_per_step'],4))"
done
