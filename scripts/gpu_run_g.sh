#!/bin/bash
# binned fill: block-size A/B on config 4 (experiments library), after the parity tests on the product library
cd $GRAFT_REPO_ROOT
O=gpurun_out/g; mkdir -p $O
timeout 600 python -m pytest tests/test_binned_gpu.py tests/test_adapter_gpu.py tests/test_reference_path.py -m gpu -q --timeout 200 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -3 $O/pytest.log
for nt in 1024 768 1024 768; do
  M3B_LIB=mach3_b200/libm3b200_exp.so M3B_BINNED_THREADS=$nt timeout 300 python bench.py --workload cfg4 --no-cpu-baseline --steps 50 > $O/bench_cfg4_t$nt.json 2> $O/bench_cfg4_t$nt.err
  python -c "
import json; d=json.load(open('$O/bench_cfg4_t$nt.json')); print('threads $nt', 'step_ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"
done
