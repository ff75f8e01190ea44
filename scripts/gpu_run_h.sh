#!/bin/bash
# batched-proposal kernel: parity tests, then config 5 timing incl. the LLH-scan leg (tight timeouts: a ring-protocol bug hangs the kernel)
cd $GRAFT_REPO_ROOT
O=gpurun_out/h; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -x --timeout 90 -k "batch or Batch or fitters or delayed or scan or toys" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -3 $O/pytest.log
grep -q "pytest rc 0" $O/pytest.log || exit 1
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err
python -c "
import json; d=json.load(open('$O/bench_cfg5.json')); print('step_ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4)); print(d['config']['llh_scan_256_points'], d['config']['first_generation_kernel_ms'])"
