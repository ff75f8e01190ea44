#!/bin/bash
# binned fill: parity, config 4 timing, ncu --set full of the fill kernel, A/B of the front depth at 1024 threads
cd $GRAFT_REPO_ROOT
O=gpurun_out/g; mkdir -p $O
timeout 600 python -m pytest tests/test_binned_gpu.py tests/test_adapter_gpu.py tests/test_reference_path.py tests/test_selection.py tests/test_shifts.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python -c "
import json; d=json.load(open('$O/bench_cfg4.json')); print('default', 'step_ms', round(d['ms_per_step'],4), d['roofline'])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:binned_fill -s 10 -c 1 -f -o $O/full_fill1024 python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $O/ncu_full.log 2>&1
