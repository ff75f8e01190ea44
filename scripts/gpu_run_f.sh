#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/f; mkdir -p $O
timeout 600 python -m pytest tests/test_binned_gpu.py tests/test_adapter_gpu.py tests/test_reference_path.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum -k regex:binned -s 20 -c 6 --csv --log-file $O/launches_cfg4.csv python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $O/ncu_cfg4.log 2>&1
tail -3 $O/pytest.log; tail -2 $O/smoke.log; cut -c1-900 $O/bench_cfg4.json
