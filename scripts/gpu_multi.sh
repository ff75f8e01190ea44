#!/bin/bash
# multi-GPU round-trip on N GPUs of one box: the group / sharding tests on distinct devices, the adapter spread over N devices,
# then the bench line at N (torchrun, one rank per GPU) with its parity check and its single-process record
N=${1:-2}
cd $GRAFT_REPO_ROOT
O=gpurun_out/multi$N; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nvidia-smi topo -m >> $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_multigpu.py tests/test_shifts.py tests/test_monolith_file.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 oracle/_ref/adapter_test 1000000 poisson $N time > $O/adapter_${N}members.log 2>&1; echo "rc $?" >> $O/adapter_${N}members.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "rc $?" >> $O/bench.err
tail -4 $O/pytest.log; tail -5 $O/adapter_${N}members.log; tail -c 1500 $O/bench.json; tail -5 $O/bench.err
