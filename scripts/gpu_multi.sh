# usage: bash scripts/gpu_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu/parity_ranks.py > gpurun_out/multi_parity_$N.log 2>&1; echo "parity exit $?" >> gpurun_out/multi_parity_$N.log
grep -v "^W\|^\[W\|warn" gpurun_out/multi_parity_$N.log | tail -14
for ex in nccl peer; do
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 --exchange $ex > gpurun_out/bench_cfg3_${N}_$ex.json 2> gpurun_out/bench_cfg3_${N}_$ex.err; echo "bench $ex exit $?"
  tail -c 3000 gpurun_out/bench_cfg3_${N}_$ex.json; tail -5 gpurun_out/bench_cfg3_${N}_$ex.err
done
