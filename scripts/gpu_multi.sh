# usage: bash scripts/gpu_multi2.sh N   (under gpurun --gpus N): parity + default-exchange bench + nccl bench
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu/parity_ranks.py > gpurun_out/multi_parity_$N.log 2>&1; echo "parity exit $?" >> gpurun_out/multi_parity_$N.log
grep "PARITY\|FAIL\|parity exit" gpurun_out/multi_parity_$N.log
for ex in auto nccl; do
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 --exchange $ex > gpurun_out/bench_cfg3_${N}_$ex.json 2> gpurun_out/bench_cfg3_${N}_$ex.err; echo "bench $ex exit $?"
  grep "^{" gpurun_out/bench_cfg3_${N}_$ex.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config']['exchange'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'GB/s', d['roofline']['achieved'], 'e2e ms', d['e2e']['ms_per_step'], 'value', d['value'])"
done
