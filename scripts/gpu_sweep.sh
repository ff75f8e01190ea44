mkdir -p gpurun_out

timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
B="timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 10"
for cfg in "M3B_VARIANT=tma" "M3B_VARIANT=tma M3B_TILE=512" "M3B_VARIANT=tma M3B_TILE=256" "M3B_VARIANT=1 M3B_TILE=256" "M3B_VARIANT=tma M3B_TMA_STAGES=4" "M3B_VARIANT=tma M3B_TMA_STAGES=3"; do
  echo "== $cfg" >> gpurun_out/sweep.log
  env $cfg $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['config']['grid_blocks'], d['config']['smem_bytes'], d['e2e']['ms_per_step'], d['llh']['last_value_step'])
    else: print(l.rstrip())
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
