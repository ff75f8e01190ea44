mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
M3B_TRACE_LAST=1 timeout 200 python scripts/trace_fill.py 2>&1 | tail -12
B="timeout 300 python bench.py --no-cpu-baseline --steps 200 --warmup 20"
rm -f gpurun_out/sweep3.log
for cfg in "M3B_GUARD_X2=4" "M3B_GUARD_X2=2" "M3B_GUARD_X2=1" "M3B_GUARD_X2=6" "M3B_GUARD_X2=4 M3B_TILE=512" "M3B_GUARD_X2=2 M3B_TILE=512" "M3B_GUARD_X2=1 M3B_TILE=512" "M3B_GUARD_X2=4 M3B_TILE=256"; do
  echo "== $cfg" >> gpurun_out/sweep3.log
  env $cfg $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['config'].get('tma_stages'), d['e2e']['ms_per_step'], d['llh']['last_value_step'])
    else: print(l.rstrip())
" >> gpurun_out/sweep3.log
done
cat gpurun_out/sweep3.log
