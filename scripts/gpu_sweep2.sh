mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --steps 200 --warmup 20"
rm -f gpurun_out/sweep2.log
for cfg in "M3B_X=0" "M3B_TMA_BLOCKS_PER_SM=2" "M3B_TMA_BLOCKS_PER_SM=2 M3B_TILE=512" "M3B_TILE=512" "M3B_TMA_STAGES=5" "M3B_X=1"; do
  echo "== $cfg" >> gpurun_out/sweep2.log
  env $cfg $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['config']['grid_blocks'], d['config']['smem_bytes'], d['config'].get('tma_stages'), d['e2e']['ms_per_step'], d['llh']['last_value_step'])
    else: print(l.rstrip())
" >> gpurun_out/sweep2.log
done
cat gpurun_out/sweep2.log
