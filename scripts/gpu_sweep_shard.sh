mkdir -p gpurun_out
B="timeout 400 python bench.py --no-cpu-baseline --workload cfg3 --events 2500608 --steps 100 --warmup 10"
rm -f gpurun_out/sweep4.log
for cfg in "M3B_TILE=1024" "M3B_TILE=512" "M3B_TILE=1024 M3B_GUARD_X2=4" "M3B_TILE=1024 M3B_GUARD_X2=10" "M3B_TILE=512 M3B_GUARD_X2=10"; do
  echo "== $cfg" >> gpurun_out/sweep4.log
  env $cfg $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['config'].get('tma_stages'), d['e2e']['ms_per_step'])
    else: print(l.rstrip()[:200])
" >> gpurun_out/sweep4.log
done
cat gpurun_out/sweep4.log
