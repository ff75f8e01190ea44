#!/bin/bash
# A/B of the binned fill kernel's scheduling / histogram placement (experiments build of the library)
cd $GRAFT_REPO_ROOT
O=gpurun_out/e; mkdir -p $O
export M3B_LIB=$GRAFT_REPO_ROOT/mach3_b200/libm3b200_exp.so
i=0
for cfg in "M3B_BINNED_GRID_STRIDE=1 M3B_BINNED_SMEM_HIST_MAX_KB=200 M3B_BINNED_MAX_BPS=2" "M3B_BINNED_GRID_STRIDE=0 M3B_BINNED_SMEM_HIST_MAX_KB=200 M3B_BINNED_MAX_BPS=2" \
           "M3B_BINNED_GRID_STRIDE=1 M3B_BINNED_SMEM_HIST_MAX_KB=48" "M3B_BINNED_GRID_STRIDE=0 M3B_BINNED_SMEM_HIST_MAX_KB=48" \
           "M3B_BINNED_GRID_STRIDE=0 M3B_BINNED_SMEM_HIST_MAX_KB=48 M3B_BINNED_MAX_BPS=3" "M3B_BINNED_GRID_STRIDE=0 M3B_BINNED_SMEM_HIST_MAX_KB=48 M3B_BINNED_MAX_BPS=2" \
           "M3B_BINNED_GRID_STRIDE=1 M3B_BINNED_SMEM_HIST_MAX_KB=48 M3B_BINNED_MAX_BPS=2"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --workload cfg4 --no-cpu-baseline --steps 30 > $O/cfg4_$i.json 2> $O/cfg4_$i.err
  echo "$cfg" > $O/cfg4_$i.cfg
  python -c "
import json,sys
j=json.loads(open('$O/cfg4_$i.json').read().strip().splitlines()[-1])
print('$cfg', 'kernel_ms', round(j['roofline']['kernel_ms'],4), 'sync', round(j['ms_per_step'],4))"
done
