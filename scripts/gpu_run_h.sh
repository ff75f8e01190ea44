#!/bin/bash
# batch kernel timing ablations (experiments library)
cd $GRAFT_REPO_ROOT
O=gpurun_out/h; mkdir -p $O
M3B_LIB=mach3_b200/libm3b200_exp.so timeout 400 python scripts/batch_ablation.py 1200000 2>&1 | tee $O/ablation.txt | tail -20
