#!/bin/bash
# One ncu --set full capture (with source correlation) of the kernels matching $2, exported to CSV next to the report.
#   gpurun -- 'bash tools/gpu_ncu_full.sh n1 pfa_search_kernel 4'
O=gpurun_out/${1:-ncu}
K=${2:-pfa_search_kernel}
SKIP=${3:-4}
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -f -o $O/k python tools/quick_acq_bench.py 32 > $O/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/k.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/k.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
ls -la $O
