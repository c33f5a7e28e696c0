#!/bin/bash
bash tools/gpu_ab.sh SGX_PFA_CFG "543 643"
SGX_PFA_CFG=643 timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py -x -q 2>&1 | tail -1
for c in 543 643; do
SGX_PFA_CFG=$c timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pfa_search_kernel -s 4 -c 1 python tools/quick_acq_bench.py 32 2>&1 | grep -E "duration|dram|hit_rate|scoreboard|issue_active"
done
