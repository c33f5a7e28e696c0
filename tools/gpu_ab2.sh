#!/bin/bash
# A/B of two search-kernel variants: timing, parity tests and an ncu metric pass for each.  $1 = "cfgA cfgB"
for c in ${1:-543}; do
  echo "== SGX_PFA_CFG=$c"
  SGX_PFA_CFG=$c SGX_ACQ_PROF=1 timeout 200 python tools/quick_acq_bench.py 32 2>&1 | tail -2
  SGX_PFA_CFG=$c timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio --clock-control none -k regex:pfa_search_kernel -s 4 -c 1 python tools/quick_acq_bench.py 32 2>&1 | grep -E "duration|inst_executed|issue_active|wavefronts|scoreboard|throttle"
done
c=$(echo $1 | awk '{print $NF}')
SGX_PFA_CFG=$c timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py tests/test_gpu_config3.py -x -q 2>&1 | tail -1
