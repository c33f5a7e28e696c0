#!/bin/bash
# Quick GPU check of an acquisition change: timing (with the per-stage stream times), the acquisition parity tests and one
# ncu metric pass on the kernels matching $2 (default: the search kernel).
#   gpurun -- 'bash tools/gpu_quick.sh q1 "fine_cols_kernel|fine_rows_kernel"'
O=gpurun_out/${1:-quick}
K=${2:-pfa_search_kernel}
mkdir -p $O
SGX_ACQ_PROF=1 timeout 200 python tools/quick_acq_bench.py 32 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py tests/test_gpu_config3.py tests/test_gpu_config2_full.py -x -q 2>&1 | tail -2
M=gpu__time_duration.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"$K" -s 6 -c 4 python tools/quick_acq_bench.py 32 2>&1 | grep -E "^  [a-z].*\(|duration|scoreboard|throttle|barrier|inst_executed|issue_active|wavefronts|conflicts" | tail -44
