O=gpurun_out/r2y; mkdir -p $O
python tools/tc_dft_experiment.py 1676 > $O/tc_dft.jsonl 2> $O/tc_dft.err; cat $O/tc_dft.jsonl | cut -c1-220
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed.sum --clock-control none -k regex:"gemm|cutlass|sm100|sm90|xmma|nvjet" -c 40 --csv --log-file $O/tc_kernels.csv python tools/tc_dft_experiment.py 1676 > /dev/null 2>&1; echo "ncu rc=$?"
cut -d, -f5,13,15 $O/tc_kernels.csv | sort | uniq -c | sort -rn | head -30
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -2 $O/pytest_gpu.log
