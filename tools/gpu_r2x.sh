O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_config2_full.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee $O/summary.txt; tail -2 $O/pytest_acq.log | tee -a $O/summary.txt
python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fine_cols_kernel" -s 2 -c 1 -o $O/cols python tools/quick_acq_bench.py 32 > $O/ncu_cols.log 2>&1; echo "ncu rc=$?"
