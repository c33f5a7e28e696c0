O=gpurun_out/r2o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py tests/test_gpu_config3.py tests/test_gpu_config2_full.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_acq.log | tee -a $O/summary.txt
python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt
SGX_ACQ_FINE2=0 python tools/quick_acq_bench.py 32 2>&1 | tail -1 | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fine_cols_kernel|fine_rows_kernel" -s 2 -c 2 -o $O/fine2 python tools/quick_acq_bench.py 32 > $O/ncu_fine.log 2>&1; echo "ncu rc=$?"
