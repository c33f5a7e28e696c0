O=gpurun_out/r2w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_config3.py tests/test_gpu_config2_full.py tests/test_gpu_configs.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee $O/summary.txt; tail -2 $O/pytest_acq.log | tee -a $O/summary.txt
python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt
SGX_ACQ_PFA_FWD=0 python tools/quick_acq_bench.py 32 2>&1 | tail -1 | tee -a $O/summary.txt
python tools/quick_acq_bench.py 1 2>&1 | tail -1 | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
