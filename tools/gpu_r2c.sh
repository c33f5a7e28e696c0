# round 2, call c: warp-slice prime-factor search kernel
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_acq.log | tee -a $O/summary.txt
for c in 3 2; do echo "--- ctas/sm<=$c" | tee -a $O/summary.txt; SGX_PFA_CTAS_PER_SM=$c python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pfa_search_kernel -s 2 -c 1 -o $O/pfa python tools/quick_acq_bench.py 32 > $O/ncu_pfa.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
