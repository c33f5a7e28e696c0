# round 2, call b: prime-factor search kernel -- parity on hardware, timing per configuration, one ncu capture
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_fft.py tests/test_gpu_configs.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_acq.log | tee -a $O/summary.txt
echo "--- old path" | tee -a $O/summary.txt; SGX_ACQ_PFA=0 python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt
for cfg in "1 4" "1 3" "1 2" "2 2" "2 1" "4 1"; do set -- $cfg; echo "--- groups=$1 ctas/sm<=$2" | tee -a $O/summary.txt; SGX_PFA_GROUPS=$1 SGX_PFA_CTAS_PER_SM=$2 python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pfa_search_kernel -s 2 -c 1 -o $O/pfa python tools/quick_acq_bench.py 32 > $O/ncu_pfa.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
