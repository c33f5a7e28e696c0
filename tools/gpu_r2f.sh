O=gpurun_out/r2f; mkdir -p $O; : > $O/summary.txt
for cfg in 62 43 121; do for pers in 0 1; do
  echo "--- cfg=$cfg persist=$pers" | tee -a $O/summary.txt
  SGX_DEBUG=1 SGX_PFA_CFG=$cfg SGX_PFA_PERSIST=$pers python tools/quick_acq_bench.py 32 2>&1 | grep -E "persisting|R=32" | tail -2 | tee -a $O/summary.txt
  SGX_PFA_CFG=$cfg SGX_PFA_PERSIST=$pers timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pfa_search_kernel -s 2 -c 1 --csv python tools/quick_acq_bench.py 32 2>/dev/null | grep -E "pfa_search" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | tr '\n' ';' | tee -a $O/summary.txt; echo | tee -a $O/summary.txt
done; done
echo "--- cfg=121 persist 48MB" | tee -a $O/summary.txt; SGX_PFA_CFG=121 SGX_PFA_PERSIST=1 SGX_PFA_PERSIST_MB=48 python tools/quick_acq_bench.py 32 2>&1 | tail -1 | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pfa_search_kernel -s 2 -c 1 -o $O/pfa python tools/quick_acq_bench.py 32 > $O/ncu_pfa.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
