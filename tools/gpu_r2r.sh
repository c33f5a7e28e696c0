O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/quick_acq_bench.py 32 2>&1 | tail -2 | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file $O/acq_launches.csv python tools/quick_acq_bench.py 32 > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pfa_search_kernel|fine_cols_kernel|fine_rows_kernel" -s 4 -c 4 -o $O/acq python tools/quick_acq_bench.py 32 > $O/ncu_acq.log 2>&1; echo "ncu rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 2 -c 1 -o $O/trk python tools/quick_track_bench.py 32 300 > $O/ncu_trk.log 2>&1; echo "ncu trk rc=$?"
