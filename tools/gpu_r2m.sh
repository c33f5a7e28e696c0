O=gpurun_out/r2m; mkdir -p $O
python tools/sweep_c5.py $O/sweep_r2.json 8 > $O/sweep.log 2>&1; tail -3 $O/sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fft_pass_kernel|fft_pass_async_kernel" -s 40 -c 8 -o $O/fine python tools/quick_acq_bench.py 32 > $O/ncu_fine.log 2>&1; echo "ncu rc=$?"
