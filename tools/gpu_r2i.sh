O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $O/summary.txt
tail -15 $O/pytest_gpu.log | tee -a $O/summary.txt
