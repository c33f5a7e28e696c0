#!/bin/bash
# One consolidated GPU session: full -m gpu suite, smoke, bench (both arms), ncu launch list and full captures.
O=gpurun_out/${1:-s4}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py > $O/bench_gpu.json 2> $O/bench_gpu.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_pass_async_kernel -s 30 -c 2 -o $O/fftpass python tools/quick_acq_bench.py 32 > $O/ncu_fft.log 2>&1; echo "ncu fft rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt
