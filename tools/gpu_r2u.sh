# round 2 final session on one B200: suite, smoke, both bench arms, launch list of the bench command
O=gpurun_out/r2u; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $O/summary.txt; tail -2 $O/pytest_gpu.log | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py > $O/bench_gpu.json 2> $O/bench_gpu.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
