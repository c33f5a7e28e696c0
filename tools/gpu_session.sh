#!/bin/bash
# One consolidated GPU session (gpurun, one B200): full -m gpu suite, smoke, both bench arms, the launch list of the
# bench command, and the ncu captures of the hot kernels.  Output: gpurun_out/<name>/ (scratch; copy what should be judged
# into profiles/).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r2'
O=gpurun_out/${1:-session}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py > $O/bench_gpu.json 2> $O/bench_gpu.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pfa_search_kernel|pfa_forward_kernel|fine_cols_kernel|fine_rows_kernel" -s 7 -c 7 -f -o $O/acq python tools/quick_acq_bench.py 32 > $O/ncu_acq.log 2>&1; echo "ncu acq rc=$?" | tee -a $O/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 2 -c 1 -o $O/trk python tools/quick_track_bench.py 32 300 > $O/ncu_trk.log 2>&1; echo "ncu trk rc=$?" | tee -a $O/summary.txt
# gpurun brings back at most 64 MiB: keep the CSV exports (raw metrics per launch, per-instruction source page of the
# search kernel), drop the reports
for r in acq trk; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ncu -i $O/acq.ncu-rep --page source --csv -k regex:pfa_search_kernel > $O/acq_search_src.csv 2>/dev/null
rm -f $O/acq.ncu-rep $O/trk.ncu-rep
du -sh $O
cat $O/summary.txt
