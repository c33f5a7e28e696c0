#!/bin/bash
# Acquisition iteration: parity tests of the acquisition path + timing (quick bench, 32 recordings) + bench line.
O=gpurun_out/${1:-acq}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py -x -q > $O/pytest_acq.log 2>&1; echo "acq tests rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_acq.log | tee -a $O/summary.txt
timeout 300 python tools/quick_acq_bench.py 32 > $O/quick32.log 2>&1; tail -3 $O/quick32.log | tee -a $O/summary.txt
for mb in 64 128; do SGX_ACQ_CHUNK_MB=$mb timeout 300 python tools/quick_acq_bench.py 32 2>&1 | tail -1 | sed "s/^/chunk $mb: /" | tee -a $O/summary.txt; done
timeout 600 python bench.py --no-e2e --no-cpu > $O/bench_noe2e.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'E' | tee -a $O/summary.txt
import json,sys,glob,os
p=sorted(glob.glob('gpurun_out/*/bench_noe2e.json'), key=os.path.getmtime)[-1]
d=json.loads(open(p).read().strip().splitlines()[-1])
print('track ms', d['ms_per_step'], 'acq ms', d['secondary']['ms_per_step'], 'cells/s %.4g' % d['secondary']['value'], 'detected', d['secondary']['config']['detected'])
E
