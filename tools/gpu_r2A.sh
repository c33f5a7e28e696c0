O=gpurun_out/r2A; mkdir -p $O
python tools/tc_dft_experiment.py 1676 > $O/tc_dft.jsonl 2> $O/tc_dft.err; cut -c1-200 $O/tc_dft.jsonl; tail -3 $O/tc_dft.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 150 -c 120 --csv --log-file $O/tc_kernels.csv python tools/tc_dft_experiment.py 1676 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2A/tc_kernels.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki][:110],{}).setdefault(r[mi],[]).append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, {m:(len(x), round(sum(x)/len(x),2)) for m,x in v.items()})
PY
