set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
cd tools/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pr2 pipe_rates2.cu && /tmp/pr2 > ../../$O/pipe_rates2.jsonl 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pr1 pipe_rates.cu && /tmp/pr1 > ../../$O/pipe_rates.txt 2>&1
cd ../..
python tools/tc_dft_experiment.py 1676 > $O/tc_dft.jsonl 2> $O/tc_dft.err
lscpu > $O/lscpu.txt; numactl -H > $O/numa.txt 2>&1; nvidia-smi topo -m > $O/topo.txt 2>&1
cat $O/pipe_rates2.jsonl; cat $O/tc_dft.jsonl; tail -3 $O/tc_dft.err
