for c in 143 543; do echo "== cfg $c"; SGX_PFA_CFG=$c SGX_ACQ_PROF=1 timeout 200 python tools/quick_acq_bench.py 32 2>&1 | tail -3; done
SGX_PFA_CFG=543 timeout 600 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_configs.py -x -q 2>&1 | tail -1
