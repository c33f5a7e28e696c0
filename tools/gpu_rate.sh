#!/bin/bash
# Rate-sweep parity tests and the config-5 sweep (prime-factor path at N = 16 368 vs the generic passes).
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_acquisition.py -x -q 2>&1 | tail -2
mkdir -p gpurun_out/rate
timeout 600 python tools/sweep_c5.py gpurun_out/rate/sweep.json 2>&1 | grep -E '"fs_mhz": 16.3676' | cut -c1-260
SGX_ACQ_PFA=0 timeout 600 python tools/sweep_c5.py gpurun_out/rate/sweep_generic.json 2>&1 | grep -E '"fs_mhz": 16.3676, "samples_per_code": 16368, "coherent_ms": 1,' | cut -c1-260
