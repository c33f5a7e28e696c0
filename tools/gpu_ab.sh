#!/bin/bash
# A/B of kernel configurations on the 32-recording batch: $1 = environment variable, $2 = values.
V=${1:-SGX_PFA_CFG}
for c in ${2:-543}; do echo "== $V=$c"; env $V=$c SGX_ACQ_PROF=1 timeout 200 python tools/quick_acq_bench.py 32 2>&1 | tail -2; done
