#!/bin/bash
# A/B of search-kernel configurations (SGX_PFA_CFG values in $1) on the 32-recording batch.
for c in ${1:-543}; do echo "== cfg $c"; SGX_PFA_CFG=$c SGX_ACQ_PROF=1 timeout 200 python tools/quick_acq_bench.py 32 2>&1 | tail -2; done
