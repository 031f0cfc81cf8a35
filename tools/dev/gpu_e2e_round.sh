#!/bin/bash
# Dev-time: one gpurun call of an end-to-end iteration: smoke, GPU tests, e2e profile.   usage: tools/dev/gpu_e2e_round.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/${tag}_tests.log
timeout 600 python tools/dev/e2e_profile.py $2 > gpurun_out/${tag}_e2e.log 2>&1; echo "e2e rc=$?"; cat gpurun_out/${tag}_e2e.log
