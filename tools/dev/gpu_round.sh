#!/bin/bash
# Dev-time: what one gpurun call of a kernel iteration runs.  Every step under its own timeout; logs under gpurun_out/.
# usage: tools/dev/gpu_round.sh <tag> [variant.so ...]
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 300 python tools/dev/quick_bench.py --all > gpurun_out/${tag}_qb.log 2>&1; echo "qb rc=$?"; cat gpurun_out/${tag}_qb.log
for v in "$@"; do
  n=$(basename $v .so)
  MP2V_B200_LIB=$PWD/tiny_mp2v_dec_b200/_lib/variants/$n.so timeout 300 python tools/dev/quick_bench.py > gpurun_out/${tag}_qb_$n.log 2>&1; echo "== variant $n"; cat gpurun_out/${tag}_qb_$n.log
done
