#!/bin/bash
# Dev-time: everything profiles/ holds for one round, in one gpurun call (one GPU).  Every ncu command runs only after the same
# command line has exited 0 without ncu.   usage: tools/dev/profiles_round.sh <tag>
t=$1; o=gpurun_out; mkdir -p $o
NCU="ncu --clock-control none"
run() { echo "+ $*" >&2; "$@"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${t}_smi.log 2>&1
# -- bench lines
run timeout 600 python bench.py > $o/${t}_bench_n1.json 2> $o/${t}_bench_n1.err || echo "bench rc=$?"
run timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $o/${t}_bench_reference_n1.json 2> $o/${t}_bench_reference_n1.err || echo "ref rc=$?"
# -- launch list of a short bench run
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra"
run timeout 300 $B > $o/${t}_bench_short.json 2> $o/${t}_bench_short.err && \
run timeout 900 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $o/${t}_launches_bench.csv $B > $o/${t}_launches.log 2>&1 || echo "launch list rc=$?"
# -- full captures
run timeout 200 python tools/dev/quick_bench.py --intra-only > $o/${t}_qb_intra.log 2>&1 && \
run timeout 600 $NCU --set full --import-source on -k regex:recon_kernel -s 1 -c 1 -o $o/${t}_intra python tools/dev/quick_bench.py --intra-only > $o/${t}_ncu_intra.log 2>&1 || echo "intra rc=$?"
run timeout 200 python tools/dev/quick_bench.py --ipb-only > $o/${t}_qb_ipb.log 2>&1 && \
run timeout 900 $NCU --set full --import-source on -k regex:recon_kernel -s 7 -c 7 -o $o/${t}_ipb python tools/dev/quick_bench.py --ipb-only > $o/${t}_ncu_ipb.log 2>&1 || echo "ipb rc=$?"
run timeout 200 python tools/dev/e2e_once.py 2 > $o/${t}_e2e_plain.log 2>&1 && \
run timeout 900 $NCU --set full --import-source on -k regex:parse_stream -s 2 -c 1 -o $o/${t}_parse python tools/dev/e2e_once.py 2 > $o/${t}_ncu_parse.log 2>&1 || echo "parse rc=$?"
run timeout 600 $NCU --set full -k regex:scan_ -c 3 -o $o/${t}_scan python tools/dev/e2e_once.py 1 > $o/${t}_ncu_scan.log 2>&1 || echo "scan rc=$?"
run timeout 200 python tools/dev/convert_bench.py > $o/${t}_convert.log 2>&1 && \
run timeout 900 $NCU --set full -k regex:convert_kernel -s 2 -c 1 -o $o/${t}_convert_nv12 python tools/dev/convert_bench.py > $o/${t}_ncu_convert.log 2>&1 || echo "convert rc=$?"
run timeout 200 python tools/dev/quick_bench.py --all > $o/${t}_quick_bench.txt 2>&1
tools/dev/trace_e2e.sh > $o/${t}_trace.log 2>&1
ls -la $o | tail -30
