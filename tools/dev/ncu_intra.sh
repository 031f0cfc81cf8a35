set -x
timeout 200 python tools/dev/quick_bench.py --intra-only > gpurun_out/${1}_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:recon_kernel -s 1 -c 1 -o gpurun_out/${1}_intra python tools/dev/quick_bench.py --intra-only > gpurun_out/${1}_ncu.log 2>&1
timeout 200 python tools/dev/quick_bench.py --ipb-only > gpurun_out/${1}_plain_ipb.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k regex:recon_kernel -s 7 -c 7 -o gpurun_out/${1}_ipb python tools/dev/quick_bench.py --ipb-only > gpurun_out/${1}_ncu_ipb.log 2>&1
cat gpurun_out/${1}_plain.log gpurun_out/${1}_plain_ipb.log
