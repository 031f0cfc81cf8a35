set -x
timeout 200 python tools/dev/e2e_once.py 2 > gpurun_out/${1}_plain.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k regex:parse_stream -s 2 -c 1 -o gpurun_out/${1}_parse python tools/dev/e2e_once.py 2 > gpurun_out/${1}_ncu.log 2>&1
tail -3 gpurun_out/${1}_plain.log
