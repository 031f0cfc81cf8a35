set -x
timeout 200 python tools/dev/e2e_once.py 2 > gpurun_out/r2h_plain.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k regex:parse_stream -s 8 -c 2 -o gpurun_out/r2h_parse python tools/dev/e2e_once.py 2 > gpurun_out/r2h_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches.csv python tools/dev/e2e_once.py 2 > gpurun_out/r2h_ncu2.log 2>&1
cat gpurun_out/r2h_plain.log
