#!/bin/bash
# Dev-time: one short bench.py line for each of the other BASELINE.json workloads.
for w in "$@"; do
  timeout 300 python bench.py --workload $w --no-extra --no-cpu-baseline --steps 5 > gpurun_out/b_$w.json 2> gpurun_out/b_$w.err
  python - "$w" <<'PY' || tail -3 gpurun_out/b_$w.err
import json, sys
w = sys.argv[1]
d = json.load(open("gpurun_out/b_%s.json" % w))
print(w, "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "d2h GB/s", d["e2e"]["d2h_gbs"],
      "no-download", d["e2e"]["without_frame_download"]["value"], "host parser", d["e2e"]["host_parser_mode"]["value"])
PY
done
