#!/bin/bash
# Dev-time: end-to-end decode time (1080p IPB texture, 120 frames) against the parser placement knobs
for dl in "--download" ""; do
  for ps in ${PSTREAMS:-2 4}; do
    for lanes in ${LANES:-0 1 2}; do
      r=$(MP2V_PARSE_STREAMS=$ps MP2V_PARSE_LANES=$lanes timeout 120 python tools/dev/e2e_once.py 6 $dl --resident 2>&1 | grep -E "^decode|^resident best" | tail -4 | tr '\n' ' ')
      echo "download='$dl' parse_streams=$ps lanes=$lanes : $r"
    done
  done
done
