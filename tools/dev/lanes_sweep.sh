#!/bin/bash
# Dev-time: slices per parser warp x parse lot size, on the 1080p IPB texture workload (no frame download)
for lanes in ${LANES:-1 2 4}; do
  for lot in ${LOTS:-64 128}; do
    r=$(MP2V_PARSE_LANES=$lanes MP2V_LOT=$lot timeout 120 python tools/dev/e2e_once.py 5 --resident 2>&1 | grep -E "^decode|^resident best" | tail -3 | tr '\n' ' ')
    echo "lanes=$lanes lot=$lot : $r"
  done
done
