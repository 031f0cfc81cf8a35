#!/bin/bash
# Dev-time: parse lot size x concurrent parse streams, with and without the frame download
for dl in "" "--download"; do
  for ps in 1 2 4; do
    for lot in 16 32 64; do
      r=$(MP2V_PARSE_STREAMS=$ps MP2V_LOT=$lot timeout 120 python tools/dev/e2e_once.py 6 $dl 2>&1 | grep "^decode" | tail -3 | tr '\n' ' ')
      echo "download='$dl' parse_streams=$ps lot=$lot : $r"
    done
  done
done
