#!/bin/bash
for v in "" cta32 cta64; do
  for dl in "" "--download"; do
    for ps in 2 4; do
      lib=""; [ -n "$v" ] && lib=$PWD/tiny_mp2v_dec_b200/_lib/variants/$v.so
      r=$(MP2V_B200_LIB=$lib MP2V_PARSE_STREAMS=$ps timeout 120 python tools/dev/e2e_once.py 6 $dl 2>&1 | grep "^decode" | tail -3 | tr '\n' ' ')
      echo "variant='$v' download='$dl' parse_streams=$ps : $r"
    done
  done
done
