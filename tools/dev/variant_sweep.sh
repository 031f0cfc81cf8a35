#!/bin/bash
# Dev-time: kernel-tuning variants (tiny_mp2v_dec_b200/_lib/variants/*.so) on the resident / end-to-end decode
for v in "" $(ls tiny_mp2v_dec_b200/_lib/variants/*.so 2>/dev/null); do
  if [ -n "$v" ]; then export MP2V_B200_LIB=$PWD/$v; else unset MP2V_B200_LIB; fi
  r=$(timeout 120 python tools/dev/e2e_once.py 6 --resident 2>&1 | grep -E "^decode|^resident best" | tail -3 | tr '\n' ' ')
  echo "variant='$(basename "$v")' : $r"
done
