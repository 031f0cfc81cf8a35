#!/bin/bash
# Dev-time: device timeline (MP2V_TRACE) + host stage profile of one end-to-end decode, with and without the frame download
for dl in "--download" ""; do
  echo "=== $dl"
  MP2V_TRACE=1 MP2V_PROFILE=1 timeout 200 python tools/dev/e2e_once.py 3 $dl 2>&1 | grep -E "^decode|mp2v trace\] (launch|parse|resident)|mp2v profile" | tail -60
done
