#!/bin/bash
for v in 2 3; do
  cp tools/_variant_cols$v.so tinyopt_b200/libtinyopt_b200.so
  echo "== cols $v"
  timeout 120 python tools/run_once.py C4 131072 3 2>&1 | tail -1
done
