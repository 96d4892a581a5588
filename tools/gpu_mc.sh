#!/bin/bash
# A/B of the cluster-multicast JtJ kernel (TOB200_LG_MC) over raw-ring depths
out=gpurun_out; mkdir -p $out
{
for raw in 2 3 4; do for mc in 0 1; do
echo "== C5 1184 problems: MC=$mc RAW=$raw"
TOB200_LG_MC=$mc TOB200_LG_RAW_STAGES=$raw timeout 120 python tools/run_once.py C5 1184 3 | tail -1
done; done
} 2>&1 | tee $out/mc_ab.txt
