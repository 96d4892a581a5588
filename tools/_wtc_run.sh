mkdir -p gpurun_out
{
for i in 1 2; do timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 2p; done
} > gpurun_out/wtc32.txt 2>&1
cat gpurun_out/wtc32.txt
