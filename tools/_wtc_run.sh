mkdir -p gpurun_out
{
timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 1,3p
} > gpurun_out/wtc20.txt 2>&1
cat gpurun_out/wtc20.txt
