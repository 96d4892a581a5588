mkdir -p gpurun_out
{
for w in 1 2 4 8 16; do ./tools/cuda/_ldlt_bench 50 $w 1; done
./tools/cuda/_ldlt_bench 50 8 148
./tools/cuda/_ldlt_bench 32 1 1
./tools/cuda/_ldlt_bench 20 1 1
} > gpurun_out/ldlt_bench.txt 2>&1
cat gpurun_out/ldlt_bench.txt
