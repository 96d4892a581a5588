mkdir -p gpurun_out
{
for rep in 1 2; do
for v in "" base; do
  if [ -n "$v" ]; then export TOB200_LIB_OVERRIDE=$PWD/tinyopt_b200/libtinyopt_b200_$v.so; else unset TOB200_LIB_OVERRIDE; fi
  echo "== variant '$v'"; timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 2p
done
done
} > gpurun_out/wtc16.txt 2>&1
cat gpurun_out/wtc16.txt
