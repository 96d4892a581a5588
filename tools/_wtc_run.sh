mkdir -p gpurun_out
{
for v in "" t6; do
  if [ -n "$v" ]; then export TOB200_LIB_OVERRIDE=$PWD/tinyopt_b200/libtinyopt_b200_$v.so; else unset TOB200_LIB_OVERRIDE; fi
  echo "== variant '$v'"; timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 2,3p
done
} > gpurun_out/wtc24.txt 2>&1
cat gpurun_out/wtc24.txt
