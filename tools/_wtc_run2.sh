mkdir -p gpurun_out
{
for cfg in "65536 100 13" "65536 200 20" "65536 200 27" "65536 300 28" "65536 300 32" "65536 400 40" "32768 1000 50" "32768 200 50" "65536 100 50" "32768 500 55"; do
  set -- $cfg
  echo "== B=$1 m=$2 n=$3"; timeout 120 python tools/wtc_check.py $1 $2 $3 2 2>&1 | sed -n 1,3p
done
} > gpurun_out/wtc_shapes.txt 2>&1
cat gpurun_out/wtc_shapes.txt
