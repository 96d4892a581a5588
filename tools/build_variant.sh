#!/bin/bash
# A/B variants of the library: recompile the named translation units with extra flags and link them with the
# other (unchanged) objects into tinyopt_b200/libtinyopt_b200_<name>.so; select with TOB200_LIB_OVERRIDE.
#   tools/build_variant.sh ctas3 "-DTOB200_WPP_CTAS=3" wpp_inst_f32_d
set -e
name=$1; flags=$2; shift 2
cd "$(dirname "$0")/../tinyopt_b200/csrc"
mkdir -p build/var_$name
objs=""
for o in build/*.o; do
  b=$(basename $o .o); use=$o
  for tu in "$@"; do
    if [ "$b" == "$tu" ]; then
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
        -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $flags -c $tu.cu -o build/var_$name/$tu.o > build/var_$name/$tu.ptxas.log 2>&1 \
        || (cat build/var_$name/$tu.ptxas.log; exit 1)
      use=build/var_$name/$tu.o
    fi
  done
  objs="$objs $use"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtinyopt_b200_$name.so $objs
echo built tinyopt_b200/libtinyopt_b200_$name.so
