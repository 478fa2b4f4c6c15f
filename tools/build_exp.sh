#!/bin/bash
# builds an experimental copy of libdctc.so with extra -D flags: tools/build_exp.sh <name> <flags...>  -> tools/exp/libdctc_<name>.so
# (timing experiments only; select it with DCTC_LIB=... in tools/bench_retarget.py / tools/time_*.py)
set -e
NAME=$1; shift
D=$(dirname $0)/..
mkdir -p $D/tools/exp/obj_$NAME
for f in $D/dct_carver_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDCTC_BUILD "$@" -c $f -o $D/tools/exp/obj_$NAME/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $D/tools/exp/libdctc_$NAME.so $D/tools/exp/obj_$NAME/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $D/tools/exp/obj_$NAME
echo built tools/exp/libdctc_$NAME.so
