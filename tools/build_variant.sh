#!/bin/bash
# tools/build_variant.sh NAME "-DX=1 -DY=0": a second libmfk with other macro settings in mfk_sgd.cu (A/B runs via MFK_LIB_PATH)
set -e
name=$1; shift
cs=matrix_factorization_b200/csrc
mkdir -p $cs/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include "$@" -c $cs/mfk_sgd.cu -o $cs/variants/sgd_$name.o
objs=$(ls $cs/build/*.o | grep -v mfk_sgd.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $cs/variants/libmfk_$name.so $cs/variants/sgd_$name.o $objs -cudart static
echo $cs/variants/libmfk_$name.so
