#!/bin/bash
# Builds kernel variants side by side: our_first_climate_model_b200/lib_v<N>/librcm_b200.so with -DRCM_VARIANT=<N>
# (selected at run time with RCM_B200_LIB=<path>).  Usage: tools/build_variants.sh 1 2 ...
cd "$(dirname "$0")/../our_first_climate_model_b200/csrc" || exit 1
for v in "$@"; do
  ( make OUT=../lib_v$v NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off -diag-suppress 128 -DRCM_VARIANT=$v" ../lib_v$v/librcm_b200.so > /tmp/build_v$v.log 2>&1 || echo "variant $v failed" ) &
done
wait
ls -la ../lib_v*/librcm_b200.so
