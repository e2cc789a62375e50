#!/bin/bash
# Builds kernel variants side by side: our_first_climate_model_b200/lib_<name>/librcm_b200.so with extra nvcc flags
# (selected at run time with RCM_B200_LIB=<path>).  Usage: tools/build_variants.sh name1:"-DFOO=1" name2:"-DBAR=2 -DBAZ" ...
cd "$(dirname "$0")/../our_first_climate_model_b200/csrc" || exit 1
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( make OUT=../lib_$name NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off -diag-suppress 128 $flags" ../lib_$name/librcm_b200.so > /tmp/build_$name.log 2>&1 || echo "variant $name failed: /tmp/build_$name.log" ) &
done
wait
ls -la ../lib_*/librcm_b200.so
