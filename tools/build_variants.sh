#!/bin/bash
# tools/build_variants.sh "<name> <-D flags>" ...  -> gpu-accel-ofdm-ls-mrc_b200/variants/lib_<name>.so
# developer helper: build tuning variants of the library to compare in one GPU session
cd "$(dirname "$0")/../gpu-accel-ofdm-ls-mrc_b200"
mkdir -p variants
for spec in "$@"; do
  name=${spec%% *}; flags=${spec#* }
  [ "$name" = "$flags" ] && flags=""
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC $flags \
       -o variants/lib_$name.so csrc/lsmrc_capi.cu &
done
wait
ls -la variants
