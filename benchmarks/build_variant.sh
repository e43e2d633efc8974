#!/bin/bash
# build_variant.sh NAME "-DFOO=1 -DBAR=2": libtvk_cuda.so with extra defines -> build_variants/libtvk_NAME.so (TVK_CUDA_LIB selects it)
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2
out=build_variants/obj_$name
mkdir -p $out
C=imagevis3d_b200/csrc
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=default -Iinclude -I$C $defs"
for f in tvk_context.cu tvk_raycast.cu tvk_visibility.cu tvk_builder.cu tvk_quantize.cu tvk_exchange.cu tvk_classic.cu tvk_rebrick.cu host/tvk_host.cpp host/tvk_uvf.cpp host/tvk_rawio.cpp; do
  o=$out/$(basename $f).o
  # only the builder depends on the defines used so far: reuse the stock objects for the rest when they are current
  if [ "$f" != "tvk_builder.cu" ] && [ -z "$ALL" ] && [ -f imagevis3d_b200/build/$(basename $f).o ]; then cp imagevis3d_b200/build/$(basename $f).o $o; else nvcc $FL -c $C/$f -o $o & fi
done
wait
nvcc -shared -o build_variants/libtvk_$name.so $out/*.o -gencode arch=compute_100a,code=sm_100a -lcudart -lz -ldl
rm -rf $out
echo build_variants/libtvk_$name.so
