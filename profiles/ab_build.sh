#!/bin/bash
# A/B builds behind the *_ab_*.txt logs: recompiles ONE source with extra -D switches and links it with the objects of the main build;
# run a script against it with GMD_AB_LIB=build_tmp/libgmd_NAME.so (the library honours the variable, see gm_diffusion_b200/_lib.py).
# usage: profiles/ab_build.sh NAME file.cu "-DFOO=1 ..."  -> build_tmp/libgmd_NAME.so (other objects from the main build)
set -e
NAME=$1; SRC=$2; DEFS=$3
C=gm_diffusion_b200/_C
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS -c gm_diffusion_b200/csrc/$SRC -o build_tmp/$NAME.$SRC.o
OBJS=""
for f in api.cu hdr.cu sched.cu gemm.cu norm.cu attn.cu; do
  if [ "$f" == "$SRC" ]; then OBJS="$OBJS build_tmp/$NAME.$SRC.o"; else OBJS="$OBJS $C/$f.o"; fi
done
nvcc -shared -o build_tmp/libgmd_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart
echo build_tmp/libgmd_$NAME.so
