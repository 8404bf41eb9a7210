#!/bin/bash
# build liborb_b200 variants with different LBA kernel tunables: tools/build_variants.sh "name:-DLBA_NT=384 -DLBA_SCH=256" ...
set -e
cd "$(dirname "$0")/.."
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
mkdir -p variants/obj
for f in orb_slam3_modified_b200/csrc/*.cu; do
  b=$(basename $f .cu); [ "$b" = lba ] && continue
  [ variants/obj/$b.o -nt $f ] || nvcc $FLAGS -c $f -o variants/obj/$b.o &
done
wait
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( nvcc $FLAGS $defs -Xptxas -v -c orb_slam3_modified_b200/csrc/lba.cu -o variants/obj/lba_$name.o 2>&1 | grep -A1 "lba_cluster_kernel" | grep spill | sed "s/^/$name: /"
    nvcc -shared -o variants/liborb_$name.so variants/obj/lba_$name.o $(ls variants/obj/*.o | grep -v "/lba_") -lcudart ) &
done
wait
ls variants/*.so
