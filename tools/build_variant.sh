#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>  -> getfem_b200/libgfgpu_<name>.so (experiments; GFGPU_LIB selects it)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../getfem_b200/csrc"
mkdir -p _obj/var_$NAME
for f in *.cu; do
  o=_obj/var_$NAME/${f%.cu}.o
  if [ "$f" == "recompute_tiles.cu" ] || [ ! -f $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $o &
  fi
done
wait
nvcc -shared -o ../libgfgpu_$NAME.so _obj/var_$NAME/*.o
echo built ../libgfgpu_$NAME.so
