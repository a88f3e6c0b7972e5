#!/bin/bash
# Build kernel variants (extra nvcc -D flags) next to the in-tree library and time them back to back on one box.
#   tools/variants.sh build "name1:-DFLAG=1" "name2:-DOTHER=2" ...     (here, no GPU needed)
#   tools/variants.sh run name1 name2 ...                               (on the GPU box; always includes the default)
set -e
cd "$(dirname "$0")/.."
LIBDIR=pytorch-simclr_b200/lib
mode=$1; shift
if [ "$mode" = build ]; then
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
      $flags -o $LIBDIR/libsimclr_v_$name.so pytorch-simclr_b200/csrc/capi.cu &
  done
  wait
  ls -la $LIBDIR/*.so
else
  for rep in 1 2; do
    for v in b200 "$@"; do
      lib=$PWD/$LIBDIR/libsimclr_v_$v.so; [ "$v" = b200 ] && lib=$PWD/$LIBDIR/libsimclr_b200.so
      echo -n "== $v: "
      SIMCLR_B200_LIB=$lib python tools/ab_step.py ${AB_ARGS:-}
    done
  done
fi
