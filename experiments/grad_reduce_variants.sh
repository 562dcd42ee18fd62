#!/bin/bash
# Builds warps-per-CTA variants of the tensor-core gradient kernel on the GPU box and times each (kernel experiments only).
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
OTHERS="api gibbs gibbs_imma cd_epoch energy_grad grad_reduce enumerate rotated fullspace"
for f in $OTHERS; do nvcc $FLAGS -c qucumber_b200/csrc/$f.cu -o /tmp/$f.o & done
VARIANTS="${VARIANTS:-8:2 4:4 4:3}"
for v in $VARIANTS; do
  w=${v%%:*}; m=${v##*:}
  nvcc $FLAGS -DQCB_GD_WARPS=$w -DQCB_GD_MIN_CTAS=$m -c qucumber_b200/csrc/grad_reduce_dmma.cu -o /tmp/gd.$w.$m.o &
done
wait
for v in $VARIANTS; do
  w=${v%%:*}; m=${v##*:}
  objs=""; for f in $OTHERS; do objs="$objs /tmp/$f.o"; done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/lib_gd_$w_$m.so $objs /tmp/gd.$w.$m.o
  echo "== warps:min_ctas $v"
  QCB_LIB_PATH=/tmp/lib_gd_$w_$m.so python tools/grad_time.py 2>&1 | grep -E "grad|energy"
  QCB_LIB_PATH=/tmp/lib_gd_$w_$m.so python -m pytest tests -m gpu -x -q -k "gradient or energy or rotated" 2>&1 | tail -1
done
