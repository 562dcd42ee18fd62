#!/bin/bash
# Builds occupancy variants of the IMMA Gibbs kernel on the GPU box and times each (kernel experiments only).
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
OTHERS="api gibbs cd_epoch energy_grad grad_reduce grad_reduce_dmma enumerate rotated fullspace"
for f in $OTHERS; do nvcc $FLAGS -c qucumber_b200/csrc/$f.cu -o /tmp/$f.o & done
VARIANTS="${VARIANTS:-256:2 576:1 640:1}"
for v in $VARIANTS; do
  t=${v%%:*}; m=${v##*:}
  nvcc $FLAGS -Xptxas -v -DQCB_GI_THREADS=$t -DQCB_GI_MIN_CTAS=$m -c qucumber_b200/csrc/gibbs_imma.cu -o /tmp/gibbs_imma.$t.o 2> /tmp/ptxas.$t.log &
done
wait
for v in $VARIANTS; do
  t=${v%%:*}
  objs=""; for f in $OTHERS; do objs="$objs /tmp/$f.o"; done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/lib_$t.so $objs /tmp/gibbs_imma.$t.o
  echo "== threads:min_ctas $v  $(grep -A2 'imma_kernelILi4ELi4ELb0' /tmp/ptxas.$t.log | grep -E -o 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
  for k in 10 100; do
    QCB_LIB_PATH=/tmp/lib_$t.so python bench.py --quick --k $k --steps 3 --warmup 2 2>&1 | grep quick | cut -c1-90
  done
done
