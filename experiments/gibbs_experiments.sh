#!/bin/bash
# Builds ablated variants of the Gibbs kernel on the GPU box and times each (kernel experiments only).
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
for f in api energy_grad grad_reduce enumerate rotated; do nvcc $FLAGS -c qucumber_b200/csrc/$f.cu -o /tmp/$f.o & done
i=0
for v in "HALFPHILOX" "BASE"; do
  nvcc $FLAGS -DQCB_EXP_$v -c qucumber_b200/csrc/gibbs.cu -o /tmp/gibbs.$i.o &
  i=$((i+1))
done
wait
i=0
for v in "HALFPHILOX" "BASE"; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/lib_$i.so /tmp/api.o /tmp/energy_grad.o /tmp/grad_reduce.o /tmp/enumerate.o /tmp/rotated.o /tmp/gibbs.$i.o
  echo "== $v"
  QCB_LIB_PATH=/tmp/lib_$i.so python bench.py --quick --k 100 --steps 3 --warmup 2 2>&1 | grep quick | cut -c1-140
  i=$((i+1))
done
