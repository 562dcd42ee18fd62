#!/bin/bash
# How many of an update's four Philox calls the tcgen05 Gibbs kernels draw BEFORE the MMA-completion wait (the wait is where
# 19 % of all warp samples sat: the join of the tile's 8 warps plus the MMA latency).  Builds one library per setting here (no
# GPU needed), times them on the box:  bash experiments/gibbs_umma_prefetch.sh build ; gpurun -- bash experiments/gibbs_umma_prefetch.sh time
# Result (round 2, 2^20 chains, N = nh = 64, k = 10 / 100, same box): 0: 1.554 / 13.36 ms, 1: 1.493 / 12.84, 2: 1.530 / 13.14,
# 3: 1.518, 4: 1.496 / 12.85 -> GU_PREFETCH = 1.  Wide kernel (128 x 128 | 128 x 64): 0: 3.82 | 2.92 ms, 1: 3.82 | 3.11, 2: 3.81 | 3.15
# (spills at 64 registers) -> GW_PREFETCH = 0.
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p experiments/_variants
  objs=""; for f in qucumber_b200/build/*.o; do case $f in *gibbs_umma.o) ;; *) objs="$objs $f";; esac; done
  for pf in 0 1 2 3 4; do
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DGU_PREFETCH=$pf -c qucumber_b200/csrc/gibbs_umma.cu -o /tmp/gu_$pf.o &&
      nvcc -shared -gencode arch=compute_100a,code=sm_100a -o experiments/_variants/lib_pf$pf.so $objs /tmp/gu_$pf.o &
  done
  wait
else
  for pf in 0 1 2 3 4; do for k in 10 100; do
    echo -n "prefetch=$pf k=$k "
    QCB_LIB_PATH=$PWD/experiments/_variants/lib_pf$pf.so python bench.py --quick --k $k --steps 10 --warmup 3 2>/dev/null | cut -c1-75
  done; done
fi
