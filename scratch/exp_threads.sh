#!/bin/bash
set -e
cd /root/repo
for T in 768 1024; do
  rm -rf scratch/var/t$T && mkdir -p scratch/var/t$T/csrc
  cp lego-loam-sr_b200/csrc/* scratch/var/t$T/csrc/
  sed -i "s/^constexpr int LM_THREADS = 512;/constexpr int LM_THREADS = $T;/; s/^constexpr int LM_CHUNK = LM_THREADS; /constexpr int LM_CHUNK = 512; /" scratch/var/t$T/csrc/lm.cuh
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -Xptxas -v -o scratch/var/lib_t$T.so scratch/var/t$T/csrc/abi.cu 2>&1 | grep -A2 "lm_iter_kernel" | tail -2 &
done
wait
