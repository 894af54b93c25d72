#!/bin/bash
# Timeline-instrumented copy of the library (-DLL_TIMELINE, see csrc/lm.cuh): scratch/var/lib_timeline.so.
# Use: LLB200_LIB=scratch/var/lib_timeline.so python scratch/time_variant.py   (prints the timeline of the last scans)
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
  -DLL_TIMELINE -o scratch/var/lib_timeline.so lego-loam-sr_b200/csrc/abi.cu
ls -la scratch/var/lib_timeline.so
