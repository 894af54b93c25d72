#!/bin/bash
# A/B builds: copy csrc/ to scratch/var/<name>/csrc, apply an optional python patch (run with cwd = the copy's root, so it
# edits csrc/<file>), build scratch/var/lib_<name>.so. Use: LLB200_LIB=scratch/var/lib_<name>.so python scratch/split_iters.py
#   bash scratch/build_variant.sh <name> [patch.py|-] [extra nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; patch=${2:--}; shift; shift || true
d=scratch/var/$name
rm -rf $d && mkdir -p $d/csrc scratch/var/include
cp lego-loam-sr_b200/csrc/* $d/csrc/
cp include/legoloam_b200.h scratch/var/include/
if [ "$patch" != "-" ]; then (cd $d && python "$OLDPWD/$patch"); fi
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
  "$@" -o scratch/var/lib_$name.so $d/csrc/abi.cu
ls -la scratch/var/lib_$name.so
