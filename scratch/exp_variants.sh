#!/bin/bash
# Build experiment variants of the CUDA library from patched COPIES of the sources (product sources untouched).
set -e
cd /root/repo
rm -rf scratch/var && mkdir -p scratch/var
for v in base nosolve nofit; do
  mkdir -p scratch/var/$v/csrc scratch/var/include
  cp lego-loam-sr_b200/csrc/* scratch/var/$v/csrc/
  cp include/legoloam_b200.h scratch/var/include/
  sed -i 's#../../include/legoloam_b200.h#../../include/legoloam_b200.h#' scratch/var/$v/csrc/abi.cu
done
# nosolve: the last block only advances the iteration counter
python - <<'PY'
p='scratch/var/nosolve/csrc/lm.cuh'
s=open(p).read()
s=s.replace("  if (threadIdx.x == 0) lm_solve(s_red, prm, S, stage_mode);", "  if (threadIdx.x == 0) { S->iter += 1; S->sel_num = (int)s_red[28]; if (S->iter >= prm.iter_count_thres) { S->done = 1; S->iter_num = S->iter; } }")
open(p,'w').write(s)
p='scratch/var/nofit/csrc/lm.cuh'
s=open(p).read()
s=s.replace("          keep = m == 0 ? corner_fit(nbx, nby, nbz, slot[0], slot[1], slot[2], A.prm.fit, coeff)\n                        : surf_fit(nbx, nby, nbz, slot[0], slot[1], slot[2], A.prm.fit, coeff);",
            "          keep = true; coeff[0] = nbx[0] * 1e-3f; coeff[1] = nby[1] * 1e-3f; coeff[2] = nbz[2] * 1e-3f; coeff[3] = 0.01f;")
open(p,'w').write(s)
PY
for v in base nosolve nofit; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static -o scratch/var/lib_$v.so scratch/var/$v/csrc/abi.cu &
done
wait
ls -la scratch/var/*.so
