#!/bin/bash
out=gpurun_out/r02q_sweep.log; : > $out
run() { echo "== $*" >> "$out"; env "$@" python tools/prof_bader.py 1024 8 0 2 2>&1 | tail -1 >> "$out"; }
run C2G_X=0
for k in 5 6 10; do run C2G_W3_K=$k; done
run C2G_W3_K=6 C2G_W3_IDLE=64
run C2G_W3_K=8 C2G_W3_IDLE=64
for k in 8 12 24; do run C2G_W3_KC=$k; done
run C2G_W3_KC=12 C2G_W3_IDLEC=96
run C2G_W3_KC=16 C2G_W3_IDLEC=160
