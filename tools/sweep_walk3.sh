#!/bin/bash
# Sweep of the walker kernels on the 1024^3 headline workload (tools/prof_bader.py): per-phase times per variant.
# usage: tools/sweep_walk3.sh out.log
out=${1:-gpurun_out/sweep_walk3.log}
: > "$out"
run() { echo "== $*" >> "$out"; env "$@" python tools/prof_bader.py 1024 8 0 2 2>&1 | tail -2 >> "$out"; }
run C2G_X=0
for k in 8 12 16 24; do for idle in 16 32 48; do run C2G_WALK3=1 C2G_W3_K=$k C2G_W3_IDLE=$idle; done; done
for k in 8 16 32; do for idle in 96 128 192 224; do run C2G_WALK3=14 C2G_W3_KC=$k C2G_W3_IDLEC=$idle; done; done
