#!/bin/bash
# One-command pin of the CPU oracle against real critic2 (needs a critic2 binary, i.e. a Fortran compiler somewhere):
#   tools/pin_with_critic2.sh [/path/to/critic2] [workdir]
# Writes the seeded parity cases as cube files + .cri inputs, runs critic2 on them (BADER and YT with WCUBE), and
# compares critic2's basins with the oracle's point by point (tools/pin_cases.py).  Exit status 0 = pinned.
set -e
C2=${1:-critic2}
DIR=${2:-pin_run}
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v "$C2" >/dev/null || { echo "critic2 binary '$C2' not found (build the reference: cmake + gfortran)"; exit 2; }
python "$HERE/tools/pin_cases.py" write "$DIR"
( cd "$DIR" && for f in *.cri; do echo "critic2 $f"; "$C2" "$f" "${f%.cri}.cro"; done )
python "$HERE/tools/pin_cases.py" compare "$DIR"
