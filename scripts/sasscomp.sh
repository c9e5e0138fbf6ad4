#!/bin/bash
# compile one .cu for sm_100a and print registers / spills + the static loop mix (scripts/sass_mix.py)
# usage: scripts/sasscomp.sh file.cu [kernel-substring] [extra nvcc flags...]
src=$1; shift; pat=${1:-}; shift || true
here=$(cd "$(dirname "$0")/.." && pwd)
o=/tmp/$(basename "$src" .cu).o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v --expt-relaxed-constexpr \
  -I "$here/fftoptionpricing_b200/csrc" "$@" -c "$src" -o "$o" 2>&1 | grep -E "Used|spill|error" | grep -v "Used [23][0-9] reg"
cuobjdump -sass "$o" > "${o%.o}.sass"
python "$here/scripts/sass_mix.py" "$pat" < "${o%.o}.sass" | cut -c1-330
