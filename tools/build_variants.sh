#!/bin/bash
# Build library variants for a same-box A/B: tools/lib_<name>.so from extra nvcc defines.
#   tools/build_variants.sh name1 "-DX=1" name2 "-DY=2" ...
set -e
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  PST_NVCC_DEFS="$defs" python -m prosstt_b200.build --force > /tmp/build_$name.log 2>&1
  cp prosstt_b200/libprosstt_b200.so tools/lib_$name.so
  grep -A2 "draw_counts_kernelILi10ELb1ELb0" /tmp/build_$name.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $name"
done
python -m prosstt_b200.build --force > /dev/null 2>&1
