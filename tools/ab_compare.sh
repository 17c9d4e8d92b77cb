#!/bin/bash
# Same-box A/B of two builds of the library (box-to-box variation is +-3 %, more than most single changes):
#   tools/ab/libA.so, tools/ab/libB.so  ->  alternating section profiles A B A B
# usage (on the GPU box): bash tools/ab_compare.sh [reps]
L=sa-toolkit_b200/csrc/libsatools_hifigan.so
cp $L /tmp/lib_keep.so
for i in $(seq 1 ${1:-2}); do
  for v in A B; do
    cp tools/ab/lib$v.so $L
    echo "== $v"; python tools/section_profile.py 64 5 x 2>&1 | tail -2 | sed "s/launches:.* 48:/48:/" | cut -c1-200
  done
done
cp /tmp/lib_keep.so $L
