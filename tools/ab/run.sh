#!/bin/bash
# same-box A/B of two builds of the library: tools/ab/lib_old.so vs tools/ab/lib_new.so
for r in 1 2; do
for v in old new; do
  cp tools/ab/lib_$v.so massive_marl_benchmark_b200/libmmb_b200.so
  tools/sweep_ten_ant.sh "MMB_AB=$v"
done
done
