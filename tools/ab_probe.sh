#!/bin/bash
# A/B of kernel builds on the probe workload: tools/ab_probe.sh <ticks> <reps> <tree> lib1.so lib2.so ...
# (KIND=fish|flows in the environment selects the sender family; default rat)
ticks=$1; reps=$2; tree=$3; shift 3
for round in 1 2; do
  for lib in "$@"; do
    echo "== $(basename $lib)"
    REMY_LIB=$lib python tools/gpu_probe.py $ticks $reps $tree 32 ${KIND:-rat} 2>&1 | tail -2
  done
done
