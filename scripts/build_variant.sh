#!/bin/bash
# Builds a variant of libqhbm_b200.so for A/B timing (QHBM_B200_LIB=<path> selects it at run time).
# usage: scripts/build_variant.sh <out.so> [-DMACRO ...]
set -e
cd "$(dirname "$0")/.."
OUT=$1; shift
OBJ=/tmp/qhbm_obj; mkdir -p $OBJ
C=qhbm-library_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in ebm measure comm; do
  if [ ! -f $OBJ/$f.o ] || [ $C/$f.cu -nt $OBJ/$f.o ]; then nvcc $FLAGS -c $C/$f.cu -o $OBJ/$f.o & fi
done
nvcc $FLAGS -c $C/plan.cpp -o $OBJ/plan_$$.o &
nvcc $FLAGS "$@" -c $C/sim_lean.cu -o $OBJ/sim_lean_$$.o &
nvcc $FLAGS "$@" -c $C/sim.cu -o $OBJ/sim_$$.o
wait
nvcc -shared -o $OUT $OBJ/sim_$$.o $OBJ/sim_lean_$$.o $OBJ/plan_$$.o $OBJ/ebm.o $OBJ/measure.o $OBJ/comm.o -ldl
rm -f $OBJ/sim_$$.o $OBJ/sim_lean_$$.o $OBJ/plan_$$.o
echo built $OUT
