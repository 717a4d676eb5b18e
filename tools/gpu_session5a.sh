#!/bin/bash
# Session 5, call A: A/B of the gather's occupancy (register cap via __launch_bounds__ min blocks) and of the
# grid size of the grid-stride particle kernels.  One line of per-stage times per variant -> gpurun_out/s5a_ab.log
mkdir -p gpurun_out
L=spacecharge.jl_b200/lib
out=gpurun_out/s5a_ab.log
: > $out
run() {  # label, env...
    label=$1; shift
    echo "## $label" >> $out
    env "$@" timeout 120 python bench.py --stages-only >> $out 2>> gpurun_out/s5a_ab.err
}
run "default" X=1
run "g6 lib" SCB_LIB=$PWD/$L/libspacecharge_b200_g6.so
run "g8 lib" SCB_LIB=$PWD/$L/libspacecharge_b200_g8.so
run "g6 lib, gather grid 6/SM" SCB_LIB=$PWD/$L/libspacecharge_b200_g6.so SCB_GATHER_PER_SM=6
run "g8 lib, gather grid 8/SM" SCB_LIB=$PWD/$L/libspacecharge_b200_g8.so SCB_GATHER_PER_SM=8
run "g6 lib, gather grid 24/SM" SCB_LIB=$PWD/$L/libspacecharge_b200_g6.so SCB_GATHER_PER_SM=24
run "default lib, gather grid 5/SM, deposit grid 4/SM" SCB_GATHER_PER_SM=5 SCB_DEPOSIT_PER_SM=4
run "default lib, gather grid 20/SM, deposit grid 16/SM" SCB_GATHER_PER_SM=20 SCB_DEPOSIT_PER_SM=16
echo "## default f32" >> $out; timeout 120 python bench.py --stages-only --dtype f32 >> $out 2>> gpurun_out/s5a_ab.err
echo "## g6 lib f32" >> $out; SCB_LIB=$PWD/$L/libspacecharge_b200_g6.so timeout 120 python bench.py --stages-only --dtype f32 >> $out 2>> gpurun_out/s5a_ab.err
echo "## g8 lib f32" >> $out; SCB_LIB=$PWD/$L/libspacecharge_b200_g8.so timeout 120 python bench.py --stages-only --dtype f32 >> $out 2>> gpurun_out/s5a_ab.err
cat $out
