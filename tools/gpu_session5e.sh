#!/bin/bash
# Session 5, call E: gather without the run-time slab filter (FL = false instantiation), six resident CTAs (default)
# against the compiler's own register allocation (g0 variant) -> gpurun_out/s5e_ab.log
mkdir -p gpurun_out
L=$PWD/spacecharge.jl_b200/lib
out=gpurun_out/s5e_ab.log
: > $out
for dt in f64 f32; do
    echo "## default lib (SCB_GATHER_MINB=6), $dt" >> $out
    timeout 120 python bench.py --stages-only --dtype $dt >> $out 2>> gpurun_out/s5e_ab.err
    echo "## g0 lib (no register cap), $dt" >> $out
    SCB_LIB=$L/libspacecharge_b200_g0.so timeout 120 python bench.py --stages-only --dtype $dt >> $out 2>> gpurun_out/s5e_ab.err
done
cat $out
