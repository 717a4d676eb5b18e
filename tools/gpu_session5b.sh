#!/bin/bash
# Session 5, call B: grid size of the grid-stride particle kernels (CTAs per SM) -> gpurun_out/s5b_grid.log
mkdir -p gpurun_out
out=gpurun_out/s5b_grid.log
: > $out
for k in 64 96 128 256 512 2048; do
    echo "## gather and deposit grids: $k CTAs per SM" >> $out
    SCB_GATHER_PER_SM=$k SCB_DEPOSIT_PER_SM=$k timeout 120 python bench.py --stages-only >> $out 2>> gpurun_out/s5b_grid.err
done
cat $out
