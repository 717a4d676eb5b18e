#!/bin/bash
# Session 5, call H: end-to-end (host-buffer) step with and without NUMA-local placement of the pinned buffers
mkdir -p gpurun_out
out=gpurun_out/s5h_numa.log
{ nvidia-smi topo -m 2>&1 | head -14; lscpu | grep -i -E "numa|socket|^CPU\(s\)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; } > $out 2>&1
for bind in 1 0; do
    echo "## SCB_NUMA_BIND=$bind" >> $out
    SCB_NUMA_BIND=$bind timeout 150 python bench.py --no-cpu-baseline --no-gpu-baseline --no-records 2>> gpurun_out/s5h_numa.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print(json.dumps({'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']}))" >> $out
done
cat $out
