#!/bin/bash
# Session 5, final call: the whole GPU suite, smoke(), the bench lines (f64 with baselines, f32), the ncu launch list
# of the bench command and --set full captures of the dominant kernels in both precisions.  Outputs: gpurun_out/s5_*.
mkdir -p gpurun_out
date +%s > gpurun_out/s5_t0
timeout 500 python -m pytest tests -q -m gpu -x > gpurun_out/s5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s5_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/s5_smoke.log
timeout 420 python bench.py > gpurun_out/s5_bench_f64.json 2> gpurun_out/s5_bench_f64.err
echo "bench rc=$?" >> gpurun_out/s5_bench_f64.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-records > gpurun_out/s5_ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_interpolate_pair2_f64|k_deposit_tiles|k_z_eo' \
    --launch-skip 3 -c 3 -o gpurun_out/s5_top3_f64 python tools/profile_step.py large f64 2 > gpurun_out/s5_ncu_full_f64.log 2>&1
timeout 150 python bench.py --dtype f32 --no-cpu-baseline --no-gpu-baseline > gpurun_out/s5_bench_f32.json 2> gpurun_out/s5_bench_f32.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_interpolate_packed_f32|k_deposit_tiles|k_z_eo' \
    --launch-skip 3 -c 3 -o gpurun_out/s5_top3_f32 python tools/profile_step.py large f32 2 > gpurun_out/s5_ncu_full_f32.log 2>&1
date +%s > gpurun_out/s5_t1
tail -3 gpurun_out/s5_pytest.log; tail -2 gpurun_out/s5_smoke.log; cut -c1-400 gpurun_out/s5_bench_f64.json
